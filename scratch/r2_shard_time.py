"""One rank's stage 2 of an N-GPU proof (its share of the five multiexps), on ONE GPU, under the current environment.
python scratch/r2_shard_time.py tag [world] [rank]   (ZA_DEBUG_TIMELINE=1 prints the per-multiexp timeline)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, za_b200
from za_b200 import synthetic
tag = sys.argv[1] if len(sys.argv) > 1 else "x"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = za_b200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
log_m = 20
nc = (1 << log_m) - 2
ni, na, ptr, var, coeff, inputs, aux = synthetic.mul_chain(nc, x0=5); counts = synthetic.pk_counts_for_mul_chain(nc)
circ = za_b200.Circuit(ctx, ni, na, ptr, var, coeff)
pk = za_b200.Parameters.synthetic(ctx, counts["ic"], counts["h"], counts["l"], counts["a"], counts["b_g1"], counts["b_g2"])
rho = 0.115
w0 = max(0.05, min(1.0, (1.0 - rho * (world - 1)) / (1.0 + rho))) if world > 1 else 1.0
plan = os.environ.get("ZA_SHARD_PLAN")          # "1": the device `rank` of za_prover_plan(world) instead of the weighted slices
if plan:
    lo, hi = za_b200.prover_plan(circ, world)[rank]
    pk.partition_ranges(circ, lo, hi)
    print("plan ranges", lo, hi, flush=True)
else:
    pk.partition(circ, rank, world, w0)
wit = torch.from_numpy(np.concatenate([inputs, aux])).cuda()
h = torch.from_numpy(synthetic.random_scalars(1 << log_m, 5)).cuda()
torch.cuda.synchronize()
ctx.profile(True)
ts = []
for i in range(8):
    torch.cuda.synchronize(); t = time.perf_counter()
    za_b200.prove_msm_enqueue(ctx, pk, circ, wit.data_ptr(), h.data_ptr(), rank, world, za_b200.MSM_WITNESS)
    za_b200.prove_msm_enqueue(ctx, pk, circ, wit.data_ptr(), h.data_ptr(), rank, world, za_b200.MSM_H)
    za_b200.prove_msm_collect(ctx)
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    if i == 2: ctx.profile_read()
prof = ctx.profile_read()
print("[%s] rank %d/%d stage 2 ms:" % (tag, rank, world), " ".join("%.2f" % x for x in ts), {k: round(v["ms"] / 5, 3) for k, v in prof.items() if v["ms"] > 0}, flush=True)
