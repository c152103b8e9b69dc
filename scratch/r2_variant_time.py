"""Round 2: stand-alone G1 / G2 multiexp (2^20 points, fixed-base table) and whole-proof timings of the library
selected with ZA_B200_SO.  python scratch/r2_variant_time.py [tag]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, za_b200
from za_b200 import synthetic
tag = sys.argv[1] if len(sys.argv) > 1 else "main"
ctx = za_b200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
n = 1 << 20
for group in (1, 2):
    bases = za_b200.Bases.generate(ctx, group, n, 1)
    bases.precompute()
    sc = torch.from_numpy(synthetic.random_scalars(n, 2)).cuda()
    for _ in range(3): res = za_b200.multiexp_device(ctx, bases, sc.data_ptr(), n)
    ctx.profile(True); ctx.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record(st)
    for _ in range(reps): res = za_b200.multiexp_device(ctx, bases, sc.data_ptr(), n)
    e1.record(st); torch.cuda.synchronize()
    p = ctx.profile_read(); ctx.profile(False)
    print("[%s] G%d msm 2^20 total %.3f ms  " % (tag, group, e0.elapsed_time(e1) / reps) + " ".join("%s %.3f" % (k, v["ms"] / reps) for k, v in p.items() if v["ms"] > 0), flush=True)
    del bases, sc
log_m = 20
nc = (1 << log_m) - 2
cs = synthetic.mul_chain(nc, x0=5); counts = synthetic.pk_counts_for_mul_chain(nc)
ni, na, ptr, var, coeff, inputs, aux = cs
circ = za_b200.Circuit(ctx, ni, na, ptr, var, coeff)
pk = za_b200.Parameters.synthetic(ctx, counts["ic"], counts["h"], counts["l"], counts["a"], counts["b_g1"], counts["b_g2"])
wit = torch.from_numpy(np.concatenate([inputs, aux])).cuda()
torch.cuda.synchronize()
proof = None
for rep in range(2):
    ts = []
    for i in range(6):
        torch.cuda.synchronize()
        t = time.perf_counter()
        proof = za_b200.create_proof_device(ctx, pk, circ, wit.data_ptr(), 11, 13)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    print("[%s] proof 2^20 ms:" % tag, " ".join("%.2f" % x for x in ts), flush=True)
import hashlib
print("[%s] proof sha256 %s" % (tag, hashlib.sha256(proof).hexdigest()[:16]), flush=True)
