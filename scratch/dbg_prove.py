import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, za_b200
from za_b200 import synthetic
log_m = 20
nc = (1 << log_m) - 2
cs = synthetic.mul_chain(nc, x0=5); counts = synthetic.pk_counts_for_mul_chain(nc)
ni, na, ptr, var, coeff, inputs, aux = cs
ctx = za_b200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
circ = za_b200.Circuit(ctx, ni, na, ptr, var, coeff)
pk = za_b200.Parameters.synthetic(ctx, counts["ic"], counts["h"], counts["l"], counts["a"], counts["b_g1"], counts["b_g2"])
wit = torch.from_numpy(np.concatenate([inputs, aux])).cuda()
torch.cuda.synchronize()
def run(tag, n, prof):
    ctx.profile(prof); ctx.profile_read()
    ts = []
    for i in range(n):
        t = time.perf_counter()
        za_b200.create_proof_device(ctx, pk, circ, wit.data_ptr(), 11, 13)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    p = ctx.profile_read()
    print(tag, ["%.1f" % x for x in ts], {k: round(v["ms"] / n, 2) for k, v in p.items() if v["ms"] > 0}, flush=True)
run("prof off", 5, False)
run("prof on ", 5, True)
run("prof off", 5, False)
run("prof on ", 5, True)
