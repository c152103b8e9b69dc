"""Whole-proof timing at 2^20 under the current env.  python scratch/r2_proof_time.py tag"""
import sys, os, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, za_b200
from za_b200 import synthetic
tag = sys.argv[1] if len(sys.argv) > 1 else "x"
ctx = za_b200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
log_m = 20
nc = (1 << log_m) - 2
ni, na, ptr, var, coeff, inputs, aux = synthetic.mul_chain(nc, x0=5); counts = synthetic.pk_counts_for_mul_chain(nc)
circ = za_b200.Circuit(ctx, ni, na, ptr, var, coeff)
pk = za_b200.Parameters.synthetic(ctx, counts["ic"], counts["h"], counts["l"], counts["a"], counts["b_g1"], counts["b_g2"])
wit = torch.from_numpy(np.concatenate([inputs, aux])).cuda()
torch.cuda.synchronize()
for rep in range(2):
    ts = []
    for i in range(6):
        torch.cuda.synchronize(); t = time.perf_counter()
        proof = za_b200.create_proof_device(ctx, pk, circ, wit.data_ptr(), 11, 13)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
print("[%s] proof 2^20 ms:" % tag, " ".join("%.2f" % x for x in ts), hashlib.sha256(proof).hexdigest()[:8], flush=True)
