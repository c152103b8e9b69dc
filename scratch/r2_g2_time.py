"""G2 multiexp 2^20 (table) timing under the current env.  python scratch/r2_g2_time.py tag"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, za_b200
from za_b200 import synthetic
tag = sys.argv[1] if len(sys.argv) > 1 else "x"
ctx = za_b200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
for logn in (20, 17):
    n = 1 << logn
    bases = za_b200.Bases.generate(ctx, 2, n, 1); bases.precompute()
    sc = torch.from_numpy(synthetic.random_scalars(n, 2)).cuda()
    for _ in range(3): res = za_b200.multiexp_device(ctx, bases, sc.data_ptr(), n)
    ctx.profile(True); ctx.profile_read()
    for _ in range(5): res = za_b200.multiexp_device(ctx, bases, sc.data_ptr(), n)
    p = ctx.profile_read(); ctx.profile(False)
    import hashlib
    print("[%s] G2 2^%d: " % (tag, logn) + " ".join("%s %.3f" % (k, v["ms"] / 5) for k, v in p.items() if v["ms"] > 0), hashlib.sha256(res).hexdigest()[:8], flush=True)
