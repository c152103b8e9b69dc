"""Short profiling target: one NTT, one G1 MSM, one G2 MSM (inputs resident).  Run under ncu with -k filters."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, za_b200
from za_b200 import synthetic
what = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = za_b200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
if what in ("all", "ntt"):
    logn = 22
    v = torch.from_numpy(synthetic.random_scalars(1 << logn, 1)).cuda()
    for _ in range(2): ctx.ntt_device(v.data_ptr(), logn, za_b200.FFT)
    torch.cuda.synchronize()
if what in ("all", "g1"):
    n = 1 << 22
    bases = za_b200.Bases.generate(ctx, 1, n, 1)
    sc = torch.from_numpy(synthetic.random_scalars(n, 2)).cuda()
    for _ in range(2): za_b200.multiexp_device(ctx, bases, sc.data_ptr(), n)
if what in ("all", "g2"):
    n = 1 << 20
    bases = za_b200.Bases.generate(ctx, 2, n, 1)
    sc = torch.from_numpy(synthetic.random_scalars(n, 3)).cuda()
    for _ in range(2): za_b200.multiexp_device(ctx, bases, sc.data_ptr(), n)
if what == "g1t":
    n = 1 << 20
    bases = za_b200.Bases.generate(ctx, 1, n, 1)
    bases.precompute()
    sc = torch.from_numpy(synthetic.random_scalars(n, 2)).cuda()
    for _ in range(2): za_b200.multiexp_device(ctx, bases, sc.data_ptr(), n)
if what == "g2t":
    n = 1 << 20
    bases = za_b200.Bases.generate(ctx, 2, n, 1)
    bases.precompute()
    sc = torch.from_numpy(synthetic.random_scalars(n, 3)).cuda()
    for _ in range(2): za_b200.multiexp_device(ctx, bases, sc.data_ptr(), n)
print("done")
if what == "hpoly":
    logn = 20
    n = 1 << logn
    a = torch.from_numpy(synthetic.random_scalars(3 * n, 3)).cuda()
    ctx.fr_convert_device(a.data_ptr(), 3 * n, False)
    p = a.data_ptr()
    for _ in range(2): ctx.h_poly_device(p, p + 32 * n, p + 64 * n, logn)
    torch.cuda.synchronize()
    print("hpoly done")
