"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, za_b200
from za_b200 import synthetic
from tests import oracle as O, circuits
ctx = za_b200.Context(0)
for lg in (2, 5, 11, 13):
    d = O.random_frs(1 << lg, lg)
    for mode in range(4):
        assert np.array_equal(ctx.ntt(d, mode), O.fft(d, lg, mode, 4)), (lg, mode)
a, b, c = (O.random_frs(3000, s) for s in (1, 2, 3))
got, ck = ctx.h_poly(a, b, c, checkpoints=True)
exp, eck = O.h_poly(a, b, c, threads=4, checkpoints=True)
assert np.array_equal(got, exp) and np.array_equal(ck, eck)
for group, n in ((1, 70), (1, 100), (1, 5000), (2, 100), (2, 4500)):
    pts = O.g1_multiples(n) if group == 1 else O.g2_multiples(n)
    bases = za_b200.Bases(ctx, group, pts)
    for s in (O.random_frs(n, n), circuits.witness_like(n, n)):
        rc, e = O.multiexp("g1" if group == 1 else "g2", pts, s, threads=4)
        assert za_b200.multiexp(ctx, bases, s) == e, (group, n)
    if n >= 4096:
        bases.precompute()
        assert za_b200.multiexp(ctx, bases, s) == e, (group, n, "table")
ni, na, ptr, var, coeff, inputs, aux = circuits.mul_chain_fast(5000, x0=3)
ocs = O.CS(ni, na, ptr, var, coeff)
prm = O.Params.generate(ocs, [3, 5, 7, 11, 13], threads=8)
rc, expect = prm.create_proof(ocs, inputs, aux, 9, 10, threads=8)
pk = za_b200.Parameters.read(ctx, prm.write(), checked=True)
circ = za_b200.Circuit(ctx, ni, na, ptr, var, coeff)
proof, tr = za_b200.create_proof(ctx, pk, circ, inputs, aux, 9, 10, trace=True)
assert proof == expect
print("sanitize target ok")
