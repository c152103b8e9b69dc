import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, za_b200
from tests import oracle as O
ctx = za_b200.Context(0)
for n in (65, 100, 1000):
    pts = O.g1_multiples(n)
    bases = za_b200.Bases(ctx, 1, pts)
    s = O.random_frs(n, 40 + n)
    got = za_b200.multiexp(ctx, bases, s)
    rc, exp = O.multiexp("g1", pts, s, threads=4)
    print(n, got == exp, flush=True)
