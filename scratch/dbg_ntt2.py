import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, za_b200
from tests import oracle as O
ctx = za_b200.Context(0)
for logn in (23, 24):
    d = O.random_frs(1 << logn, 5)
    for mode in (0, 1):
        got = ctx.ntt(d, mode)
        exp = O.fft(d, logn, mode, threads=16)
        bad = np.nonzero((got != exp).any(axis=1))[0]
        print(logn, mode, "mismatches", len(bad), bad[:16], [bin(x) for x in bad[:4]], flush=True)
