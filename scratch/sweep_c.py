import sys, os, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, za_b200
from za_b200 import synthetic
lg = int(sys.argv[1]); group = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = za_b200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
n = 1 << lg
bases = za_b200.Bases.generate(ctx, group, n, 1)
c = bases.precompute()
sc = torch.from_numpy(synthetic.random_scalars(n, lg)).cuda()
def timed(fn, reps):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t = timed(lambda: za_b200.multiexp_device(ctx, bases, sc.data_ptr(), n), 5)
print(f"group {group} log n {lg} table c={c} env={os.environ.get('ZA_MSM_TABLE')}: {t:.3f} ms")
