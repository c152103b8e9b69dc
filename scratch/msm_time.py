"""Accumulation time of one 2^20-point table-mode multiexp (G1 and G2) for several pair-round settings.
python scratch/msm_time.py [g1|g2|both] [rounds list]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, za_b200
from za_b200 import synthetic
what = sys.argv[1] if len(sys.argv) > 1 else "both"
rounds_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 3, 4]
ctx = za_b200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
n = 1 << 20
for group in ([1] if what == "g1" else [2] if what == "g2" else [1, 2]):
    bases = za_b200.Bases.generate(ctx, group, n, 1)
    bases.precompute()
    sc = torch.from_numpy(synthetic.random_scalars(n, 2)).cuda()
    ref = None
    for r in rounds_list:
        os.environ["ZA_MSM_ROUNDS"] = str(r)
        for _ in range(2): res = za_b200.multiexp_device(ctx, bases, sc.data_ptr(), n)
        ctx.profile(True); ctx.profile_read()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(st)
        for _ in range(reps): res = za_b200.multiexp_device(ctx, bases, sc.data_ptr(), n)
        e1.record(st); torch.cuda.synchronize()
        p = ctx.profile_read(); ctx.profile(False)
        if ref is None: ref = res
        print("G%d rounds=%d total %.3f ms  " % (group, r, e0.elapsed_time(e1) / reps) + " ".join("%s %.3f" % (k, v["ms"] / reps) for k, v in p.items() if v["ms"] > 0) + ("  OK" if res == ref else "  MISMATCH"), flush=True)
