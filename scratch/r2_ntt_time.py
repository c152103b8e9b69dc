"""Round 2: NTT / H pipeline timings.  python scratch/r2_ntt_time.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, za_b200
from za_b200 import synthetic
ctx = za_b200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
tag = os.environ.get("ZA_NTT_MAXK", "default")
peak = za_b200.imad_peak(ctx)
def ev(fn, reps):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for log_n in (16, 18, 20, 22, 24):
    n = 1 << log_n
    v = torch.from_numpy(synthetic.random_scalars(n, 1)).cuda()
    ms = ev(lambda: ctx.ntt_device(v.data_ptr(), log_n, za_b200.FFT), 10)
    print("[maxk %s] fft 2^%d: %.3f ms  imad_frac %.3f  %.2f GElem/s" % (tag, log_n, ms, 264 * (n / 2) * log_n / (ms * 1e-3) / peak, n / ms / 1e6), flush=True)
    if log_n <= 22:
        v3 = torch.from_numpy(synthetic.random_scalars(3 * n, 2)).cuda()
        ms3 = ev(lambda: ctx.ntt_device(v3.data_ptr(), log_n, za_b200.IFFT, 3), 10)
        print("[maxk %s] ifft x3 batch 2^%d: %.3f ms  imad_frac %.3f" % (tag, log_n, ms3, 3 * 264 * (n / 2) * log_n / (ms3 * 1e-3) / peak), flush=True)
        a = torch.from_numpy(synthetic.random_scalars(3 * n, 3)).cuda()
        ctx.fr_convert_device(a.data_ptr(), 3 * n, False)
        p = a.data_ptr()
        msh = ev(lambda: ctx.h_poly_device(p, p + 32 * n, p + 64 * n, log_n), 10)
        print("[maxk %s] h_poly 2^%d: %.3f ms  imad_frac %.3f" % (tag, log_n, msh, 264 * (7 * (n / 2) * log_n + 6 * n) / (msh * 1e-3) / peak), flush=True)
    del v
