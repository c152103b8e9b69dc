import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, za_b200
from za_b200 import synthetic
ctx = za_b200.Context(0)
st = torch.cuda.current_stream(); ctx.set_stream(st.cuda_stream)
for logn in (16, 20, 22, 24, 26):
    n = 1 << logn
    v = torch.from_numpy(synthetic.random_scalars(n, 1)).cuda()
    ref = v.clone()
    ctx.profile(True); ctx.profile_read()
    for _ in range(2): ctx.ntt_device(v.data_ptr(), logn, za_b200.FFT)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(3): ctx.ntt_device(v.data_ptr(), logn, za_b200.FFT)
    e1.record(st); torch.cuda.synchronize()
    p = ctx.profile_read()["ntt"]
    print(logn, "ms/transform", e0.elapsed_time(e1) / 3, "prof", p["ms"] / max(p["spans"], 1), p["spans"], "GElem/s", n / (e0.elapsed_time(e1) / 3) / 1e6, flush=True)
    # roundtrip check: ifft(fft(x)) == x in Montgomery domain (any 253-bit value is a valid element)
    w = ref.clone()
    ctx.ntt_device(w.data_ptr(), logn, za_b200.FFT); ctx.ntt_device(w.data_ptr(), logn, za_b200.IFFT)
    torch.cuda.synchronize()
    print("  roundtrip equal:", bool(torch.equal(w, ref)))
