"""Config 5 sweeps: G1 MSM 2^16..2^26 (uniform + witness-like scalars, with and without the fixed-base
table) and Fr NTT 2^16..2^26 on one GPU, inputs resident.  Prints a markdown table."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, za_b200
from za_b200 import synthetic
max_log = int(sys.argv[1]) if len(sys.argv) > 1 else 26
ctx = za_b200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
imad = za_b200.imad_peak(ctx)
def timed(fn, reps):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print(f"IMAD peak measured: {imad/1e12:.2f} T/s\n")
print("| log2 n | G1 MSM ms (table) | Mpts/s | IMAD frac (accumulate) | G1 MSM ms (no table) | Mpts/s | witness-like ms (table) | Mpts/s | G2 MSM ms (table) | Mpts/s |")
print("|---|---|---|---|---|---|---|---|---|---|")
for lg in range(16, max_log + 1, 2):
    n = 1 << lg
    reps = 5 if lg <= 22 else 2
    bases = za_b200.Bases.generate(ctx, 1, n, 1)
    sc = torch.from_numpy(synthetic.random_scalars(n, lg)).cuda()
    scw = torch.from_numpy(synthetic.witness_like_scalars(n, lg + 100)).cuda()
    plain = timed(lambda: za_b200.multiexp_device(ctx, bases, sc.data_ptr(), n), reps)
    c = bases.precompute()
    ctx.profile(True); ctx.profile_read()
    tab = timed(lambda: za_b200.multiexp_device(ctx, bases, sc.data_ptr(), n), reps)
    p = ctx.profile_read()["msm_accumulate_g1"]; ctx.profile(False)
    frac = (p["work"] * 264 / (p["ms"] * 1e-3)) / imad if p["ms"] else 0      # work is reported in Fq products
    tabw = timed(lambda: za_b200.multiexp_device(ctx, bases, scw.data_ptr(), n), reps)
    del bases
    g2 = ""
    if lg <= 24:
        b2 = za_b200.Bases.generate(ctx, 2, n, 1); b2.precompute()
        t2 = timed(lambda: za_b200.multiexp_device(ctx, b2, sc.data_ptr(), n), reps)
        g2 = f"{t2:.3f} | {n/t2/1e3:.1f}"
        del b2
    else:
        g2 = "- | -"
    print(f"| {lg} | {tab:.3f} (c={c}) | {n/tab/1e3:.1f} | {frac:.3f} | {plain:.3f} | {n/plain/1e3:.1f} | {tabw:.3f} | {n/tabw/1e3:.1f} | {g2} |", flush=True)
    del sc, scw
    torch.cuda.empty_cache()
print()
print("| log2 N | Fr NTT ms (forward, natural->natural) | GElem/s | HBM frac (64 B/elem) | IMAD frac |")
print("|---|---|---|---|---|")
hbm = 6558.1          # MEASURED_PEAKS.json hbm_gbs
for lg in range(16, max_log + 1, 2):
    n = 1 << lg
    v = torch.from_numpy(synthetic.random_scalars(n, lg)).cuda()
    t = timed(lambda: ctx.ntt_device(v.data_ptr(), lg, za_b200.FFT), 5 if lg <= 22 else 3)
    print(f"| {lg} | {t:.4f} | {n/t/1e6:.3f} | {64.0*n/(t*1e-3)/1e9/hbm:.4f} | {264*(n/2)*lg/(t*1e-3)/imad:.3f} |", flush=True)
    del v
