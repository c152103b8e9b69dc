"""Config 5 on several GPUs of one box: G1 MSM 2^16..2^26 over N = 1, 2, 4, 8 devices.
One process, one host thread and one context per device; device k owns the point range [k n / N, (k+1) n / N) with its
own fixed-base table and its slice of the scalars; every device returns an XYZZ partial sum and the host adds them
(the scheme of za_prover: no collective on the data path).  Time = max over the devices of the CUDA-event time of its
multiexp (events on the device's own stream), average of `reps` runs after two warm-ups; the correctness of the sharded
sum is checked at the small sizes against the single-device result.
python scratch/sweep_config5_multi.py [max_log] [Ns e.g. 1,2,4,8] [min_log]"""
import sys, os, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, za_b200
from za_b200 import synthetic

max_log = int(sys.argv[1]) if len(sys.argv) > 1 else 26
Ns = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8]
min_log = int(sys.argv[3]) if len(sys.argv) > 3 else 16
ndev = torch.cuda.device_count()
Ns = [n for n in Ns if n <= ndev]
ctxs, streams = [], []
for k in range(max(Ns)):
    with torch.cuda.device(k):
        c = za_b200.Context(k)
        s = torch.cuda.Stream(device=k)
        c.set_stream(s.cuda_stream)
        ctxs.append(c); streams.append(s)


def run_all(fns):
    errs = []
    def wrap(f):
        try: f()
        except Exception as e: errs.append(e)
    th = [threading.Thread(target=wrap, args=(f,)) for f in fns]
    for t in th: t.start()
    for t in th: t.join()
    if errs: raise errs[0]


def sweep(N, lg, reference):
    n = 1 << lg
    reps = 5 if lg <= 22 else 2
    full = synthetic.random_scalars(n, lg)
    fullw = synthetic.witness_like_scalars(n, lg + 100)
    st = [None] * N
    def setup(k):
        with torch.cuda.device(k):
            lo, hi = n * k // N, n * (k + 1) // N
            bases = za_b200.Bases.generate(ctxs[k], 1, hi - lo, lo + 1)
            c = bases.precompute()
            sc = torch.from_numpy(full[lo:hi]).cuda(k)
            scw = torch.from_numpy(fullw[lo:hi]).cuda(k)
            torch.cuda.synchronize(k)
            st[k] = dict(bases=bases, c=c, sc=sc, scw=scw, n=hi - lo)
    run_all([lambda k=k: setup(k) for k in range(N)])
    out = {}
    for name in ("sc", "scw"):
        ms = [0.0] * N
        parts = [None] * N
        def work(k):
            with torch.cuda.device(k):
                d = st[k]
                f = lambda: za_b200.multiexp_device(ctxs[k], d["bases"], d[name].data_ptr(), d["n"], partial=True)
                for _ in range(2): f()
                torch.cuda.synchronize(k)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(streams[k])
                for _ in range(reps): parts[k] = f()
                e1.record(streams[k]); torch.cuda.synchronize(k)
                ms[k] = e0.elapsed_time(e1) / reps
        run_all([lambda k=k: work(k) for k in range(N)])
        out[name] = (max(ms), za_b200.point_sum(1, parts))
    for k in range(N):
        st[k]["bases"].close()
    c = st[0]["c"]
    del st
    for k in range(N):
        with torch.cuda.device(k): torch.cuda.empty_cache()
    ok = ""
    if reference is not None:
        ok = "yes" if (out["sc"][1] == reference[0] and out["scw"][1] == reference[1]) else "NO"
    return out, c, ok


print("| log2 n | GPUs | table c (per device) | G1 MSM ms (uniform 253-bit) | Mpts/s | witness-like ms | Mpts/s | sum equals 1-GPU result |")
print("|---|---|---|---|---|---|---|---|")
sizes = [int(x) for x in os.environ['ZA_SWEEP_SIZES'].split(',')] if os.environ.get('ZA_SWEEP_SIZES') else list(range(min_log, max_log + 1, 2))
for lg in sizes:
    ref = None
    for N in Ns:
        out, c, ok = sweep(N, lg, ref)
        if N == Ns[0] and lg <= 22: ref = (out["sc"][1], out["scw"][1])
        n = 1 << lg
        t, tw = out["sc"][0], out["scw"][0]
        print(f"| {lg} | {N} | {c} | {t:.3f} | {n / t / 1e3:.1f} | {tw:.3f} | {n / tw / 1e3:.1f} | {ok if N != Ns[0] else ('reference' if ref else '-')} |", flush=True)
