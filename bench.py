#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200 (BASELINE.json):
  "Groth16 prove ms @2^20 constraints; G1 MSM Mpts/s; Fr NTT GElem/s".

One step = one create_proof over the synthetic 2^20-domain multiplication-chain circuit (SURVEY §8d
config 4).  `value` = device-timed ms per proof with the witness resident in HBM (per-class profiling
events off); `e2e` = the same through the public host-buffer call (za_prover_create_proof: witness H2D
and the partial results D2H inside the timed region).  Every line carries `proof_sha256` and, from the
oracle run after the timed region, `cpu_baseline.proof_matches_gpu`.  Rooflines: `roofline` (dominant
kernel, IMAD bound, stand-alone launch time, ncu DRAM traffic), `roofline_g2`, `roofline_ntt` (against
HBM, IMAD fraction beside it), `roofline_step` (all field products of a proof at their IMAD cost / step);
`kernel_ms_per_step` are the overlapping in-step class times, `kernel_ms_alone_per_step` the stand-alone
ones whose serial sum the step is.  Sub-metrics: 2^24-point G1 MSM and Fr NTT (config 5), a real key at
2^20 through Parameters::read(checked), the EdDSA-MiMC statement (config 3).
N > 1: torchrun launches one rank per GPU as the contract demands; rank 0 drives ALL N devices through ONE
za_prover_create_proof (a host thread per device inside the library, planned point ranges, h slices
stored into the peers' memory by the last NTT pass, host combine: no NCCL on the data path), the other
ranks only join the barriers — total work is fixed, so scaling is "strong".

`--impl reference`: the CPU restatement of bellman's algorithm (oracle/, multi-threaded) on the same
workload; it is also what `cpu_baseline` reports.  The oracle is never on the GPU arm's timed path.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Groth16 prove ms @2^20 constraints; G1 MSM Mpts/s; Fr NTT GElem/s"
IMAD_PER_MODMUL = 264          # 8x32-bit-limb Montgomery product, IMAD-class instructions (64 a.b + 64 m.q limb products x 2, + 8)
IMAD_PER_SQR = 208             # dedicated squaring: 36 limb products + the same 64-product reduction
# per "Fq product" unit the accumulation kernels report as work (G1: 10 per XYZZ mixed addition = 8 M + 2 S;
# G2: an Fq2 product = 3 units, an Fq2 squaring = 2; lazily reduced Karatsuba = 3 wide products + 2 reductions = 656 IMAD)
IMAD_PER_G1_UNIT = (8 * IMAD_PER_MODMUL + 2 * IMAD_PER_SQR) / 10.0
IMAD_PER_G2_UNIT = (5 * 656 + 2 * IMAD_PER_MODMUL) / 17.0       # batched-affine addition over Fq2: 5 M + 1 S
# The accumulation kernels report their algorithmic work in Fq products (za_ctx_profile_read): an XYZZ mixed
# addition is 8M + 2S = 10 (over Fq2: 8 x 3 + 2 x 2 = 28), a batched-affine addition of the G2 pair rounds
# 5M + 1S over Fq2 = 17.
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617   # compiler/src/algebra/fs.rs:15-16
R_FIXED = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF1234567890AB
S_FIXED = 0x0FEDCBA9876543210FEDCBA9876543210FEDCBA9876543210FEDCBA98765


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_prove_workload(log_m):
    """The synthetic circuit and proving-key sizes, shared by both arms."""
    from za_b200 import synthetic
    nc = (1 << log_m) - 2
    t = time.time()
    cs = synthetic.mul_chain(nc, x0=0x5A410004)
    counts = synthetic.pk_counts_for_mul_chain(nc)
    return nc, cs, counts, time.time() - t


def run_reference(args):
    """The CPU restatement of bellman's create_proof on this host's cores (oracle/), same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from tests import oracle as O
    cores = os.cpu_count() or 1
    total_steps = args.steps + args.warmup
    # the TRUE workload for every step (8-9 s per 2^20 proof on 16 threads: steps + warmup = 25 is ~3.5 minutes);
    # --ref-log-m shrinks the sample explicitly (then the line says so: sample_log_m, extrapolated)
    log_m = args.log_m if args.ref_log_m <= 0 else min(args.log_m, args.ref_log_m)
    nc, cs, counts, _ = cpu_prove_workload(log_m)
    ni, na, ptr, var, coeff, inputs, aux = cs
    ocs = O.CS(ni, na, ptr, var, coeff)
    prm = O.Params.synthetic(counts["ic"], counts["h"], counts["l"], counts["a"], counts["b_g1"], counts["b_g2"], threads=cores)
    times = []
    proof = None
    for i in range(total_steps):
        t = time.perf_counter()
        rc, proof = prm.create_proof(ocs, inputs, aux, R_FIXED, S_FIXED, threads=cores)
        dt = (time.perf_counter() - t) * 1e3
        assert rc == 0
        if i >= args.warmup:
            times.append(dt)
    ms = sum(times) / len(times)
    scale = float(1 << (args.log_m - log_m))
    value = ms * scale
    sample = f"create_proof, mul-chain circuit, domain 2^{log_m}, {cores} threads" + \
             (f"; linearly extrapolated x{int(scale)} to 2^{args.log_m}" if scale != 1 else "")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": value, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64-limb Montgomery (CPU)", "data": "synthetic",
            "config": {"workload": f"synthetic mul-chain R1CS, domain 2^{args.log_m}: 7 NTT + 4 G1 MSM + 1 G2 MSM + 3 input MSMs",
                       "log_m": args.log_m, "sample_log_m": log_m},
            "cpu_baseline": {"value": value, "unit": "ms", "cores": cores, "kind": "port", "sample": sample,
                             "note": "bellman-algorithm restatement (oracle/), not bellman itself"},
            "e2e": {"value": value, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "proof_sha256": hashlib.sha256(proof).hexdigest() if proof is not None and scale == 1 else None}
    emit(line)
    return 0


def run_gpu(args):
    import torch
    import za_b200
    from za_b200 import synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        # torchrun is the LAUNCHER only (the driver's contract): barriers and the max over ranks of the timing.  The proof
        # itself is one call into the library on rank 0 — za_prover_create_proof drives all N GPUs of the box from one
        # process (host thread per device, cudaMemcpyPeerAsync for the h slices): no torch / NCCL on the data path.
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # Barriers go through a second, CPU-side (gloo) group: an NCCL barrier is a kernel that SPINS on the waiting rank's
        # GPU, and ranks > 0 reach the closing barrier of a timed region at once — their spinning kernel would time-slice
        # with the kernels rank 0's process runs on that same GPU (measured: N = 2 proofs 30 ms instead of 14 ms).
        cpu_group = dist.new_group(backend="gloo")
        t_hello = torch.ones(1, device=torch.device("cuda", local_rank))
        dist.all_reduce(t_hello)                 # NCCL sees all N ranks
        assert int(t_hello.item()) == world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(device=dev)      # rank 0: device 0's library stream, so that torch's events bracket its work
    torch.cuda.set_stream(stream)

    log_m = args.log_m
    nc, cs, counts, _ = cpu_prove_workload(log_m)
    ni, na, ptr, var, coeff, inputs, aux = cs
    m = 1 << log_m
    prover = ctx = None
    if rank == 0:
        prover = za_b200.Prover(list(range(world)))
        prover.synthetic_pk(counts["ic"], counts["h"], counts["l"], counts["a"], counts["b_g1"], counts["b_g2"])
        prover.set_circuit(ni, na, ptr, var, coeff)
        ctx = prover.ctx(0)
        ctx.set_stream(stream.cuda_stream)
    else:
        ctx = za_b200.Context(local_rank)       # the sharded G1 multiexp sub-metric runs one context per rank
        ctx.set_stream(stream.cuda_stream)
    inputs_pin = torch.from_numpy(inputs.copy()).pin_memory()
    aux_pin = torch.from_numpy(aux.copy()).pin_memory()
    if rank == 0:
        prover.upload_witness(inputs_pin.numpy(), aux_pin.numpy())
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(group=cpu_group)
        torch.cuda.synchronize()

    def prove_device_step():
        """One proof, witness resident on the devices (every device holds the span its point ranges read)."""
        return prover.create_proof(None, None, R_FIXED, S_FIXED) if rank == 0 else None

    def prove_e2e_step():
        """Through the host-buffer call: witness H2D (pinned host memory; every device pulls its span over its own PCIe
        link) and the read-back of the partial sums inside."""
        return prover.create_proof(inputs_pin.numpy(), aux_pin.numpy(), R_FIXED, S_FIXED) if rank == 0 else None

    def timed(fn, steps, warmup, after_warmup=None):
        for _ in range(warmup):
            fn()
        barrier()
        if after_warmup:
            after_warmup()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = None
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    W = max(args.warmup, 3)
    K = args.steps
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- headline: prove ms, witness resident
    # per-kernel-class CUDA-event timing covers the K timed steps only (the warm-up steps allocate and page in code)
    l0 = [0]
    def start_profile():
        if rank == 0:
            ctx.profile(True)
            ctx.profile_read()
            l0[0] = prover.launch_count()
    # the headline is timed with the per-class profiling OFF (its events cost ~0.5 ms per proof: an event pair per kernel
    # class and multiexp); a second pass of K steps with the profiling on gives the per-class times and the launch count
    prove_ms, proof = timed(prove_device_step, K, W)
    prove_profiled_ms, proof_p = timed(prove_device_step, K, 0, start_profile)
    launches = (prover.launch_count() - l0[0]) if rank == 0 else 0
    prof = ctx.profile_read() if rank == 0 else None
    ctx.profile(False)
    if rank == 0:
        assert proof == proof_p, "profiled and unprofiled proofs differ"
    steps_profiled = K
    # ---- e2e: host buffers
    e2e_ms, proof_e2e = timed(prove_e2e_step, K, W)
    if rank == 0:
        assert proof == proof_e2e, "device-resident and host-buffer proofs differ"
    clocks = sampler.summary() if sampler else None

    # The G2 multiexp runs on its own stream NEXT TO the G1 multiexps (they share the SMs by design), so the per-class
    # event times of the step overlap and are not per-kernel durations.  For the roofline objects the two accumulation
    # kernels are therefore also timed ALONE, in this process, on the same workload size (2^log_m points, fixed-base
    # table, uniform scalars): that duration is what `achieved` uses; the in-step figure is reported beside it.
    iso_other = {}      # per group: digit sort and bucket reduction of one query-sized multiexp, alone (ms)
    def isolated_accumulation(group):
        n_iso = 1 << log_m
        bases_iso = za_b200.Bases.generate(ctx, group, n_iso, 1)
        bases_iso.precompute()
        sc_iso = torch.from_numpy(synthetic.random_scalars(n_iso, 0x5A410008 + group)).to(dev)
        for _ in range(3):
            za_b200.multiexp_device(ctx, bases_iso, sc_iso.data_ptr(), n_iso)
        ctx.profile(True)
        ctx.profile_read()
        for _ in range(5):
            za_b200.multiexp_device(ctx, bases_iso, sc_iso.data_ptr(), n_iso)
        p_iso = ctx.profile_read()
        ctx.profile(False)
        del bases_iso, sc_iso
        iso_other[group] = {k: p_iso[k]["ms"] / 5.0 for k in ("msm_sort", "msm_reduce")}
        return p_iso["msm_accumulate_g1" if group == 1 else "msm_accumulate_g2"]

    def isolated_ntt():
        """The seven transforms of one proof as the H pipeline runs them (a, b, c batched; scalings and the pointwise
        step fused into loads / stores), alone on the GPU: the NTT class of za_ctx_profile over five pipelines."""
        mm = 1 << log_m
        v_iso = torch.from_numpy(synthetic.random_scalars(3 * mm, 0x5A41000B)).to(dev)
        base = v_iso.data_ptr()
        for _ in range(3):
            ctx.h_poly_device(base, base + 32 * mm, base + 64 * mm, log_m)
        ctx.profile(True)
        ctx.profile_read()
        for _ in range(5):
            ctx.h_poly_device(base, base + 32 * mm, base + 64 * mm, log_m)
        p_iso = ctx.profile_read()
        ctx.profile(False)
        del v_iso
        return p_iso["ntt"]

    iso1 = iso2 = iso_ntt = None
    if rank == 0 and log_m <= 22:
        try:
            iso1, iso2, iso_ntt = isolated_accumulation(1), isolated_accumulation(2), isolated_ntt()
        except Exception as e:          # fall back to the in-step event times rather than lose the line
            print(f"isolated kernel timing failed: {e!r}", file=sys.stderr)
            ctx.profile(False)
            iso1 = iso2 = iso_ntt = None

    line = None
    if rank == 0:
        hbm_peak, peak_src = peaks()
        imad_peak = za_b200.imad_peak(ctx)
        acc1, acc2, nttp = prof["msm_accumulate_g1"], prof["msm_accumulate_g2"], prof["ntt"]
        in_step = {"g1": acc1["ms"] / max(acc1["spans"], 1), "g2": acc2["ms"] / max(acc2["spans"], 1)}
        n_g1_launches = acc1["spans"] / max(steps_profiled, 1)
        g1_work_per_step = acc1["work"] / max(steps_profiled, 1)
        g2_work_per_step = acc2["work"] / max(steps_profiled, 1)
        ntt_in_step_ms = nttp["ms"] / max(steps_profiled, 1)
        if iso1 is not None:
            acc1, acc2, nttp = iso1, iso2, iso_ntt
        # dominant kernel of the step: the G1/G2 bucket-accumulation kernels (integer-pipe bound)
        per1, per2 = acc1["ms"] / max(acc1["spans"], 1), acc2["ms"] / max(acc2["spans"], 1)
        # time per step of the G1 accumulation = its launch timed alone x the number of such launches the step's G1 work amounts
        # to (the merged B/L/A multiexp is ONE launch over three queries: 4 query-sized launches per proof at N = 1)
        g1_equiv = g1_work_per_step / max(acc1["work"] / max(acc1["spans"], 1), 1.0) if iso1 is not None else n_g1_launches
        in_step["g1"] = prof["msm_accumulate_g1"]["ms"] / max(steps_profiled, 1) / max(g1_equiv, 1.0)      # per query-sized launch
        # the same for G2: device 0's share of the B (G2) query at N > 1 (0 when the plan gives it none)
        g2_equiv = g2_work_per_step / max(acc2["work"] / max(acc2["spans"], 1), 1.0) if iso1 is not None else 1.0
        dom = acc1 if per1 * g1_equiv >= per2 * g2_equiv else acc2
        dom_name = "msm_accumulate_g1_sm_kernel (G1 bucket accumulation, XYZZ mixed additions)" if dom is acc1 else "msm_pair_round_kernel<Fq2> x3 + msm_accumulate_kernel<Fq2>"
        traffic = ntt_traffic = None
        tfile = {}
        for name in ("r01_traffic.json", "r02_traffic.json"):          # later rounds override
            tpath = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tpath):
                tfile.update(json.load(open(tpath)))
        if dom is acc1 and log_m == 20 and tfile.get("msm_accumulate_g1_sm_kernel"):
            t = tfile["msm_accumulate_g1_sm_kernel"]
            traffic = int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
        if log_m == 20 and tfile.get("ntt_pass_kernel"):
            t = tfile["ntt_pass_kernel"]
            ntt_traffic = int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
        imads = dom["work"] * (IMAD_PER_G1_UNIT if dom is acc1 else IMAD_PER_G2_UNIT)
        achieved = imads / (dom["ms"] * 1e-3) / 1e12 if dom["ms"] > 0 else 0.0
        roofline = {"bound": "imad", "kernel": dom_name, "achieved": achieved, "peak": imad_peak / 1e12, "unit": "TIMAD/s",
                    "frac": achieved / (imad_peak / 1e12) if imad_peak else None, "traffic": traffic,
                    "traffic_note": "DRAM read+write bytes per launch from the ncu --set full capture of the same 2^20 workload (profiles/r0N_traffic.json); null for other sizes",
                    "peak_source": "measured in this run (za_imad_peak: dependency-free mad.lo.u32 on all SMs)",
                    "algorithmic": f"{int(dom['work'] / max(dom['spans'], 1))} Fq products (" + ("10 per XYZZ mixed addition: 8 products x 264 IMAD + 2 squarings x 208 IMAD" if dom is acc1 else "17 per batched-affine addition over Fq2: 5 lazily reduced Fq2 products x 656 IMAD + 1 Fq2 squaring x 528 IMAD") + ") per launch (rank 0's share)",
                    "launch_ms": dom["ms"] / max(dom["spans"], 1),
                    "timing": ("kernel timed alone in this process (CUDA events), same size and table as in the proof; in the proof it shares the SMs with the G2 multiexp's kernels by design" if iso1 is not None else "in-step CUDA events"),
                    "in_step_launch_ms": in_step["g1" if dom is acc1 else "g2"],
                    "share_of_step": (dom["ms"] / max(dom["spans"], 1)) * (g1_equiv if dom is acc1 else g2_equiv) / prove_ms,
                    "launches_per_step": (g1_equiv if dom is acc1 else g2_equiv),
                    "note": "tensor cores not applicable (multiprecision integer); HBM needs 96 B/point, two orders below compute"}
        g2_t = acc2["work"] * IMAD_PER_G2_UNIT / (acc2["ms"] * 1e-3) / 1e12 if acc2["ms"] > 0 else 0.0
        roofline_g2 = {"bound": "imad", "kernel": "G2 bucket accumulation: msm_pair_round_kernel<Fq2> (batched-affine rounds) + msm_accumulate_kernel<Fq2>",
                       "achieved": g2_t, "peak": imad_peak / 1e12, "unit": "TIMAD/s", "frac": g2_t / (imad_peak / 1e12) if imad_peak else None,
                       "algorithmic": f"{int(acc2['work'] / max(acc2['spans'], 1))} Fq products (17 per batched-affine addition, 28 per XYZZ mixed addition over Fq2) x {IMAD_PER_G2_UNIT:.0f} IMAD (an Fq2 product = 3 units = 656 IMAD lazily reduced, an Fq2 squaring = 2 units = 528) per multiexp",
                       "launch_ms": acc2["ms"] / max(acc2["spans"], 1), "in_step_launch_ms": in_step["g2"],
                       "share_of_step": acc2["ms"] / max(acc2["spans"], 1) * g2_equiv / prove_ms, "launches_per_step": g2_equiv}
        ntt_bytes = 64.0 * nttp["work"]
        ntt_gbs = ntt_bytes / (nttp["ms"] * 1e-3) / 1e9 if nttp["ms"] > 0 else 0.0
        roofline_ntt = {"bound": "hbm", "kernel": "ntt_pass_kernel (one transform = 2 passes at 2^20; a, b, c batched)", "achieved": ntt_gbs, "peak": hbm_peak, "unit": "GB/s",
                        "frac": ntt_gbs / hbm_peak, "traffic": ntt_traffic, "peak_source": peak_src,
                        "traffic_note": "DRAM read+write bytes of ONE transform (its two passes) from the ncu capture of the H pipeline at 2^20 (profiles/r02_traffic.json); algorithmic 64 B x 2^20 = 67 MB",
                        "algorithmic": "64 B per element per transform (one 32 B read + one 32 B write)",
                        "timing": ("the seven transforms of one proof as the H pipeline runs them (a, b, c batched), alone on the GPU; in a single-GPU proof the pipeline runs next to the witness multiexps and stretches" if iso_ntt is not None else "in-step CUDA events"),
                        "in_step_ms_per_proof": ntt_in_step_ms,
                        "imad_frac": (IMAD_PER_MODMUL * (nttp["work"] / 2) * log_m / (nttp["ms"] * 1e-3)) / imad_peak if nttp["ms"] > 0 else None}
        # the same classes timed ALONE (one query-sized multiexp per group, the H pipeline): their serial sum is what the step
        # costs, because every hot kernel fills the SMs by itself and concurrent streams only time-slice
        alone = None
        if iso1 is not None and world == 1 and 1 in iso_other and 2 in iso_other:
            alone = {"msm_accumulate_g1": round(per1 * g1_equiv, 4), "msm_accumulate_g2": round(per2, 4),
                     "ntt": round(nttp["ms"] / 5.0, 4),
                     "msm_sort": round(iso_other[1]["msm_sort"] * g1_equiv, 4),
                     "msm_reduce": round(iso_other[1]["msm_reduce"] * g1_equiv + iso_other[2]["msm_reduce"], 4),
                     "note": "stand-alone CUDA-event times scaled to one proof: G1 figures x the number of query-sized G1 multiexps of a proof "
                             "(the G2 multiexp shares the B digit sort); an upper bound for the merged B/L/A multiexp, whose three bucket "
                             "spaces share one reduction chain"}
            alone["sum"] = round(sum(v for k, v in alone.items() if k != "note"), 4)
        breakdown = {k: round(v["ms"] / steps_profiled, 4) for k, v in prof.items() if v["ms"] > 0}
        breakdown["note"] = "event time per kernel class per step; the classes run on concurrent streams, so the sum exceeds the step"
        # uploaded per proof: the whole witness on device 0 and, on every other device, the span its point ranges read;
        # read back: six partial results of the bucket reduction per multiexp and device (4 x G1 of 128 B points, 1 x G2 of
        # 256 B points) and the 8-byte verdict of the witness range check; the proof itself is assembled on the host
        pinfo = prover.info()
        h2d, D2H_BYTES, rank0_weight = pinfo["h2d_bytes_per_proof"], pinfo["d2h_bytes_per_proof"], pinfo["device0_weight"]
        # whole-step utilisation of the integer-multiply pipe: every Fq / Fr product of one proof (device 0's share at
        # N > 1) x 264 IMAD against the measured peak over the measured step
        p_acc = (prof["msm_accumulate_g1"]["work"] + prof["msm_accumulate_g2"]["work"]) / steps_profiled
        p_ntt = prof["ntt"]["work"] / steps_profiled / 2 * log_m
        n_g1_red = prof["msm_accumulate_g1"]["spans"] / steps_profiled
        n_g2_red = prof["msm_accumulate_g2"]["spans"] / steps_profiled
        buckets = prof["msm_reduce"]["work"] / steps_profiled / max(n_g1_red + n_g2_red, 1)      # per multiexp
        p_red = buckets * 2 * (14 * n_g1_red + 42 * n_g2_red)        # ~2 XYZZ additions per bucket, 12M + 2S each (x3 over Fq2)
        p_other = (prof["h_pointwise"]["work"] * 3 + prof["r1cs_eval"]["work"] + (ni + na) * steps_profiled) / steps_profiled
        step_products = p_acc + p_ntt + p_red + p_other
        step_imads = (prof["msm_accumulate_g1"]["work"] * IMAD_PER_G1_UNIT + prof["msm_accumulate_g2"]["work"] * IMAD_PER_G2_UNIT) / steps_profiled \
            + (p_ntt + p_red + p_other) * IMAD_PER_MODMUL
        roofline_step = {"bound": "imad", "achieved": step_imads / (prove_ms * 1e-3) / 1e12, "peak": imad_peak / 1e12, "unit": "TIMAD/s",
                         "frac": step_imads / (prove_ms * 1e-3) / imad_peak if imad_peak else None,
                         "products_per_proof": {"bucket_accumulation": int(p_acc), "ntt_butterflies": int(p_ntt), "bucket_reduction_estimate": int(p_red), "other": int(p_other)},
                         "note": "all field products of one proof on device 0 (its share of the multiexps at N > 1) x their IMAD cost (product 264, squaring 208, Fq2 product 656) / ms_per_step / measured IMAD peak"}
        line = {"metric": METRIC, "value": prove_ms, "unit": "ms", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": prove_ms,
                "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": {"workload": f"synthetic mul-chain R1CS, {nc} constraints (domain 2^{log_m}), {na} aux: "
                                       "7 NTT + 4 G1 MSM + 1 G2 MSM + 3 input MSMs per proof; synthetic proving key "
                                       "(bases = known multiples of the generators), r and s fixed",
                           "log_m": log_m, "parallelism": f"one process drives {world} GPUs (za_prover_create_proof): msm point-range x{world}, device 0 (H pipeline) takes {rank0_weight:.3f} of a share of the witness multiexps; torchrun ranks > 0 only take part in the barriers" if world > 1 else "single GPU",
                           "l2": "inputs larger than L2: proving key 470 MB + witness 32 MB per step"},
                "roofline": roofline, "roofline_g2": roofline_g2, "roofline_ntt": roofline_ntt, "roofline_step": roofline_step, "kernel_ms_per_step": breakdown, "kernel_ms_alone_per_step": alone,
                "proof_sha256": hashlib.sha256(proof).hexdigest(),
                "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": D2H_BYTES},
                "gpu_launches": int(launches), "clocks": clocks, "imad_peak_timads": imad_peak / 1e12,
                "profiled_ms_per_step": prove_profiled_ms,
                "timing_note": "value / ms_per_step: K steps with the per-class event profiling off; kernel_ms_per_step, gpu_launches and the in-step figures of the roofline objects come from a second pass of K steps with it on (profiled_ms_per_step)"}

    # ---- sub-metrics: 2^log_msm-point G1 MSM (point range sharded over the ranks, fixed-base table per shard,
    #      partial sums combined on the host) and, at N = 1, the 2^log_ntt Fr NTT; inputs resident
    if not args.no_sub:
        sub = {}
        n_msm = 1 << args.log_msm
        lo, hi = za_b200.share(n_msm, rank, world)
        n_loc = hi - lo
        bases = za_b200.Bases.generate(ctx, 1, n_loc, 1 + lo)
        tab_c = bases.precompute()
        sc_all = synthetic.random_scalars(n_msm, 0x5A410005)
        sc = torch.from_numpy(sc_all[lo:hi].copy()).to(dev)
        rec = torch.zeros(128, dtype=torch.uint8, device=dev)
        recs = [torch.zeros_like(rec) for _ in range(world)] if world > 1 else None
        torch.cuda.synchronize()

        def msm_step(d_scalars):
            if world == 1:
                return za_b200.multiexp_device(ctx, bases, d_scalars, n_loc)
            part = za_b200.multiexp_device(ctx, bases, d_scalars, n_loc, partial=True)
            rec.copy_(torch.from_numpy(np.frombuffer(part, np.uint8).copy()))
            dist.all_gather(recs, rec)
            return za_b200.point_sum(1, [r.cpu().numpy().tobytes() for r in recs]) if rank == 0 else None

        ctx.profile(True)
        ctx.profile_read()
        msm_ms, msm_pt = timed(lambda: msm_step(sc.data_ptr()), max(2, min(K, 5)), 3)
        p2 = ctx.profile_read()
        ctx.profile(False)
        a = p2["msm_accumulate_g1"]
        ipk = za_b200.imad_peak(ctx)
        sub["g1_msm"] = {"log_n": args.log_msm, "n_gpus": world, "ms": msm_ms, "mpts_s": n_msm / msm_ms / 1e3, "scalars": "uniform 253-bit",
                         "fixed_base_table_c": tab_c, "accumulate_ms": a["ms"] / max(a["spans"], 1),
                         "sort_ms": p2["msm_sort"]["ms"] / max(a["spans"], 1),
                         "imad_frac": (a["work"] * IMAD_PER_G1_UNIT / (a["ms"] * 1e-3)) / ipk if a["ms"] else None}
        if rank == 0 and args.log_msm <= 20:
            # linearity check of the full-size result: bases are (1+i) G, so the sum is (sum s_i (1+i) mod r) G
            from tests import oracle as O
            from tests import pyref as P
            v = sc_all.view(np.uint64).reshape(n_msm, 4)
            tot = sum(sum(int(x) * (i + 1) for i, x in enumerate(v[:, limb])) << (64 * limb) for limb in range(4)) % P.R_MOD
            sub["g1_msm"]["matches_closed_form"] = bool(O.g1_tuple(msm_pt) == O.g1_mul(P.G1_GEN, tot))
        scw = torch.from_numpy(synthetic.witness_like_scalars(n_msm, 0x5A410006)[lo:hi].copy()).to(dev)
        msm_w_ms, _ = timed(lambda: msm_step(scw.data_ptr()), 2, 3)
        sub["g1_msm_witness_like"] = {"log_n": args.log_msm, "n_gpus": world, "ms": msm_w_ms, "mpts_s": n_msm / msm_w_ms / 1e3,
                                      "scalars": "40% 0, 30% 1, 30% uniform"}
        del bases, sc, scw
        if world == 1:
            n_ntt = 1 << args.log_ntt
            v = torch.from_numpy(synthetic.random_scalars(n_ntt, 0x5A410007)).to(dev)
            torch.cuda.synchronize()
            ntt_ms, _ = timed(lambda: ctx.ntt_device(v.data_ptr(), args.log_ntt, za_b200.FFT), max(2, min(K, 5)), 3)
            hbm_peak, _ = peaks()
            sub["fr_ntt"] = {"log_n": args.log_ntt, "ms": ntt_ms, "gelem_s": n_ntt / ntt_ms / 1e6,
                             "hbm_frac": 64.0 * n_ntt / (ntt_ms * 1e-3) / 1e9 / hbm_peak,
                             "imad_frac": IMAD_PER_MODMUL * (n_ntt / 2) * args.log_ntt / (ntt_ms * 1e-3) / ipk,
                             "order": "natural in, natural out, forward"}
            del v
        if world == 1 and args.log_setup > 0:
            try:
                # the rows either side of create_proof on a REAL key: generate_parameters on the
                # GPU -> Parameters::write stream -> Parameters::read(checked) -> create_proof (host buffers) -> verify_proof
                log_s = min(log_m, args.log_setup)
                nc_s = (1 << log_s) - 2
                ni2, na2, ptr2, var2, coeff2, in2, aux2 = synthetic.mul_chain(nc_s, x0=0x5A410002)
                circ2 = za_b200.Circuit(ctx, ni2, na2, ptr2, var2, coeff2)
                t0 = time.perf_counter()
                blob = za_b200.generate_parameters(ctx, circ2, 0x5A410011, 0x5A410012, 0x5A410013, 0x5A410014, 0x5A410015)
                t1 = time.perf_counter()
                pk2 = za_b200.Parameters.read(ctx, blob, checked=True)
                t2 = time.perf_counter()
                pr2 = za_b200.create_proof(ctx, pk2, circ2, in2, aux2, R_FIXED, S_FIXED)      # first call: buffers are allocated
                t3 = time.perf_counter()
                pr2 = za_b200.create_proof(ctx, pk2, circ2, in2, aux2, R_FIXED, S_FIXED)
                t4 = time.perf_counter()
                pub = [int.from_bytes(in2[i].tobytes(), "little") for i in range(1, ni2)]
                ok = za_b200.verify_proof(pk2.vk(), pr2, pub)
                t5 = time.perf_counter()
                bad = za_b200.verify_proof(pk2.vk(), pr2, [(pub[0] + 1) % R_MOD] + pub[1:])
                sub["real_key_pipeline"] = {"log_m": log_s, "parameters_bytes": len(blob), "generate_parameters_ms": (t1 - t0) * 1e3,
                                            "parameters_read_checked_ms": (t2 - t1) * 1e3, "create_proof_host_buffers_ms": (t4 - t3) * 1e3,
                                            "verify_proof_host_ms": (t5 - t4) * 1e3, "proof_verifies": bool(ok),
                                            "wrong_public_input_rejected": bool(not bad), "timing": "host wall clock, one call each"}
                del pk2, circ2, blob
            except Exception as e:      # a sub-metric must never cost the headline line
                sub["real_key_pipeline"] = {"error": repr(e)[:200]}
        if world == 1 and not args.no_config3:
            # config 3 (BASELINE.json): the EdDSA-MiMC verification statement of circomlib as a hand-built R1CS
            # (tests/eddsa_circuit.py, 7 429 constraints, bit-heavy witness) through setup -> read -> prove -> verify
            try:
                from tests import eddsa_circuit as E3
                (ni3, na3, ptr3, var3, coeff3, in3, aux3), info3 = E3.eddsa_mimc_verifier(**E3.KAT)
                circ3 = za_b200.Circuit(ctx, ni3, na3, ptr3, var3, coeff3)
                pk3 = za_b200.Parameters.read(ctx, za_b200.generate_parameters(ctx, circ3, 0x5A410031, 0x5A410032, 0x5A410033, 0x5A410034, 0x5A410035), checked=True)
                for _ in range(3):
                    pr3 = za_b200.create_proof(ctx, pk3, circ3, in3, aux3, R_FIXED, S_FIXED)
                t0 = time.perf_counter()
                reps3 = 10
                for _ in range(reps3):
                    pr3 = za_b200.create_proof(ctx, pk3, circ3, in3, aux3, R_FIXED, S_FIXED)
                t1 = time.perf_counter()
                pub3 = [int.from_bytes(in3[i].tobytes(), "little") for i in range(1, ni3)]
                sub["config3_eddsa_mimc"] = {"constraints": info3["constraints"], "num_aux": na3, "create_proof_host_buffers_ms": (t1 - t0) * 1e3 / reps3,
                                             "proof_verifies": bool(za_b200.verify_proof(pk3.vk(), pr3, pub3)),
                                             "wrong_message_rejected": bool(not za_b200.verify_proof(pk3.vk(), pr3, pub3[:-1] + [1235])),
                                             "note": "statement of circomlib EdDSAMiMCVerifier on the vector of za_test/eddsamimc.za, R1CS built by hand (not za's compiler output)"}
                del pk3, circ3
            except Exception as e:      # a sub-metric must never cost the headline line
                sub["config3_eddsa_mimc"] = {"error": repr(e)[:200]}
        if line is not None:
            line["submetrics"] = sub

    # ---- CPU baseline beside it (rank 0, N = 1): bounded sample of the same workload; also the full-size checker
    if rank == 0 and not args.no_cpu_baseline:
        from tests import oracle as O
        cores = os.cpu_count() or 1
        log_s = min(log_m, args.cpu_log_m)
        nc_s, cs_s, counts_s, _ = cpu_prove_workload(log_s)
        ocs = O.CS(cs_s[0], cs_s[1], cs_s[2], cs_s[3], cs_s[4])
        prm = O.Params.synthetic(counts_s["ic"], counts_s["h"], counts_s["l"], counts_s["a"], counts_s["b_g1"], counts_s["b_g2"], threads=cores)
        t = time.perf_counter()
        rc, cpu_proof = prm.create_proof(ocs, cs_s[5], cs_s[6], R_FIXED, S_FIXED, threads=cores)
        cpu_ms = (time.perf_counter() - t) * 1e3
        assert rc == 0
        if log_s == log_m:
            parity = cpu_proof == proof
        else:
            pr2 = za_b200.Prover(list(range(world)))
            pr2.synthetic_pk(counts_s["ic"], counts_s["h"], counts_s["l"], counts_s["a"], counts_s["b_g1"], counts_s["b_g2"])
            pr2.set_circuit(cs_s[0], cs_s[1], cs_s[2], cs_s[3], cs_s[4])
            parity = cpu_proof == pr2.create_proof(cs_s[5], cs_s[6], R_FIXED, S_FIXED)
            pr2.close()
        scale = float(1 << (log_m - log_s))
        line["cpu_baseline"] = {"value": cpu_ms * scale, "unit": "ms", "cores": cores, "kind": "port",
                                "sample": f"one create_proof, domain 2^{log_s}, {cores} threads" +
                                          (f", extrapolated x{int(scale)}" if scale != 1 else ""),
                                "note": "bellman-algorithm restatement (oracle/), not bellman itself",
                                "proof_matches_gpu": bool(parity)}
        if not parity:
            print("PARITY FAILURE: CPU oracle proof differs from the GPU proof", file=sys.stderr)

    if rank == 0:
        if os.environ.get("ZA_BENCH_PRINT_PROOF"):
            line["proof_hex"] = proof.hex()
        emit(line)
    if dist is not None:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()
    ctx.close()
    if prover is not None:
        prover.close()
    return 0


_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else (NCCL banners, library chatter) to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="za_b200", choices=["za_b200", "reference"])
    ap.add_argument("--log-m", dest="log_m", type=int, default=20, help="log2 of the evaluation domain of the prove workload")
    ap.add_argument("--log-msm", dest="log_msm", type=int, default=24)
    ap.add_argument("--log-ntt", dest="log_ntt", type=int, default=24)
    ap.add_argument("--no-config3", dest="no_config3", action="store_true", help="skip the config-3 (EdDSA-MiMC) sub-metric")
    ap.add_argument("--log-setup", dest="log_setup", type=int, default=20, help="domain of the real-key pipeline sub-metric (0 = skip)")
    ap.add_argument("--cpu-log-m", dest="cpu_log_m", type=int, default=20, help="domain of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the MSM / NTT sub-metrics")
    ap.add_argument("--ref-log-m", dest="ref_log_m", type=int, default=0, help="--impl reference: domain of the timed sample (0 = the true workload)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
