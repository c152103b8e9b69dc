"""Summarise ncu captures (gpurun_out/*.ncu-rep, launch list csv) into profiles/<tag>_*.md / .csv"""
import csv, subprocess, sys, io, re, collections, os
tag = sys.argv[1]
out = []
KEYS = [("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of max"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy (IMAD) pipe active %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("smsp__sass_inst_executed_op_local_ld.sum", "local loads"), ("smsp__sass_inst_executed_op_local_st.sum", "local stores"),
        ("smsp__average_warp_latency_per_inst_issued.ratio", "cycles between issues per warp")]
def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]
for rep, title in ((f"gpurun_out/{tag}_acc_g1.ncu-rep", "msm_accumulate_kernel<Fq> — 2^22-point G1 MSM, c=16 (prof_target.py g1)"),
                   (f"gpurun_out/{tag}_acc_g2.ncu-rep", "msm_accumulate_kernel<Fq2> — 2^20-point G2 MSM (prof_target.py g2)"),
                   (f"gpurun_out/{tag}_ntt.ncu-rep", "ntt_pass_kernel — 2^22 forward NTT, 3 passes (prof_target.py ntt)")):
    if not os.path.exists(rep): continue
    hdr, units, rows = raw(rep)
    out.append(f"### {title}\n")
    out.append("`ncu --set full --clock-control none --import-source on` (one GPU, cold-cache replays)\n")
    for r in rows:
        name = r[hdr.index("Kernel Name")]
        out.append(f"**{name[:110]}**\n\n| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k); out.append(f"| {label} (`{k}`) | {r[i]} {units[i]} |")
        stalls = []
        for i, h in enumerate(hdr):
            m = re.match(r"smsp__pcsamp_warps_issue_stalled_(\w+)$", h)
            if m and not h.endswith("_not_issued"):
                try: stalls.append((float(r[i]), m.group(1)))
                except ValueError: pass
        tot = sum(s for s, _ in stalls) or 1
        top = sorted(stalls, reverse=True)[:6]
        out.append("| top stall reasons (pc samples) | " + ", ".join(f"{n} {100*s/tot:.0f}%" for s, n in top) + " |")
        out.append("")
# launch list
ll = f"gpurun_out/{tag}_launches.csv"
if os.path.exists(ll):
    rows = list(csv.reader(open(ll)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi: continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("za::", "")
        v = float(r[vi].replace(",", "")); u = r[ui]
        v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out.append("### Launch list — `bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub` (5 proofs at 2^20 + setup)\n")
    out.append("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised: compare SHARES.\n")
    out.append("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
        out.append(f"| {k[:80]} | {v[0]} | {v[1]:.3f} | {100*v[1]/tot:.1f}% | {v[1]/v[0]*1e3:.1f} |")
    out.append(f"| **total** | | {tot:.3f} | | |\n")
    os.makedirs("profiles", exist_ok=True)
    with open(f"profiles/{tag}_launches.csv", "w") as f:
        f.write(open(ll).read())
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out)[:6000])
