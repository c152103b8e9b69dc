"""Brief view of an ncu capture: python scratch/ncu_brief.py raw.csv [source.csv]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:90])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k); print("   %-70s %s %s" % (k, r[i], units[i]))
    st = []
    for i, h in enumerate(hdr):
        m = re.match(r"smsp__pcsamp_warps_issue_stalled_(\w+)$", h)
        if m and not h.endswith("_not_issued"):
            try: st.append((float(r[i]), m.group(1)))
            except ValueError: pass
    tot = sum(s for s, _ in st) or 1
    print("   stalls:", ", ".join("%s %.1f%%" % (n, 100 * s / tot) for s, n in sorted(st, reverse=True)[:9]))
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2])))
    hdr = rows[1]
    si = hdr.index("# Samples"); src = hdr.index("Source"); ie = hdr.index("Instructions Executed"); lsb = hdr.index("stall_long_sb")
    data = rows[2:]
    T = sum(int(r[si]) for r in data) or 1
    bounds = [-1] + [i for i, r in enumerate(data) if 'RET' in r[src] or 'EXIT' in r[src]]
    for k in range(len(bounds) - 1):
        a, b = bounds[k] + 1, bounds[k + 1] + 1
        s = sum(int(r[si]) for r in data[a:b]); e = sum(int(r[ie]) for r in data[a:b]); l = sum(int(r[lsb]) for r in data[a:b])
        print("fn %5d-%5d samples %6d (%.1f%%) long_sb %d executed %d" % (a, b, s, 100 * s / T, l, e))
    top = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]
    for i in sorted(top):
        r = data[i]; print(i, r[src][:64].ljust(64), r[si], "long_sb", r[lsb], "exec", r[ie])
