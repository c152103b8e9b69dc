"""Hot spots of one kernel section of an `ncu --page source --csv` export.
python scratch/src_hot.py src.csv [section] [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
hdr_idx = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
hi = hdr_idx[sec]
end = hdr_idx[sec + 1] - 1 if sec + 1 < len(hdr_idx) else len(rows)
hdr = rows[hi]
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
si = hdr.index("# Samples"); src = hdr.index("Source"); ie = hdr.index("Instructions Executed")
cols = {k: hdr.index(k) for k in ("stall_long_sb", "stall_wait", "stall_math", "stall_dispatch", "stall_short_sb", "stall_selected", "stall_not_selected", "stall_no_inst", "stall_branch_resolving", "stall_lg")}
def I(x):
    try: return int(x)
    except ValueError: return 0
T = sum(I(r[si]) for r in data) or 1
print("kernel:", rows[hi - 1][:2] if hi else "", "instructions", len(data), "samples", T)
bounds = [-1] + [i for i, r in enumerate(data) if r[src].strip().startswith(("RET", "EXIT")) or " RET" in r[src] or "EXIT" in r[src]]
for k in range(len(bounds) - 1):
    a, b = bounds[k] + 1, bounds[k + 1] + 1
    s = sum(I(r[si]) for r in data[a:b])
    if s * 200 < T: continue
    e = sum(I(r[ie]) for r in data[a:b])
    print("region %5d-%5d samples %7d (%.1f%%) executed %d  " % (a, b, s, 100 * s / T, e) + " ".join("%s %.1f%%" % (n[6:], 100 * sum(I(r[c]) for r in data[a:b]) / max(s, 1)) for n, c in cols.items()))
top = sorted(range(len(data)), key=lambda i: -I(data[i][si]))[:top_n]
for i in sorted(top):
    r = data[i]
    print(i, r[src][:70].ljust(70), r[si], "lsb", r[cols["stall_long_sb"]], "wait", r[cols["stall_wait"]], "exec", r[ie])
