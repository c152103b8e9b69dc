"""Instruction mix of the kernels in an `ncu --page source --csv` export: python scratch/inst_mix.py src.csv"""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr_idx = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
for n, hi in enumerate(hdr_idx):
    hdr = rows[hi]
    end = hdr_idx[n + 1] - 1 if n + 1 < len(hdr_idx) else len(rows)
    data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
    src = hdr.index("Source"); ie = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
    ops = collections.Counter(); samp = collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[src])
        if not m: continue
        op = m.group(2)
        key = op if op.startswith("IMAD") else op.split(".")[0]
        try: ops[key] += int(r[ie]); samp[key] += int(r[si])
        except ValueError: pass
    tot = sum(ops.values()); ts = sum(samp.values()) or 1
    print("kernel", n, rows[hi - 1][:1] if hi else "", "total warp insts", tot)
    for k, v in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
        print("   %-24s %12d %5.1f%%  samples %5.1f%%" % (k, v, 100 * v / tot, 100 * samp[k] / ts))
