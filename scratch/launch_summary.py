"""Per-kernel durations of the LAST multiexp in an ncu launch-list csv (gpu__time_duration.sum)."""
import csv, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
    idx = [i for i, r in enumerate(rows) if 'digits_hist' in r[ki]]
    tot = {}
    order = []
    for r in rows[idx[-1]:]:
        name = r[ki].split('(')[0].replace('void ', '')[:70]
        if name not in tot: order.append(name)
        tot.setdefault(name, []).append(float(r[vi]) / 1e3)
    print(f)
    for n in order:
        print("  %-72s n=%2d  sum %9.1f us   %s" % (n, len(tot[n]), sum(tot[n]), " ".join("%.0f" % x for x in tot[n][:6])))
