"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total, average, share."""
import csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = {}
for r in rows[1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    name = re.sub(r"\(.*", "", r[ki]).replace("plslam::<unnamed>::", "").replace("plslam::", "")
    a = acc.setdefault(name[:100], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in acc.values())
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.1f | %.1f%% |" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
