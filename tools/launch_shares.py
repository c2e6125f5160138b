"""Per-kernel share of the time of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`).
usage: launch_shares.py launches.csv out.csv"""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).strip()
    tot[name] += float(r[iv].replace(",", "")) * scale[r[iu]]
    cnt[name] += 1
total = sum(tot.values())
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_ms", "share"])
    for k in sorted(tot, key=tot.get, reverse=True):
        w.writerow([k, cnt[k], f"{tot[k]:.3f}", f"{tot[k] / total:.4f}"])
print(open(sys.argv[2]).read())
