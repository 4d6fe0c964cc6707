"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launch count per kernel name."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iu = hdr.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    v_us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v)
    name = r[ik].split("(")[0][:90]
    tot[name] += v_us
    cnt[name] += 1
total = sum(tot.values())
print("total %.1f ms over %d launches" % (total / 1e3, sum(cnt.values())))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%9.2f ms %5.1f%% %6d x  %s" % (v / 1e3, 100 * v / total, cnt[k], k))
