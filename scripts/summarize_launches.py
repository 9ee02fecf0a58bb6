"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total, average, share.
    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows:
    if r is hdr or len(r) <= max(ik, iv) or r[ik] == "Kernel Name":
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    unit = r[iu]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    name = re.sub(r"\(.*$", "", r[ik]).replace("void ", "").replace("mcx::<unnamed>::", "").replace("mcx::", "").replace("unnamed>::", "").replace("(int)", "").replace("(bool)", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.1f | %.1f%% |" % (name, n, us, us / n, 100 * us / tot))
