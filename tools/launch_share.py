#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time and share
of ONE steady-state step (the launches between two consecutive gray-conversion kernels)."""
import collections, csv, re, sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
iname, ival, imet = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
launches = [(re.sub(r"\(.*", "", r[iname]), float(r[ival].replace(",", ""))) for r in rows[1:] if r[imet] == "gpu__time_duration.sum"]
unit = rows[1][hdr.index("Metric Unit")]
scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
marks = [i for i, (n, _) in enumerate(launches) if "gray_from_bgr" in n]
if len(marks) >= 3:
    lo, hi = marks[-2], marks[-1]          # the last complete step
else:
    lo, hi = 0, len(launches)
step = launches[lo:hi]
agg, cnt = collections.OrderedDict(), collections.Counter()
for n, v in step:
    agg[n] = agg.get(n, 0.0) + v * scale
    cnt[n] += 1
total = sum(agg.values())
print(f"launches captured {len(launches)}; one step = launches [{lo}, {hi}) = {len(step)} launches, {total:.1f} us (cold-cache, serialised)")
for n, v in sorted(agg.items(), key=lambda t: -t[1]):
    print(f"  {v:9.1f} us  {v / total * 100:5.1f}%  x{cnt[n]:<3d} {n}")
