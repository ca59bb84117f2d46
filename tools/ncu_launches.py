#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total/avg time, share.
Usage: ncu_launches.py launches.csv   (times are cold-cache and serialised: compare SHARES, not absolutes)"""
import collections, csv, re, sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    name = r["Kernel Name"]
    m = re.search(r"pmm_forward_kernel<(\w+), (\d+), (\d+), (\w+), (\w+)>", name)
    if m:
        short = f"pmm_forward_kernel<{m.group(1)},K={m.group(2)},W={m.group(3)},striped={m.group(4)},flush={m.group(5)}>"
    else:
        short = re.sub(r"\(.*", "", name).split("::")[-1]
    d = agg.setdefault(short, [0, 0.0, r["Grid Size"], r["Block Size"]])
    d[0] += 1
    d[1] += float(r["Metric Value"].replace(",", ""))
ours = lambda k: "elementwise" not in k and "probe" not in k
tot = sum(v[1] for k, v in agg.items() if ours(k))
print("time unit:", rows[0]["Metric Unit"], "| share = share of our kernels' total time")
for k, v in agg.items():
    share = f"{v[1] / tot * 100:5.1f}%" if ours(k) else "   (not part of the step)"
    print(f"{k:72s} launches={v[0]:3d} total={v[1]:12.0f} avg={v[1] / v[0]:11.0f} grid={v[2]:>12s} block={v[3]:>11s} share={share}")
