#!/usr/bin/env python
"""Per-kernel totals and shares of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv python bench.py
--no-graph ...`): python scripts/launch_shares.py X.csv [steps]  ->  markdown table (times are cold-cache and serialised:
shares, not absolutes)."""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].replace("void ", "").split("(")[0][:80]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches/step | us/step | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0] / steps:.1f} | {v[1] / 1e3 / steps:.1f} | {100 * v[1] / tot:.1f} % |")
print(f"| total | {sum(v[0] for v in agg.values()) / steps:.1f} | {tot / 1e3 / steps:.1f} | |")
