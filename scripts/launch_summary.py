"""Aggregates an ncu launch list (csv of gpu__time_duration.sum) per kernel: python scripts/launch_summary.py in.csv title > out.md"""
import collections
import csv
import sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
agg = collections.OrderedDict()
for r in rows:
    name = r["Kernel Name"]
    short = name.split("(")[0]
    short = short[-70:]
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"])
ours = {k: v for k, v in agg.items() if "serb::" in k or k.startswith("void serb") or "serb" in k}
tot_ours = sum(v[1] for v in ours.values()) or 1.0
print(f"# {sys.argv[2] if len(sys.argv) > 2 else 'ncu launch list'}\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: shares matter, not absolutes).\n")
print("| kernel (libser_b200) | launches | total us | share of libser_b200 time |\n|---|---|---|---|")
for k, a in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1] / 1e3:.1f} | {100 * a[1] / tot_ours:.1f}% |")
others = sum(v[1] for k, v in agg.items() if k not in ours)
print(f"\nOther kernels in the process (torch: synthetic input generation, outside the timed region): {others / 1e3:.1f} us in "
      f"{sum(v[0] for k, v in agg.items() if k not in ours)} launches.")
