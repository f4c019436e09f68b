"""Per-kernel share of a bench step from an ncu launch list (gpu__time_duration.sum CSV):
    python tools/launch_summary.py gpurun_out/launches_x.csv [n_steps]"""
import csv, re, sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").strip()
    if not (name.startswith("jd::") or name.startswith("tc") or name.startswith("fft::") or name.startswith("conv") or "jd" in name):
        continue
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[-1]) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | avg us | share of our kernel time |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t / n:.2f} | {100 * t / tot:.1f}% |")
print(f"\ntotal {tot:.1f} us over all launches listed")
