"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py`: the launches of the
last complete step (between two logmel launches), per-launch microseconds and share, plus per-kernel totals.
usage: python tools/launch_summary.py gpurun_out/launches_final.csv > profiles/<name>_summary.txt"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, mi, gi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Metric Unit")
seq = []
for r in rows[hi + 1:]:
    if len(r) <= mi:
        continue
    t = float(r[mi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    seq.append((re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("bt::", ""), r[gi], t))
idx = [i for i, (k, g, t) in enumerate(seq) if "logmel" in k]
st, en = idx[-2], idx[-1]
step = [x for x in seq[st:en] if not x[0].startswith("at::")]
tot = sum(t for _, _, t in step)
print("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 2 --no-cpu-baseline")
print(f"# one bench step (64 x 30 s clips, 128 chunks, one wave), {len(step)} launches, sum {tot / 1e3:.2f} ms (cold-cache, serialised)")
print("# idx  kernel  grid  us  share")
for i, (k, g, t) in enumerate(step):
    print(f"{i:3d} {k:46s} {g:18s} {t:9.1f} {100 * t / tot:5.1f}%")
agg = collections.OrderedDict()
for k, g, t in step:
    k2 = re.sub(r"<.*", "", k)
    agg[k2] = agg.get(k2, 0.0) + t
print("# per kernel")
for k, t in sorted(agg.items(), key=lambda x: -x[1]):
    print(f"{k:30s} {t / 1e3:8.2f} ms {100 * t / tot:5.1f}%")
