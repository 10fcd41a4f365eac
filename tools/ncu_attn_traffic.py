"""profiles/attn_traffic.json from an `ncu --set full` capture of the time-attention kernel.

  ncu --set full --clock-control none --import-source on -k regex:attn_tc48_kernel -s 2 -c 2 \
      -o gpurun_out/r2_attn_full python tools/prof_step.py 64 1
  ncu -i gpurun_out/r2_attn_full.ncu-rep --page raw --csv > gpurun_out/r2_attn_full_raw.csv
  python tools/ncu_attn_traffic.py gpurun_out/r2_attn_full_raw.csv

With 64 x 30 s clips in one wave a step has 3 frontend launches (grid 49152 CTAs) followed by 6 main-layer
launches (24576 CTAs); `-s 2 -c 2` captures the last frontend launch and the first main-layer launch.
bench.py reads dram_bytes_per_launch_mean for the `roofline.traffic` field."""
import csv
import json
import os
import sys

raw = sys.argv[1]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}


def num(r, k):
    return float(r[idx[k]].replace(",", ""))


def to_bytes(r, k):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[idx[k]]]
    return num(r, k) * scale


launches = []
for r in rows[2:]:
    if "attn_tc" not in r[idx["Kernel Name"]]:
        continue
    t_unit = units[idx["gpu__time_duration.sum"]]
    t = num(r, "gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(t_unit, 1e-6)
    launches.append({
        "kernel": r[idx["Kernel Name"]][:60],
        "grid": r[idx["launch__grid_size"]],
        "time_ms": t,
        "dram_read": to_bytes(r, "dram__bytes_read.sum"),
        "dram_write": to_bytes(r, "dram__bytes_write.sum"),
        "tensor_pct": num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "xu_pct": num(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "issue_pct": num(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        "regs": r[idx["launch__registers_per_thread"]],
        "inst": num(r, "smsp__inst_executed.sum"),
        "cycles": num(r, "sm__cycles_elapsed.max"),
    })
front = max(launches, key=lambda d: float(d["grid"]))
main = min(launches, key=lambda d: float(d["grid"]))
mean = (3 * (front["dram_read"] + front["dram_write"]) + 6 * (main["dram_read"] + main["dram_write"])) / 9
out = {
    "dram_bytes_per_launch_mean": mean,
    "frontend_launch": front,
    "main_launch": main,
    "source": "ncu --set full --clock-control none -k regex:attn_tc48_kernel -s 2 -c 2 python tools/prof_step.py 64 1 "
              "(64 x 30 s clips, one 128-chunk wave): dram__bytes_read.sum + dram__bytes_write.sum, mean over the "
              "3 frontend + 6 main-layer launches of a step",
    "algorithmic_bytes_per_launch": {"frontend": 1572864000, "main": 786432000},
}
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "attn_traffic.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
