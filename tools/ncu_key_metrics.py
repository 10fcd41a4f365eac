"""One line per captured launch from `ncu --page raw --csv` files: the metrics profiles/*_ncu_key_metrics.txt quotes.
usage: python tools/ncu_key_metrics.py "<section title>" raw.csv [kernel-substring] [max-rows]  >> profiles/..._key_metrics.txt"""
import csv
import sys

title, raw = sys.argv[1], sys.argv[2]
want = sys.argv[3] if len(sys.argv) > 3 else ""
limit = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
COLS = [
    ("time", "gpu__time_duration.sum"),
    ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("issue%", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("dram_rd", "dram__bytes_read.sum"),
    ("dram_wr", "dram__bytes_write.sum"),
    ("l1_lsu_wavefronts%", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("l2%", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
]
print(f"## {title}")
print("kernel | " + " | ".join(c for c, _ in COLS))
n = 0
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if want not in name:
        continue
    cells = []
    for c, k in COLS:
        if k not in idx:
            cells.append("-")
            continue
        v, u = r[idx[k]], units[idx[k]]
        try:
            v = f"{float(v.replace(',', '')):.4g}"
        except ValueError:
            pass
        cells.append(v + (f" {u}" if u and u not in ("%", "") and c in ("time", "dram_rd", "dram_wr") else ""))
    print(name[:44] + " | " + " | ".join(cells))
    n += 1
    if n >= limit:
        break
print()
