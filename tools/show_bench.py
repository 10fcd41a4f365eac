"""Print the headline numbers and per-kernel-class ms/step of bench JSON lines: python tools/show_bench.py file..."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.load(open(path))
    except Exception as e:
        print(path, "unreadable:", e)
        continue
    r = d.get("roofline", {})
    print(f"{path}: value {d['value']:.0f} clips/s, {d['ms_per_step']:.2f} ms/step, e2e {d['e2e']['value']:.0f}, "
          f"attn {r.get('avg_launch_ms') or 0:.3f} ms/launch frac {r.get('frac') or 0:.3f}")
    print("   " + ", ".join(f"{k} {v['ms_per_step']}/{v['launches_per_step']}" for k, v in d.get("kernel_time_shares", {}).items()))
