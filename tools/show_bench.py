"""Print the key numbers of a bench.py JSON line: python tools/show_bench.py file.json"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("n_gpus", d.get("n_gpus"), "value", d.get("value"), "ms/step", d.get("ms_per_step"))
print("latency", d.get("latency", {}).get("ms_per_forward"), "e2e", d.get("e2e", {}).get("value"),
      "e2e one-at-a-time", d.get("e2e", {}).get("one_at_a_time", {}).get("value"))
print("clocks", d.get("clocks"), "launches/step", d.get("gpu_launches_per_step"), "overflows", d.get("plan_overflows"))
for k in d.get("kernels", [])[:8]:
    print("  ", k)
