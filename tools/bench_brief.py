#!/usr/bin/env python
"""Runs bench.py with the given extra arguments / environment and prints a one-line digest of its JSON line."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "5", "--no-cpu-baseline",
                    *sys.argv[2:]], capture_output=True, text=True)
try:
    d = json.loads(r.stdout.strip().splitlines()[-1])
    rf = d.get("roofline") or {}
    print(tag, "fps", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "attn_ms", rf.get("ms_per_launch_by_layer"),
          "frac", rf.get("frac"), "launches/step", d["gpu_launches"] // d["steps"])
except Exception as e:  # noqa: BLE001
    print(tag, "FAILED", e, r.stderr[-800:])
