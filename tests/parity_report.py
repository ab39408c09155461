"""Collects the achieved parity numbers of the GPU tests in gpurun_out/r02_parity.json (the round script copies the file to
profiles/): pytest -q swallows prints, the judge wants the numbers."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def report(key, **rec):
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    p = os.path.join(out, "r02_parity.json")
    try:
        prev = json.load(open(p))
    except Exception:  # noqa: BLE001
        prev = {}
    prev[key] = rec
    json.dump(prev, open(p, "w"), indent=1, sort_keys=True)
