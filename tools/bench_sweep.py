#!/usr/bin/env python
"""Runs bench.py several times with different env/flags and prints one compact line each.

    python tools/bench_sweep.py "KD_EXP=0" "KD_EXP=1 --lanes 148" ...
Each argument: space separated; NAME=VALUE tokens go to the environment, the rest to bench.py.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["--steps", "1", "--warmup", "1", "--no-e2e", "--no-cpu-baseline"]

for spec in sys.argv[1:]:
    env = dict(os.environ)
    extra = []
    for tok in spec.split():
        if "=" in tok and not tok.startswith("--"):
            k, v = tok.split("=", 1)
            env[k] = v
        else:
            extra.append(tok)
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + BASE + extra,
                             env=env, capture_output=True, text=True, timeout=150)
        d = json.loads(out.stdout.strip().splitlines()[-1])
        c = d["roofline"]["counters"]
        f = max(1, c["frames"])
        per = {k: int(c[k] / f) for k in c if k.startswith("cycles") or k in
               ("tokens_out", "emit_arcs", "slots_claimed", "candidates", "arcs_evaluated")}
        print(f"[{spec}] fps={int(d['value'])} kernel_ms={d['roofline']['kernel_ms']:.1f} "
              f"GB/s={d['roofline']['achieved']:.0f} per-frame={per}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"[{spec}] FAILED: {e!r}", flush=True)
        if 'out' in dir():
            print(out.stderr[-600:], flush=True)
