#!/usr/bin/env python3
"""End-to-end (host pinned buffers -> TextToIdsBatchCsr -> host) throughput of cfg 2 for a few settings of the host
pipeline's tuning knobs; one subprocess per setting (the knobs are read once per process).
   python tools/e2e_sweep.py            # on the GPU box
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SETTINGS = [dict(), dict(BLINGFIRE_B200_SLOTS="6", BLINGFIRE_B200_AHEAD="3"), dict(BLINGFIRE_B200_SLOTS="8", BLINGFIRE_B200_AHEAD="4"),
            dict(BLINGFIRE_B200_CHUNK_MB="16"), dict(BLINGFIRE_B200_CHUNK_MB="16", BLINGFIRE_B200_SLOTS="8", BLINGFIRE_B200_AHEAD="4"),
            dict(BLINGFIRE_B200_CHUNK_MB="8", BLINGFIRE_B200_SLOTS="8", BLINGFIRE_B200_AHEAD="5"),
            dict(BLINGFIRE_B200_CHUNK_MB="64", BLINGFIRE_B200_SLOTS="4", BLINGFIRE_B200_AHEAD="2"),
            dict(BLINGFIRE_B200_CHUNK_MB="16", BLINGFIRE_B200_SLOTS="6", BLINGFIRE_B200_AHEAD="2")]
for st in SETTINGS:
    env = dict(os.environ, **st)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--configs", "cfg2", "--no-cpu", "--no-parity", "--steps", "10",
                        "--warmup", "3"], capture_output=True, text=True, env=env)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(json.dumps({"settings": st, "value": round(d["value"], 1), "e2e": round(d["e2e"]["value"], 2), "e2e_ms": round(d["e2e"]["ms_per_step"], 2),
                          "e2e_u16": round(d["e2e_u16"]["value"], 2), "e2e_pageable": round(d["e2e_pageable"]["value"], 2)}), flush=True)
    except Exception as e:
        print("failed", st, e, r.stderr[-500:], flush=True)
