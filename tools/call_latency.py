#!/usr/bin/env python3
"""Latency of the per-document drop-in calls (the reference's own symbols) on short strings: TextToIds for three models,
TextToWords with the default model.  Context: the reference does a 42-byte line in ~2 us on one CPU thread; a GPU call
costs a launch and two copies.  python tools/call_latency.py"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import blingfire_b200 as bf  # noqa: E402
from _common import model_path, read_lines  # noqa: E402

torch.cuda.set_device(0)
torch.zeros(1, device="cuda")
L = bf.lib()
lines = [l for l in read_lines("test.txt") if 20 <= len(l) <= 120][:2000]
ids = np.zeros(128, np.int32)
for name, unk in (("bert_base_tok.bin", 100), ("gpt2.bin", 0), ("xlm_roberta_base.bin", 3)):
    h = bf.load_model(model_path(name))
    for l in lines[:200]:
        L.TextToIds(ctypes.c_void_p(h), l, len(l), ids.ctypes.data, 128, unk)
    t0 = time.perf_counter()
    for l in lines:
        L.TextToIds(ctypes.c_void_p(h), l, len(l), ids.ctypes.data, 128, unk)
    dt = time.perf_counter() - t0
    print(f"TextToIds {name}: {1e6 * dt / len(lines):.1f} us per call, {sum(map(len, lines)) / dt / 1e6:.2f} MB/s")
    bf.free_model(h)
out = ctypes.create_string_buffer(1024)
for l in lines[:200]:
    L.TextToWords(l, len(l), out, 1024)
t0 = time.perf_counter()
for l in lines:
    L.TextToWords(l, len(l), out, 1024)
dt = time.perf_counter() - t0
print(f"TextToWords (default model): {1e6 * dt / len(lines):.1f} us per call, {sum(map(len, lines)) / dt / 1e6:.2f} MB/s")
