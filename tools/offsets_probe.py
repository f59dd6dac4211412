#!/usr/bin/env python3
"""Where the time of TextToIdsWithOffsetsBatchCsr goes: pageable against pinned caller buffers, per model.
python tools/offsets_probe.py [lines] [once]   ("once": a single call per model, for a launch list under ncu)"""
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
once = len(sys.argv) > 2
base = [l for l in read_lines("test.txt") if l]
lines = [base[i % len(base)] for i in range(n)]
buf, offs = bf.make_csr(lines)
mb = len(buf) / 1e6
L = bf.lib()
cap = int(np.minimum(np.diff(offs) + 1, 128).sum())


def run(h, unk, text, out, off):
    r = L.TextToIdsWithOffsetsBatchCsr(ctypes.c_void_p(h), text.ctypes.data, offs.ctypes.data, n, out[0].ctypes.data, out[1].ctypes.data,
                                       out[2].ctypes.data, cap, off.ctypes.data, 128, unk)
    assert r > 0, (r, bf.last_error())
    return r


for name, unk in (("bert_base_tok.bin", 100), ("xlm_roberta_base.bin", 3)):
    h = bf.load_model(model_path(name))
    off = np.zeros(n + 1, np.int64)
    out = np.zeros((3, cap), np.int32)                     # pageable, already touched
    if once:
        print(name, run(h, unk, buf, out, off), "ids")
        bf.free_model(h)
        continue
    run(h, unk, buf, out, off)
    t0 = time.perf_counter()
    for _ in range(3):
        total = run(h, unk, buf, out, off)
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {total} ids, pageable touched buffers: {mb / dt:.1f} MB/s ({dt * 1e3:.1f} ms)")
    p_text = torch.from_numpy(buf.copy()).pin_memory().numpy()
    p_out = torch.zeros((3, cap), dtype=torch.int32).pin_memory().numpy()
    run(h, unk, p_text, p_out, off)
    t0 = time.perf_counter()
    for _ in range(3):
        run(h, unk, p_text, p_out, off)
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: pinned buffers: {mb / dt:.1f} MB/s ({dt * 1e3:.1f} ms)")
    assert (p_out[:, :total] == out[:, :total]).all()
    bf.free_model(h)
