#!/usr/bin/env python3
"""Throughput of the additive batch calls served by the generic lexer engine (thread per document) and of the offsets
batch call, on corpus lines; beside them the per-document symbols on the same lines (a sample).
python tools/words_bench.py [lines]"""
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
base = [l for l in read_lines("test.txt") if l]
lines = [base[i % len(base)] for i in range(n)]
buf, offs = bf.make_csr(lines)
mb = len(buf) / 1e6


def timed(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


dt = timed(lambda: bf.text_to_words_batch((buf, offs), raw=True))
print(f"TextToWordsBatch (default wbd.bin), {n} lines, {mb:.1f} MB: {mb / dt:.1f} MB/s")
dt = timed(lambda: bf.text_to_sentences_batch((buf, offs), raw=True))
print(f"TextToSentencesBatch (default sbd.bin): {mb / dt:.1f} MB/s")
for name, unk in (("bert_base_tok.bin", 100), ("xlm_roberta_base.bin", 3)):
    h = bf.load_model(model_path(name))
    dt = timed(lambda: bf.text_to_ids_with_offsets_batch(h, (buf[: offs[100000]], offs[:100001]), 128, unk), reps=2)
    print(f"TextToIdsWithOffsetsBatch {name}, 100000 lines: {offs[100000] / 1e6 / dt:.1f} MB/s")
    dt = timed(lambda: bf.text_to_ids_with_offsets_batch_csr(h, (buf, offs), 128, unk), reps=2)
    print(f"TextToIdsWithOffsetsBatchCsr {name}, {n} lines: {mb / dt:.1f} MB/s")
    bf.free_model(h)
L = bf.lib()
out = ctypes.create_string_buffer(4096)
sample = lines[:3000]
t0 = time.perf_counter()
for l in sample:
    L.TextToWords(l, len(l), out, 4096)
dt = time.perf_counter() - t0
print(f"TextToWords per document (3000 lines): {sum(map(len, sample)) / 1e6 / dt:.2f} MB/s")
