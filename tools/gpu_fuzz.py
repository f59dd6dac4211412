#!/usr/bin/env python3
"""Differential fuzz of the GPU batch path against the oracle, at a scale the pytest suite does not take: random joins of
corpus lines, truncations, injected bytes, whitespace-free runs, runs of one character, random bytes; several models, two
passes per model (the second one meets the words / segments the first one added to the run-time tables), several UnkIds and
output caps.   python tools/gpu_fuzz.py [seed] [docs per model]      (on a GPU box)"""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import blingfire_b200 as bf  # noqa: E402
from _common import Oracle, model_path, read_lines  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ndocs = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
rng = random.Random(seed)
lines = read_lines("test.multi.txt") + read_lines("test.txt")


def docs(n):
    out = []
    for _ in range(n):
        k = rng.choice([1, 1, 2, 3, 6, 12, 30])
        d = b" ".join(rng.choice(lines) for _ in range(k))
        r = rng.random()
        if r < 0.15:
            d = d[:rng.randint(0, len(d))]
        elif r < 0.25:
            p = rng.randint(0, len(d)); d = d[:p] + bytes([rng.randint(0, 255)]) + d[p:]
        elif r < 0.30:
            d = bytes(rng.randint(0, 255) for _ in range(rng.randint(1, 80)))
        elif r < 0.33:
            d = d.replace(b" ", b"")
        elif r < 0.36:
            d = bytes([rng.choice(b"=-.a ")]) * rng.randint(1, 1500)
        elif r < 0.40:
            w = rng.choice(lines).split()
            d = b" ".join(rng.choice(w) if w else b"x" for _ in range(rng.randint(1, 200)))      # few distinct words, many times
        out.append(d)
    return out


torch.cuda.set_device(0)
torch.zeros(1, device="cuda")
o = Oracle()
MODELS = [("bert_base_tok.bin", 100), ("bert_multi_cased.bin", 100), ("bert_chinese.bin", 100), ("gpt2.bin", 0), ("roberta.bin", 3),
          ("xlm_roberta_base.bin", 3), ("xlnet.bin", 0)]
t0 = time.time()
total = 0
for name, unk in MODELS:
    h = bf.load_model(model_path(name))
    ho = o.load(model_path(name))
    D = docs(ndocs)
    buf, offs = bf.make_csr(D)
    for max_ids, u in ((1024, unk), (1024, unk), (rng.choice([1, 3, 17, 64]), 7777)):
        ids, counts = bf.text_to_ids_batch(h, (buf, offs), max_ids, u)
        _, oids, ocounts = o.batch(ho, buf, offs, max_ids, u, threads=64)
        bad = np.nonzero(counts != ocounts)[0]
        assert len(bad) == 0, (name, "count", int(bad[0]), D[bad[0]][:80], int(counts[bad[0]]), int(ocounts[bad[0]]))
        mask = np.arange(max_ids)[None, :] < counts[:, None]
        diff = np.nonzero(((ids != oids) & mask).any(axis=1))[0]
        assert len(diff) == 0, (name, "ids", int(diff[0]), D[diff[0]][:80])
        total += int(counts.sum())
    bf.free_model(h)
    o.free(ho)
    print(f"{name}: {len(D)} documents x 3 passes ok ({time.time() - t0:.0f} s)", flush=True)
print(f"OK: {len(MODELS)} models, {total} ids compared, seed {seed}")
