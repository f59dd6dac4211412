#!/usr/bin/env python3
"""A small workload for compute-sanitizer (memcheck / racecheck / initcheck) on the GPU box:
   compute-sanitizer --tool memcheck python tools/sanitize_small.py [model ...]
Runs a few thousand documents (corpus lines, edge cases, long and invalid documents) through the batch C ABI of each
model, twice (the second pass meets the words the first one added to the word table), and checks them against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import blingfire_b200 as bf  # noqa: E402
from _common import Oracle, model_path, read_lines  # noqa: E402

MODELS = sys.argv[1:] or ["bert_base_tok.bin", "gpt2.bin", "xlm_roberta_base.bin"]
UNK = {"bert_base_tok.bin": 100, "gpt2.bin": 0, "xlm_roberta_base.bin": 3}

torch.cuda.set_device(0)
torch.zeros(1, device="cuda")
lines = read_lines("test.txt", drop_empty=False)[:1500] + read_lines("test.multi.txt", drop_empty=False)[:1500]
docs = [b" ".join(lines[i:i + 4]) for i in range(0, len(lines), 4)]
docs += [b"", b" ", b"\xef\xbb\xbf", b"\xef\xbb\xbfhello", b"abc \xff def", b"\xe6\x88", b"hello\x00world", b"a" * 400,
         b"a" * 1000 + b" " + b"b" * 700, "我".encode() * 900, b"." * 700, ("word " * 300).encode(), b"x" * 573,
         b"supercalifragilisticexpialidocious antidisestablishmentarianism " * 20, b" ".join(lines[:200])]
o = Oracle()
for name in MODELS:
    h = bf.load_model(model_path(name))
    ho = o.load(model_path(name))
    buf, offs = bf.make_csr(docs)
    unk = UNK.get(name, 100)
    for rep in range(2):
        ids, counts = bf.text_to_ids_batch(h, (buf, offs), 512, unk)
        _, oids, ocounts = o.batch(ho, buf, offs, 512, unk, threads=8)
        assert (counts == ocounts).all(), name
        mask = np.arange(512)[None, :] < counts[:, None]
        assert (ids[mask] == oids[mask]).all(), name
    bf.free_model(h)
    print(name, "ok:", len(docs), "documents,", int(counts.sum()), "ids")
