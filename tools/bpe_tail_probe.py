#!/usr/bin/env python3
"""Where does the gpt2 (cfg 3) kernel time go: the same corpus with and without the documents that
hold a very long whitespace-free run (those leave the streaming BPE path).  Development probe."""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)


def longest_run(text, offs):
    white = (text <= 0x20) | (text == 0xA0)
    idx = np.arange(len(text), dtype=np.int64)
    last_white = np.where(white, idx, -1)
    np.maximum.accumulate(last_white, out=last_white)
    run = idx - last_white                       # length of the current non-white run ending here
    # runs do not cross documents: clamp by the distance to the document start
    doc_of = np.searchsorted(offs, idx, side="right") - 1
    run = np.minimum(run, idx - offs[doc_of] + 1)
    out = np.zeros(len(offs) - 1, dtype=np.int64)
    np.maximum.at(out, doc_of, run)
    return out


def main():
    import torch
    import blingfire_b200 as bf
    import corpus
    from _common import model_path
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
    L = bf.lib()
    L.BlingFireB200LastKernelMs.restype = ctypes.c_double
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    text, offs = corpus.gen_docs("EN", n, seed=3, fixed_len=0)
    runs = longest_run(text, offs)
    h = bf.load_model(model_path("gpt2.bin"))
    for name, keep in [("all", np.ones(n, bool)), ("run<=380", runs <= 380), ("run<=120", runs <= 120), ("run<=60", runs <= 60)]:
        docs = [bytes(text[offs[i]:offs[i + 1]]) for i in np.nonzero(keep)[0]]
        t2, o2 = bf.make_csr(docs)
        nb = int(o2[-1])
        h_text = torch.empty(nb + 64, dtype=torch.uint8, pin_memory=True)
        h_text[:nb].copy_(torch.from_numpy(t2))
        h_offs = torch.from_numpy(o2).pin_memory()
        cap = int(np.minimum(np.diff(o2), 4096).sum())
        h_ids = torch.empty(cap + 1, dtype=torch.int32, pin_memory=True)
        h_idoffs = torch.zeros(len(docs) + 1, dtype=torch.int64, pin_memory=True)
        best = 1e9
        for _ in range(3):
            r = L.TextToIdsBatchCsr(ctypes.c_void_p(h), h_text.data_ptr(), h_offs.data_ptr(), len(docs), h_ids.data_ptr(), cap,
                                    h_idoffs.data_ptr(), 4096, 0)
            assert r >= 0
            best = min(best, L.BlingFireB200LastKernelMs())
        print(f"{name}: docs={len(docs)} bytes={nb} kernel_ms={best:.3f} GB/s={nb / best / 1e6:.2f}", flush=True)


if __name__ == "__main__":
    main()
