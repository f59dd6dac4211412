#!/usr/bin/env python3
"""Side measurements for BASELINE.md's results table: the configurations other than the bench line
(cfg 1 TextToWords, cfg 3 gpt2, cfg 4 xlm-r; cfg 2 through the same path for comparison), through
the host-pointer C ABI (TextToIdsBatchCsr) with PINNED host buffers, next to the reference's CPU path
on a sample.  Reports, per configuration, the end-to-end GB/s (copies included) and the kernel-only
GB/s (device time of the tokenization kernels, CUDA events inside the library).
Not the contract benchmark (that is bench.py on cfg 2)."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)


def ref_cpu(model, text, offs, n, max_ids, unk, threads):
    L = ctypes.CDLL(os.path.join(ROOT, "oracle", "librefdriver.so"))
    L.ref_time_batch.restype = ctypes.c_double
    L.ref_time_batch.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    ref = os.path.join(ROOT, "oracle", "_ref", "libblingfiretokdll.so")
    tok = ctypes.c_int64(0)
    so = np.ascontiguousarray(offs[: n + 1])
    L.ref_time_batch(ref.encode(), model.encode(), text.ctypes.data, so.ctypes.data, min(n, 5000), max_ids, unk, threads, ctypes.byref(tok), None)
    secs = L.ref_time_batch(ref.encode(), model.encode(), text.ctypes.data, so.ctypes.data, n, max_ids, unk, threads,
                            ctypes.byref(tok), None)
    return int(so[-1]) / secs / 1e9, tok.value / secs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=500_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", default="", help="comma-separated configuration names (default: all)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the reference CPU timing")
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x)
    import torch
    import blingfire_b200 as bf
    import corpus
    from _common import model_path, read_lines
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
    cores = len(os.sched_getaffinity(0))
    L = bf.lib()
    L.BlingFireB200LastKernelMs.restype = ctypes.c_double
    out = {"cores": cores}

    # cfg 1: default TextToWords on 10k short ASCII lines (per-call API)
    if not only or "cfg1" in only:
        lines = [l for l in read_lines("test.txt") if len(l) <= 120 and all(c < 128 for c in l)][:10000]
        buf = ctypes.create_string_buffer(1024)
        L.TextToWords(lines[0], len(lines[0]), buf, 1024)
        t0 = time.perf_counter()
        for l in lines:
            L.TextToWords(l, len(l), buf, 1024)
        dt = time.perf_counter() - t0
        nb = sum(len(l) for l in lines)
        out["cfg1_TextToWords_10k_lines"] = {"MB_per_s": nb / dt / 1e6, "us_per_call": dt / len(lines) * 1e6, "bytes": nb}

    for name, model, unk, max_ids, gen in [
        ("cfg3_gpt2", "gpt2.bin", 0, 4096, lambda n: corpus.gen_docs("EN", n, seed=3, fixed_len=0)),
        ("cfg4_xlmr", "xlm_roberta_base.bin", 3, 512, lambda n: corpus.gen_docs("MULTI", n, seed=4, fixed_len=512, emoji_every=16)),
        ("cfg2_bert_same_path", "bert_base_tok.bin", 100, 512, lambda n: corpus.gen_docs("EN", n, seed=2, fixed_len=512)),
        # not a BASELINE configuration: the Unigram model on cfg 3's long documents (only with --only)
        ("extra_xlmr_long_docs", "xlm_roberta_base.bin", 3, 4096, lambda n: corpus.gen_docs("EN", n, seed=3, fixed_len=0)),
    ]:
        if (only and name not in only) or (not only and name.startswith("extra_")):
            continue
        text, offs = gen(args.docs)
        n, nbytes = len(offs) - 1, int(offs[-1])
        h_text = torch.empty(nbytes + 64, dtype=torch.uint8, pin_memory=True)
        h_text[:nbytes].copy_(torch.from_numpy(text))
        h_offs = torch.from_numpy(offs).pin_memory()
        cap = int(np.minimum(np.diff(offs), max_ids).sum())
        h_ids = torch.empty(cap + 1, dtype=torch.int32, pin_memory=True)
        h_idoffs = torch.zeros(n + 1, dtype=torch.int64, pin_memory=True)
        h = bf.load_model(model_path(model))

        def call():
            r = L.TextToIdsBatchCsr(ctypes.c_void_p(h), h_text.data_ptr(), h_offs.data_ptr(), n, h_ids.data_ptr(), cap,
                                    h_idoffs.data_ptr(), max_ids, unk)
            assert r >= 0, bf.last_error()
            return r

        call()
        best, kms, tot = 1e9, 0.0, 0
        for _ in range(args.reps):
            t0 = time.perf_counter()
            tot = call()
            dt = time.perf_counter() - t0
            if dt < best:
                best, kms = dt, L.BlingFireB200LastKernelMs()
        ns = min(n, 50000)
        cpu_gbs, cpu_tps = (None, None) if args.no_cpu else ref_cpu(model_path(model), text, offs, ns, max_ids, unk, cores)
        out[name] = {"docs": n, "bytes": nbytes, "tokens": int(tot), "e2e_GB_per_s": nbytes / best / 1e9,
                     "e2e_ms": best * 1e3, "kernel_ms": kms, "kernel_GB_per_s": nbytes / (kms * 1e-3) / 1e9 if kms > 0 else None,
                     "tokens_per_s_kernel": tot / (kms * 1e-3) if kms > 0 else None,
                     "cpu_ref_GB_per_s": cpu_gbs, "cpu_ref_tokens_per_s": cpu_tps, "cpu_threads": cores, "cpu_sample_docs": ns,
                     "api": "TextToIdsBatchCsr, pinned host buffers"}
        bf.free_model(h)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
