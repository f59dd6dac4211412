#!/usr/bin/env python3
"""Generates tests/golden/golden_r2.json by running THE REFERENCE ITSELF (oracle/_ref/libblingfiretokdll.so,
compiled from the reference's own sources by oracle/Makefile): IdsToText (blingfiretokdll.cpp:1689-1745) over
the shipped *.i2w arrays, and TextToIds / TextToIdsWithOffsets after SetNoDummyPrefix (:1669-1679).
Run in the build container; the JSON is committed so that the oracle and the product can be pinned on the
GPU box, where the reference tree does not exist.

    python tests/golden/make_golden_r2.py
"""
import base64
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _common import Ref, model_path, read_lines  # noqa: E402


def b64(b):
    return base64.b64encode(b).decode()


def main():
    r = Ref()
    out = {"generator": "tests/golden/make_golden_r2.py over oracle/_ref (the reference built from its own sources)",
           "ids_to_text": [], "no_dummy_prefix": []}
    # ---- IdsToText: ids produced by the reference's own TextToIds, plus special / out-of-range ids ----
    pairs = [("bert_base_tok.bin", "bert_base_tok.i2w", 100), ("bert_base_cased_tok.bin", "bert_base_cased_tok.i2w", 100),
             ("gpt2.bin", "gpt2.i2w", 0), ("roberta.bin", "roberta.i2w", 3), ("xlm_roberta_base.bin", "xlm_roberta_base.i2w", 3),
             ("bert_chinese.bin", "bert_chinese.i2w", 100), ("laser100k.bin", "laser100k.i2w", 0), ("uri100k.bin", "uri100k.i2w", 0)]
    lines = read_lines("test.txt")[:12] + read_lines("test.multi.txt")[:12] + [b"  leading spaces", b"x", b"Hello, World!  \xf0\x9f\x98\x80"]
    for tok_model, i2w, unk in pairs:
        ht = r.load(model_path(tok_model))
        hw = r.load(model_path(i2w))
        cases = []
        for data in lines:
            n, ids = r.text_to_ids(ht, data, 128, unk)
            cases.append(ids[:max(n, 0)].tolist())
        cases += [[], [0], [1, 2, 3], [0, 1, 2, 3, 4, 5, 100, 101, 102, 103, 104], [5, 10 ** 9], [-1], [7, -5, 9], [250001, 3], [50256, 50257]]
        for ids in cases:
            for skip in (False, True):
                for max_out in (4096, 8, 0):
                    n, text = r.ids_to_text(hw, np.array(ids, np.int32), max_out, skip)
                    out["ids_to_text"].append({"i2w": i2w, "ids": ids, "skip_special": skip, "max_out": max_out, "ret": int(n),
                                               "out": b64(text)})
        # a tokenizer model without [i2w] answers 0 (:1700-1702)
        n, text = r.ids_to_text(ht, np.array([1, 2, 3], np.int32), 64, False)
        out["ids_to_text"].append({"i2w": tok_model, "ids": [1, 2, 3], "skip_special": False, "max_out": 64, "ret": int(n), "out": b64(text)})
        r.free(ht); r.free(hw)
    # ---- SetNoDummyPrefix: both settings, ids and offsets ----
    texts = read_lines("test.txt")[:25] + read_lines("test.multi.txt")[:25] + [b"", b" ", b"  two  spaces ", b"a", b"\xe2\x96\x81x", b"\xef\xbb\xbfbom"]
    for m, unk in (("xlm_roberta_base.bin", 3), ("xlnet.bin", 0), ("gpt2.bin", 0), ("roberta.bin", 3), ("laser100k.bin", 0), ("bert_base_tok.bin", 100)):
        h = r.load(model_path(m))
        for flag in (True, False):
            ret = r.set_no_dummy_prefix(h, flag)
            for data in texts:
                n, ids, st, en = r.text_to_ids_with_offsets(h, data, 128, unk)
                n2, ids2 = r.text_to_ids(h, data, 128, unk)
                assert n2 == n and (ids2[:n] == ids[:n]).all()
                out["no_dummy_prefix"].append({"model": m, "unk": unk, "flag": flag, "set_ret": int(ret), "input": b64(data), "count": int(n),
                                               "ids": ids[:max(n, 0)].tolist(), "starts": st[:max(n, 0)].tolist(), "ends": en[:max(n, 0)].tolist()})
        r.free(h)
    with open(os.path.join(HERE, "golden_r2.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote golden_r2.json", os.path.getsize(os.path.join(HERE, "golden_r2.json")), "bytes;",
          len(out["ids_to_text"]), "IdsToText cases,", len(out["no_dummy_prefix"]), "no-dummy-prefix cases")


if __name__ == "__main__":
    main()
