#!/usr/bin/env python3
"""Generates tests/golden/golden.json by running THE REFERENCE ITSELF (oracle/_ref/
libblingfiretokdll.so, compiled from the reference's own sources by oracle/Makefile) on
inputs taken from the reference's own corpora.  Run in the build container (where
/root/reference exists); the JSON is committed so the oracle and the CUDA path can be pinned
on the GPU box, where the reference tree does not exist.

    python tests/golden/make_golden.py
"""
import base64
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _common import Ref, fnv1a64_ids, model_path, read_lines  # noqa: E402

EDGE = [
    b"", b"a", b" ", b"   ", b"abc \xff def", b"\xe6\x88", b"\xef\xbb\xbf", b"\xef\xbb\xbfhello", b"hello\x00world",
    b"a" * 400, b"a" * 1000 + b" " + b"b" * 700, "我爱北京".encode(), b"[unk] [UNK] [cls][sep] [mask] [unused0] [mas",
    b"x" * 299 + b"y", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\xc0\xaf", "é".encode() * 700, b"\xf0\x9f\x98\x80 smile",
    b"tab\tnew\nline\r\n", b"don't stop-me now!!! (ok?)", "ÀÉÎÕÜ ñandú".encode(), b"\xe2\x80\x8b zero width",
    b"1234567890 3.14159 1,000,000", b"http://www.microsoft.com/a/b?c=d&e=f#g", b"\x01\x02\x03 ctrl",
    b"end with partial \xe2\x82", b"\x80 leading continuation", "Ünïcödé ｆｕｌｌｗｉｄｔｈ".encode(),
    ("word " * 300).encode(), b"." * 700, b"\xef\xbb\xbf\xef\xbb\xbfdouble bom",
]
MODELS = [("bert_base_tok.bin", 100), ("bert_base_cased_tok.bin", 100), ("bert_chinese.bin", 100),
          ("gpt2.bin", 0), ("xlm_roberta_base.bin", 3), ("xlnet.bin", 0), ("roberta.bin", 3)]
DIGESTS = [
    # (model, unk, corpus, n_lines (None = all), lines per doc, max_ids)
    ("bert_base_tok.bin", 100, "test.txt", None, 1, 65536),
    ("bert_base_tok.bin", 100, "test.txt", None, 8, 65536),
    ("bert_base_tok.bin", 100, "test.multi.txt", 20000, 1, 100),
    ("bert_base_cased_tok.bin", 100, "test.txt", 20000, 1, 65536),
    ("bert_chinese.bin", 100, "test.multi.txt", 20000, 1, 65536),
    ("gpt2.bin", 0, "test.txt", None, 8, 65536),
    ("xlm_roberta_base.bin", 3, "test.txt", None, 8, 65536),
    ("xlm_roberta_base.bin", 3, "test.multi.txt", 20000, 1, 65536),
]


def b64(b):
    return base64.b64encode(b).decode()


def main():
    r = Ref()
    out = {"generator": "tests/golden/make_golden.py over oracle/_ref (the reference built from its own sources)",
           "edge_cases": [], "digests": [], "words": [], "ids_with_offsets": [], "split_with_offsets": [], "py_offsets": []}
    handles = {}
    for m, unk in MODELS:
        handles[m] = r.load(model_path(m))
        for data in EDGE:
            for max_ids in (512, 3):
                n, ids = r.text_to_ids(handles[m], data, max_ids, unk)
                out["edge_cases"].append({"model": m, "unk": unk, "max_ids": max_ids, "input": b64(data),
                                          "count": int(n), "ids": ids[:max(n, 0)].tolist(),
                                          "tail_untouched": bool((ids[max(n, 0):] == -7).all())})
    for m, unk, corpus, nl, group, max_ids in DIGESTS:
        lines = read_lines(corpus, drop_empty=False)
        if nl:
            lines = lines[:nl]
        h = handles.get(m) or r.load(model_path(m))
        handles[m] = h
        dig = 0xcbf29ce484222325
        tokens = 0
        for i in range(0, len(lines), group):
            doc = b" ".join(lines[i:i + group])
            n, ids = r.text_to_ids(h, doc, max_ids, unk)
            tokens += n
            dig = fnv1a64_ids(ids[:n], dig)
        out["digests"].append({"model": m, "unk": unk, "corpus": corpus, "lines": len(lines), "group": group,
                               "max_ids": max_ids, "tokens": tokens, "fnv1a64": f"{dig:016x}"})
        print(out["digests"][-1])
    # default word breaker (embedded wbd.bin) on the first short ASCII lines + edge inputs
    lines = [l for l in read_lines("test.txt") if len(l) <= 120 and all(c < 128 for c in l)][:300]
    for data in lines + EDGE:
        n, s = r.text_to_words(data)
        out["words"].append({"input": b64(data), "ret": int(n), "out": b64(s)})
    # offsets variant on a few lines (bert + xlm-r)
    for m, unk in (("bert_base_tok.bin", 100), ("xlm_roberta_base.bin", 3), ("gpt2.bin", 0)):
        for data in read_lines("test.multi.txt")[:40] + EDGE[:12]:
            n, ids, st, en = r.text_to_ids_with_offsets(handles[m], data, 256, unk)
            out["ids_with_offsets"].append({"model": m, "unk": unk, "input": b64(data), "count": int(n),
                                            "ids": ids[:n].tolist(), "starts": st[:n].tolist(), "ends": en[:n].tolist()})
    # words / sentences with offsets: default models (embedded wbd / sbd) and the files in ldbsrc
    para = [b" ".join(read_lines("test.txt")[i:i + 6]) for i in range(0, 240, 6)] + read_lines("test.multi.txt")[:40]
    para += [b"Hello world! How are you?  I am fine.\nThanks. ", b"  \n ", b"No terminator", b"A.\x00B. C", b"\xef\xbb\xbfBom. Second one.",
             "Dr. Smith went to Washington. He arrived at 5 p.m. It was late!".encode(), "一。二。三".encode()]
    for kind in ("words", "sentences"):
        for data in para + EDGE[:16]:
            for max_out in (None, 3):
                n, text, st, en = r.split(kind, data, None, max_out)
                k = max_out if max_out is not None else 2 * len(data) + 16
                out["split_with_offsets"].append({"kind": kind, "input": b64(data), "max_out": k, "ret": int(n), "out": b64(text),
                                                  "starts": st[:max(k, 1)].tolist()[:64], "ends": en[:max(k, 1)].tolist()[:64]})
    # the reference's own Python wrapper (dist-pypi/blingfire/__init__.py:170-227): code-point offsets
    sys.path.insert(0, "/root/reference/dist-pypi")
    import blingfire as ref_py
    out["py_offsets"] = []
    texts = [l.decode("utf-8") for l in read_lines("test.multi.txt")[:60] + read_lines("test.txt")[:60]]
    texts += [" ".join(texts[i:i + 4]) for i in range(0, 40, 4)] + ["naïve café. Hello 我爱北京!", "x", "^", "Hello world! How are you?  Fine. Ünïcode ok."]
    for t in texts:
        w, wo = ref_py.text_to_words_with_offsets(t)
        sn, so = ref_py.text_to_sentences_and_offsets(t)
        out["py_offsets"].append({"text": t, "words": w, "word_offsets": [list(x) for x in wo], "sentences": sn,
                                  "sentence_offsets": [list(x) for x in so]})
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote golden.json", os.path.getsize(os.path.join(HERE, "golden.json")), "bytes")


if __name__ == "__main__":
    main()
