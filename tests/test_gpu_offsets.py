"""TextToIdsWithOffsets for lexer models through the C ABI: ids, start offsets and end
offsets against the golden fixtures (the reference itself) and the oracle."""
import base64
import json
import os

import numpy as np
import pytest

from _common import GOLDEN, Oracle, have_data, model_path, read_lines

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_data(), reason="data/ not staged")]


@pytest.fixture(scope="module")
def bf():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
    import blingfire_b200
    return blingfire_b200


def test_offsets_vs_golden(bf):
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        golden = json.load(f)
    h = bf.load_model(model_path("bert_base_tok.bin"))
    n_checked = 0
    for c in golden["ids_with_offsets"]:
        if c["model"] != "bert_base_tok.bin":
            continue
        data = base64.b64decode(c["input"])
        ids, st, en = bf.utf8text_to_ids_with_offsets(h, data, 256, c["unk"], no_padding=True)
        assert len(ids) == c["count"], data[:40]
        assert ids.astype(np.int64).tolist() == c["ids"] and st.tolist() == c["starts"] and en.tolist() == c["ends"], data[:40]
        n_checked += 1
    assert n_checked > 30
    bf.free_model(h)


@pytest.mark.parametrize("name", ["bert_base_tok.bin", "bert_base_cased_tok.bin", "bert_chinese.bin", "wbd.bin"])
def test_offsets_vs_oracle(bf, name):
    h = bf.load_model(model_path(name))
    o = Oracle()
    ho = o.load(model_path(name))
    docs = read_lines("test.multi.txt")[:400] + read_lines("test.txt")[:400] + [
        b"\xef\xbb\xbfbom first", b"abc \xff def", b"a" * 400, "naïve café 我爱北京 [unk] qwrtzx".encode(), b" ", b"x"]
    for d in docs:
        for max_ids in (200, 3):
            n, oi, os_, oe = o.text_to_ids_with_offsets(ho, d, max_ids, 100)
            ids, st, en = bf.utf8text_to_ids_with_offsets(h, d, max_ids, 100, no_padding=True)
            assert len(ids) == n, d[:40]
            assert (ids.astype(np.int32) == oi[:n]).all() and (st == os_[:n]).all() and (en == oe[:n]).all(), d[:40]
    # untouched tails
    import ctypes
    L = bf.lib()
    a = np.full(64, -7, np.int32); b = np.full(64, -7, np.int32); c = np.full(64, -7, np.int32)
    n = L.TextToIdsWithOffsets(ctypes.c_void_p(h), b"hello world", 11, a.ctypes.data, b.ctypes.data, c.ctypes.data, 64, 100)
    assert n >= 1 and (a[n:] == -7).all() and (b[n:] == -7).all() and (c[n:] == -7).all()
    bf.free_model(h)
