"""[pos-dict] host logic on the CPU: the flattened segmentation tables (double-array Mealy automaton,
symbol map, I2Info, charmap; blingfire_b200/csrc/seg_tables.cpp) driven by a sequential twin, against
the oracle, which reads the packed image.  Covers Unigram-LM (xlm_roberta_base, xlnet), BPE-opt (gpt2),
BPE-opt-with-merges (roberta) and a plain cp-mode BPE (bpe_example)."""
import ctypes
import os
import random

import numpy as np
import pytest

from _common import ROOT, Oracle, have_data, model_path, read_lines

pytestmark = pytest.mark.skipif(not have_data(), reason="data/ not staged (run __graft_entry__.build())")


@pytest.fixture(scope="module")
def sptwin():
    L = ctypes.CDLL(os.path.join(ROOT, "tests", "twin", "libsp_twin.so"))
    L.sptwin_load.restype = ctypes.c_void_p
    L.sptwin_load.argtypes = [ctypes.c_char_p]
    L.sptwin_free.argtypes = [ctypes.c_void_p]
    L.sptwin_error.restype = ctypes.c_char_p
    L.sptwin_error.argtypes = [ctypes.c_void_p]
    L.sptwin_info.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.sptwin_check_bpe_order.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.sptwin_text_to_ids.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    L.sptwin_text_to_ids_streamed.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int]
    return L


_twins = {}


def twin_model(sptwin, name):
    """One flattened model per file for the whole module: building the double-array of the 250k-token
    xlm-r vocabulary takes tens of seconds."""
    if name not in _twins:
        _twins[name] = sptwin.sptwin_load(model_path(name).encode())
        assert sptwin.sptwin_error(_twins[name]) == b"", sptwin.sptwin_error(_twins[name])
    return _twins[name]


def docs_for_fuzz(seed, n):
    rng = random.Random(seed)
    lines = read_lines("test.multi.txt")[:3000] + read_lines("test.txt")[:3000]
    docs = [b"", b" ", b"  a  b  ", b"\xef\xbb\xbf", b"\xef\xbb\xbfhello", b"abc \xff def", b"hello\x00world", b"a" * 400,
            b"-" * 300, b"\xc2\xa0nbsp\xc2\xa0", "▁already▁marked ▁".encode(), "我爱北京".encode(), b"\t\ttabs\n\nnl  ",
            b"x", b"!", "é".encode() * 300]
    for _ in range(n):
        d = b" ".join(rng.choice(lines) for _ in range(rng.randint(1, 5)))
        r = rng.random()
        if r < 0.15:
            d = d[: rng.randint(0, len(d))]
        elif r < 0.25:
            p = rng.randint(0, len(d))
            d = d[:p] + bytes([rng.randint(0, 255)]) + d[p:]
        elif r < 0.30:
            d = bytes(rng.randint(0, 255) for _ in range(rng.randint(1, 40)))
        docs.append(d)
    return docs


@pytest.mark.parametrize("name,unks", [("xlm_roberta_base.bin", (3,)), ("xlnet.bin", (0,)), ("gpt2.bin", (0, 50256)),
                                       ("roberta.bin", (3,)), ("bpe_example.bin", (0, 1))])
def test_segmentation_tables_match_oracle(sptwin, name, unks):
    h = twin_model(sptwin, name)
    o = Oracle()
    ho = o.load(model_path(name))
    out = np.zeros(4096, np.int32)
    for unk in unks:
        for d in docs_for_fuzz(11, 6000):
            for max_ids in (4096, 5):
                n1, a = o.text_to_ids(ho, d, max_ids, unk)
                out[:] = -7
                n2 = sptwin.sptwin_text_to_ids(h, d, len(d), out.ctypes.data, max_ids, unk)
                assert n1 == n2 and (a[:n1] == out[:n1]).all(), (name, unk, d[:60])


@pytest.mark.parametrize("name,unk", [("xlm_roberta_base.bin", 3), ("xlnet.bin", 0), ("gpt2.bin", 0), ("roberta.bin", 3)])
def test_streaming_decompositions_are_exact(sptwin, name, unk):
    """What the streaming kernels rely on, run sequentially on the CPU against the oracle: a Unigram
    document cut at the last U+2581 of every window with the best score carried over; a BPE document
    segment by segment, segments split at the positions no token spans."""
    h = twin_model(sptwin, name)
    o = Oracle()
    ho = o.load(model_path(name))
    out = np.zeros(8192, np.int32)
    lines = read_lines("test.multi.txt")[:1500] + read_lines("test.txt")[:1500]
    docs = docs_for_fuzz(23, 2000) + [b" ".join(lines[i:i + 40]) for i in range(0, 1200, 40)] + [
        b"http://www.example.com/a/very/long/path/without/any/space/" * 6, "我爱北京天安门".encode() * 60, b"=" * 500 + b" x"]
    uncut = 0
    for d in docs:
        n1, a = o.text_to_ids(ho, d, 8192, unk)
        for window in (16, 24, 97, 576):
            n2 = sptwin.sptwin_text_to_ids_streamed(h, d, len(d), out.ctypes.data, 8192, unk, window)
            if n2 == -4:          # a run without U+2581 longer than the window: the kernel takes the general path
                uncut += 1
                continue
            assert n1 == n2 and (a[:n1] == out[:n1]).all(), (name, window, d[:60])
    assert uncut < len(docs)      # (only the tiny windows and the pathological documents)


def test_model_properties(sptwin):
    for name, algo, raw in [("gpt2.bin", 4, 1), ("roberta.bin", 5, 1), ("xlm_roberta_base.bin", 0, 0)]:
        h = twin_model(sptwin, name)
        assert sptwin.sptwin_info(h, 5) == algo and sptwin.sptwin_info(h, 7) == raw
        assert sptwin.sptwin_info(h, 4) == 0     # no token has U+2581 past its first symbol
        if algo == 0:
            assert sptwin.sptwin_info(h, 12) == 1    # "U+2581" alone is a token: cuts at U+2581 are exact
    

def test_bpe_order_is_one_integer(sptwin):
    """The streaming BPE kernel sorts arcs by one integer: the ordinal of (rank, id) must order keys
    exactly like the reference comparator.  gpt2 (ids = merge order) puts the one-symbol tokens first;
    roberta sorts by rank first and gives them the lowest rank, so they come last."""
    for name, singles_first in [("gpt2.bin", 1), ("roberta.bin", 0), ("bpe_example.bin", None)]:
        h = twin_model(sptwin, name)
        assert sptwin.sptwin_info(h, 9) == 1
        if singles_first is not None:
            assert sptwin.sptwin_info(h, 10) == singles_first
        assert sptwin.sptwin_info(h, 11) > 0
        assert sptwin.sptwin_check_bpe_order(h, 37) == 0
        h = twin_model(sptwin, "xlm_roberta_base.bin")
    assert sptwin.sptwin_info(h, 9) == 0 and sptwin.sptwin_check_bpe_order(h, 37) == -1   # Unigram: no such table
