"""Host logic of the product, checked on the CPU:
  * the flattened HBM tables (blingfire_b200/csrc/lexer_tables.cpp) against the oracle's readers of
    the packed image, exhaustively (every state x every class) -- the analogue of the reference's
    own `fa_fsm2fsm_pack --auto-test` (ldbsrc/Makefile.gnu:104-203);
  * the warp algorithm (sync points, chunks, windows with carry, tiling rule) through the host
    twin, which runs the product's own __host__ __device__ routine, against the oracle.
CPU only; no product compute path is exercised here (that is tests/test_gpu_parity.py)."""
import ctypes
import os
import random

import numpy as np
import pytest

from _common import ROOT, Oracle, have_data, model_path, read_lines

pytestmark = pytest.mark.skipif(not have_data(), reason="data/ not staged (run __graft_entry__.build())")


class Twin:
    def __init__(self):
        self.lib = L = ctypes.CDLL(os.path.join(ROOT, "tests", "twin", "libwp_twin.so"))
        L.twin_load.restype = ctypes.c_void_p
        L.twin_load.argtypes = [ctypes.c_char_p]
        L.twin_free.argtypes = [ctypes.c_void_p]
        L.twin_load_sparse.restype = ctypes.c_void_p
        L.twin_load_sparse.argtypes = [ctypes.c_char_p]
        L.twin_blob_digest.restype = ctypes.c_uint64
        L.twin_blob_digest.argtypes = [ctypes.c_void_p]
        L.twin_dense_on_host.argtypes = [ctypes.c_void_p]
        L.twin_error.restype = ctypes.c_char_p
        L.twin_error.argtypes = [ctypes.c_void_p]
        L.twin_fast_ok.argtypes = [ctypes.c_void_p]
        L.twin_fast_why.restype = ctypes.c_char_p
        L.twin_fast_why.argtypes = [ctypes.c_void_p]
        L.twin_info.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.twin_ptr.restype = ctypes.c_void_p
        L.twin_ptr.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.twin_text_to_ids.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int]
        L.twin_text_to_ids_ex.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]

    def load(self, name):
        h = self.lib.twin_load(model_path(name).encode())
        assert self.lib.twin_error(h) == b"", self.lib.twin_error(h)
        return h

    def arr(self, h, what, dtype, n):
        p = self.lib.twin_ptr(h, what)
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,))

    def ids_ex(self, h, data, max_ids, unk, window, use_memo, stats=None):
        out = np.full(max_ids, -7, np.int32)
        n = self.lib.twin_text_to_ids_ex(h, data, len(data), out.ctypes.data, max_ids, unk, window, use_memo,
                                         stats.ctypes.data if stats is not None else None)
        return n, out

    def ids(self, h, data, max_ids, unk, window):
        out = np.full(max_ids, -7, np.int32)
        n = self.lib.twin_text_to_ids(h, data, len(data), out.ctypes.data, max_ids, unk, window)
        return n, out


@pytest.fixture(scope="module")
def twin():
    return Twin()


@pytest.fixture(scope="module")
def oracle():
    o = Oracle()
    L = o.lib
    L.bfo_dfa_row.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.bfo_iwmap_many.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.bfo_state_info_many.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p]
    L.bfo_lexer_class_many.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.bfo_act_get.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    return o


@pytest.mark.parametrize("name", ["bert_base_tok.bin", "wbd.bin", "bert_chinese.bin"])
def test_flattened_tables_match_packed_readers(twin, oracle, name):
    h = twin.load(name)
    ho = oracle.load(model_path(name))
    info = lambda i: twin.lib.twin_info(h, i)
    NS, NC, first_final, dead, wide, n_iw = info(0), info(1), info(2), info(4), info(5), info(6)
    orig = twin.arr(h, 1, np.int32, NS)
    cls_of_iw = twin.arr(h, 2, np.uint16, n_iw)
    trans = twin.arr(h, 0, np.uint32 if wide else np.uint16, NS * (NC + 1)).reshape(NS, NC + 1)
    none = 0xFFFFFFFF if wide else 0xFFFF

    # class map == FAIwMap_pack::GetNewIw for every Iw of the map's domain
    ref_cls = np.empty(n_iw, np.int32)
    oracle.lib.bfo_iwmap_many(ho, 0, n_iw, ref_cls.ctypes.data)
    assert (np.where(ref_cls < 0, NC, ref_cls) == cls_of_iw).all()
    any_cls = ref_cls[0]   # IW_ANY = 0

    # one representative Iw per class
    rep = np.full(NC, -1, np.int64)
    first = np.unique(cls_of_iw, return_index=True)
    for c, i in zip(*first):
        if c < NC:
            rep[c] = i
    assert (rep >= 0).all(), "a class without any Iw"
    rep32 = rep.astype(np.int32)

    # every state x every class: same destination as FARSDfa_pack_triv::GetDest (+ IW_ANY fallback)
    new_of_orig = {int(o): s for s, o in enumerate(orig) if o >= 0}
    rng = random.Random(7)
    states = list(range(NS)) if NS <= 30000 else sorted(rng.sample(range(NS), 30000))
    row = np.empty(NC, np.int32)
    for s in states:
        if s == dead:
            assert (trans[s] == none).all()
            continue
        oracle.lib.bfo_dfa_row(ho, int(orig[s]), rep32.ctypes.data, NC, row.ctypes.data)
        if any_cls >= 0:
            row = np.where(row == -1, row[any_cls], row)
        exp = np.array([none if d == -1 else (dead if d == -2 else new_of_orig[int(d)]) for d in row], dtype=np.int64)
        assert (trans[s, :NC].astype(np.int64) == exp).all(), f"state {s}"
        assert int(trans[s, NC]) == (int(exp[any_cls]) if any_cls >= 0 else none)

    # finality, rule ids, tags
    fin = np.empty(NS, np.int32)
    ows = np.empty(NS, np.int32)
    o32 = orig.astype(np.int32).copy()
    o32[dead] = int(orig[0])
    oracle.lib.bfo_state_info_many(ho, o32.ctypes.data, NS, fin.ctypes.data, ows.ctypes.data)
    fin[dead] = 0
    ows[dead] = -1
    assert ((np.arange(NS) >= first_final) == (fin != 0)).all()
    assert (twin.arr(h, 4, np.int32, NS) == ows).all()
    tags = twin.arr(h, 5, np.int32, NS)
    for s in np.nonzero(fin)[0][:: max(1, NS // 4000)]:
        p = ctypes.POINTER(ctypes.c_int)()
        n = oracle.lib.bfo_act_get(ho, int(ows[s]), ctypes.byref(p))
        assert n >= 3 and tags[s] == p[2]

    # function initial states (FAWbdConfKeeper::CalcFnIniStates)
    nfn = info(7)
    fn = twin.arr(h, 6, np.uint32, nfn)
    for f in range(nfn):
        o = oracle.lib.bfo_fn_ini(ho, f)
        assert (fn[f] == 0xFFFFFFFF) if o < 0 else (int(orig[fn[f]]) == o)

    # combined code point -> class table (charmap + clamp + class map), when the charmap is 1->1
    if twin.lib.twin_fast_ok(h):
        cc = twin.arr(h, 3, np.uint16, 0x110000)
        ref = np.empty(0x110000, np.int32)
        oracle.lib.bfo_lexer_class_many(ho, 0, 0x110000, ref.ctypes.data)
        assert (ref != -2).all()
        assert (np.where(ref < 0, NC, ref) == cc).all()
    twin.lib.twin_free(h)


def test_fast_path_classification(twin):
    for name, ok in [("bert_base_tok.bin", 1), ("bert_base_cased_tok.bin", 1), ("bert_chinese.bin", 1),
                     ("wbd.bin", 0), ("sbd.bin", 0)]:
        h = twin.lib.twin_load(model_path(name).encode())
        assert twin.lib.twin_error(h) == b""
        assert twin.lib.twin_fast_ok(h) == ok, (name, twin.lib.twin_fast_why(h))
        twin.lib.twin_free(h)


EDGE = [b"", b"abc \xff def", b"\xe6\x88", b"   ", b"\xef\xbb\xbfhello", b"hello\x00world", b"a" * 400,
        b"a" * 1000 + b" " + b"b" * 700, "我爱北京".encode(), b"[unk] [UNK] [cls][sep] [mask] [unused0] [mas",
        b"\xef\xbb\xbf", b"x" * 299 + b"y", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\xc0\xaf", "é".encode() * 700,
        ("word " * 300).encode(), b"." * 700, "我".encode() * 900, b"ab" * 700 + b"\xff"]


@pytest.mark.parametrize("name,corpus,nlines,group,window,max_ids", [
    ("bert_base_tok.bin", "test.txt", 30000, 1, 640, 512),
    ("bert_base_tok.bin", "test.txt", 20000, 40, 340, 4096),     # long documents, tiny window: carry logic
    ("bert_base_tok.bin", "test.multi.txt", 8000, 3, 640, 100),
    ("bert_base_cased_tok.bin", "test.txt", 8000, 2, 640, 512),
    ("bert_chinese.bin", "test.multi.txt", 8000, 2, 400, 512),
    # 232k states x 10 004 classes: the 32-bit table variant (9.3 GB on the host, filled by several threads)
    ("bert_multi_cased.bin", "test.multi.txt", 3000, 2, 640, 512),
])
def test_twin_matches_oracle(twin, oracle, name, corpus, nlines, group, window, max_ids):
    h = twin.load(name)
    ho = oracle.load(model_path(name))
    lines = read_lines(corpus, drop_empty=False)[:nlines]
    docs = [b" ".join(lines[i:i + group]) for i in range(0, len(lines), group)] + EDGE
    for d in docs:
        n1, a = oracle.text_to_ids(ho, d, max_ids, 100)
        n2, b = twin.ids(h, d, max_ids, 100, window)
        assert n1 == n2 and (a[:n1] == b[:n1]).all(), d[:80]
    twin.lib.twin_free(h)


def test_twin_fuzz(twin, oracle):
    h = twin.load("bert_base_tok.bin")
    ho = oracle.load(model_path("bert_base_tok.bin"))
    rng = random.Random(99)
    lines = read_lines("test.multi.txt")[:2000] + read_lines("test.txt")[:2000]
    for _ in range(4000):
        d = b" ".join(rng.choice(lines) for _ in range(rng.randint(1, 12)))
        r = rng.random()
        if r < 0.2:
            d = d[: rng.randint(0, len(d))]
        elif r < 0.3:
            p = rng.randint(0, len(d))
            d = d[:p] + bytes([rng.randint(0, 255)]) + d[p:]
        elif r < 0.35:
            d = bytes(rng.randint(0, 255) for _ in range(rng.randint(1, 60)))
        n1, a = oracle.text_to_ids(ho, d, 300, 100)
        n2, b = twin.ids(h, d, 300, 100, rng.choice([340, 400, 640]))
        assert n1 == n2 and (a[:n1] == b[:n1]).all(), d[:80]
    twin.lib.twin_free(h)


MEMO_WORDS = [b"unaffable", b"antidisestablishmentarianism", b"supercalifragilisticexpialidocious", b"qwrtzxqwrtzx",
              b"internationalization", b"a" * 24, b"a" * 25, b"b" * 12, b"b" * 13, "naïveté".encode(), b"electroencephalography",
              b"x", b"zzzzzzzzzzzzzzzzzzzzzzzz", b"tokenization's", b"[unk]", b"[UNK]x", b"don't", b"www.example-site.com/path",
              b"1234567890123", b"hydroxychloroquine", b"pneumonoultramicroscopicsilicovolcanoconiosis"]


@pytest.mark.parametrize("name,corpus", [("bert_base_tok.bin", "test.txt"), ("bert_base_cased_tok.bin", "test.txt"),
                                         ("bert_chinese.bin", "test.multi.txt")])
def test_memo_never_changes_an_id(twin, name, corpus):
    """The load-time memo (class groups, whole-word table) and the words learned at run time are abbreviations of the lexer
    loops: with the memo off (0), with the table as built at load time (2) and with learning on (1, twice, so that the
    second pass finds what the first one added) every document yields the same ids."""
    h = twin.load(name)
    assert twin.lib.twin_info(h, 13) > 1000 and twin.lib.twin_info(h, 15) >= 8      # vocabulary words in the table; key length
    lines = read_lines(corpus, drop_empty=False)[:6000]
    docs = [b" ".join(lines[i:i + 3]) for i in range(0, len(lines), 3)] + EDGE
    docs += [b" ".join(MEMO_WORDS), b" ".join(reversed(MEMO_WORDS)), b"".join(MEMO_WORDS)] + MEMO_WORDS
    want = [twin.ids_ex(h, d, 512, 100, 640, 0) for d in docs]
    stats = [np.zeros(8, np.int64) for _ in range(3)]
    for k, mode in enumerate((2, 1, 1)):
        for d, (n0, a0) in zip(docs, want):
            n, a = twin.ids_ex(h, d, 512, 100, 640, mode, stats[k])
            assert n == n0 and (a[:max(n, 0)] == a0[:max(n, 0)]).all(), (mode, d[:80])
    # the table serves most words; what it did not hold at load time it holds after one pass
    assert stats[0][0] > 3 * stats[0][1] or name == "bert_chinese.bin"
    assert stats[0][3] == 0 and stats[2][3] > 0 and stats[2][1] < stats[1][1]
    # another UnkId: a learned "this word is one UnkId" entry does not remember the id
    for d in docs[:300] + MEMO_WORDS:
        n0, a0 = twin.ids_ex(h, d, 512, 7777, 640, 0)
        n, a = twin.ids_ex(h, d, 512, 7777, 640, 1)
        assert n == n0 and (a[:max(n, 0)] == a0[:max(n, 0)]).all(), d[:80]
    twin.lib.twin_free(h)


def test_wide_model_without_the_dense_host_table(twin):
    """LoadModel does not stage a table with 32-bit entries on the host (9.3 GB for bert_multi_cased): the grammar analysis,
    the class groups and the word table come out of the stored arcs -- bit for bit what the dense table gives."""
    import time
    t0 = time.time()
    hs = twin.lib.twin_load_sparse(model_path("bert_multi_cased.bin").encode())
    t1 = time.time()
    assert twin.lib.twin_error(hs) == b"" and twin.lib.twin_dense_on_host(hs) == 0 and twin.lib.twin_fast_ok(hs) == 1
    hd = twin.load("bert_multi_cased.bin")
    t2 = time.time()
    assert twin.lib.twin_dense_on_host(hd) == 1
    assert twin.lib.twin_info(hs, 13) == twin.lib.twin_info(hd, 13) > 50000
    assert twin.lib.twin_blob_digest(hs) == twin.lib.twin_blob_digest(hd)
    print(f"bert_multi_cased: host tables from the stored arcs {t1 - t0:.2f} s, with the dense table {t2 - t1:.2f} s")
    twin.lib.twin_free(hs)
    twin.lib.twin_free(hd)
    # a narrow table is always dense on the host
    hn = twin.lib.twin_load_sparse(model_path("bert_base_tok.bin").encode())
    assert twin.lib.twin_dense_on_host(hn) == 1
    twin.lib.twin_free(hn)
