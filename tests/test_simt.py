"""The SOURCE of the kernels (blingfire_b200/csrc/wp_kernel.cu: wp_tokenize_kernel; sp_kernel.cu: sp_unigram_kernel,
sp_bpe_kernel) compiled for the host over the SIMT shim (tests/simt: one OS thread per lane, full-mask warp intrinsics as
rendezvous) and run on the CPU against the oracle: the warp-level logic itself -- shuffles, ballots, the
register window of the Unigram relaxation, the streaming windows, the lane re-dealing of the BPE hard pass,
the general path with offsets -- not a twin of it.  A kernel whose lanes do not reach the same intrinsics
in the same order deadlocks here, hence the timeouts.  The GPU suite remains the parity test proper."""
import ctypes
import os
import random

import numpy as np
import pytest

from _common import ROOT, Oracle, have_data, model_path, read_lines

pytestmark = [pytest.mark.skipif(not have_data(), reason="data/ not staged (run __graft_entry__.build())"),
              pytest.mark.timeout(600)]


@pytest.fixture(scope="module")
def sim():
    L = ctypes.CDLL(os.path.join(ROOT, "tests", "simt", "libsp_simt.so"))
    L.spsim_load.restype = ctypes.c_void_p
    L.spsim_load.argtypes = [ctypes.c_char_p]
    L.spsim_error.restype = ctypes.c_char_p
    L.spsim_error.argtypes = [ctypes.c_void_p]
    L.spsim_batch.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return L


_models = {}


def sim_model(sim, name):
    if name not in _models:
        _models[name] = sim.spsim_load(model_path(name).encode())
        assert sim.spsim_error(_models[name]) == b"", sim.spsim_error(_models[name])
    return _models[name]


def run_batch(sim, name, docs, max_ids, unk, warps=3, offsets=False):
    h = sim_model(sim, name)
    offs = np.zeros(len(docs) + 1, np.int64)
    np.cumsum([len(d) for d in docs], out=offs[1:])
    buf = b"".join(docs) + b"\0"
    ids = np.full((len(docs), max_ids), -7, np.int32)
    counts = np.full(len(docs), -7, np.int32)
    st = np.full((len(docs), max_ids), -7, np.int32)
    en = np.full((len(docs), max_ids), -7, np.int32)
    flag = sim.spsim_batch(h, buf, offs.ctypes.data, len(docs), ids.ctypes.data, counts.ctypes.data,
                           st.ctypes.data if offsets else None, en.ctypes.data if offsets else None, max_ids, unk, warps)
    assert flag == 0, f"kernel error flag {flag}"
    return ids, counts, st, en


def check(sim, name, docs, max_ids, unk, offsets=False):
    o = Oracle()
    ho = o.load(model_path(name))
    ids, counts, st, en = run_batch(sim, name, docs, max_ids, unk, offsets=offsets)
    for i, d in enumerate(docs):
        n, a, s, e = o.text_to_ids_with_offsets(ho, d, max_ids, unk)
        assert counts[i] == n, (name, i, d[:60], int(counts[i]), n)
        assert (ids[i, :n] == a[:n]).all(), (name, i, d[:60])
        # (a streamed document that turns out to be invalid has count 0 after earlier windows were emitted:
        # the device row is scratch, the host API copies `count` entries only)
        assert n == 0 or (ids[i, n:] == -7).all(), "a row was written beyond its count"
        if offsets:
            assert (st[i, :n] == s[:n]).all() and (en[i, :n] == e[:n]).all(), (name, i, d[:60])
    o.free(ho)


def corpus_docs(seed, n):
    rng = random.Random(seed)
    lines = read_lines("test.multi.txt")[:3000] + read_lines("test.txt")[:3000]
    docs = [b"", b" ", b"  a  b  ", b"\xef\xbb\xbf", b"\xef\xbb\xbfhello", b"abc \xff def", b"hello\x00world", b"a" * 400, b"-" * 300,
            b"\xc2\xa0nbsp\xc2\xa0", "▁already▁marked ▁".encode(), "我爱北京".encode() * 40, b"\t\ttabs\n\nnl  ", b"x", b"!",
            "é".encode() * 300, "ﬁ ½ ™ ｶﾞ ＡＢＣ".encode(), b"http://www.example.com/a/long/path/without/space?x=1&y=2" * 3]
    for _ in range(n):
        d = b" ".join(rng.choice(lines) for _ in range(rng.randint(1, 4)))
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


SP_MODELS = [("xlm_roberta_base.bin", 3), ("xlnet.bin", 0), ("laser100k.bin", 1), ("gpt2.bin", 0), ("roberta.bin", 3), ("bpe_example.bin", 1)]


@pytest.mark.parametrize("name,unk", SP_MODELS)
def test_kernel_source_matches_oracle(sim, name, unk):
    docs = corpus_docs(31, 110)
    check(sim, name, docs, 512, unk)
    check(sim, name, docs[:60], 5, unk)              # truncation, incl. invalid bytes after MaxIds was reached


@pytest.mark.parametrize("name", ["gpt2.bin", "roberta.bin"])
def test_bpe_segment_memo(sim, name):
    """The streaming BPE path keeps resolved U+2581-delimited segments in a table (sp_bpe.cuh): the same documents again
    (now served from the table), then with other UnkIds -- one of them a real token id -- on the same model."""
    docs = corpus_docs(77, 60) + [b"supercalifragilisticexpialidocious antidisestablishmentarianism " * 6, b" a" * 300, b"the the the the"]
    for unk in (0, 0, 50256, 7, 262):
        check(sim, name, docs, 512, unk)


@pytest.mark.parametrize("name,unk", [("xlm_roberta_base.bin", 3), ("xlnet.bin", 0), ("gpt2.bin", 0), ("roberta.bin", 3)])
def test_kernel_source_long_documents(sim, name, unk):
    """Several windows per document: the streamed Unigram form (cuts at U+2581 and inside U+2581-free runs, the
    score carried over) and the sliding BPE window (split pass, cooperative pass, the general path behind)."""
    lines = read_lines("test.multi.txt")[:1200] + read_lines("test.txt")[:1200]
    docs = [b" ".join(lines[i:i + 25]) for i in range(0, 100, 25)] + [b" ".join(lines[1200 + i:1200 + i + 25]) for i in range(0, 100, 25)]
    docs += [b" ".join(lines[:40]) + b"\xff", b"\x80" + b" ".join(lines[60:100]), ("word " * 500).encode(), b"ab" * 700,
             "我爱北京天安门".encode() * 90, b"=" * 700 + b" x", ("ﬁ ½ ™ " * 300).encode(),
             b"https://homedepot.ugc.bazaarvoice.com/answers/submit/1999aa/product/205299499/question/4069427/undohelpfulness.djs?authsourcetype=__AUTHTYPE__&format=" * 5]
    check(sim, name, docs, 4096, unk)
    check(sim, name, docs[::2], 37, unk)


@pytest.mark.parametrize("name,unk", [("xlm_roberta_base.bin", 3), ("xlnet.bin", 0), ("laser100k.bin", 1), ("gpt2.bin", 0), ("roberta.bin", 3),
                                      ("bpe_example.bin", 1), ("xlnet_nonorm.bin", 0)])
def test_kernel_source_offsets(sim, name, unk):
    """TextToIdsWithOffsets_sp: the byte offsets carried through -- sp_unigram_offsets_kernel (one-window fast path, the
    general path behind it for documents of more than a window), sp_bpe_offsets_kernel (the sliding window: offsets slide
    with the symbols; cuts at U+2581 and inside a run without one), the general path for the models neither serves."""
    docs = corpus_docs(7, 70)
    lines = read_lines("test.txt")[:400]
    docs += [b" ".join(lines[i:i + 30]) for i in range(0, 120, 30)] + [b"ab" * 700, b" a" * 300, b"=" * 700 + b" x",
             b"supercalifragilisticexpialidocious antidisestablishmentarianism " * 12, b"\xef\xbb\xbf" + b"word " * 300]
    # around the window's capacity (576 symbols, the dummy prefix included), with and without a charmap expansion in it
    docs += [b"a" * k for k in (573, 574, 575, 576, 577)] + [("ab " * 191 + "ﬁ").encode(), ("é" * 574).encode(), ("é" * 577).encode(),
             ("word " * 500).encode(), b"\xef\xbb\xbf" + b"x y " * 140, b"  lead and trail  ", "\u3000ideographic\u3000space".encode(),
             "½".encode(), " ½".encode(), "▁".encode(), " ".encode() * 50]
    check(sim, name, docs, 700, unk, offsets=True)
    check(sim, name, docs[::3], 5, unk, offsets=True)


# ---- the fused WordPiece kernel ----------------------------------------------------------------------------

@pytest.fixture(scope="module")
def wpsim():
    L = ctypes.CDLL(os.path.join(ROOT, "tests", "simt", "libwp_simt.so"))
    L.wpsim_load.restype = ctypes.c_void_p
    L.wpsim_load.argtypes = [ctypes.c_char_p]
    L.wpsim_error.restype = ctypes.c_char_p
    L.wpsim_error.argtypes = [ctypes.c_void_p]
    L.wpsim_free.argtypes = [ctypes.c_void_p]
    L.wpsim_batch.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                              ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return L


def check_wp(wpsim, name, docs, max_ids, unk=100, warps=4, handle=None):
    h = handle or wpsim.wpsim_load(model_path(name).encode())
    assert wpsim.wpsim_error(h) == b"", wpsim.wpsim_error(h)
    o = Oracle()
    ho = o.load(model_path(name))
    offs = np.zeros(len(docs) + 1, np.int64)
    np.cumsum([len(d) for d in docs], out=offs[1:])
    buf = b"".join(docs) + b"\0"
    ids = np.full((len(docs), max_ids), -7, np.int32)
    counts = np.full(len(docs), -7, np.int32)
    assert wpsim.wpsim_batch(h, buf, offs.ctypes.data, len(docs), ids.ctypes.data, counts.ctypes.data, max_ids, unk, warps) == 0
    for i, d in enumerate(docs):
        n, a = o.text_to_ids(ho, d, max_ids, unk)
        assert counts[i] == n, (name, i, d[:60], int(counts[i]), n)
        assert (ids[i, :n] == a[:n]).all(), (name, i, d[:60])
        assert (ids[i, n:] == -7).all(), "a row was written beyond its count"
    o.free(ho)
    if handle is None:
        wpsim.wpsim_free(h)


def wp_docs(seed, n):
    rng = random.Random(seed)
    lines = read_lines("test.multi.txt")[:3000] + read_lines("test.txt")[:3000]
    docs = [b"", b" ", b"\xef\xbb\xbf", b"\xef\xbb\xbfhello", b"abc \xff def", b"\xe6\x88", b"hello\x00world", b"a" * 400,
            b"a" * 1000 + b" " + b"b" * 700, "我".encode() * 900, b"ab" * 700 + b"\xff", b"." * 700, ("word " * 300).encode(),
            b"x" * 573, b"y" * 571 + "é".encode(), ("é" * 700).encode(), b"\xf0\x9f\x98\x80" * 300, b"\x80" + b"a" * 600,
            b"[unk] [UNK] [cls][sep] [mask] [unused0] [mas", b"don't stop-me now!!! (ok?)", b" ".join(lines[:80])]
    for _ in range(n):
        d = b" ".join(rng.choice(lines) for _ in range(rng.randint(1, 14)))
        r = rng.random()
        if r < 0.2:
            d = d[: rng.randint(0, len(d))]
        elif r < 0.3:
            p = rng.randint(0, len(d))
            d = d[:p] + bytes([rng.randint(0, 255)]) + d[p:]
        elif r < 0.35:
            d = bytes(rng.randint(0, 255) for _ in range(rng.randint(1, 60)))
        docs.append(d)
    return docs


@pytest.mark.parametrize("name", ["bert_base_tok.bin", "bert_base_cased_tok.bin", "bert_chinese.bin"])
def test_wp_kernel_source_matches_oracle(wpsim, name):
    """wp_tokenize_kernel<uint16_t>: decode and validation, sync points, chunk ordering, walks, emission, and the
    multi-window carry (documents beyond 576 code points), ragged and invalid inputs, truncation."""
    docs = wp_docs(5, 320)
    check_wp(wpsim, name, docs, 512)
    check_wp(wpsim, name, docs[:120], 7, unk=7777)


def test_wp_kernel_source_learned_words(wpsim):
    """The run-time side of the word table: several warps add words concurrently (same words in different documents), a
    second pass over the same model finds them, a third one with another UnkId and a small output cap does too."""
    name = "bert_base_tok.bin"
    words = [b"unaffable", b"antidisestablishmentarianism", b"supercalifragilisticexpialidocious", b"qwrtzxqwrtzx", b"a" * 24,
             b"a" * 25, b"b" * 13, "naïveté".encode(), b"electroencephalography", b"zzzzzzzzzzzzzzzzzzzzzzzz", b"hydroxychloroquine",
             b"internationalization", b"1234567890123", b"tokenization's"]
    docs = [b" ".join(words[i:] + words[:i]) for i in range(len(words))] * 3 + wp_docs(21, 60)
    h = wpsim.wpsim_load(model_path(name).encode())
    check_wp(wpsim, name, docs, 512, warps=8, handle=h)
    check_wp(wpsim, name, docs, 512, warps=3, handle=h)
    check_wp(wpsim, name, docs, 9, unk=4242, warps=5, handle=h)
    wpsim.wpsim_free(h)


def test_wp_kernel_source_wide_table(wpsim):
    """wp_tokenize_kernel<uint32_t>: bert_multi_cased (232k states x 10 004 classes, 32-bit table entries)."""
    check_wp(wpsim, "bert_multi_cased.bin", wp_docs(9, 120), 512)


# ---- the generic lexer engine ------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def lexsim():
    L = ctypes.CDLL(os.path.join(ROOT, "tests", "simt", "liblex_simt.so"))
    L.lexsim_load.restype = ctypes.c_void_p
    L.lexsim_load.argtypes = [ctypes.c_char_p]
    L.lexsim_error.restype = ctypes.c_char_p
    L.lexsim_error.argtypes = [ctypes.c_void_p]
    L.lexsim_free.argtypes = [ctypes.c_void_p]
    L.lexsim_ids.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.lexsim_triples.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_int]
    return L


def lex_docs(n):
    lines = read_lines("test.multi.txt")[:n] + read_lines("test.txt")[:n]
    return lines + [b" ".join(lines[i:i + 5]) for i in range(0, 2 * n, 40)] + [
        b"\xef\xbb\xbfbom first", b"abc \xff def", b"a" * 400, "naïve café 我爱北京 [unk] qwrtzx".encode(), b" ", b"x", b"hello\x00world",
        b"Dr. Smith went to Washington. He arrived at 5 p.m. It was late!", b"tab\tnew\nline\r\n"]


@pytest.mark.parametrize("name", ["wbd.bin", "wbd_chuni.bin", "bert_base_tok.bin", "bert_chinese.bin"])
def test_lexer_kernel_source_ids_and_offsets(lexsim, name):
    """lex_decode_kernel -> lex_run_kernel -> lex_wp[_offsets]_kernel: TextToIds[WithOffsets]_wp for any [wbd] grammar."""
    h = lexsim.lexsim_load(model_path(name).encode())
    assert lexsim.lexsim_error(h) == b"", lexsim.lexsim_error(h)
    o = Oracle()
    ho = o.load(model_path(name))
    docs = [d for d in lex_docs(150) if d]
    offs = np.zeros(len(docs) + 1, np.int64)
    np.cumsum([len(d) for d in docs], out=offs[1:])
    buf = b"".join(docs) + b"\0"
    for max_ids, with_offsets in ((256, True), (256, False), (3, True)):
        ids = np.full((len(docs), max_ids), -7, np.int32)
        counts = np.full(len(docs), -7, np.int32)
        st = np.full((len(docs), max_ids), -7, np.int32)
        en = np.full((len(docs), max_ids), -7, np.int32)
        assert lexsim.lexsim_ids(h, buf, offs.ctypes.data, len(docs), ids.ctypes.data, counts.ctypes.data,
                                 st.ctypes.data if with_offsets else None, en.ctypes.data if with_offsets else None, max_ids, 100, 2) == 0
        for i, d in enumerate(docs):
            n, a, s, e = o.text_to_ids_with_offsets(ho, d, max_ids, 100)
            assert counts[i] == n and (ids[i, :n] == a[:n]).all() and (ids[i, n:] == -7).all(), (name, d[:60])
            if with_offsets:
                assert (st[i, :n] == s[:n]).all() and (en[i, :n] == e[:n]).all(), (name, d[:60])
    o.free(ho)
    lexsim.lexsim_free(h)


@pytest.mark.parametrize("name", ["wbd.bin", "sbd.bin", "wbd_chuni.bin"])
def test_lexer_kernel_source_triples(lexsim, name):
    """The TextToWords / TextToSentences view (no charmap, U+0000 -> U+0020): the (Tag, From, To) triples of the lexer
    kernels against FALexTools_t::Process in the oracle."""
    h = lexsim.lexsim_load(model_path(name).encode())
    assert lexsim.lexsim_error(h) == b"", lexsim.lexsim_error(h)
    o = Oracle()
    ho = o.load(model_path(name))
    o.lib.bfo_lex_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    docs = [d for d in lex_docs(120) if d]
    offs = np.zeros(len(docs) + 1, np.int64)
    np.cumsum([len(d) for d in docs], out=offs[1:])
    buf = b"".join(docs) + b"\0"
    ncps = np.zeros(len(docs), np.int32)
    tri = np.zeros(3 * int(offs[-1]) + 8, np.int32)
    tri_count = np.zeros(len(docs), np.int32)
    assert lexsim.lexsim_triples(h, buf, offs.ctypes.data, len(docs), ncps.ctypes.data, tri.ctypes.data, tri_count.ctypes.data, 2) == 0
    for i, d in enumerate(docs):
        try:
            body = d[3:] if d[:3] == b"\xef\xbb\xbf" else d
            cps = np.array([0x20 if ord(ch) == 0 else ord(ch) for ch in body.decode("utf-8")], np.int32)
        except UnicodeDecodeError:
            assert ncps[i] <= 0, d[:40]
            continue
        if len(cps) == 0:
            assert ncps[i] <= 0
            continue
        assert ncps[i] == len(cps), d[:40]
        exp = np.zeros(3 * len(cps) + 3, np.int32)
        rn = o.lib.bfo_lex_process(ho, cps.ctypes.data, len(cps), exp.ctypes.data, 3 * len(cps))
        got = tri[3 * int(offs[i]): 3 * int(offs[i]) + int(tri_count[i])]
        assert tri_count[i] == rn and (got == exp[:rn]).all(), (name, d[:60])
    o.free(ho)
    lexsim.lexsim_free(h)
