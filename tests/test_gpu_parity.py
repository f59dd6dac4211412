"""Parity tests proper: the CUDA path, called through the C ABI (blingfire_b200 -> libblingfiretokdll.so),
against the oracle on the same inputs, against the committed golden fixtures (produced by the
reference itself), and -- at BASELINE.json's full sizes -- through size-independent properties.
Bit-exact: ids and counts are integers."""
import base64
import json
import os
import random

import numpy as np
import pytest

from _common import GOLDEN, Oracle, fnv1a64_ids, have_data, model_path, read_lines

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_data(), reason="data/ not staged")]

BERT_TEXT = ("Эpple pie. How do I renew my virtual smart card?: /Microsoft IT/ 'virtual' smart card certificates for "
             "DirectAccess are valid for one year. In order to get to microsoft.com we need to type pi@1.2.1.2.")
BERT_IDS = [1208, 9397, 2571, 11345, 1012, 2129, 2079, 1045, 20687, 2026, 7484, 6047, 4003, 1029, 1024, 1013, 7513,
            2009, 1013, 1005, 7484, 1005, 6047, 4003, 17987, 2005, 3622, 6305, 9623, 2015, 2024, 9398, 2005, 2028,
            2095, 1012, 1999, 2344, 2000, 2131, 2000, 7513, 1012, 4012, 2057, 2342, 2000, 2828, 14255, 1030, 1015,
            1012, 1016, 1012, 1015, 1012, 1016, 1012]


@pytest.fixture(scope="module")
def bf():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
    import blingfire_b200
    return blingfire_b200


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


_models = {}


def gpu_model(bf, name):
    if name not in _models:
        _models[name] = bf.load_model(model_path(name))
        assert bf.lib().BlingFireB200ModelEngine(_models[name]) == 1, f"{name}: no GPU engine"
    return _models[name]


def check_batch(bf, oracle, name, docs, max_ids, unk):
    h = gpu_model(bf, name)
    ho = oracle.load(model_path(name))
    buf, offs = bf.make_csr(docs)
    ids, counts = bf.text_to_ids_batch(h, (buf, offs), max_ids, unk)
    _, oids, ocounts = oracle.batch(ho, buf if len(buf) else np.zeros(1, np.uint8), offs, max_ids, unk, threads=8)
    bad = np.nonzero(counts != ocounts)[0]
    assert len(bad) == 0, f"{name}: count mismatch at docs {bad[:5]} gpu={counts[bad[:5]]} oracle={ocounts[bad[:5]]} {docs[bad[0]][:80]!r}"
    mask = np.arange(max_ids)[None, :] < counts[:, None]
    diff = np.nonzero(((ids != oids) & mask).any(axis=1))[0]
    assert len(diff) == 0, f"{name}: id mismatch at docs {diff[:5]} {docs[diff[0]][:80]!r}"
    assert (ids[~mask] == 0).all(), "tail of a row was written"
    # the compact entry point agrees with the row-major one
    cids, coffs = bf.text_to_ids_batch_csr(h, (buf, offs), max_ids, unk)
    assert (np.diff(coffs) == counts).all()
    assert (cids == ids[mask]).all()
    oracle.free(ho)
    return ids, counts


def test_known_answer_bert(bf):
    h = gpu_model(bf, "bert_base_tok.bin")
    ids = bf.text_to_ids(h, BERT_TEXT, 128, 100, no_padding=True)   # README.md:111,131-135
    assert ids.tolist() == BERT_IDS
    padded = bf.text_to_ids(h, BERT_TEXT, 128, 100)
    assert padded[:58].tolist() == BERT_IDS and (padded[58:] == 0).all()


def test_edge_cases_vs_golden(bf, golden):
    """Single-document TextToIds through the C ABI, incl. the untouched-tail contract."""
    import ctypes
    L = bf.lib()
    for case in golden["edge_cases"]:
        if case["model"] not in ("bert_base_tok.bin", "bert_base_cased_tok.bin", "bert_chinese.bin"):
            continue
        h = gpu_model(bf, case["model"])
        data = base64.b64decode(case["input"])
        out = np.full(case["max_ids"], -7, np.int32)
        n = L.TextToIds(ctypes.c_void_p(h), data, len(data), out.ctypes.data, case["max_ids"], case["unk"])
        assert n == case["count"], (case["model"], data[:40])
        assert out[:n].tolist() == case["ids"], (case["model"], data[:40])
        assert (out[n:] == -7).all(), "ids beyond the returned count must stay untouched"


def test_corpus_digests_vs_golden(bf, golden):
    for d in golden["digests"]:
        if d["model"] not in ("bert_base_tok.bin", "bert_base_cased_tok.bin", "bert_chinese.bin"):
            continue
        lines = read_lines(d["corpus"], drop_empty=False)[: d["lines"]]
        docs = [b" ".join(lines[i:i + d["group"]]) for i in range(0, len(lines), d["group"])]
        h = gpu_model(bf, d["model"])
        cids, coffs = bf.text_to_ids_batch_csr(h, docs, d["max_ids"], d["unk"])
        assert int(coffs[-1]) == d["tokens"], d
        assert f"{fnv1a64_ids(cids):016x}" == d["fnv1a64"], d


@pytest.mark.parametrize("name,corpus,nlines,group,max_ids", [
    ("bert_base_tok.bin", "test.txt", 40000, 1, 128),
    ("bert_base_tok.bin", "test.txt", 40000, 12, 512),
    ("bert_base_tok.bin", "test.txt", 30000, 60, 4096),        # multi-window documents
    ("bert_base_tok.bin", "test.multi.txt", 20000, 3, 512),    # 2-4 byte code points
    ("bert_base_cased_tok.bin", "test.txt", 20000, 5, 512),
    ("bert_chinese.bin", "test.multi.txt", 20000, 2, 512),
])
def test_batch_matches_oracle_on_corpora(bf, oracle, name, corpus, nlines, group, max_ids):
    lines = read_lines(corpus, drop_empty=False)[:nlines]
    docs = [b" ".join(lines[i:i + group]) for i in range(0, len(lines), group)]
    check_batch(bf, oracle, name, docs, max_ids, 100)


def test_ragged_edge_and_invalid_documents(bf, oracle):
    rng = random.Random(5)
    lines = read_lines("test.multi.txt")[:3000] + read_lines("test.txt")[:3000]
    docs = [b"", b" ", b"\xef\xbb\xbf", b"\xef\xbb\xbfhello", b"abc \xff def", b"\xe6\x88", b"hello\x00world", b"a" * 400,
            b"a" * 1000 + b" " + b"b" * 700, "我".encode() * 900, b"ab" * 700 + b"\xff", b"." * 700, ("word " * 300).encode(),
            b"x" * 513, b"y" * 511 + "é".encode(), ("é" * 700).encode(), b"\xf0\x9f\x98\x80" * 300, b"\x80" + b"a" * 600]
    for _ in range(6000):
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
    for max_ids in (300, 7):
        check_batch(bf, oracle, "bert_base_tok.bin", docs, max_ids, 100)


def test_learned_words_on_a_fresh_model(bf, oracle):
    """The run-time side of the word table on the device: thousands of warps add the same words at once on a model that has
    just been loaded; the following passes (same ids, then another UnkId and a small cap) find them in the table."""
    name = "bert_base_tok.bin"
    words = [b"unaffable", b"antidisestablishmentarianism", b"supercalifragilisticexpialidocious", b"qwrtzxqwrtzx", b"a" * 24,
             b"a" * 25, b"b" * 13, "naïveté".encode(), b"electroencephalography", b"zzzzzzzzzzzzzzzzzzzzzzzz", b"hydroxychloroquine",
             b"internationalization", b"1234567890123", b"tokenization's", b"pneumonoultramicroscopicsilicovolcanoconiosis"]
    lines = read_lines("test.txt", drop_empty=False)[60000:90000]
    docs = [b" ".join(words[i % len(words):] + words[:i % len(words)]) for i in range(20000)]
    docs += [b" ".join(lines[i:i + 6]) for i in range(0, len(lines), 6)]
    saved = _models.pop(name, None)
    try:
        _models[name] = bf.load_model(model_path(name))
        for max_ids, unk in ((512, 100), (512, 100), (9, 4242)):
            check_batch(bf, oracle, name, docs, max_ids, unk)
        bf.free_model(_models.pop(name))
    finally:
        if saved is not None:
            _models[name] = saved


@pytest.mark.timeout(900)
def test_wide_table_model_on_the_device(bf, oracle):
    """bert_multi_cased: 232k states x 10 004 classes -- the 32-bit table entries (9.3 GB in HBM), 14-bit classes in the
    word keys (two per key word), 119k-word vocabulary."""
    name = "bert_multi_cased.bin"
    lines = read_lines("test.multi.txt", drop_empty=False)[:30000]
    docs = [b" ".join(lines[i:i + 3]) for i in range(0, len(lines), 3)]
    docs += [b"", b"abc \xff def", "我爱北京".encode() * 200, b"a" * 700, "é".encode() * 700, b"[UNK] [CLS]x [unused1]"]
    check_batch(bf, oracle, name, docs, 512, 100)
    check_batch(bf, oracle, name, docs[:3000], 512, 100)
    bf.free_model(_models.pop(name))


def test_unaligned_offsets_and_unk_id(bf, oracle):
    """Documents starting at every byte alignment; a non-default UnkId."""
    docs = [b"x" * k + b" unaffable qwrtzx " + "naïve café".encode() for k in range(0, 9)] * 50
    check_batch(bf, oracle, "bert_base_tok.bin", docs, 64, 7777)


def test_device_entry_point_matches_host(bf, oracle):
    import torch
    import corpus
    h = gpu_model(bf, "bert_base_tok.bin")
    text, offs = corpus.gen_docs("EN", 20000, seed=2, fixed_len=512)
    ids, counts = bf.text_to_ids_batch(h, (text, offs), 512, 100)
    d_text = torch.empty(len(text) + 64, dtype=torch.uint8, device="cuda")
    d_text[: len(text)].copy_(torch.from_numpy(text))
    d_offs = torch.from_numpy(offs).cuda()
    d_ids = torch.zeros((len(offs) - 1, 512), dtype=torch.int32, device="cuda")
    d_counts = torch.zeros(len(offs) - 1, dtype=torch.int32, device="cuda")
    bf.text_to_ids_batch_device(h, d_text.data_ptr(), d_offs.data_ptr(), len(offs) - 1, int(offs[-1]),
                                d_ids.data_ptr(), d_counts.data_ptr(), 512, 100, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (d_counts.cpu().numpy() == counts).all()
    assert (d_ids.cpu().numpy() == ids).all()


def test_full_size_properties(bf, oracle):
    """cfg 2 at full size (1 M docs x ~512 B): size-independent properties (every document against the reference is
    tests/test_gpu_fullsize.py).
      * replica / rotation invariance: rotating the document order rotates the result rows
      * concatenation: per-document counts sum to the CSR total
      * a strided sample is checked id-for-id against the oracle port (the full-size test uses oracle/_ref)."""
    import corpus
    n = 1_000_000
    h = gpu_model(bf, "bert_base_tok.bin")
    text, offs = corpus.cfg2(n)
    cids, coffs = bf.text_to_ids_batch_csr(h, (text, offs), 512, 100)
    counts = np.diff(coffs)
    assert counts.min() > 0 and counts.max() <= 512
    assert int(coffs[-1]) == len(cids)
    # rotation by 15 625 docs (cfg 5's replica rule): same multiset of rows, rotated
    rot = 15625
    lens = np.diff(offs)
    order = np.roll(np.arange(n), -rot)
    r_offs = np.zeros(n + 1, np.int64)
    np.cumsum(lens[order], out=r_offs[1:])
    r_text = np.concatenate([text[offs[rot]:], text[:offs[rot]]])
    rids, roffs = bf.text_to_ids_batch_csr(h, (r_text, r_offs), 512, 100)
    assert (np.diff(roffs) == counts[order]).all()
    assert (rids == np.concatenate([cids[coffs[rot]:], cids[:coffs[rot]]])).all()
    # strided sample against the oracle
    ho = oracle.load(model_path("bert_base_tok.bin"))
    for d in range(0, n, 997):
        doc = bytes(text[offs[d]:offs[d + 1]])
        k, oids = oracle.text_to_ids(ho, doc, 512, 100)
        assert k == counts[d] and (oids[:k] == cids[coffs[d]:coffs[d + 1]]).all(), d


# ---- generic lexer engine: TextToWords (cfg 1) and TextToIds on non-FastPath grammars ----

def test_text_to_words_vs_golden(bf, golden):
    """Default word breaker through the C ABI against outputs of the reference itself."""
    import ctypes
    L = bf.lib()
    for w in golden["words"]:
        data = base64.b64decode(w["input"])
        cap = 2 * len(data) + 16
        out = ctypes.create_string_buffer(cap)
        n = L.TextToWords(data, len(data), out, cap)
        assert n == w["ret"], data[:40]
        if n > 0:
            assert out.raw[:n] == base64.b64decode(w["out"]), data[:40]


def test_cfg1_text_to_words_10k_ascii_lines(bf, oracle):
    """BASELINE cfg 1: default TextToWords on the first 10 000 pure-ASCII lines <= 120 B, byte for
    byte against the oracle; plus the too-small-buffer and README examples."""
    import ctypes
    L = bf.lib()
    ho = oracle.load(model_path("wbd.bin"))
    lines = [l for l in read_lines("test.txt") if len(l) <= 120 and all(c < 128 for c in l)][:10000]
    assert len(lines) == 10000
    out = ctypes.create_string_buffer(1024)
    for l in lines:
        n = L.TextToWords(l, len(l), out, 1024)
        n2, s2 = oracle.text_to_words(ho, l, 1024)
        assert n == n2 and out.raw[:n] == s2, l
    # README.md:72-74
    s = "I saw a girl with a telescope.".encode()
    n = L.TextToWords(s, len(s), out, 1024)
    assert out.raw[:n - 1] == b"I saw a girl with a telescope ."
    # output does not fit: the required size is returned, nothing is copied (blingfiretokdll.cpp:560-565)
    small = ctypes.create_string_buffer(b"\x7f" * 8, 8)
    assert L.TextToWords(s, len(s), small, 8) == n and small.raw == b"\x7f" * 8
    assert L.TextToWords(b"", 0, out, 1024) == 0
    assert L.TextToWords(b"\xff\xfe", 2, out, 1024) == -1


def test_text_to_words_batch(bf, oracle):
    """The additive batch form (lexer AND string building on the GPU) against the per-document reference semantics:
    cfg 1's 10 000 lines with the default model, multilingual and edge documents with wbd.bin / wbd_chuni.bin handles,
    a too-small buffer."""
    import ctypes
    import time
    L = bf.lib()
    lines = [l for l in read_lines("test.txt") if len(l) <= 120 and all(c < 128 for c in l)][:10000]
    ho = oracle.load(model_path("wbd.bin"))
    want = [oracle.text_to_words(ho, l, 1024) for l in lines]
    buf, offs = bf.make_csr(lines)
    out, out_offs, results = bf.text_to_words_batch((buf, offs), raw=True)       # warm-up + result
    t0 = time.perf_counter()
    out, out_offs, results = bf.text_to_words_batch((buf, offs), raw=True)
    dt = time.perf_counter() - t0
    data = out.tobytes()
    for i, (n2, s2) in enumerate(want):
        assert results[i] == n2 and data[out_offs[i]:out_offs[i + 1]] == s2, lines[i]
    assert len(buf) / dt / 1e6 > 19.6, f"batch TextToWords at {len(buf) / dt / 1e6:.1f} MB/s: the reference does 19.6 MB/s on one CPU thread"
    assert bf.text_to_words_batch(["I saw a girl with a telescope.", ""]) == ["I saw a girl with a telescope .", ""]
    docs = read_lines("test.multi.txt", drop_empty=False)[:4000] + [b"", b"\xff\xfe", b"\xef\xbb\xbf", b"\xef\xbb\xbfbom", b"a\x00b c", b"x" * 5000,
                                                                    b"can't won't cannot U.S.A. 3.14 e-mail", b" ", b"\xe6\x88"]
    for name in ("wbd.bin", "wbd_chuni.bin"):
        h = bf.load_model(model_path(name))
        ho2 = oracle.load(model_path(name))
        out, out_offs, results = bf.text_to_words_batch(bf.make_csr(docs), h=h, raw=True)
        data = out.tobytes()
        for i, d in enumerate(docs):
            n2, s2 = oracle.text_to_words(ho2, d, 2 * len(d) + 16)
            assert results[i] == n2, (name, d[:40], int(results[i]), n2)
            assert data[out_offs[i]:out_offs[i + 1]] == (s2 if n2 > 0 else b""), (name, d[:40])
        bf.free_model(h)
    # capacity too small: -total, offsets and results complete
    b2, o2 = bf.make_csr(lines[:100])
    oo = np.zeros(101, np.int64); rr = np.zeros(100, np.int32); small = np.zeros(8, np.uint8)
    r = L.TextToWordsBatch(None, b2.ctypes.data, o2.ctypes.data, 100, small.ctypes.data, 8, oo.ctypes.data, rr.ctypes.data)
    assert r == -int(sum(w[0] for w in want[:100])) and oo[100] == -r and (rr == [w[0] for w in want[:100]]).all()


def test_text_to_sentences_batch(bf, oracle):
    """The additive batch form of TextToSentences against the per-document reference semantics: default model and an
    sbd.bin handle, paragraphs of several sentences, leading white space, embedded newlines and NULs, invalid input."""
    lines = read_lines("test.txt", drop_empty=False)[:3000]
    docs = [b" ".join(lines[i:i + 5]) for i in range(0, len(lines), 5)]
    docs += [b"", b"  Hello there.  How are you?\nFine, thanks!  ", b"Dr. Smith went to Washington. He arrived at 5 p.m. It was late!", b"\xff\xfe", b"a\x00b. c\nd.",
             b"   ", b"No end", "Это первое. А это второе? Да!".encode(), b"x" * 3000] + read_lines("test.multi.txt")[:800]
    ho = oracle.load(model_path("sbd.bin"))
    h = bf.load_model(model_path("sbd.bin"))
    for handle in (None, h):
        out, out_offs, results = bf.text_to_sentences_batch(bf.make_csr(docs), h=handle, raw=True)
        data = out.tobytes()
        for i, d in enumerate(docs):
            n2, s2, _, _ = oracle.split("sentences", ho, d, 2 * len(d) + 16)
            assert results[i] == n2, (d[:40], int(results[i]), n2)
            assert data[out_offs[i]:out_offs[i + 1]] == (s2 if n2 > 0 else b""), d[:40]
    assert bf.text_to_sentences_batch(["Hello there. How are you?"]) == ["Hello there.\nHow are you?"]
    bf.free_model(h)


def test_text_to_words_with_model_multilingual(bf, oracle):
    import ctypes
    L = bf.lib()
    for name in ("wbd.bin", "wbd_chuni.bin", "sbd.bin"):
        h = bf.load_model(model_path(name))
        ho = oracle.load(model_path(name))
        out = ctypes.create_string_buffer(1 << 16)
        for l in read_lines("test.multi.txt")[:1500] + [b"a\x00b c", b"can't won't cannot U.S.A. 3.14 e-mail", b"x" * 700]:
            n = L.TextToWordsWithModel(l, len(l), out, 1 << 16, ctypes.c_void_p(h))
            n2, s2 = oracle.text_to_words(ho, l, 1 << 16)
            assert n == n2 and (n <= 0 or out.raw[:n] == s2), (name, l[:60])
        bf.free_model(h)


def test_generic_engine_text_to_ids(bf, oracle):
    """TextToIds on a lexer model outside the FastPath shape (wbd.bin: contexts, tag-less actions,
    nested calls) goes through the generic engine + exact post-pass."""
    h = bf.load_model(model_path("wbd.bin"))
    assert bf.lib().BlingFireB200ModelEngine(h) == 2
    ho = oracle.load(model_path("wbd.bin"))
    docs = read_lines("test.txt")[:3000] + read_lines("test.multi.txt")[:1000] + [b"", b"\xff", b"a" * 900]
    buf, offs = bf.make_csr(docs)
    ids, counts = bf.text_to_ids_batch(h, (buf, offs), 64, 100)
    _, oids, ocounts = oracle.batch(ho, buf, offs, 64, 100, threads=4)
    assert (counts == ocounts).all()
    mask = np.arange(64)[None, :] < counts[:, None]
    assert (ids[mask] == oids[mask]).all()
    bf.free_model(h)
