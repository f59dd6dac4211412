"""TextToIdsWithOffsets through the C ABI, for lexer models (the _wp branch) and
[pos-dict] models (the _sp branch): ids, start offsets and end offsets against the golden fixtures
(the reference itself) and the oracle."""
import base64
import ctypes
import json
import os
import random

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


@pytest.mark.parametrize("name", ["bert_base_tok.bin", "xlm_roberta_base.bin", "gpt2.bin"])
def test_offsets_vs_golden(bf, name):
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        golden = json.load(f)
    h = bf.load_model(model_path(name))
    n_checked = 0
    for c in golden["ids_with_offsets"]:
        if c["model"] != name:
            continue
        data = base64.b64decode(c["input"])
        ids, st, en = bf.utf8text_to_ids_with_offsets(h, data, 256, c["unk"], no_padding=True)
        assert len(ids) == c["count"], data[:40]
        assert ids.astype(np.int64).tolist() == c["ids"] and st.tolist() == c["starts"], data[:40]
        # a token that is only the dummy prefix has start -1; its end offset is an out-of-bounds read
        # in the reference (blingfiretokdll.cpp:1527) and is excluded from the comparison
        keep = [k for k in range(len(ids)) if c["starts"][k] >= 0 or c["ends"][k] == -1]
        assert [int(en[k]) for k in keep] == [c["ends"][k] for k in keep], data[:40]
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
    L = bf.lib()
    a = np.full(64, -7, np.int32); b = np.full(64, -7, np.int32); c = np.full(64, -7, np.int32)
    n = L.TextToIdsWithOffsets(ctypes.c_void_p(h), b"hello world", 11, a.ctypes.data, b.ctypes.data, c.ctypes.data, 64, 100)
    assert n >= 1 and (a[n:] == -7).all() and (b[n:] == -7).all() and (c[n:] == -7).all()
    bf.free_model(h)


def _sp_docs(seed):
    rng = random.Random(seed)
    lines = read_lines("test.multi.txt")[:1500] + read_lines("test.txt")[:1500]
    docs = rng.sample(lines, 250)
    docs += [b" ".join(rng.choice(lines) for _ in range(6)) for _ in range(20)]          # beyond the smem window
    docs += [b"\xef\xbb\xbfbom first", b"abc \xff def", b"a" * 700, b"." * 300, b" ", b"x", b"  lead and trail  ",
             "ﬁne ＡＢＣ ½ ™ ｶﾞ".encode(), "a b c　d".encode(), "🙂🙂 emoji 👩‍👩‍👧".encode(),
             b"tab\tnew\nline\r\n", "é combining".encode()]
    return docs


# one model of every [pos-dict] kind: Unigram + charmap, BPE-opt raw bytes, Unigram w/o normalisation,
# BPE with merges ranks, plain BPE, Unigram 100k
@pytest.mark.parametrize("name,unk", [("xlm_roberta_base.bin", 3), ("gpt2.bin", 0), ("xlnet_nonorm.bin", 0), ("roberta.bin", 3),
                                      ("bpe_example.bin", 1), ("laser100k.bin", 1), ("xlnet.bin", 0)])
def test_sp_offsets_vs_oracle(bf, name, unk):
    h = bf.load_model(model_path(name))
    assert bf.lib().BlingFireB200ModelEngine(ctypes.c_void_p(h)) == 3
    o = Oracle()
    ho = o.load(model_path(name))
    for d in _sp_docs(11):
        for max_ids in (1024, 5):
            n, oi, os_, oe = o.text_to_ids_with_offsets(ho, d, max_ids, unk)
            ids, st, en = bf.utf8text_to_ids_with_offsets(h, d, max_ids, unk, no_padding=True)
            assert len(ids) == n, d[:40]
            assert (ids.astype(np.int32) == oi[:n]).all(), d[:40]
            assert (st == os_[:n]).all() and (en == oe[:n]).all(), d[:40]
            # the ids agree with the offset-less entry point
            plain = bf.text_to_ids(h, d, max_ids, unk, no_padding=True)
            assert (plain == ids).all(), d[:40]
    L = bf.lib()
    L.TextToIdsWithOffsets_sp.restype = ctypes.c_int
    L.TextToIdsWithOffsets_sp.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    a = np.full(64, -7, np.int32); b = np.full(64, -7, np.int32); c = np.full(64, -7, np.int32)
    n = L.TextToIdsWithOffsets_sp(ctypes.c_void_p(h), b"hello world", 11, a.ctypes.data, b.ctypes.data, c.ctypes.data, 64, unk)
    assert n >= 1 and (a[n:] == -7).all() and (b[n:] == -7).all() and (c[n:] == -7).all()
    # the _wp spelling refuses a [pos-dict] model, like the reference's dispatch would never route it there
    L.TextToIdsWithOffsets_wp.restype = ctypes.c_int
    L.TextToIdsWithOffsets_wp.argtypes = L.TextToIdsWithOffsets_sp.argtypes
    assert L.TextToIdsWithOffsets_wp(ctypes.c_void_p(h), b"hello world", 11, a.ctypes.data, b.ctypes.data, c.ctypes.data, 64, unk) == 0
    bf.free_model(h)


@pytest.mark.parametrize("name,unk", [("bert_base_tok.bin", 100), ("bert_chinese.bin", 100), ("xlm_roberta_base.bin", 3), ("gpt2.bin", 0),
                                      ("xlnet.bin", 0)])
def test_offsets_batch_vs_oracle(bf, name, unk):
    """The additive batch form of TextToIdsWithOffsets: row-major ids / starts / ends with untouched tails, per-document
    counts, against the per-document oracle; a small cap; more documents than one chunk holds."""
    h = bf.load_model(model_path(name))
    o = Oracle()
    ho = o.load(model_path(name))
    docs = _sp_docs(5) + [b""] + read_lines("test.txt")[:1500]
    for max_ids in (300, 4):
        ids, st, en, counts = bf.text_to_ids_with_offsets_batch(h, docs, max_ids, unk)
        for i, d in enumerate(docs):
            n, oi, os_, oe = o.text_to_ids_with_offsets(ho, d, max_ids, unk)
            assert counts[i] == n, (name, d[:40], int(counts[i]), n)
            assert (ids[i, :n] == oi[:n]).all() and (st[i, :n] == os_[:n]).all(), (name, d[:40])
            keep = [k for k in range(n) if os_[k] >= 0 or oe[k] == -1]         # (blingfiretokdll.cpp:1527, see test_offsets_vs_golden)
            assert (en[i, keep] == oe[keep]).all(), (name, d[:40])
            assert (ids[i, n:] == 0).all() and (st[i, n:] == 0).all() and (en[i, n:] == 0).all()
        # the compact form returns the same entries back to back
        cid, cst, cen, off = bf.text_to_ids_with_offsets_batch_csr(h, docs, max_ids, unk)
        assert off[0] == 0 and (np.diff(off) == counts).all() and len(cid) == off[-1] == counts.sum()
        mask = np.arange(max_ids)[None, :] < counts[:, None]
        assert (cid == ids[mask]).all() and (cst == st[mask]).all() and (cen == en[mask]).all()
    # more documents than one chunk holds (16 M cells): the same rows, repeated
    reps = (16 << 20) // (len(docs) * 300) + 2
    bid, bst, ben, boff = bf.text_to_ids_with_offsets_batch_csr(h, docs * reps, 300, unk)
    sid, sst, sen, soff = bf.text_to_ids_with_offsets_batch_csr(h, docs, 300, unk)
    assert len(bid) == reps * len(sid) and (np.diff(boff) == np.tile(np.diff(soff), reps)).all()
    assert (bid == np.tile(sid, reps)).all() and (bst == np.tile(sst, reps)).all() and (ben == np.tile(sen, reps)).all()
    # more chunks than the call keeps in flight (a wide row makes a chunk hold few documents): the slots are reused
    reps = 5 * (16 << 20) // (len(docs) * 2048) + 2
    bid, bst, ben, boff = bf.text_to_ids_with_offsets_batch_csr(h, docs * reps, 2048, unk)
    sid, sst, sen, soff = bf.text_to_ids_with_offsets_batch_csr(h, docs, 2048, unk)
    assert len(bid) == reps * len(sid) and (np.diff(boff) == np.tile(np.diff(soff), reps)).all()
    assert (bid == np.tile(sid, reps)).all() and (bst == np.tile(sst, reps)).all() and (ben == np.tile(sen, reps)).all()
    # too small a capacity: the need comes back negated, the offsets are complete
    L = bf.lib()
    buf, offs = bf.make_csr(docs)
    small = np.full((3, 16), -7, np.int32)
    off2 = np.zeros(len(docs) + 1, np.int64)
    r = L.TextToIdsWithOffsetsBatchCsr(ctypes.c_void_p(h), buf.ctypes.data, offs.ctypes.data, len(docs), small[0].ctypes.data,
                                       small[1].ctypes.data, small[2].ctypes.data, 16, off2.ctypes.data, 4, unk)
    assert r == -int(off[-1]) and (off2 == off).all()
    bf.free_model(h)


def _split_via_capi(bf, fn_name, data, model, max_out):
    from _common import split_call
    L = bf.lib()
    f = getattr(L, fn_name)
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    return split_call(f, data, model, max_out)


def test_words_and_sentences_with_offsets_vs_golden(bf):
    """TextTo{Words,Sentences}WithOffsetsWithModel with the default models (wbd.bin / sbd.bin next to the
    library; the reference embeds the same bytes) against what the reference itself returned."""
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        golden = json.load(f)
    n_checked = 0
    for c in golden["split_with_offsets"]:
        data = base64.b64decode(c["input"])
        fn = "TextToWordsWithOffsetsWithModel" if c["kind"] == "words" else "TextToSentencesWithOffsetsWithModel"
        n, text, st, en = _split_via_capi(bf, fn, data, None, c["max_out"])
        assert n == c["ret"], (c["kind"], data[:40], bf.last_error())
        assert text == base64.b64decode(c["out"]), (c["kind"], data[:40])
        assert st[:64].tolist() == c["starts"] and en[:64].tolist() == c["ends"], (c["kind"], data[:40])
        n_checked += 1
    assert n_checked > 300


@pytest.mark.parametrize("kind,name", [("sentences", "sbd.bin"), ("words", "wbd.bin"), ("words", "wbd_chuni.bin")])
def test_words_and_sentences_vs_oracle(bf, kind, name):
    """The same calls with a model handle, on multilingual paragraphs, against the oracle; and the plain
    spellings (no offsets, default model) through the Python wrapper."""
    h = bf.load_model(model_path(name))
    o = Oracle()
    ho = o.load(model_path(name))
    lines = read_lines("test.multi.txt")[:600] + read_lines("test.txt")[:600]
    docs = [b" ".join(lines[i:i + 5]) for i in range(0, len(lines), 5)] + [b"\xef\xbb\xbfBom. Next!", b"abc \xff def", b" \n ", b"x",
                                                                          b"One.\nTwo?\r\nThree!  ", "句子一。句子二！".encode()]
    fn = "TextToWordsWithOffsetsWithModel" if kind == "words" else "TextToSentencesWithOffsetsWithModel"
    for d in docs:
        for max_out in (None, 2):
            n1, t1, s1, e1 = o.split(kind, ho, d, max_out)
            n2, t2, s2, e2 = _split_via_capi(bf, fn, d, h, max_out)
            assert n1 == n2 and t1 == t2, (kind, d[:40])
            assert (s1 == s2).all() and (e1 == e2).all(), (kind, d[:40])
    if kind == "sentences":
        assert bf.text_to_sentences("Hello world! How are you?  Fine.") == bf.text_to_sentences_with_model(h, "Hello world! How are you?  Fine.")
        text, st, en = bf.utf8text_to_sentences_with_offsets(b"Hello world! How are you?")
        assert text == b"Hello world!\nHow are you?" and st.tolist() == [0, 13] and en.tolist() == [11, 24]
    elif name == "wbd.bin":
        text, st, en = bf.utf8text_to_words_with_offsets("naïve café.".encode())
        assert text == "naïve café .".encode() and st.tolist() == [0, 7, 12] and en.tolist() == [5, 11, 12]
    bf.free_model(h)


def test_python_wrapper_string_offsets_vs_golden(bf):
    """text_to_words_with_offsets / text_to_sentences_and_offsets (code-point offsets) against the
    reference's own Python wrapper."""
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        golden = json.load(f)
    for c in golden["py_offsets"]:
        w, wo = bf.text_to_words_with_offsets(c["text"])
        assert w == c["words"] and [list(x) for x in wo] == c["word_offsets"], c["text"][:40]
        sn, so = bf.text_to_sentences_and_offsets(c["text"])
        assert sn == c["sentences"] and [list(x) for x in so] == c["sentence_offsets"], c["text"][:40]
