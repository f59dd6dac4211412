"""Shared test helpers: ctypes bindings for the CHECKERS (oracle/liboracle.so = plain-C
restatement, oracle/_ref/libblingfiretokdll.so = the reference compiled from its own
sources) and corpus/digest utilities.  Nothing here is product code."""
import ctypes
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "data")
LDB = os.path.join(DATA, "ldb")
CORPUS = os.path.join(DATA, "corpus")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def model_path(name):
    return os.path.join(LDB, name)


def have_data():
    return os.path.exists(model_path("bert_base_tok.bin")) and os.path.exists(os.path.join(CORPUS, "test.txt"))


def fnv1a64_ids(ids, h=0xcbf29ce484222325):
    """SURVEY 8c digest: FNV-1a-64 over the uint32 id stream (h ^= id; h *= prime)."""
    for v in ids:
        h ^= int(v) & 0xFFFFFFFF
        h = (h * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def fnv1a64_ids_np(ids_u32):
    """Same digest, vectorised in chunks via Python ints (exact)."""
    h = 0xcbf29ce484222325
    for v in np.asarray(ids_u32, dtype=np.uint32).tolist():
        h ^= v
        h = (h * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


class Oracle:
    """Plain-C restatement (oracle/bf_oracle.c)."""

    def __init__(self):
        p = os.path.join(ROOT, "oracle", "liboracle.so")
        self.lib = L = ctypes.CDLL(p)
        L.bfo_load_model.restype = ctypes.c_void_p
        L.bfo_load_model.argtypes = [ctypes.c_char_p]
        L.bfo_free_model.argtypes = [ctypes.c_void_p]
        L.bfo_text_to_ids.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.bfo_text_to_ids_with_offsets.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.bfo_text_to_words.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        for fn in ("bfo_text_to_words_with_offsets", "bfo_text_to_sentences_with_offsets"):
            getattr(L, fn).argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_int]
        L.bfo_set_no_dummy_prefix.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.bfo_ids_to_text.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.bfo_text_to_ids_digests.restype = ctypes.c_int64
        L.bfo_text_to_ids_digests.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.bfo_csr_digests.restype = None
        L.bfo_csr_digests.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int]
        L.bfo_fold_digests.restype = ctypes.c_uint64
        L.bfo_fold_digests.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        L.bfo_text_to_ids_batch.restype = ctypes.c_int64
        L.bfo_text_to_ids_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        for f in ("bfo_dfa_initial", "bfo_has_seg"):
            getattr(L, f).argtypes = [ctypes.c_void_p]
        L.bfo_dfa_get_dest.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.bfo_dfa_is_final.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.bfo_dfa_get_ow.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.bfo_iwmap_new_iw.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.bfo_fn_ini.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.bfo_charmap_get.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]

    def load(self, path):
        h = self.lib.bfo_load_model(path.encode())
        assert h, f"oracle failed to load {path}"
        return h

    def free(self, h):
        self.lib.bfo_free_model(h)

    def set_no_dummy_prefix(self, h, flag):
        return self.lib.bfo_set_no_dummy_prefix(h, int(bool(flag)))

    def ids_to_text(self, h, ids, max_out, skip_special):
        return ids_to_text_call(self.lib.bfo_ids_to_text, h, ids, max_out, skip_special)

    def digests(self, h, buf, offsets, max_ids, unk, threads=None):
        """Per-document (FNV-1a-64 digest, count) of the oracle's TextToIds over a CSR batch, threaded in C."""
        n = len(offsets) - 1
        dig = np.zeros(n, np.uint64)
        counts = np.zeros(n, np.int32)
        buf = np.ascontiguousarray(buf)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        tot = self.lib.bfo_text_to_ids_digests(h, buf.ctypes.data, offsets.ctypes.data, n, dig.ctypes.data, counts.ctypes.data,
                                               max_ids, unk, threads or (os.cpu_count() or 1))
        return int(tot), dig, counts

    def csr_digests(self, ids, id_offsets, threads=None):
        """The same per-document digests of ids already in CSR form (int32 or uint16)."""
        ids = np.ascontiguousarray(ids)
        id_offsets = np.ascontiguousarray(id_offsets, dtype=np.int64)
        n = len(id_offsets) - 1
        dig = np.zeros(n, np.uint64)
        self.lib.bfo_csr_digests(ids.ctypes.data, ids.dtype.itemsize, id_offsets.ctypes.data, n, dig.ctypes.data,
                                 threads or (os.cpu_count() or 1))
        return dig

    def fold(self, dig, id_offsets=None, counts=None):
        n = len(dig)
        return int(self.lib.bfo_fold_digests(dig.ctypes.data, id_offsets.ctypes.data if id_offsets is not None else None,
                                             counts.ctypes.data if counts is not None else None, n))

    def text_to_ids(self, h, data: bytes, max_ids=512, unk=0):
        ids = np.full(max_ids, -7, np.int32)
        n = self.lib.bfo_text_to_ids(h, data, len(data), ids.ctypes.data, max_ids, unk)
        return n, ids

    def text_to_ids_with_offsets(self, h, data: bytes, max_ids=512, unk=0):
        ids = np.full(max_ids, -7, np.int32)
        st = np.full(max_ids, -7, np.int32)
        en = np.full(max_ids, -7, np.int32)
        n = self.lib.bfo_text_to_ids_with_offsets(h, data, len(data), ids.ctypes.data, st.ctypes.data, en.ctypes.data, max_ids, unk)
        return n, ids, st, en

    def split(self, kind, h, data: bytes, max_out=None):
        """(ret, text, starts, ends) of the words / sentences call with offsets."""
        fn = self.lib.bfo_text_to_words_with_offsets if kind == "words" else self.lib.bfo_text_to_sentences_with_offsets
        max_out = max_out if max_out is not None else 2 * len(data) + 16
        out = ctypes.create_string_buffer(max(max_out, 1))
        st = np.full(max(max_out, 1), -7, np.int32)
        en = np.full(max(max_out, 1), -7, np.int32)
        n = fn(h, data, len(data), out, st.ctypes.data, en.ctypes.data, max_out)
        return n, (out.raw[: max(n, 0)] if n <= max_out else b""), st, en

    def text_to_words(self, h, data: bytes, max_out=None):
        max_out = max_out if max_out is not None else 2 * len(data) + 16
        out = ctypes.create_string_buffer(max(max_out, 1))
        n = self.lib.bfo_text_to_words(h, data, len(data), out, max_out)
        return n, out.raw[: max(n, 0)] if n <= max_out else b""

    def batch(self, h, buf: np.ndarray, offsets: np.ndarray, max_ids, unk, threads=1):
        nd = len(offsets) - 1
        ids = np.zeros((nd, max_ids), np.int32)
        counts = np.zeros(nd, np.int32)
        tot = self.lib.bfo_text_to_ids_batch(h, buf.ctypes.data, offsets.ctypes.data, nd, ids.ctypes.data,
                                             counts.ctypes.data, max_ids, unk, threads)
        return tot, ids, counts


class Ref:
    """The reference library itself (oracle/_ref), built by oracle/Makefile."""

    PATH = os.path.join(ROOT, "oracle", "_ref", "libblingfiretokdll.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        self.lib = L = ctypes.CDLL(self.PATH)
        L.LoadModel.restype = ctypes.c_void_p
        L.LoadModel.argtypes = [ctypes.c_char_p]
        L.FreeModel.argtypes = [ctypes.c_void_p]
        L.TextToIds.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.TextToIdsWithOffsets.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.TextToWords.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.TextToWordsWithModel.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        for fn in ("TextToWordsWithOffsetsWithModel", "TextToSentencesWithOffsetsWithModel"):
            getattr(L, fn).argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_void_p]

        L.SetNoDummyPrefix.argtypes = [ctypes.c_void_p, ctypes.c_bool]
        L.IdsToText.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_bool]

    def split(self, kind, data: bytes, model=None, max_out=None):
        return split_call(getattr(self.lib, "TextToWordsWithOffsetsWithModel" if kind == "words" else
                                  "TextToSentencesWithOffsetsWithModel"), data, model, max_out)

    def load(self, path):
        h = self.lib.LoadModel(path.encode())
        assert h
        return h

    def set_no_dummy_prefix(self, h, flag):
        return self.lib.SetNoDummyPrefix(h, bool(flag))

    def ids_to_text(self, h, ids, max_out, skip_special):
        return ids_to_text_call(self.lib.IdsToText, h, ids, max_out, skip_special)

    def free(self, h):
        self.lib.FreeModel(h)

    def text_to_ids(self, h, data: bytes, max_ids=512, unk=0):
        ids = np.full(max_ids, -7, np.int32)
        n = self.lib.TextToIds(h, data, len(data), ids.ctypes.data, max_ids, unk)
        return n, ids

    def text_to_ids_with_offsets(self, h, data: bytes, max_ids=512, unk=0):
        ids = np.full(max_ids, -7, np.int32)
        st = np.full(max_ids, -7, np.int32)
        en = np.full(max_ids, -7, np.int32)
        n = self.lib.TextToIdsWithOffsets(h, data, len(data), ids.ctypes.data, st.ctypes.data, en.ctypes.data, max_ids, unk)
        return n, ids, st, en

    def text_to_words(self, data: bytes, model=None, max_out=None):
        max_out = max_out if max_out is not None else 2 * len(data) + 16
        out = ctypes.create_string_buffer(max(max_out, 1))
        if model is None:
            n = self.lib.TextToWords(data, len(data), out, max_out)
        else:
            n = self.lib.TextToWordsWithModel(data, len(data), out, max_out, ctypes.c_void_p(model))
        return n, out.raw[: max(n, 0)] if n <= max_out else b""


def ids_to_text_call(fn, h, ids, max_out, skip_special):
    """Calls an IdsToText-shaped function: (ret, bytes written incl. the NUL when it fitted, else b"")."""
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    out = ctypes.create_string_buffer(max(max_out, 1))
    n = fn(ctypes.c_void_p(h), ids.ctypes.data if len(ids) else None, len(ids), out, max_out, bool(skip_special))
    return n, (out.raw[:n] if 0 < n <= max_out else b"")


def split_call(fn, data: bytes, model, max_out=None):
    """Calls a TextTo{Words,Sentences}WithOffsetsWithModel-shaped function: (ret, text, starts, ends) with the
    offset arrays sized max_out like the reference wants them (it zero-fills max_out entries)."""
    max_out = max_out if max_out is not None else 2 * len(data) + 16
    out = ctypes.create_string_buffer(max(max_out, 1))
    st = np.full(max(max_out, 1), -7, np.int32)
    en = np.full(max(max_out, 1), -7, np.int32)
    n = fn(data, len(data), out, st.ctypes.data, en.ctypes.data, max_out, ctypes.c_void_p(model) if model else None)
    text = out.raw[: max(n, 0)] if n <= max_out else b""
    return n, text, st, en


def read_lines(name, limit=None, drop_empty=True):
    with open(os.path.join(CORPUS, name), "rb") as f:
        lines = f.read().split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    if drop_empty:
        lines = [l for l in lines if l]
    return lines[:limit] if limit else lines
