"""blingfire_b200 -- Python binding of the B200-native drop-in for BlingFire's TextToIds path.

Mirrors the reference's own ctypes wrapper for this path (dist-pypi/blingfire/__init__.py:
229-253: load_model / free_model / text_to_ids, :85-122: text_to_words[_with_model]) with the
same names, argument meaning and return conventions, over the C ABI declared in
include/blingfiretokdll_b200.h.  Additive: the batch calls (text_to_ids_batch*), which the
reference does not have.

The CUDA library is mandatory: importing works anywhere, but the first call that needs it
raises if lib/libblingfiretokdll.so is missing -- there is no CPU path to fall back to.
"""
import ctypes
import os
from ctypes import c_bool, c_char_p, c_int, c_int32, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libblingfiretokdll.so")

_lib = None


def lib():
    """The loaded C-ABI library (ctypes.CDLL).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "blingfire_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.GetBlingFireTokVersion.restype = c_int
        L.LoadModel.restype = c_void_p
        L.LoadModel.argtypes = [c_char_p]
        L.SetModel.restype = c_void_p
        L.SetModel.argtypes = [c_char_p, c_int]
        L.FreeModel.restype = c_int
        L.FreeModel.argtypes = [c_void_p]
        for name in ("TextToIds", "TextToIds_wp", "TextToIds_sp"):
            f = getattr(L, name)
            f.restype = c_int
            f.argtypes = [c_void_p, c_char_p, c_int, c_void_p, c_int, c_int]
        L.TextToWords.restype = c_int
        L.TextToWords.argtypes = [c_char_p, c_int, c_void_p, c_int]
        L.TextToWordsWithModel.restype = c_int
        L.TextToWordsWithModel.argtypes = [c_char_p, c_int, c_void_p, c_int, c_void_p]
        L.TextToIdsBatch.restype = c_int64
        L.TextToIdsBatch.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int]
        L.TextToIdsBatchCsr.restype = c_int64
        L.TextToIdsBatchCsr.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int]
        L.TextToIdsWithOffsetsBatch.restype = c_int64
        L.TextToIdsWithOffsetsBatch.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]
        L.TextToIdsWithOffsetsBatchCsr.restype = c_int64
        L.TextToIdsWithOffsetsBatchCsr.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                                   c_int, c_int]
        L.TextToWordsBatch.restype = c_int64
        L.TextToWordsBatch.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]
        L.TextToSentencesBatch.restype = c_int64
        L.TextToSentencesBatch.argtypes = L.TextToWordsBatch.argtypes
        L.TextToIdsBatchCsrU16.restype = c_int64
        L.TextToIdsBatchCsrU16.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int]
        L.TextToIdsBatchDevice.restype = c_int
        L.TextToIdsBatchDevice.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p]
        L.TextToIdsBatchDeviceSized.restype = c_int
        L.TextToIdsBatchDeviceSized.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p]
        L.BlingFireB200DeviceStatus.restype = c_int
        L.BlingFireB200DeviceStatus.argtypes = [c_void_p, c_void_p]
        L.BlingFireB200CompactDevice.restype = c_int64
        L.BlingFireB200CompactDevice.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]
        L.SetNoDummyPrefix.restype = c_int
        L.SetNoDummyPrefix.argtypes = [c_void_p, c_bool]
        L.IdsToText.restype = c_int
        L.IdsToText.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_bool]
        L.BlingFireB200LastError.restype = c_char_p
        L.BlingFireB200KernelLaunches.restype = c_int64
        L.BlingFireB200ModelEngine.restype = c_int
        L.BlingFireB200ModelEngine.argtypes = [c_void_p]
        _lib = L
    return _lib


def last_error():
    return lib().BlingFireB200LastError().decode("utf-8", "replace")


def kernel_launches():
    return int(lib().BlingFireB200KernelLaunches())


def get_blingfiretok_version():
    return lib().GetBlingFireTokVersion()


def load_model(file_name):
    """dist-pypi/blingfire/__init__.py:229-234.  Returns a handle; raises on failure
    (the reference aborts the process on a missing file; a None handle would only defer the error)."""
    h = lib().LoadModel(file_name.encode("utf-8"))
    if not h:
        raise RuntimeError(f"LoadModel({file_name!r}) failed: {last_error()}")
    return h


def free_model(h):
    lib().FreeModel(c_void_p(h))


def text_to_ids(h, s, max_len, unk=0, no_padding=False):
    """dist-pypi/blingfire/__init__.py:243-253: zero-padded uint32 array of max_len ids
    (or just the produced ids with no_padding=True)."""
    s_bytes = s.encode("utf-8") if isinstance(s, str) else bytes(s)
    o = np.zeros(max_len, dtype=np.int32)
    t_count = lib().TextToIds(c_void_p(h), s_bytes, len(s_bytes), o.ctypes.data, max_len, unk)
    if t_count == 0 and s_bytes and last_error():
        # 0 is also a legitimate result (empty / invalid input); a GPU-side failure must not hide behind it
        raise RuntimeError(f"TextToIds failed: {last_error()}")
    out_count = min(max_len, t_count) if no_padding else max_len
    return o.view(np.uint32)[:out_count]


def ids_to_text(h, ids, skip_special_tokens=True, output_buffer_size=None):
    """dist-pypi/blingfire/__init__.py:256-270: the text of a sequence of ids ('' on any failure)."""
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    if output_buffer_size is None:
        output_buffer_size = len(ids) * 32
    o = ctypes.create_string_buffer(max(output_buffer_size, 1))
    o_len = lib().IdsToText(c_void_p(h), ids.ctypes.data if len(ids) else None, len(ids), o, output_buffer_size,
                            bool(skip_special_tokens))
    if o_len == -1 or o_len > output_buffer_size:
        return ""
    return o.value.decode("utf-8")


def change_settings_dummy_prefix(h, add_prefix):
    """dist-pypi/blingfire/__init__.py:287-288."""
    lib().SetNoDummyPrefix(c_void_p(h), not add_prefix)


def utf8text_to_ids_with_offsets(h, s_bytes, max_len, unk=0, no_padding=False):
    """dist-pypi/blingfire/__init__.py:272-285: (ids uint32, start offsets int32, end offsets int32)."""
    L = lib()
    L.TextToIdsWithOffsets.restype = c_int
    L.TextToIdsWithOffsets.argtypes = [c_void_p, c_char_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int]
    ids = np.zeros(max_len, dtype=np.int32)
    starts = np.zeros(max_len, dtype=np.int32)
    ends = np.zeros(max_len, dtype=np.int32)
    t_count = L.TextToIdsWithOffsets(c_void_p(h), s_bytes, len(s_bytes), ids.ctypes.data, starts.ctypes.data, ends.ctypes.data,
                                     max_len, unk)
    if t_count == 0 and s_bytes and last_error():
        raise RuntimeError(f"TextToIdsWithOffsets failed: {last_error()}")
    out_count = min(max_len, t_count) if no_padding else max_len
    return ids.view(np.uint32)[:out_count], starts[:out_count], ends[:out_count]


def text_to_words_with_model(h, s):
    """dist-pypi/blingfire/__init__.py:105-122."""
    s_bytes = s.encode("utf-8")
    o_cap = len(s_bytes) * 3 + 1
    o = ctypes.create_string_buffer(o_cap)
    n = lib().TextToWordsWithModel(s_bytes, len(s_bytes), o, o_cap, c_void_p(h) if h else None)
    if n < 0 or n > o_cap:
        return ""
    return o.value.decode("utf-8")


def text_to_words(s):
    """dist-pypi/blingfire/__init__.py:85-102 (default word-breaking model)."""
    return text_to_words_with_model(None, s)


def text_to_ids_with_offsets_batch(h, docs, max_len, unk=0):
    """ADDITIVE: text_to_ids_with_offsets for many documents in one call.  Returns (ids, starts, ends, counts): three
    zero-filled [n, max_len] int32 arrays with the first counts[i] entries of row i set, and counts [n]."""
    buf, offs = make_csr([d.encode("utf-8") if isinstance(d, str) else d for d in docs]) if not isinstance(docs, tuple) else docs
    n = len(offs) - 1
    ids = np.zeros((n, max_len), np.int32)
    starts = np.zeros((n, max_len), np.int32)
    ends = np.zeros((n, max_len), np.int32)
    counts = np.zeros(n, np.int32)
    r = lib().TextToIdsWithOffsetsBatch(c_void_p(h), buf.ctypes.data if len(buf) else None, offs.ctypes.data, n, ids.ctypes.data,
                                        starts.ctypes.data, ends.ctypes.data, counts.ctypes.data, max_len, unk)
    if r < 0:
        raise RuntimeError(f"TextToIdsWithOffsetsBatch failed: {last_error()}")
    return ids, starts, ends, counts


def text_to_ids_with_offsets_batch_csr(h, docs, max_len, unk=0, capacity=None):
    """ADDITIVE: the compact form (TextToIdsWithOffsetsBatchCsr).  Returns (ids, starts, ends, id_offsets): the ids and byte
    offsets of document i are entries id_offsets[i]:id_offsets[i+1] of the three int32 arrays."""
    buf, offs = make_csr([d.encode("utf-8") if isinstance(d, str) else d for d in docs]) if not isinstance(docs, tuple) else docs
    n = len(offs) - 1
    buf = np.ascontiguousarray(buf)
    offs = np.ascontiguousarray(offs, dtype=np.int64)
    if capacity is None:
        capacity = int(np.minimum(np.diff(offs) + 1, max_len).clip(min=0).sum())
    id_offsets = np.zeros(n + 1, np.int64)
    for _ in range(2):
        out = np.empty((3, max(capacity, 1)), np.int32)
        r = lib().TextToIdsWithOffsetsBatchCsr(c_void_p(h), buf.ctypes.data if len(buf) else None, offs.ctypes.data, n, out[0].ctypes.data,
                                               out[1].ctypes.data, out[2].ctypes.data, capacity, id_offsets.ctypes.data, max_len, unk)
        if r >= 0:
            return out[0, :r], out[1, :r], out[2, :r], id_offsets
        if r == -1 and last_error():
            break
        capacity = -r            # a [pos-dict] model can emit more ids than bytes
    raise RuntimeError(f"TextToIdsWithOffsetsBatchCsr failed ({r}): {last_error()}")


def text_to_sentences_batch(docs, h=None, raw=False):
    """ADDITIVE: text_to_sentences[_with_model] for many documents in one call (TextToSentencesBatch)."""
    return text_to_words_batch(docs, h, raw, _fn="TextToSentencesBatch")


def text_to_words_batch(docs, h=None, raw=False, _fn="TextToWordsBatch"):
    """ADDITIVE: text_to_words[_with_model] for many documents in one call (TextToWordsBatch: lexer and string building on
    the GPU).  `docs`: str / bytes items, or a (uint8 buffer, int64 offsets) pair.  Returns the list of strings ("" where
    the per-document call returns "": bad UTF-8, empty input); with raw=True, (buffer, offsets, results) as numpy arrays."""
    buf, offs = make_csr([d.encode("utf-8") if isinstance(d, str) else d for d in docs]) if not isinstance(docs, tuple) else docs
    n = len(offs) - 1
    out_offs = np.zeros(n + 1, np.int64)
    results = np.zeros(n, np.int32)
    cap = int(2 * len(buf) + n + 16)
    out = np.empty(cap, np.uint8)
    fn = getattr(lib(), _fn)
    r = fn(c_void_p(h) if h else None, buf.ctypes.data if len(buf) else None, offs.ctypes.data, n, out.ctypes.data, cap,
           out_offs.ctypes.data, results.ctypes.data)
    if r < 0 and -r > cap:
        cap = int(-r)
        out = np.empty(cap, np.uint8)
        r = fn(c_void_p(h) if h else None, buf.ctypes.data if len(buf) else None, offs.ctypes.data, n, out.ctypes.data, cap,
               out_offs.ctypes.data, results.ctypes.data)
    if r < 0:
        raise RuntimeError(f"{_fn} failed: {last_error()}")
    if raw:
        return out[:r], out_offs, results
    data = out[:r].tobytes()
    return [data[out_offs[i]:out_offs[i + 1] - 1].decode("utf-8") if results[i] > 0 else "" for i in range(n)]


def text_to_sentences_with_model(h, s):
    """dist-pypi/blingfire/__init__.py:45-62."""
    s_bytes = s.encode("utf-8")
    o_cap = len(s_bytes) * 2 + 1
    o = ctypes.create_string_buffer(o_cap)
    L = lib()
    L.TextToSentencesWithModel.restype = c_int
    L.TextToSentencesWithModel.argtypes = [c_char_p, c_int, c_void_p, c_int, c_void_p]
    n = L.TextToSentencesWithModel(s_bytes, len(s_bytes), o, o_cap, c_void_p(h) if h else None)
    if n < 0 or n > o_cap:
        return ""
    return o.value.decode("utf-8")


def text_to_sentences(s):
    """dist-pypi/blingfire/__init__.py:25-42 (default sentence-breaking model)."""
    return text_to_sentences_with_model(None, s)


def _utf8_split_with_offsets(fn_name, s_bytes, h=None):
    """Raw form of the WithOffsets calls: (text bytes without the NUL, starts, ends) with BYTE offsets of
    the first byte of each token's first character and the last byte of its last character."""
    L = lib()
    f = getattr(L, fn_name)
    f.restype = c_int
    f.argtypes = [c_char_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]
    o_cap = len(s_bytes) * 2 + 1
    o = ctypes.create_string_buffer(o_cap)
    starts = np.zeros(o_cap, dtype=np.int32)
    ends = np.zeros(o_cap, dtype=np.int32)
    n = f(s_bytes, len(s_bytes), o, starts.ctypes.data, ends.ctypes.data, o_cap, c_void_p(h) if h else None)
    if n <= 0 or n > o_cap:
        return b"", starts[:0], ends[:0]
    text = o.raw[: n - 1]
    sep = b" " if "Words" in fn_name else b"\n"
    k = text.count(sep) + 1          # like the reference wrapper (:188): an empty output still counts one token
    return text, starts[:k], ends[:k]


def _string_offsets(s, s_bytes, starts, ends):
    """The reference wrapper's mapping (dist-pypi/blingfire/__init__.py:192-219) from the byte offsets the C
    ABI returns (first byte of the first character, last byte of the last one) to (begin, end) pairs in
    code points of the Python string, end exclusive."""
    utf8_offsets = [o for pair in zip(starts, ends) for o in pair]
    string_offsets = []
    string_offset = 0
    is_end_offset = False
    for utf8_offset, b in enumerate(s_bytes):
        if b & 0xC0 != 0x80:
            while len(string_offsets) < len(utf8_offsets) and utf8_offsets[len(string_offsets)] + is_end_offset == utf8_offset:
                string_offsets.append(string_offset)
                is_end_offset = not is_end_offset
            string_offset += 1
    if len(string_offsets) < len(utf8_offsets):
        string_offsets.append(len(s))
    assert len(string_offsets) == len(utf8_offsets), "%s != %s" % (len(string_offsets), len(utf8_offsets))
    return list(zip(string_offsets[::2], string_offsets[1::2]))


def text_to_words_with_offsets(s):
    """dist-pypi/blingfire/__init__.py:222-223: (' '-joined words, [(begin, end)] in code points)."""
    s_bytes = s.encode("utf-8")
    text, st, en = utf8text_to_words_with_offsets(s_bytes)
    if len(st) == 0:
        return "", []
    return text.decode("utf-8"), _string_offsets(s, s_bytes, [int(x) for x in st], [int(x) for x in en])


def text_to_sentences_and_offsets(s):
    """dist-pypi/blingfire/__init__.py:225-226: ('\\n'-joined sentences, [(begin, end)] in code points)."""
    s_bytes = s.encode("utf-8")
    text, st, en = utf8text_to_sentences_with_offsets(s_bytes)
    if len(st) == 0:
        return "", []
    return text.decode("utf-8"), _string_offsets(s, s_bytes, [int(x) for x in st], [int(x) for x in en])


def utf8text_to_words_with_offsets(s_bytes, h=None):
    """TextToWordsWithOffsets[WithModel] (blingfiretokdll.cpp:415-566) on bytes, byte offsets."""
    return _utf8_split_with_offsets("TextToWordsWithOffsetsWithModel", s_bytes, h)


def utf8text_to_sentences_with_offsets(s_bytes, h=None):
    """TextToSentencesWithOffsets[WithModel] (blingfiretokdll.cpp:163-355) on bytes, byte offsets."""
    return _utf8_split_with_offsets("TextToSentencesWithOffsetsWithModel", s_bytes, h)


# ---- additive batch API --------------------------------------------------------------------

def make_csr(docs):
    """Concatenates an iterable of bytes/str documents into (uint8 buffer, int64 offsets)."""
    bs = [d.encode("utf-8") if isinstance(d, str) else bytes(d) for d in docs]
    offsets = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        np.cumsum([len(b) for b in bs], out=offsets[1:])
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8) if bs else np.zeros(0, np.uint8)
    return buf, offsets


def text_to_ids_batch(h, docs, max_len, unk=0):
    """Row-major batch: returns (ids[n_docs, max_len] int32 zero-padded, counts[n_docs] int32).
    Row i / counts[i] equal the reference's TextToIds on document i."""
    buf, offsets = docs if isinstance(docs, tuple) else make_csr(docs)
    n = len(offsets) - 1
    ids = np.zeros((n, max_len), dtype=np.int32)
    counts = np.zeros(n, dtype=np.int32)
    buf = np.ascontiguousarray(buf)
    r = lib().TextToIdsBatch(c_void_p(h), buf.ctypes.data, offsets.ctypes.data, n, ids.ctypes.data,
                             counts.ctypes.data, max_len, unk)
    if r < 0:
        raise RuntimeError(f"TextToIdsBatch failed: {last_error()}")
    return ids, counts


def text_to_ids_batch_csr(h, docs, max_len, unk=0, capacity=None, dtype=np.int32):
    """Compact batch: returns (ids[total], id_offsets int64[n_docs+1]).  dtype=np.uint16 asks for 16-bit ids
    (models whose ids all fit; halves the bytes that come back over PCIe).  The caller's arrays may be ordinary
    (pageable) numpy memory: the library stages them through its own pinned buffers."""
    buf, offsets = docs if isinstance(docs, tuple) else make_csr(docs)
    n = len(offsets) - 1
    if capacity is None:
        # a first guess that fits natural text; the call reports the exact need if it does not
        capacity = int(np.minimum(np.diff(offsets) + 1, max_len).clip(min=0).sum())
    fn = lib().TextToIdsBatchCsrU16 if np.dtype(dtype) == np.uint16 else lib().TextToIdsBatchCsr
    id_offsets = np.zeros(n + 1, dtype=np.int64)
    buf = np.ascontiguousarray(buf)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    for _ in range(2):
        ids = np.empty(max(capacity, 1), dtype=dtype)
        r = fn(c_void_p(h), buf.ctypes.data, offsets.ctypes.data, n, ids.ctypes.data, capacity, id_offsets.ctypes.data, max_len, unk)
        if r >= 0:
            return ids[:r], id_offsets
        if r == -1 and last_error():
            break
        capacity = -r            # the batch needs this many ids: a [pos-dict] model can emit more ids than bytes
    raise RuntimeError(f"TextToIdsBatchCsr failed ({r}): {last_error()}")


def text_to_ids_batch_device(h, d_text, d_offsets, n_docs, total_bytes, d_ids, d_counts, max_len, unk=0, stream=0,
                             max_doc_bytes=0):
    """Device-pointer entry: all arguments are raw device addresses (ints).  Asynchronous; for [pos-dict]
    models pass max_doc_bytes (>= the longest document) to keep it so."""
    r = lib().TextToIdsBatchDeviceSized(c_void_p(h), c_void_p(d_text), c_void_p(d_offsets), n_docs, total_bytes, max_doc_bytes,
                                        c_void_p(d_ids), c_void_p(d_counts), max_len, unk, c_void_p(stream))
    if r != 0:
        raise RuntimeError(f"TextToIdsBatchDevice failed: {last_error()}")


def device_status(h, stream=0):
    """Synchronises `stream`; raises if a kernel of the device-pointer calls on it reported an error."""
    r = lib().BlingFireB200DeviceStatus(c_void_p(h), c_void_p(stream))
    if r != 0:
        raise RuntimeError(f"device batch failed ({r}): {last_error()}")


def compact_device(d_ids, d_counts, n_docs, max_len, d_csr, d_row_off, stream=0):
    """Row-major device ids -> CSR on the device (raw device addresses).  Asynchronous."""
    r = lib().BlingFireB200CompactDevice(c_void_p(d_ids), c_void_p(d_counts), n_docs, max_len, c_void_p(d_csr), c_void_p(d_row_off),
                                         c_void_p(stream))
    if r != 0:
        raise RuntimeError(f"BlingFireB200CompactDevice failed: {last_error()}")
