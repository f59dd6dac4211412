"""Synthetic workloads of BASELINE.json / SURVEY 8d, built from the staged corpora in data/corpus.
Bench/test tooling (not product code).  The heavy lifting is tools/corpusgen.c."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CORPUS = os.path.join(ROOT, "data", "corpus")
_SO = os.path.join(ROOT, "tools", "libcorpusgen.so")
_gen = None


def build_corpusgen():
    src = os.path.join(ROOT, "tools", "corpusgen.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


def _lib():
    global _gen
    if _gen is None:
        build_corpusgen()
        L = ctypes.CDLL(_SO)
        L.gen_docs.restype = ctypes.c_int64
        L.gen_docs.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int64,
                               ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        _gen = L
    return _gen


def _pool_from_lines(lines):
    offs = np.zeros(len(lines) + 1, dtype=np.int64)
    np.cumsum([len(l) for l in lines], out=offs[1:])
    return np.frombuffer(b"".join(lines), dtype=np.uint8).copy(), offs


_pools = {}


def pool(name):
    """'EN': non-empty lines of test.txt.  'MULTI': lines of test.multi.txt whose non-ASCII byte
    fraction is >= 0.5."""
    if name in _pools:
        return _pools[name]
    if name == "EN":
        with open(os.path.join(CORPUS, "test.txt"), "rb") as f:
            lines = [l for l in f.read().split(b"\n") if l]
    elif name == "MULTI":
        with open(os.path.join(CORPUS, "test.multi.txt"), "rb") as f:
            raw = f.read().split(b"\n")
        lines = []
        for l in raw:
            if not l:
                continue
            a = np.frombuffer(l, dtype=np.uint8)
            if (a >= 0x80).sum() * 2 >= len(a):
                lines.append(l.rstrip(b"\r"))
        lines = [l for l in lines if l]
    else:
        raise KeyError(name)
    _pools[name] = _pool_from_lines(lines)
    return _pools[name]


def gen_docs(pool_name, n_docs, seed, fixed_len=512, emoji_every=0, out=None):
    """Returns (uint8 text buffer, int64 offsets[n_docs+1]).  fixed_len <= 0 selects the
    log-uniform 64..4096 length law of cfg 3."""
    buf, offs = pool(pool_name)
    cap = int(n_docs) * (fixed_len if fixed_len > 0 else 4096) + 4096
    if out is None:
        out = np.empty(cap, dtype=np.uint8)
    doc_off = np.zeros(n_docs + 1, dtype=np.int64)
    r = _lib().gen_docs(buf.ctypes.data, offs.ctypes.data, len(offs) - 1, seed, n_docs, fixed_len, emoji_every,
                        out.ctypes.data, len(out), doc_off.ctypes.data)
    if r < 0:
        raise RuntimeError("corpusgen: output buffer too small")
    return out[:r], doc_off


def cfg2(n_docs=1_000_000):
    """bert_base_tok.bin TextToIds, English docs ~512 B, seed 2."""
    return gen_docs("EN", n_docs, seed=2, fixed_len=512)


def cfg3(n_docs=1_000_000):
    """gpt2.bin, log-uniform 64..4096 B, seed 3."""
    return gen_docs("EN", n_docs, seed=3, fixed_len=0)


def cfg4(n_docs=1_000_000):
    """xlm_roberta_base.bin, multilingual ~512 B, seed 4, a 4-byte code point every 16th doc."""
    return gen_docs("MULTI", n_docs, seed=4, fixed_len=512, emoji_every=16)
