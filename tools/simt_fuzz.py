#!/usr/bin/env python3
"""Long-running differential fuzz of the kernel SOURCES over the SIMT shim (tests/simt) against the oracle, on the CPU:
random joins of corpus lines, truncations, injected bytes, whitespace-free runs, runs of one character, random
output caps, all [pos-dict] models and two WordPiece models.  Not part of the pytest suite (a round of 400
documents takes minutes).  Usage: python tools/simt_fuzz.py [seed] [rounds] [offsets]
("offsets": the [pos-dict] models through sp_unigram_offsets_kernel / sp_bpe_offsets_kernel, ids + starts + ends)"""
import sys, random, time
import os as _os
sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'tests'))
import numpy as np, ctypes, os
import test_simt as T
from _common import Oracle, model_path, read_lines, ROOT
class Sim: pass
def load_sp():
    L = ctypes.CDLL(os.path.join(ROOT, "tests", "simt", "libsp_simt.so"))
    L.spsim_load.restype = ctypes.c_void_p; L.spsim_load.argtypes = [ctypes.c_char_p]
    L.spsim_error.restype = ctypes.c_char_p; L.spsim_error.argtypes = [ctypes.c_void_p]
    L.spsim_batch.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return L
def load_wp():
    L = ctypes.CDLL(os.path.join(ROOT, "tests", "simt", "libwp_simt.so"))
    L.wpsim_load.restype = ctypes.c_void_p; L.wpsim_load.argtypes = [ctypes.c_char_p]
    L.wpsim_error.restype = ctypes.c_char_p; L.wpsim_error.argtypes = [ctypes.c_void_p]
    L.wpsim_free.argtypes = [ctypes.c_void_p]
    L.wpsim_batch.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return L
rng=random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 1)
lines=read_lines("test.multi.txt")+read_lines("test.txt")
def docs(n):
    out=[]
    for _ in range(n):
        k=rng.choice([1,1,2,3,6,12,30])
        d=b" ".join(rng.choice(lines) for _ in range(k))
        r=rng.random()
        if r<0.15: d=d[:rng.randint(0,len(d))]
        elif r<0.25:
            p=rng.randint(0,len(d)); d=d[:p]+bytes([rng.randint(0,255)])+d[p:]
        elif r<0.30: d=bytes(rng.randint(0,255) for _ in range(rng.randint(1,80)))
        elif r<0.33: d=d.replace(b" ",b"")          # no whitespace: long runs
        elif r<0.36: d=bytes([rng.choice(b"=-.a ")]) * rng.randint(1,1500)
        out.append(d)
    return out
sp=load_sp(); wp=load_wp()
with_offsets = "offsets" in sys.argv[3:]
t0=time.time(); nd=0
for rnd in range(int(sys.argv[2]) if len(sys.argv)>2 else 3):
    D=docs(400)
    for name,unk in T.SP_MODELS:
        for max_ids in (2048, rng.choice([1,3,17,64])):
            T.check(sp, name, D, max_ids, unk, offsets=with_offsets)
    for name in ([] if with_offsets else ["bert_base_tok.bin","bert_chinese.bin"]):
        T.check_wp(wp, name, D, 1024); T.check_wp(wp, name, D[:150], rng.choice([1,2,9]))
    nd+=len(D); print("round",rnd,"docs",nd,"%.0fs"%(time.time()-t0), flush=True)
print("OK")
