#!/usr/bin/env python3
"""Stage the reference's *data* artefacts (compiled .bin models and test corpora)
into data/ so that tests, smoke() and bench.py can run on the GPU box, where
/root/reference does not exist.

These are inputs the user hands to LoadModel (the north_star: "C++ host code
loads the existing .bin LDB"), not reference sources; nothing here is compiled.
data/ is git-ignored (kept out of history) but travels with the gpurun snapshot,
exactly like oracle/_ref/ and the built .so files.  Run from build() whenever
/root/reference is present; a no-op otherwise.
"""
import os
import shutil
import sys
import zipfile

REF = os.environ.get("BLINGFIRE_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "data")

MODELS = [
    "bert_base_tok.bin", "bert_base_cased_tok.bin", "bert_multi_cased.bin", "bert_chinese.bin",
    "wbd.bin", "wbd_chuni.bin", "sbd.bin", "gpt2.bin", "roberta.bin", "xlm_roberta_base.bin",
    "xlnet.bin", "xlnet_nonorm.bin", "bpe_example.bin", "laser100k.bin", "uri100k.bin",
    # id -> text arrays for IdsToText (separate LDB files with only an [i2w] section)
    "bert_base_tok.i2w", "bert_base_cased_tok.i2w", "bert_chinese.i2w", "gpt2.i2w", "roberta.i2w",
    "xlm_roberta_base.i2w", "laser100k.i2w", "uri100k.i2w",
]
CORPORA = [("ldbsrc/bert_base_tok/test.txt", "test.txt")]
ZIPS = [("ldbsrc/bert_multi_cased/test.multi.txt.zip", "test.multi.txt")]


def stage(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print(f"[stage_data] {REF} not present; keeping whatever is in {DATA}")
        return False
    os.makedirs(os.path.join(DATA, "ldb"), exist_ok=True)
    os.makedirs(os.path.join(DATA, "corpus"), exist_ok=True)
    for m in MODELS:
        src = os.path.join(REF, "ldbsrc", "ldb", m)
        dst = os.path.join(DATA, "ldb", m)
        if os.path.exists(src) and (not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src)):
            shutil.copyfile(src, dst)
    for rel, name in CORPORA:
        src = os.path.join(REF, rel)
        dst = os.path.join(DATA, "corpus", name)
        if os.path.exists(src) and (not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src)):
            shutil.copyfile(src, dst)
    for rel, name in ZIPS:
        src = os.path.join(REF, rel)
        dst = os.path.join(DATA, "corpus", name)
        if os.path.exists(src) and not os.path.exists(dst):
            with zipfile.ZipFile(src) as z:
                member = [n for n in z.namelist() if n.endswith(name)][0]
                with z.open(member) as fi, open(dst, "wb") as fo:
                    shutil.copyfileobj(fi, fo)
    if verbose:
        print(f"[stage_data] staged models and corpora into {DATA}")
    return True


if __name__ == "__main__":
    stage()
    sys.exit(0)
