"""Full-size parity ("parity is checked on EVERY doc of cfgs 2-4 against the oracle, not sampled"):
1 000 000 documents per BASELINE config through the C ABI, every document's ids compared -- by its FNV-1a-64
digest (SURVEY 8c recipe) and its count -- with the reference's on the same batch.  The reference side is
oracle/_ref (the reference compiled from its own sources) when it travelled with the snapshot, else the oracle
port; both are threaded in C and never materialise an [ndocs][max_ids] matrix."""
import numpy as np
import pytest

from _common import Oracle, have_data, model_path

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_data(), reason="data/ not staged")]

N_DOCS = 1_000_000


@pytest.fixture(scope="module")
def bf():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
    import blingfire_b200
    return blingfire_b200


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
def test_every_document_of_the_baseline_config(bf, name):
    import bench
    cfg = bench.CONFIGS[name]
    text, offs = bench.make_batch(cfg, N_DOCS)
    want_dig, want_counts, src, secs = bench.oracle_digests(cfg, text, offs)
    h = bf.load_model(model_path(cfg["model"]))
    ids, io = bf.text_to_ids_batch_csr(h, (text, offs), cfg["max_ids"], cfg["unk"])     # pageable numpy in and out
    o = Oracle()
    dig = o.csr_digests(ids, io)
    counts = np.diff(io)
    bad = np.nonzero((dig != want_dig) | (counts != want_counts))[0]
    assert len(bad) == 0, (f"{name}: {len(bad)} of {N_DOCS} documents differ from {src}; first: doc {bad[0]} "
                           f"gpu count {counts[bad[0]]} vs {want_counts[bad[0]]}: {bytes(text[offs[bad[0]]:offs[bad[0] + 1]])[:120]!r}")
    assert int(io[-1]) == int(want_counts.sum())
    assert o.fold(dig, id_offsets=io) == o.fold(want_dig, counts=want_counts)
    if name == "cfg2":
        # the 16-bit form carries the same ids
        ids16, io16 = bf.text_to_ids_batch_csr(h, (text, offs), cfg["max_ids"], cfg["unk"], dtype=np.uint16)
        assert (io16 == io).all() and (o.csr_digests(ids16, io16) == want_dig).all()
    bf.free_model(h)
