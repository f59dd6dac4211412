"""The N>1 host logic on the CPU: world_size-2 gloo process group.  Each rank takes its byte-balanced
document shard (and its rotated replica), tokenizes it with the ORACLE (test infrastructure -- there is
no GPU here), and the all-reduced {docs, bytes, tokens} must equal the single-process totals."""
import os
import socket

import numpy as np
import pytest

from _common import have_data, model_path

pytestmark = pytest.mark.skipif(not have_data(), reason="data/ not staged")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests"), os.path.join(root, "tools")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from _common import Oracle
    import corpus
    from blingfire_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    text, offs = corpus.gen_docs("EN", 4000, seed=2, fixed_len=0)   # ragged lengths
    o = Oracle()
    h = o.load(model_path("bert_base_tok.bin"))
    # (1) byte-balanced contiguous shards
    lo, hi = sharding.shard_bounds(offs, rank, world)
    sub_offs = np.ascontiguousarray(offs[lo:hi + 1])
    tot, _, counts = o.batch(h, text, sub_offs, 512, 100, threads=2)
    stats = sharding.all_reduce_stats(hi - lo, int(sub_offs[-1] - sub_offs[0]), tot)
    # (2) rotated replicas: every rank the whole set, rotated by rank * 625 documents
    rtext, roffs = sharding.rotate_replica(text, offs, rank * 625)
    rtot, _, rcounts = o.batch(h, rtext, roffs, 512, 100, threads=2)
    rstats = sharding.all_reduce_stats(len(roffs) - 1, int(roffs[-1]), rtot)
    q.put((rank, lo, hi, stats, rstats, int(rtot), rcounts[:5].tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_stats():
    import torch.multiprocessing as mp
    import corpus
    from _common import Oracle
    from blingfire_b200 import sharding
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    text, offs = corpus.gen_docs("EN", 4000, seed=2, fixed_len=0)
    o = Oracle()
    h = o.load(model_path("bert_base_tok.bin"))
    tot, _, counts = o.batch(h, text, offs, 512, 100, threads=4)
    # shards are contiguous, cover everything, and are balanced by bytes
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 4000
    b0 = offs[res[0][2]] - offs[0]
    assert abs(int(b0) - int(offs[-1]) // 2) <= 4096
    for r in res:
        assert r[3] == (4000, int(offs[-1]), int(tot))                 # all-reduced shard stats
        assert r[4] == (8000, 2 * int(offs[-1]), 2 * int(tot))         # two full replicas
        assert r[5] == int(tot)
    # rotation really rotates: rank 1's first documents are documents 625.. of the set
    assert res[1][6] == counts[625:630].tolist()
    # unit checks of the helpers
    for w in (1, 3, 8):
        cuts = [sharding.shard_bounds(offs, r, w) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == 4000 and all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
