"""Multi-GPU host logic for the batch TextToIds path.

Documents are independent and the model tables are read-only, so the path shards by document
with no data-path collective: one process per GPU, every rank tokenizes its own
documents, and the only exchange is an all-reduce of three int64 counters.  The helpers here are
backend-agnostic (NCCL on the GPU box, gloo in the CPU tests)."""
import numpy as np


def shard_bounds(offsets, rank, world):
    """Contiguous document range [lo, hi) for `rank`, balanced by BYTES (prefix sum of document
    lengths), so ranks finish together on ragged corpora."""
    n = len(offsets) - 1
    total = int(offsets[-1] - offsets[0])
    targets = offsets[0] + (np.arange(world + 1, dtype=np.int64) * total) // world
    cuts = np.searchsorted(offsets, targets, side="left").astype(np.int64)
    cuts[0], cuts[-1] = 0, n
    cuts = np.maximum.accumulate(np.clip(cuts, 0, n))
    return int(cuts[rank]), int(cuts[rank + 1])


def rotate_replica(text, offsets, rot):
    """The document set rotated left by `rot` documents (replica rule of SURVEY 8d cfg 5)."""
    n = len(offsets) - 1
    rot %= max(n, 1)
    if rot == 0:
        return text, offsets
    lens = np.diff(offsets)
    order = np.roll(np.arange(n), -rot)
    new_offs = np.zeros(n + 1, np.int64)
    np.cumsum(lens[order], out=new_offs[1:])
    cut = int(offsets[rot] - offsets[0])
    base = int(offsets[0])
    body = text[base: base + int(offsets[-1] - offsets[0])]
    return np.concatenate([body[cut:], body[:cut]]), new_offs


def all_reduce_stats(docs, nbytes, tokens, device=None):
    """Sum of per-rank {docs, bytes, tokens} over the default process group (no-op when not
    initialised).  Returns python ints."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(docs), int(nbytes), int(tokens)], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return tuple(int(x) for x in t.tolist())
