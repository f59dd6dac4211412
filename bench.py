#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on BASELINE.json's configs.

  metric    input GB/s of UTF-8 through TextToIds (+ tokens/s)
  workload  the headline line is configs[1] (SURVEY 8d cfg 2): bert_base_tok.bin, 1 M synthetic English documents
            of ~512 B, seed 2.  At N=1 the same line carries `other_configs` = cfg 3 (gpt2.bin, 1 M documents of
            64..4096 B) and cfg 4 (xlm_roberta_base.bin, 1 M multilingual documents of ~512 B), each with its own
            value / e2e / roofline / cpu_baseline / parity.
  step      one pass of the hot path over the whole batch (0.5 .. 1 GB of input, larger than L2, so no L2 flush is
            needed between timed steps)
  value     device-resident: the CSR batch is already in HBM, ids/counts stay in HBM; CUDA events on the
            launching stream
  e2e       the same batch through the C-ABI call a user makes (TextToIdsBatchCsr) with HOST (pinned) buffers:
            host->device and device->host copies inside the timed region.  `e2e_pageable` (cfg 2) is the same
            call from ordinary numpy memory, i.e. what blingfire_b200.text_to_ids_batch_csr does.
  parity    outside the timed region: EVERY document's ids (FNV-1a-64 per document, SURVEY 8c recipe) from the
            device-resident run and from the end-to-end run against the reference (oracle/_ref, else the
            oracle port) on the same batch

`--impl reference` times the reference's own CPU implementation (oracle/_ref, built from the reference's sources)
on the box's host cores on the same batch, at the best thread count of a sweep.

Multi-GPU (torchrun, one process per GPU): the path shards by document with no data-path collective; every rank
tokenizes its own rotated replica of the cfg-2 set (SURVEY 8d cfg 5), per-rank {docs, bytes, tokens} are
all-reduced over NCCL inside the step, and the step time is the max over ranks.  scaling = "weak".  Every rank
(and its pinned buffers) is bound to the CPUs of its GPU's NUMA node.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)

METRIC = "input GB/s (UTF-8), bert_base_tok TextToIds"

# algorithmic bytes per input byte.  cfg 2 (uint16 table entries): 1.99 compulsory stream (input + ids +
# offsets/counts) + 7.85 charmap + 6.17 class map + 6.17 transitions.  cfg 3 / cfg 4 "the same formula with
# (2.16 Mealy hops + 1.43 I2Info x 12 B)/byte and (3.96 hops + 2.68 x 12 B + 20 B charmap)/cp": a hop is one 16-byte
# double-array entry here (seg_tables.h), the compulsory stream is input + 4 B x ids/byte + offsets/counts.
CONFIGS = {
    "cfg2": dict(model="bert_base_tok.bin", unk=100, max_ids=512, pool="EN", seed=2, fixed_len=512, emoji=0,
                 kernel="wp_tokenize_kernel", what="TextToIds, synthetic English docs ~512 B (seed 2)"),
    "cfg3": dict(model="gpt2.bin", unk=0, max_ids=4096, pool="EN", seed=3, fixed_len=0, emoji=0,
                 kernel="sp_bpe_kernel", what="byte-BPE TextToIds, docs log-uniform 64..4096 B (seed 3)"),
    "cfg4": dict(model="xlm_roberta_base.bin", unk=3, max_ids=512, pool="MULTI", seed=4, fixed_len=512, emoji=16,
                 kernel="sp_unigram_kernel", what="Unigram-LM TextToIds, multilingual docs ~512 B (seed 4)"),
}


def algo_bytes_per_input_byte(cfg_name, nbytes, ndocs, tokens, cps=None):
    stream = 1.0 + 4.0 * tokens / nbytes + 12.0 * ndocs / nbytes
    if cfg_name == "cfg2":
        return 22.18
    if cfg_name == "cfg3":
        return stream + 2.16 * 16 + 1.43 * 12
    cp_per_byte = (cps / nbytes) if cps else 0.55
    return stream + cp_per_byte * (3.96 * 16 + 2.68 * 12 + 20)


def model_file(name):
    return os.path.join(ROOT, "data", "ldb", name)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------- NUMA

def gpu_numa_cpus(local_rank):
    """CPUs of the NUMA node GPU `local_rank` hangs off (sysfs), or None."""
    try:
        q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                           capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bus = q[-12:]                       # sysfs uses a 4-digit domain
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        return (node, cpus) if cpus else None
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Summary over the samples taken inside [t0, t1] (the timed region); falls back to every
        sample since start() (warm-up + timed region, all under load) if the region was too short."""
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for (ts, r) in self.rows if t0 is not None and t0 <= ts <= t1]
        for r in (inside if len(inside) >= 3 else [r for (_, r) in self.rows]):
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def one_shot(self, under_load):
        """Fallback when the sampling process delivered nothing (a slow nvidia-smi): one query while `under_load()` keeps the GPU
        busy -- outside the timed region, same kernel, same clocks."""
        box = {}

        def q():
            try:
                box["out"] = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                            capture_output=True, text=True, timeout=30).stdout.strip()
            except Exception:
                box["out"] = ""
        t = threading.Thread(target=q)
        t.start()
        t0 = time.time()
        while t.is_alive() and time.time() - t0 < 30:
            under_load()
        t.join()
        self.rows = [(time.time(), box.get("out", ""))]
        r = self.stop()
        r["note"] = "one query under load right after the timed region (the sampler delivered no rows in time)"
        return r


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------- the CPU reference

_refdrv = None


def refdrv():
    global _refdrv
    if _refdrv is None:
        L = ctypes.CDLL(os.path.join(ROOT, "oracle", "librefdriver.so"))
        L.ref_digest_batch.restype = ctypes.c_double
        L.ref_digest_batch.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _refdrv = L
    return _refdrv


REF_SO = os.path.join(ROOT, "oracle", "_ref", "libblingfiretokdll.so")


def reference_cpu(cfg, text, offs, n_sample, threads, counts=None, digests=None):
    """The reference's CPU TextToIds (oracle/_ref) on the first n_sample documents with `threads` pinned host
    threads: (seconds, bytes, tokens)."""
    tok = ctypes.c_int64(0)
    sub = np.ascontiguousarray(offs[: n_sample + 1])
    secs = refdrv().ref_digest_batch(REF_SO.encode(), model_file(cfg["model"]).encode(), text.ctypes.data, sub.ctypes.data, n_sample,
                                     cfg["max_ids"], cfg["unk"], threads, ctypes.byref(tok),
                                     counts.ctypes.data if counts is not None else None,
                                     digests.ctypes.data if digests is not None else None)
    if secs <= 0:
        raise RuntimeError("reference driver failed (oracle/_ref missing?)")
    return secs, int(sub[-1]), int(tok.value)


def reference_sweep(cfg, text, offs, min_seconds=3.0, thread_counts=None):
    """Throughput of the reference at several thread counts (each point >= min_seconds of work on a prefix of the
    batch sized for it): {threads: GB/s}, best thread count."""
    n = len(offs) - 1
    cores = host_cores()
    tcs = thread_counts or sorted({1, min(32, cores), min(64, cores), cores})
    reference_cpu(cfg, text, offs, min(n, 2000), 1)                     # model load, page-in
    res = {}
    for t in tcs:
        ns = min(n, max(20000, t * 8000))
        secs = nb = 0.0
        tk = 0
        reference_cpu(cfg, text, offs, min(ns, 4000 * t), t)            # thread start-up, first touch
        while secs < min_seconds:
            s, b, k = reference_cpu(cfg, text, offs, ns, t)
            secs += s; nb += b; tk += k
        res[t] = {"GB_per_s": nb / secs / 1e9, "tokens_per_s": tk / secs, "docs_per_pass": ns, "seconds": secs}
    best = max(res, key=lambda t: res[t]["GB_per_s"])
    return res, best


def cpu_baseline_record(cfg, text, offs, min_seconds):
    try:
        sweep, best = reference_sweep(cfg, text, offs, min_seconds)
        return {"value": sweep[best]["GB_per_s"], "unit": "GB/s", "cores": best, "kind": "reference",
                "tokens_per_s": sweep[best]["tokens_per_s"],
                "sample": (f"oracle/_ref TextToIds on a prefix of the same batch ({sweep[best]['docs_per_pass']} docs per pass, "
                           f">= {min_seconds:g} s per point), threads pinned one per CPU, best of the sweep"),
                "one_thread": sweep[1]["GB_per_s"] if 1 in sweep else None,
                "sweep": {str(t): round(v["GB_per_s"], 4) for t, v in sweep.items()}, "host_cores": host_cores()}
    except Exception as e:   # the checker is optional for the GPU number
        return {"value": None, "unit": "GB/s", "cores": host_cores(), "kind": "reference", "sample": f"unavailable: {e}"}


def oracle_digests(cfg, text, offs):
    """Per-document (digest, count) of the reference on the whole batch: oracle/_ref when it is there, else the port."""
    n = len(offs) - 1
    dig = np.zeros(n, np.uint64)
    counts = np.zeros(n, np.int32)
    t0 = time.time()
    if os.path.exists(REF_SO):
        reference_cpu(cfg, text, offs, n, host_cores(), counts, dig)
        np.minimum(counts, cfg["max_ids"], out=counts)
        src = "oracle/_ref (the reference built from its own sources)"
    else:
        from _common import Oracle
        o = Oracle()
        h = o.load(model_file(cfg["model"]))
        _, dig, counts = o.digests(h, text, offs, cfg["max_ids"], cfg["unk"], host_cores())
        o.free(h)
        src = "oracle/bf_oracle.c (port)"
    return dig, counts, src, time.time() - t0


def make_batch(cfg, n):
    import corpus
    return corpus.gen_docs(cfg["pool"], n, seed=cfg["seed"], fixed_len=cfg["fixed_len"], emoji_every=cfg["emoji"])


def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference's CPU path on the same batch, best thread count of a sweep."""
    if rank != 0:
        return
    cfg = CONFIGS["cfg2"]
    n = args.docs
    text, offs = make_batch(cfg, n)
    sweep, best = reference_sweep(cfg, text, offs, min_seconds=3.0)
    for _ in range(min(args.warmup, 1)):
        reference_cpu(cfg, text, offs, n, best)
    t = []
    nbytes = tokens = 0
    for _ in range(args.steps):
        secs, nbytes, tokens = reference_cpu(cfg, text, offs, n, best)
        t.append(secs)
    ms = 1e3 * float(np.mean(t))
    gbs = nbytes / (ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": gbs, "unit": "GB/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32", "data": "synthetic",
        "tokens_per_s": tokens / (ms * 1e-3),
        "config": workload_config("cfg2", cfg, n, nbytes, world),
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": best, "kind": "reference", "host_cores": host_cores(),
                         "one_thread": sweep[1]["GB_per_s"] if 1 in sweep else None,
                         "sweep": {str(k): round(v["GB_per_s"], 4) for k, v in sweep.items()},
                         "sample": f"all {n} docs of the batch per step, {best} pinned threads (best of the sweep)"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(name, cfg, n, nbytes, world):
    return {"workload": f"{name}: {cfg['model']} {cfg['what']}, {n} docs, unk={cfg['unk']}, max_ids={cfg['max_ids']}",
            "docs_per_step_per_gpu": n, "bytes_per_step_per_gpu": nbytes,
            "l2": f"input per step ({nbytes / 1e6:.0f} MB) is larger than L2; no flush needed",
            "sharding": "one rotated replica of the set per rank; NCCL all-reduce of {docs,bytes,tokens}" if world > 1 else "single GPU"}


# ---------------------------------------------------------------------------------------------- one configuration

def run_config(name, args, rank, local_rank, world, dev, full):
    """Times one configuration on this rank.  Returns the record (rank 0) or None."""
    import torch
    import torch.distributed as dist
    import blingfire_b200 as bf

    cfg = CONFIGS[name]
    n = args.docs
    unk, max_ids = cfg["unk"], cfg["max_ids"]
    h = bf.load_model(model_file(cfg["model"]))
    L = bf.lib()

    text, offs = make_batch(cfg, n)
    if rank > 0:
        from blingfire_b200 import sharding
        text, offs = sharding.rotate_replica(text, offs, (rank * 15625) % n)
    nbytes = int(offs[-1])
    max_doc = int(np.diff(offs).max())

    # pinned host copies (the e2e leg reads these), device-resident copies (the kernel-only leg)
    h_text = torch.empty(nbytes + 64, dtype=torch.uint8, pin_memory=True)
    h_text[:nbytes].copy_(torch.from_numpy(text))
    h_offs = torch.from_numpy(offs).pin_memory()
    d_text = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
    d_text.copy_(h_text, non_blocking=True)
    d_offs = h_offs.to(dev, non_blocking=True)
    d_ids = torch.empty((n, max_ids), dtype=torch.int32, device=dev)
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    stats = torch.zeros(3, dtype=torch.int64, device=dev)
    stats_base = torch.tensor([n, nbytes, 0], dtype=torch.int64, device=dev)   # (device-side: a Python scalar written into a
                                                                               # CUDA tensor is a pageable copy that waits for the stream)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()

    def step():
        bf.text_to_ids_batch_device(h, d_text.data_ptr(), d_offs.data_ptr(), n, nbytes, d_ids.data_ptr(),
                                    d_counts.data_ptr(), max_ids, unk, stream.cuda_stream, max_doc_bytes=max_doc)
        if world > 1:
            # the path's only exchange: per-rank {docs, bytes, tokens}
            stats.copy_(stats_base)
            stats[2:3] += d_counts.sum(dtype=torch.int64)
            dist.all_reduce(stats)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(1.0)   # let nvidia-smi come up; it samples through warm-up and the timed region
    # the first pass after LoadModel meets every word for the first time (the word / segment tables are filled at run
    # time): timed on its own, reported as `cold_first_step`; the steady state is what `value` is
    cold = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for k in range(args.warmup):
        if k == 0:
            cold[0].record(stream)
        step()
        if k == 0:
            cold[1].record(stream)
    torch.cuda.synchronize()
    cold_ms = cold[0].elapsed_time(cold[1]) if (args.warmup > 0 and world == 1) else None   # (N > 1: the first step also sets NCCL up)
    if world > 1:
        dist.barrier()
    launches0 = bf.kernel_launches()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    torch.cuda.synchronize()
    t_wall0 = time.time()
    ev[0].record(stream)
    for i in range(args.steps):
        step()
        ev[i + 1].record(stream)
    torch.cuda.synchronize()
    t_wall1 = time.time()
    if world > 1:
        dist.barrier()
    bf.device_status(h, stream.cuda_stream)
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[args.steps])
    launches = bf.kernel_launches() - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    if rank == 0 and not clocks.get("samples"):
        def busy():      # the kernel alone (no collective: the other ranks have moved on)
            bf.text_to_ids_batch_device(h, d_text.data_ptr(), d_offs.data_ptr(), n, nbytes, d_ids.data_ptr(), d_counts.data_ptr(), max_ids, unk,
                                        stream.cuda_stream, max_doc_bytes=max_doc)
            torch.cuda.synchronize()
        clocks = sampler.one_shot(busy)
    tokens = int(d_counts.sum().item())

    # max over ranks of the timed region; total units over all ranks
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    agg = torch.tensor([nbytes, tokens, launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg)
    total_ms = float(t.item())
    all_bytes, all_tokens, all_launches = (int(x) for x in agg.tolist())
    ms_per_step = total_ms / args.steps
    gbs = all_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers ----
    def time_e2e(call, reps):
        call()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        tot = 0
        for _ in range(reps):
            tot = call()
        torch.cuda.synchronize()
        e_ms = (time.perf_counter() - t0) * 1e3 / reps
        te = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item()), tot

    e2e = e2e16 = e2e_pg = None
    h_ids = h_idoffs = None
    if not args.no_e2e:
        lens = np.diff(offs)
        per_doc = lens + 1 if name == "cfg3" else (2 * lens + 4 if name == "cfg4" else lens)
        cap = int(np.minimum(per_doc, max_ids).sum())
        h_ids = torch.empty(cap + 1, dtype=torch.int32, pin_memory=True)
        h_idoffs = torch.zeros(n + 1, dtype=torch.int64, pin_memory=True)

        def e2e_step():
            r = L.TextToIdsBatchCsr(ctypes.c_void_p(h), h_text.data_ptr(), h_offs.data_ptr(), n, h_ids.data_ptr(), cap,
                                    h_idoffs.data_ptr(), max_ids, unk)
            assert r >= 0, bf.last_error()
            return r

        reps = max(3, args.steps // 2) if full else args.steps
        e_ms, tot = time_e2e(e2e_step, reps)
        assert tot == tokens, "e2e path and device path disagree on the token count"
        e2e = {"value": all_bytes / (e_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": nbytes + 8 * (n + 1), "d2h_bytes_per_step": 4 * tot + 8 * (n + 1),
               "api": "TextToIdsBatchCsr (host pinned buffers)", "steps": reps}
        if name == "cfg2":
            # additive 16-bit form: half the bytes come back
            h_ids16 = h_ids.view(torch.int16)

            def e2e16_step():
                r = L.TextToIdsBatchCsrU16(ctypes.c_void_p(h), h_text.data_ptr(), h_offs.data_ptr(), n, h_ids16.data_ptr(), 2 * cap,
                                           h_idoffs.data_ptr(), max_ids, unk)
                assert r >= 0, bf.last_error()
                return r

            e16_ms, tot16 = time_e2e(e2e16_step, reps)
            assert tot16 == tokens
            e2e16 = {"value": all_bytes / (e16_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": e16_ms,
                     "h2d_bytes_per_step": nbytes + 8 * (n + 1), "d2h_bytes_per_step": 2 * tot16 + 8 * (n + 1),
                     "api": "TextToIdsBatchCsrU16 (host pinned buffers, 16-bit ids)", "steps": reps}
            e2e_step()   # leave the 32-bit ids in h_ids for the parity pass
            if full:
                # pageable caller memory: what the Python wrapper passes (numpy arrays)
                p_ids = np.empty(cap + 1, dtype=np.int32)
                p_ids[:] = 0     # touch the pages once: page faults of a fresh allocation are the caller's, not the call's
                p_off = np.zeros(n + 1, dtype=np.int64)

                def pg_step():
                    r = L.TextToIdsBatchCsr(ctypes.c_void_p(h), text.ctypes.data, offs.ctypes.data, n, p_ids.ctypes.data, cap,
                                            p_off.ctypes.data, max_ids, unk)
                    assert r >= 0, bf.last_error()
                    return r

                pg_ms, totp = time_e2e(pg_step, max(3, args.steps // 4))
                assert totp == tokens and (p_off == h_idoffs.numpy()).all()
                e2e_pg = {"value": all_bytes / (pg_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": pg_ms,
                          "api": "TextToIdsBatchCsr (pageable numpy buffers in and out, staged by the library)"}
                del p_ids

    rec = None
    if rank == 0:
        peak, peak_src = load_peaks()
        kern_ms = float(np.mean(step_ms))
        cps = None
        if name == "cfg4":
            cps = int((text & 0xC0 != 0x80).sum())
        A = algo_bytes_per_input_byte(name, nbytes, n, tokens, cps)
        achieved = A * nbytes / (kern_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(cfg["kernel"], {}).get("dram_bytes_per_launch")
        rec = {
            "value": gbs, "unit": "GB/s", "ms_per_step": ms_per_step, "tokens_per_s": all_tokens / (ms_per_step * 1e-3),
            "tokens_per_step": all_tokens, "config": workload_config(name, cfg, n, nbytes, world),
            "e2e": e2e, "gpu_launches": all_launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "profiles/r02_traffic.json (ncu --set full of this kernel)" if traffic else None,
                         "peak_source": peak_src, "kernel": cfg["kernel"],
                         "algorithmic_bytes_per_input_byte": A, "kernel_ms": kern_ms},
            "clocks": clocks,
            "cold_first_step": ({"ms": cold_ms, "value": nbytes / (cold_ms * 1e-3) / 1e9, "unit": "GB/s",
                                 "note": "first pass over the batch on a freshly loaded model (run-time tables empty, first launch of the kernel); not part of the timed region"}
                                if cold_ms else None),
        }
        if e2e16:
            rec["e2e_u16"] = e2e16
        if e2e_pg:
            rec["e2e_pageable"] = e2e_pg

    # ---- parity on EVERY document, outside the timed region ----
    if rank == 0 and not args.no_parity:
        from _common import Oracle
        o = Oracle()
        want_dig, want_counts, src, secs = oracle_digests(cfg, text, offs)
        want_tokens = int(want_counts.sum())
        checked = {}
        # (a) the device-resident rows: compact on the device, bring the CSR back
        d_csr = torch.empty(tokens + 1, dtype=torch.int32, device=dev)
        d_row = torch.empty(n + 1, dtype=torch.int64, device=dev)
        bf.compact_device(d_ids.data_ptr(), d_counts.data_ptr(), n, max_ids, d_csr.data_ptr(), d_row.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        row = d_row.cpu().numpy()
        dig_dev = o.csr_digests(d_csr.cpu().numpy(), row)
        bad_dev = int(((dig_dev != want_dig) | (np.diff(row) != want_counts)).sum())
        checked["device_resident"] = bad_dev
        del d_csr, d_row
        fold = o.fold(dig_dev, id_offsets=row)
        # (b) what the end-to-end call left in the caller's buffers
        if e2e is not None:
            io = h_idoffs.numpy()
            dig_e = o.csr_digests(h_ids.numpy(), io)
            checked["e2e"] = int(((dig_e != want_dig) | (np.diff(io) != want_counts)).sum())
        want_fold = o.fold(want_dig, counts=want_counts)
        rec["parity"] = {"docs": n, "tokens": tokens, "oracle_tokens": want_tokens, "digest": f"{fold:016x}",
                         "oracle_digest": f"{want_fold:016x}", "mismatching_docs": checked,
                         "match": tokens == want_tokens and fold == want_fold and all(v == 0 for v in checked.values()),
                         "oracle": src, "oracle_seconds": round(secs, 2),
                         "recipe": "FNV-1a-64 per document over its uint32 ids; folded over (count, digest) in document order"}

    if rank == 0 and not args.no_cpu:
        rec["cpu_baseline"] = cpu_baseline_record(cfg, text, offs, args.cpu_seconds)

    bf.free_model(h)
    del d_ids, d_text, h_text, h_ids
    torch.cuda.empty_cache()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--docs", type=int, default=1_000_000)
    ap.add_argument("--configs", default="", help="comma-separated subset of cfg2,cfg3,cfg4 (default: all at 1 GPU, cfg2 otherwise)")
    ap.add_argument("--cpu-seconds", type=float, default=3.0, help="work per point of the cpu_baseline thread sweep")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-numa", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # bind this rank (and the pinned memory it is about to allocate: first touch) to its GPU's NUMA node
    all_cpus = os.sched_getaffinity(0)
    numa = None if args.no_numa else gpu_numa_cpus(local_rank)
    if numa:
        os.sched_setaffinity(0, numa[1])

    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py needs a GPU: blingfire_b200 has no CPU path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.zeros(1, device=dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    names = [c for c in args.configs.split(",") if c] or (["cfg2", "cfg3", "cfg4"] if world == 1 else ["cfg2"])
    recs = {}
    for name in names:
        # the CPU legs (parity oracle, cpu_baseline) use every core; the GPU legs run on the GPU's NUMA node
        rec = run_config_with_affinity(name, args, rank, local_rank, world, dev, numa, all_cpus)
        if rec is not None:
            recs[name] = rec

    if rank == 0:
        head = recs.get("cfg2") or next(iter(recs.values()))
        line = {
            "metric": METRIC, "value": head["value"], "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int32", "data": "synthetic", "tokens_per_s": head["tokens_per_s"], "config": head["config"],
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
            "cpu_baseline": head.get("cpu_baseline"), "clocks": head["clocks"], "parity": head.get("parity"),
            "numa": {"node": numa[0], "cpus": len(numa[1])} if numa else None,
        }
        for k in ("e2e_u16", "e2e_pageable", "cold_first_step"):
            if k in head:
                line[k] = head[k]
        others = {k: v for k, v in recs.items() if v is not head}
        if others:
            line["other_configs"] = others
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_config_with_affinity(name, args, rank, local_rank, world, dev, numa, all_cpus):
    """run_config() with the CPU legs on every core: the affinity is widened inside through this hook."""
    if not numa:
        return run_config(name, args, rank, local_rank, world, dev, full=(world == 1))
    global oracle_digests, cpu_baseline_record
    narrow = numa[1]
    od, cb = oracle_digests, cpu_baseline_record

    def wide(fn):
        def inner(*a, **k):
            os.sched_setaffinity(0, all_cpus)
            try:
                return fn(*a, **k)
            finally:
                os.sched_setaffinity(0, narrow)
        return inner

    oracle_digests, cpu_baseline_record = wide(od), wide(cb)
    try:
        return run_config(name, args, rank, local_rank, world, dev, full=(world == 1))
    finally:
        oracle_digests, cpu_baseline_record = od, cb


if __name__ == "__main__":
    main()
