#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

  metric    input GB/s of UTF-8 through TextToIds (+ tokens/s), bert_base_tok.bin
  workload  configs[1]: 1 M synthetic English documents of ~512 B (SURVEY 8d cfg 2, seed 2)
  step      one pass of the hot path over the whole 1 M-document batch (~512 MB of input,
            larger than L2, so no L2 flush is needed between timed steps)
  value     device-resident: the CSR batch is already in HBM, ids/counts stay in HBM; CUDA events
            on the launching stream
  e2e       the same batch through the C-ABI call a user makes (TextToIdsBatchCsr) with HOST
            (pinned) buffers: host->device and device->host copies inside the timed region

`--impl reference` times the reference's own CPU implementation (oracle/_ref, built from the
reference's sources) on the box's host cores on a bounded sample of the same workload.

Multi-GPU (torchrun, one process per GPU): the path shards by document with no data-path
collective; every rank tokenizes its own rotated replica of the cfg-2 set (SURVEY 8d cfg 5),
per-rank {docs, bytes, tokens} are all-reduced over NCCL inside the step, and the step time is
the max over ranks.  scaling = "weak".
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
MODEL = "bert_base_tok.bin"
UNK = 100
MAX_IDS = 512
# algorithmic bytes per input byte for cfg 2 with uint16 table entries
#   1.99 compulsory stream (input + ids + offsets/counts) + 7.85 charmap + 6.17 class map + 6.17 transitions
ALGO_BYTES_PER_INPUT_BYTE = 22.18


def model_file():
    return os.path.join(ROOT, "data", "ldb", MODEL)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


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


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def reference_cpu(text, offs, n_sample, threads):
    """The reference's CPU TextToIds on the first n_sample documents, `threads` host threads."""
    L = ctypes.CDLL(os.path.join(ROOT, "oracle", "librefdriver.so"))
    L.ref_time_batch.restype = ctypes.c_double
    L.ref_time_batch.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    ref = os.path.join(ROOT, "oracle", "_ref", "libblingfiretokdll.so")
    tok = ctypes.c_int64(0)
    sub_offs = np.ascontiguousarray(offs[: n_sample + 1])
    secs = L.ref_time_batch(ref.encode(), model_file().encode(), text.ctypes.data, sub_offs.ctypes.data, n_sample,
                            MAX_IDS, UNK, threads, ctypes.byref(tok), None)
    if secs <= 0:
        raise RuntimeError("reference driver failed (oracle/_ref missing?)")
    return secs, int(sub_offs[-1]), int(tok.value)


def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference's CPU path, all host threads, bounded sample per step."""
    if rank != 0:
        return
    import corpus
    n_docs = args.docs
    n_sample = min(n_docs, args.ref_sample)
    text, offs = corpus.gen_docs("EN", n_sample, seed=2, fixed_len=512)   # prefix of the cfg-2 stream
    cores = host_cores()
    for _ in range(args.warmup):
        reference_cpu(text, offs, min(n_sample, 20000), cores)
    t = []
    nbytes = tokens = 0
    for _ in range(args.steps):
        secs, nbytes, tokens = reference_cpu(text, offs, n_sample, cores)
        t.append(secs)
    ms = 1e3 * float(np.mean(t))
    gbs = nbytes / (ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": gbs, "unit": "GB/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32", "data": "synthetic",
        "tokens_per_s": tokens / (ms * 1e-3),
        "config": {"workload": f"cfg2: {MODEL} TextToIds, synthetic English docs ~512 B (seed 2), unk={UNK}, max_ids={MAX_IDS}",
                   "docs_per_step": n_sample, "bytes_per_step": nbytes},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "reference",
                         "sample": f"first {n_sample} docs of the cfg-2 stream per step"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--docs", type=int, default=1_000_000)
    ap.add_argument("--ref-sample", type=int, default=500_000, help="documents per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=1_000_000, help="documents of the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import blingfire_b200 as bf
    import corpus

    assert torch.cuda.is_available(), "bench.py needs a GPU: blingfire_b200 has no CPU path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.zeros(1, device=dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    h = bf.load_model(model_file())
    assert bf.lib().BlingFireB200ModelEngine(h) == 1

    # cfg 2 document set; rank r works on the replica rotated by r * 15625 documents (cfg 5 rule)
    n = args.docs
    text, offs = corpus.cfg2(n) if n == 1_000_000 else corpus.gen_docs("EN", n, seed=2, fixed_len=512)
    if rank > 0:
        from blingfire_b200 import sharding
        text, offs = sharding.rotate_replica(text, offs, (rank * 15625) % n)
    nbytes = int(offs[-1])

    # pinned host copies (the e2e leg reads these), device-resident copies (the kernel-only leg)
    h_text = torch.empty(nbytes + 64, dtype=torch.uint8, pin_memory=True)
    h_text[:nbytes].copy_(torch.from_numpy(text))
    h_offs = torch.from_numpy(offs).pin_memory()
    d_text = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
    d_text.copy_(h_text, non_blocking=True)
    d_offs = h_offs.to(dev, non_blocking=True)
    d_ids = torch.empty((n, MAX_IDS), dtype=torch.int32, device=dev)
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    stats = torch.zeros(3, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    stream = torch.cuda.current_stream()

    def step():
        bf.text_to_ids_batch_device(h, d_text.data_ptr(), d_offs.data_ptr(), n, nbytes, d_ids.data_ptr(),
                                    d_counts.data_ptr(), MAX_IDS, UNK, stream.cuda_stream)
        if world > 1:
            # the path's only exchange: per-rank {docs, bytes, tokens}
            stats[0] = n; stats[1] = nbytes; stats[2] = d_counts.sum()
            dist.all_reduce(stats)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)   # let nvidia-smi come up; it samples through warm-up and the timed region
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
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
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[args.steps])
    launches = bf.kernel_launches() - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
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

    # ---- end to end through the C ABI with host (pinned) buffers ----
    e2e = None
    if not args.no_e2e:
        cap = nbytes   # at most one id per input byte
        h_ids = torch.empty(cap, dtype=torch.int32, pin_memory=True)
        h_idoffs = torch.zeros(n + 1, dtype=torch.int64, pin_memory=True)
        L = bf.lib()

        def e2e_step():
            r = L.TextToIdsBatchCsr(ctypes.c_void_p(h), h_text.data_ptr(), h_offs.data_ptr(), n, h_ids.data_ptr(), cap,
                                    h_idoffs.data_ptr(), MAX_IDS, UNK)
            assert r >= 0, bf.last_error()
            return r

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        tot = 0
        for _ in range(args.steps):
            tot = e2e_step()
        torch.cuda.synchronize()
        e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        te = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_ms = float(te.item())
        assert tot == tokens, "e2e path and device path disagree on the token count"
        e2e = {"value": all_bytes / (e_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": nbytes + 8 * (n + 1), "d2h_bytes_per_step": 4 * tot + 8 * (n + 1),
               "api": "TextToIdsBatchCsr (host pinned buffers)"}

    if rank == 0:
        peak, peak_src = load_peaks()
        # dominant kernel = wp_tokenize_kernel (the only kernel in the device-resident step)
        kern_ms = float(np.mean(step_ms))
        achieved = ALGO_BYTES_PER_INPUT_BYTE * nbytes / (kern_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        cpu = None
        if not args.no_cpu:
            try:
                cores = host_cores()
                ns = min(n, args.cpu_sample)
                reference_cpu(text, offs, min(ns, 50000), cores)      # warm-up: model load, thread start
                runs = [reference_cpu(text, offs, ns, cores) for _ in range(3)]
                secs, b, tk = sorted(runs)[1]                        # median of 3
                cpu = {"value": b / secs / 1e9, "unit": "GB/s", "cores": cores, "kind": "reference",
                       "sample": f"first {ns} docs of the same cfg-2 batch, oracle/_ref TextToIds, {cores} threads",
                       "tokens_per_s": tk / secs}
            except Exception as e:   # the checker is optional for the GPU number
                cpu = {"value": None, "unit": "GB/s", "cores": host_cores(), "kind": "reference", "sample": f"unavailable: {e}"}
        line = {
            "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int32", "data": "synthetic",
            "tokens_per_s": all_tokens / (ms_per_step * 1e-3),
            "config": {"workload": f"cfg2: {MODEL} TextToIds, 1M synthetic English docs ~512 B (seed 2), unk={UNK}, max_ids={MAX_IDS}",
                       "docs_per_step_per_gpu": n, "bytes_per_step_per_gpu": nbytes,
                       "l2": "input per step (512 MB) is larger than L2; no flush needed",
                       "sharding": "one rotated replica of the set per rank; NCCL all-reduce of {docs,bytes,tokens}" if world > 1 else "single GPU"},
            "e2e": e2e, "gpu_launches": all_launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "wp_tokenize_kernel",
                         "algorithmic_bytes_per_input_byte": ALGO_BYTES_PER_INPUT_BYTE, "kernel_ms": kern_ms},
            "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line))
    bf.free_model(h)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
