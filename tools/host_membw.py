#!/usr/bin/env python3
"""Host DRAM bandwidth of one socket, measured the blunt way: T threads bound to the CPUs of NUMA node 0 copy private 256 MB
arrays (numpy releases the GIL); reported as read + write traffic.  Context for the multi-GPU end-to-end numbers:
four GPUs of one socket move ~190 GB/s of DMA through that socket's memory."""
import os
import threading
import time

import numpy as np


def node_cpus(node):
    spec = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    cpus = set()
    for part in spec.split(","):
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus & os.sched_getaffinity(0)


def run(threads, node=0, mb=256, reps=6):
    cpus = sorted(node_cpus(node))
    os.sched_setaffinity(0, set(cpus))
    srcs = [np.ones(mb << 20, np.uint8) for _ in range(threads)]
    dsts = [np.empty(mb << 20, np.uint8) for _ in range(threads)]
    for s, d in zip(srcs, dsts):
        d[:] = s                                   # first touch on this node
    bar = threading.Barrier(threads + 1)

    def work(i):
        os.sched_setaffinity(0, {cpus[i % len(cpus)]})
        bar.wait()
        for _ in range(reps):
            np.copyto(dsts[i], srcs[i])
        bar.wait()

    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ts:
        t.start()
    bar.wait()
    t0 = time.perf_counter()
    bar.wait()
    dt = time.perf_counter() - t0
    for t in ts:
        t.join()
    return 2.0 * threads * (mb << 20) * reps / dt / 1e9


if __name__ == "__main__":
    for t in (1, 4, 8, 16, 32):
        print(f"node 0, {t:2d} threads: {run(t):7.1f} GB/s (read + write)")
