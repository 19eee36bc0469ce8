#!/usr/bin/env python
"""What the box's host links carry when N GPUs ship one 3840x2160 RGBA frame into ONE page-locked host frame at the same
time, by the shape of each GPU's share: horizontal bands (contiguous rows) against vertical stripes (strided rows, what
svo_multi's host leg copies). Copy engines only (torch non_blocking copies = cudaMemcpy[2D]Async).

    python tools/d2h_aggregate.py [N]
"""
import sys
import time

import torch

N = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
H, W, ITERS = 2160, 3840, 200
host = torch.empty((H, W), dtype=torch.int32).pin_memory()
dev = [torch.zeros((H, W), dtype=torch.int32, device=f"cuda:{i}") for i in range(N)]
streams = [torch.cuda.Stream(device=i) for i in range(N)]


def run(name, views):
    for _ in range(3):
        for i in range(N):
            with torch.cuda.stream(streams[i]):
                for hv, dv in views(i):
                    hv.copy_(dv, non_blocking=True)
    for i in range(N):
        torch.cuda.synchronize(i)
    t0 = time.perf_counter()
    for _ in range(ITERS):
        for i in range(N):
            with torch.cuda.stream(streams[i]):
                for hv, dv in views(i):
                    hv.copy_(dv, non_blocking=True)
    for i in range(N):
        torch.cuda.synchronize(i)
    s = time.perf_counter() - t0
    print(f"N={N} {name}: {H * W * 4 * ITERS / s / 1e9:.1f} GB/s aggregate, {s / ITERS * 1e3:.3f} ms per frame", flush=True)


def bands(i):
    r0, r1 = H * i // N, H * (i + 1) // N
    return [(host[r0:r1], dev[i][r0:r1])]


def stripes(run_cols):
    px = run_cols * 8
    def f(i):
        return [(host[:, x0:x0 + px], dev[i][:, x0:x0 + px]) for x0 in range(i * px, W, N * px)]
    return f


def interleaved_bands(rows):
    def f(i):
        return [(host[r0:r0 + rows], dev[i][r0:r0 + rows]) for r0 in range(i * rows, H, N * rows)]
    return f


run("one GPU alone, whole frame", lambda i: [(host, dev[0])] if i == 0 else [])
run("contiguous band per GPU", bands)
run("interleaved bands of 64 rows", interleaved_bands(64))
run("stripes 30 tile columns (960 B rows)", stripes(30))
run("stripes 60 tile columns (1920 B rows)", stripes(60))
run("stripes 15 tile columns (480 B rows)", stripes(15))
