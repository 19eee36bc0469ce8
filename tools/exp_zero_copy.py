"""Experiment: device -> host throughput of svo_frame_copy_owned_tiles writing into mapped page-locked host memory
(SM-issued PCIe writes, 128 B per warp) against the copy engine (cudaMemcpyAsync), one GPU, 3840x2160 frame."""
import sys
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/sparse-voxel-octrees_b200')
import torch
import pysvo

W, H, S = 3840, 2160, 16
nbytes = W * H * 4
dev = 0
torch.cuda.set_device(dev)
fb = pysvo.DeviceBuffer(dev, nbytes)
fb.from_host(np.arange(W * H, dtype=np.uint32))
host = np.zeros((H, W), np.uint32)
mapped = pysvo.host_register(dev, host)
pinned = pysvo.PinnedArray((H, W), np.uint32)
stream = torch.cuda.current_stream().cuda_stream


def timed(fn, n=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for world, run in ((1, 4), (2, 4), (4, 4), (8, 4), (2, 8), (2, 16), (4, 15), (8, 15), (2, 60)):
    pysvo.frame_set_tile_run(run)
    ms = timed(lambda: pysvo.frame_copy_owned_tiles(dev, W, H, S, 0, world, fb.ptr, mapped, stream))
    print(f"zero-copy kernel, rank 0 of {world}, stripes of {run} tile columns ({run * 32} B): {ms:.4f} ms, {nbytes / world / ms / 1e6:.1f} GB/s")
pysvo.frame_set_tile_run(0)
assert np.array_equal(host[:, :32], np.arange(W * H, dtype=np.uint32).reshape(H, W)[:, :32])
ms = timed(lambda: pysvo.device_to_host_async(dev, pinned.array, fb.ptr, nbytes, stream))
print(f"copy engine, whole frame: {ms:.4f} ms, {nbytes / ms / 1e6:.1f} GB/s")
ms = timed(lambda: pysvo.device_to_host_async(dev, host, fb.ptr, nbytes, stream))
print(f"copy engine into the registered array: {ms:.4f} ms, {nbytes / ms / 1e6:.1f} GB/s")
pysvo.host_unregister(host)
