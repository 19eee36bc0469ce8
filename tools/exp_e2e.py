"""One-GPU experiment: where does the host-buffer pipeline spend its time?"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "sparse-voxel-octrees_b200"))
import numpy as np
import bench, pysvo
w = bench.pick_workload(sys.argv[1] if len(sys.argv) > 1 else "c2_sdf2048_4k")
tree = pysvo.VoxelOctree(w["path"])
W, H = w["width"], w["height"]
cams = [pysvo.orbit_camera(*c) for c in bench.cameras(None, w, 100)]
hosts = [pysvo.PinnedArray((H, W), np.uint32), pysvo.PinnedArray((H, W), np.uint32)]
for rep in range(3):
    pending = [None, None]
    t_issue = t_wait = 0.0
    pysvo.device_synchronize(0)
    t0 = time.perf_counter()
    for k in range(100):
        s = k & 1
        if pending[s] is not None:
            a = time.perf_counter(); tree.frame_wait(pending[s]); t_wait += time.perf_counter() - a
        a = time.perf_counter()
        pending[s] = tree.render_frame_async(cams[k], W, H, hosts[s].array, strips=16, flavour=1)
        t_issue += time.perf_counter() - a
    for p in pending:
        tree.frame_wait(p)
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {dt*10:.3f} ms/frame; issue {t_issue*10:.3f} ms/frame, wait {t_wait*10:.3f} ms/frame", flush=True)
# synchronous API
t0 = time.perf_counter()
for k in range(50):
    tree.render_frame(cams[k], W, H, strips=16, flavour=1, rgba=hosts[0].array, want_stats=False)
print(f"sync api: {(time.perf_counter()-t0)*20:.3f} ms/frame")
# device-only frames on one stream
buf = pysvo.DeviceBuffer(0, W*H*4)
for k in range(10): tree.render_frame_device(cams[k], W, H, buf.ptr, strips=16, flavour=1)
pysvo.device_synchronize(0)
t0 = time.perf_counter()
for k in range(100): tree.render_frame_device(cams[k], W, H, buf.ptr, strips=16, flavour=1)
pysvo.device_synchronize(0)
print(f"device api, one stream: {(time.perf_counter()-t0)*10:.3f} ms/frame")
# raw D2H speed
import ctypes as C
t0 = time.perf_counter()
for k in range(20):
    pysvo._check(pysvo.lib().svo_device_to_host(0, C.c_void_p(hosts[0].array.ctypes.data), C.c_void_p(buf.ptr), W*H*4))
dt = (time.perf_counter()-t0)/20
print(f"D2H {W*H*4/dt/1e9:.1f} GB/s ({dt*1e3:.3f} ms/frame)")
