#!/usr/bin/env python
"""Renders a few frames of a bench workload; the thing to wrap in ncu (one GPU, short).

    ncu --set full --clock-control none --import-source on -k regex:finePass -s 3 -c 1 -o gpurun_out/prof \
        python tools/profile_frames.py --workload c2_sdf2048_4k --frames 6
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "sparse-voxel-octrees_b200"))

import bench  # noqa: E402
import pysvo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default=None)
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--validation", action="store_true")
    ap.add_argument("--cam", type=int, default=0)
    a = ap.parse_args()
    w = bench.pick_workload(a.workload)
    tree = pysvo.VoxelOctree(w["path"])
    buf = pysvo.DeviceBuffer(0, w["width"] * w["height"] * 4)
    cams = [pysvo.orbit_camera(*c) for c in bench.cameras(None, w, bench.ORBIT)]
    fl = pysvo.FLAVOUR_VALIDATION if a.validation else pysvo.FLAVOUR_FAST
    for k in range(a.frames):
        st = tree.render_frame_device(cams[(a.cam + k) % bench.ORBIT], w["width"], w["height"], buf.ptr, strips=bench.STRIPS,
                                      flavour=fl, want_stats=True)
        print(f"frame {k}: coarse {st.coarse_ms:.4f} ms, fine {st.fine_ms:.4f} ms, rays {st.rays}")
    pysvo.device_synchronize(0)


if __name__ == "__main__":
    main()
