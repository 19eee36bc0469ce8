"""One-GPU experiment: cost of one rank's share of a frame under different tile interleaves."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "sparse-voxel-octrees_b200"))
import numpy as np
import bench, pysvo
w = bench.pick_workload(sys.argv[1] if len(sys.argv) > 1 else "c2_sdf2048_4k")
tree = pysvo.VoxelOctree(w["path"])
W, H = w["width"], w["height"]
buf = pysvo.DeviceBuffer(0, W * H * 4)
cams = [pysvo.orbit_camera(*c) for c in bench.cameras(None, w, 20)]
for world in (1, 2, 4, 8):
    res = []
    for rank in range(min(world, 2)):
        for rep in range(2):
            ms = []
            for cam in cams:
                st = tree.render_frame_device(cam, W, H, buf.ptr, strips=16, flavour=1, tile_rank=rank, tile_world=world, want_stats=True)
                ms.append((st.coarse_ms, st.fine_ms, st.coarse_rays, st.fine_rays))
        a = np.array(ms)
        res.append(a.mean(0))
    print(f"run={os.environ.get('SVO_TILE_RUN','4')} world={world}: " + " | ".join(f"beam {r[0]:.4f} ms ({int(r[2])} rays) fine {r[1]:.4f} ms ({int(r[3])} rays)" for r in res), flush=True)
