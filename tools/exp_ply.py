import sys, time, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/sparse-voxel-octrees_b200')
import pysvo
from tools import make_scenes
scene = sys.argv[1] if len(sys.argv) > 1 else "ico8192"
kind, res, freq = make_scenes.SCENES[scene]
ply = f"/tmp/{scene}.ply"
make_scenes.gen_lib().svo_scene_icosphere_ply(ply.encode(), freq, make_scenes.SEED)
os.environ["SVO_BUILD_DEBUG"] = "1"
for i in range(3):
    t = time.perf_counter(); tree = pysvo.VoxelOctree.build_from_ply(ply, res, mem_budget=12884901888, threads=8); dt = time.perf_counter() - t
    vs, bs = pysvo.VoxelOctree.last_voxelize_stats(), pysvo.VoxelOctree.last_build_stats()
    print(i, "wall %.3f s" % dt, "vox %.1f/%.1f/%.1f ms" % (vs.overlap_ms, vs.sort_ms, vs.fold_ms), "build %.1f/%.1f/%.1f/%.1f ms" % (bs.gather_ms, bs.sort_ms, bs.levels_ms, bs.emit_ms), flush=True)
    tree.close()
