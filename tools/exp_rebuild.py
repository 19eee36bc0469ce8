import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/sparse-voxel-octrees_b200')
import numpy as np, pysvo
from tools import make_scenes
scene = sys.argv[1] if len(sys.argv) > 1 else "sdf2048"
words, center = make_scenes.load_scene(scene)
tree = pysvo.VoxelOctree(words=words, center=center)
side = 1 << tree.depth
dims = tuple(int(round(float(c) * 2 * side)) for c in center)
for i in range(6):
    t = time.perf_counter(); again = tree.rebuild(dims); dt = time.perf_counter() - t
    st = pysvo.VoxelOctree.last_build_stats()
    print(i, 'wall %.1f ms' % (dt * 1e3), 'gather %.2f sort %.2f levels %.2f emit %.2f' % (st.gather_ms, st.sort_ms, st.levels_ms, st.emit_ms))
    again.close()
