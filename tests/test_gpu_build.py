"""GPU: octree construction -- the node array built in HBM must be, word for
word, the array the reference's VoxelOctree(VoxelData*) builds (reference src/VoxelOctree.cpp:125-205).
Checked against the plain-C restatement (oracle/svo_oracle.c, pinned to the reference in
tests/test_oracle_pins.py), against the reference's own object code (oracle/_ref) through raw .voxel
files, and against the reference-built scenes the benchmark uses."""
import struct

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def random_volume(rng, w, h, d, p):
    vox = (rng.random((d, h, w)) < p) * rng.integers(1, 2**32, (d, h, w), dtype=np.uint64)
    vox = vox.astype(np.uint32)
    if vox.max() == 0:
        vox[0, 0, 0] = 7
    return vox


def write_voxel_file(path, vox):
    d, h, w = vox.shape
    with open(path, "wb") as fp:
        fp.write(struct.pack("<iii", w, h, d))
        fp.write(np.ascontiguousarray(vox, np.uint32).tobytes())


SHAPES = [(2, 2, 2, 1.0), (3, 2, 2, 0.7), (8, 8, 8, 0.3), (16, 16, 16, 0.05), (32, 32, 32, 0.5), (20, 12, 6, 0.2),
          (33, 17, 10, 0.1), (7, 9, 5, 0.5), (64, 64, 64, 1.0), (100, 60, 30, 0.03), (128, 128, 128, 0.2), (1, 1, 64, 0.5),
          (256, 2, 2, 0.9)]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s[:3])))
def test_build_equals_oracle(pysvo, port, shape):
    """Dense host grid -> HBM. Covers non-power-of-two and odd dimensions (the reference's blind last
    z-plane, VoxelData.cpp:160-163), a completely filled volume (far words on every level) and slivers."""
    w, h, d, p = shape
    vox = random_volume(np.random.default_rng(w * 1000003 + h * 1009 + d), w, h, d, p)
    want, wcenter = port.build_octree(vox)
    tree = pysvo.VoxelOctree.build_from_voxels(vox)
    got = tree.words()
    st = pysvo.VoxelOctree.last_build_stats()
    assert got.size == want.size and np.array_equal(got, want)
    assert np.array_equal(tree.center(), wcenter)
    side = 1 << tree.depth
    assert side >= max(w, h, d) and side // 2 < max(w, h, d, 2)
    assert st.words == want.size and st.voxels > 0
    tree.close()


def test_build_from_voxel_file_equals_reference_builder(pysvo, ref, tmp_path):
    """The reference's own pipeline on the same file: VoxelData(path, mem) + VoxelOctree(VoxelData*)."""
    rng = np.random.default_rng(77)
    for (w, h, d, p) in [(64, 48, 40, 0.08), (96, 96, 96, 0.6)]:
        vox = random_volume(rng, w, h, d, p)
        path = tmp_path / f"v_{w}.voxel"
        write_voxel_file(path, vox)
        hnd = ref.tree_build_voxel_file(path, 1 << 30)
        want = ref.tree_words(hnd)
        wcenter = ref.tree_center(hnd)
        ref.tree_destroy(hnd)
        tree = pysvo.VoxelOctree.build_from_voxel_file(path)
        assert np.array_equal(tree.words(), want) and np.array_equal(tree.center(), wcenter)
        tree.close()


def test_build_sparse_list(pysvo, port):
    rng = np.random.default_rng(5)
    w, h, d = 70, 40, 52
    vox = random_volume(rng, w, h, d, 0.04)
    want, _ = port.build_octree(vox)
    z, y, x = np.nonzero(vox)
    order = rng.permutation(x.size)                       # any order
    xyz = np.stack([x, y, z], 1).astype(np.uint32)[order]
    vals = vox[z, y, x][order]
    # entries the dense path would never see: zero words and coordinates outside the volume
    xyz = np.concatenate([xyz, np.array([[w, 0, 0], [0, h + 3, 0], [1, 1, d]], np.uint32)], 0)
    vals = np.concatenate([vals, np.array([9, 9, 9], np.uint32)])
    zero_at = np.array([[2, 3, 4]], np.uint32)
    if vox[4, 3, 2] == 0:
        xyz = np.concatenate([xyz, zero_at], 0)
        vals = np.concatenate([vals, np.zeros(1, np.uint32)])
    tree = pysvo.VoxelOctree.build_from_sparse(xyz, vals, (w, h, d))
    assert np.array_equal(tree.words(), want)
    tree.close()
    with pytest.raises(pysvo.SvoError) as e:              # a coordinate named twice
        pysvo.VoxelOctree.build_from_sparse(np.concatenate([xyz, xyz[:1]], 0), np.concatenate([vals, vals[:1]]), (w, h, d))
    assert e.value.status == 3 and "more than once" in str(e.value)


def test_build_errors(pysvo):
    with pytest.raises(pysvo.SvoError) as e:
        pysvo.VoxelOctree.build_from_voxels(np.zeros((8, 8, 8), np.uint32))
    assert e.value.status == 3 and "nothing to build" in str(e.value)
    with pytest.raises(pysvo.SvoError) as e:              # only the blind plane of an odd-depth volume is filled
        v = np.zeros((3, 4, 4), np.uint32)
        v[2] = 5
        pysvo.VoxelOctree.build_from_voxels(v)
    assert e.value.status == 3
    with pytest.raises(pysvo.SvoError) as e:
        pysvo.VoxelOctree.build_from_voxels(np.ones((1, 1, 1), np.uint32))
    assert e.value.status == 1
    with pytest.raises(pysvo.SvoError) as e:
        pysvo.VoxelOctree.build_from_voxel_file("/nonexistent/x.voxel")
    assert e.value.status == 2


@pytest.mark.parametrize("scene", ["sdf256", "sdf512"])
def test_build_sdf_scene_equals_reference_built_scene(pysvo, port, tmp_path, scene):
    """The benchmark's SDF scene (tools/scene_gen.c raw .voxel file): the tree built on the GPU equals the
    cached tree the reference builder made from the same file (512^3 = 128 Mi voxels: multi-chunk
    streaming), and renders the same frame."""
    from tools import make_scenes
    cached = make_scenes.scene_path(scene)
    if not cached.exists():
        pytest.skip(f"{cached} not present")
    res = int(scene[3:])
    raw = tmp_path / f"{scene}.voxel"
    filled = make_scenes.gen_lib().svo_scene_sdf_voxel_file(str(raw).encode(), res, make_scenes.SEED)
    assert filled > 0
    want, wcenter = pysvo.oct_read(cached)
    tree = pysvo.VoxelOctree.build_from_voxel_file(raw)
    st = pysvo.VoxelOctree.last_build_stats()
    assert st.voxels == filled
    got = tree.words()
    assert got.size == want.size and np.array_equal(got, want) and np.array_equal(tree.center(), wcenter)
    cam = pysvo.orbit_camera(20.0, 40.0, 0.9)
    loaded = pysvo.VoxelOctree(cached)
    a, _, _ = tree.render_frame(cam, 640, 360, strips=4, flavour=pysvo.FLAVOUR_VALIDATION)
    b, _, _ = loaded.render_frame(cam, 640, 360, strips=4, flavour=pysvo.FLAVOUR_VALIDATION)
    assert np.array_equal(a, b) and (a != 0xFF000000).any()
    tree.close()
    loaded.close()


def test_build_multi_chunk_dense_host_grid(pysvo, port):
    """More than one 16 Mi-voxel upload chunk from host memory, chunk boundary in the middle of a z-plane."""
    rng = np.random.default_rng(9)
    w, h, d = 300, 280, 210                               # 17.6 Mi voxels
    vox = np.zeros((d, h, w), np.uint32)
    idx = rng.integers(0, vox.size, 400000)
    vox.reshape(-1)[idx] = rng.integers(1, 2**32, idx.size, dtype=np.uint64).astype(np.uint32)
    want, _ = port.build_octree(vox)
    tree = pysvo.VoxelOctree.build_from_voxels(vox)
    assert np.array_equal(tree.words(), want)
    tree.close()


def _dims_from_center(center, depth):
    side = 1 << depth                                      # VoxelData::getCenter: dims * 0.5f / side
    return tuple(int(round(float(c) * 2 * side)) for c in center)


def test_extract_voxels_and_rebuild_small(pysvo, port):
    rng = np.random.default_rng(21)
    w, h, d = 90, 64, 48
    vox = random_volume(rng, w, h, d, 0.07)
    tree = pysvo.VoxelOctree.build_from_voxels(vox)
    xyz, vals = tree.extract_voxels()
    z, y, x = np.nonzero(vox)
    assert xyz.shape[0] == x.size
    got = np.zeros_like(vox)
    got[xyz[:, 2], xyz[:, 1], xyz[:, 0]] = vals
    assert np.array_equal(got, vox)
    # Morton order, x lowest: the builder's own order
    def spread(v):
        r = np.zeros(v.shape, np.uint64)
        for b in range(8):
            r |= ((v.astype(np.uint64) >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
        return r
    keys = spread(xyz[:, 0]) | (spread(xyz[:, 1]) << np.uint64(1)) | (spread(xyz[:, 2]) << np.uint64(2))
    assert (np.diff(keys.astype(np.int64)) > 0).all()
    again = tree.rebuild((w, h, d))
    assert np.array_equal(again.words(), tree.words()) and np.array_equal(again.center(), tree.center())
    again.close()
    tree.close()


@pytest.mark.parametrize("name", ["dragon", "ico256", "sdf512", "sdf2048", "ico8192"])
def test_rebuild_reproduces_reference_built_tree(pysvo, name):
    """Round trip at full size: trees made by the REFERENCE builder (the sample Dragon and the ico* scenes
    through PlyLoader, the sdf* scenes through raw .voxel files; up to 8192^3 = 400 M words) are taken apart
    into voxels and built again in HBM -- the result must be the same array, far words and all."""
    from tools import make_scenes
    if name == "dragon":
        words, center = pysvo.oct_read(ROOT / "tests" / "golden" / "XYZRGB-Dragon.oct")
    elif make_scenes.scene_available(name):
        words, center = make_scenes.load_scene(name)
    else:
        pytest.skip(f"scene {name} not cached")
    tree = pysvo.VoxelOctree(words=words, center=center)
    dims = _dims_from_center(center, tree.depth)
    again = tree.rebuild(dims)
    st = pysvo.VoxelOctree.last_build_stats()
    got = again.words()
    assert got.size == words.size and np.array_equal(got, words)
    assert np.array_equal(again.center(), center)
    assert st.words == words.size
    again.close()
    tree.close()


def test_headless_builder_mode(pysvo, port, tmp_path):
    """`svo_headless -builder in.voxel out.oct`: the reference's -builder mode for a raw volume
    (reference src/Main.cpp:313-319) through the C++ facade's VoxelOctree::fromVoxelFile + save."""
    import subprocess
    pkg = ROOT / "sparse-voxel-octrees_b200"
    subprocess.check_call(["make", "-C", str(pkg), "headless"], stdout=subprocess.DEVNULL)
    vox = random_volume(np.random.default_rng(31), 40, 36, 30, 0.1)
    raw, out = tmp_path / "in.voxel", tmp_path / "out.oct"
    write_voxel_file(raw, vox)
    res = subprocess.run([str(pkg / "svo_headless"), "-builder", str(raw), str(out)], capture_output=True, text=True,
                         timeout=300)
    assert res.returncode == 0, res.stderr
    want, wcenter = port.build_octree(vox)
    words, center = pysvo.oct_read(out)
    assert np.array_equal(words, want) and np.array_equal(center, wcenter)
    assert f"-> {want.size} words" in res.stdout
    res = subprocess.run([str(pkg / "svo_headless"), "-builder", str(tmp_path / "nope.voxel"), str(out)],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 1 and "cannot open" in res.stderr
    # the reference's long form with a mesh: -builder --resolution r --mode m in.ply out.oct
    ply, out2 = _ico_ply(tmp_path / "ico10.ply", 10), tmp_path / "mesh.oct"
    res = subprocess.run([str(pkg / "svo_headless"), "-builder", "--resolution", "96", "--mode", "0", str(ply), str(out2)],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    tree = pysvo.VoxelOctree.build_from_ply(ply, 96)
    words2, _ = pysvo.oct_read(out2)
    assert np.array_equal(words2, tree.words())
    tree.close()


def _ico_ply(path, freq):
    from tools import make_scenes
    assert make_scenes.gen_lib().svo_scene_icosphere_ply(str(path).encode(), freq, make_scenes.SEED) > 0
    return path


def _lowpoly_ply(path):
    from ply_meshes import write_lowpoly
    return write_lowpoly(path)


def test_build_from_low_poly_ply_large_triangles(pysvo, ref, tmp_path):
    """Triangles whose bounding boxes hold 10^5 .. 10^7 cells go through the block-per-triangle path of the voxeliser
    (largeTrianglesKernel): the tree must still be the reference builder's, word for word -- one cache block and several."""
    ply = _lowpoly_ply(tmp_path / "lowpoly.ply")
    threads = ref.hardware_threads()
    for res, mem in ((96, 1 << 30), (256, 1 << 30), (200, 1 << 22)):
        h = ref.tree_build_ply(ply, res, mem)
        want = ref.tree_words(h)
        ref.tree_destroy(h)
        tree = pysvo.VoxelOctree.build_from_ply(ply, res, mem_budget=mem, threads=threads)
        st = pysvo.VoxelOctree.last_voxelize_stats()
        got = tree.words()
        tree.close()
        assert st.large_triangles >= 8, (res, st.large_triangles)
        assert got.size == want.size and np.array_equal(got, want), (res, mem, list(st.sub_block), st.cache_block)


def test_build_from_ply_equals_reference(pysvo, port, ref, tmp_path):
    """Row f3 + f2 end to end: PLY -> voxels (GPU) -> tree (GPU) against the reference's own
    PlyLoader + VoxelData + VoxelOctree run (in-memory -builder path, Main.cpp:320-325) with the same pool size:
    identical node arrays for the benchmark's icosphere at three resolutions and for the ASCII / normals /
    colours / polygon variants; a small memory budget (several cache blocks) must give the same tree too."""
    from ply_meshes import write_variants
    threads = ref.hardware_threads()
    cases = [(_ico_ply(tmp_path / "ico6.ply", 6), 64, 1 << 30), (_ico_ply(tmp_path / "ico20.ply", 20), 128, 1 << 30),
             (_ico_ply(tmp_path / "ico50.ply", 50), 256, 1 << 30), (tmp_path / "ico20.ply", 128, 1 << 20)]
    cases += [(p, res, 1 << 30) for name, p in write_variants(tmp_path) if not name.startswith("be_") for res in (48, 128)]
    for ply, res, mem in cases:
        h = ref.tree_build_ply(ply, res, mem)
        want, wcenter = ref.tree_words(h), ref.tree_center(h)
        ref.tree_destroy(h)
        tree = pysvo.VoxelOctree.build_from_ply(ply, res, mem_budget=mem, threads=threads)
        st = pysvo.VoxelOctree.last_voxelize_stats()
        got = tree.words()
        assert got.size == want.size and np.array_equal(got, want), (ply.name, res, mem, list(st.sub_block), st.cache_block)
        assert np.array_equal(tree.center(), wcenter)
        assert st.voxels > 0 and st.cell_records >= st.voxels and st.triangles > 0
        tree.close()
    # the oracle's volume, for a pool size other than this host's
    for other in (1, 3, 64):
        vol, _ = port.voxelize_ply(tmp_path / "ico20.ply", 128, other)
        want, _ = port.build_octree(vol)
        tree = pysvo.VoxelOctree.build_from_ply(tmp_path / "ico20.ply", 128, threads=other)
        assert np.array_equal(tree.words(), want), other
        tree.close()
    with pytest.raises(pysvo.SvoError) as e:
        pysvo.VoxelOctree.build_from_ply(tmp_path / "missing.ply", 64)
    assert e.value.status == 2
