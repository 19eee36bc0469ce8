"""CPU: pins the plain-C oracle (oracle/svo_oracle.c) to the reference.

Two anchors: (1) the committed golden vectors, minted from the reference's own object code by
tests/golden/make_golden.py; (2) that object code itself (oracle/_ref), when it is present.
"""
import hashlib
import json

import numpy as np
import pytest

from conftest import CAMERAS, DRAGON, GOLDEN
from oracle.pyoracle import pixel_rays

T_MISS = np.float32(1e10)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def pins():
    return json.loads((GOLDEN / "dragon_pins.json").read_text())


@pytest.fixture(scope="module")
def small():
    return np.load(GOLDEN / "dragon_small.npz")


def test_tree_walk_matches_survey_pins(port, dragon_words, pins):
    words, center = dragon_words
    assert words.size == pins["n_words"] == 119887
    assert int(words[0]) == pins["root_word"] == 0x00052323
    assert sha(words) == pins["words_sha256"]
    assert [float(x) for x in center] == pins["center"] == [0.5, 0.2265625, 0.33203125]
    st = port.tree_walk(words)
    tw = pins["tree_walk"]
    assert (st.descriptors, st.leaves, st.far_words, st.far_blocks, st.depth) == (29156, 90707, 24, 4, 8)
    assert list(st.per_level)[:8] == tw["per_level"] == [1, 3, 15, 68, 289, 1237, 5352, 22191]
    assert st.descriptors + st.leaves + st.far_words == words.size
    assert st.max_index == words.size - 1


@pytest.mark.parametrize("ci", [0, 1, 3])
def test_port_small_frames_and_rays_equal_golden(port, dragon_words, small, ci):
    words, center = dragon_words
    k = f"cam{ci}_"
    f = port.frame_constants(small[k + "model"], small[k + "view"], center, 160, 90, 4)
    rgba, depth, cc, cf = port.render_frame(words, f, threads=2, want_depth=True)
    assert np.array_equal(rgba, small[k + "rgba"])
    assert np.array_equal(depth.view(np.uint32), small[k + "depth"].view(np.uint32))
    o, d = pixel_rays(f)
    assert np.array_equal(o.view(np.uint32), small[k + "o"].view(np.uint32))
    assert np.array_equal(d.view(np.uint32), small[k + "d"].view(np.uint32))
    res = port.raymarch_batch(words, o, d, 0.0, threads=2, t_sentinel=float(T_MISS))
    assert np.array_equal(res["hit"] > 0, small[k + "hit"] > 0)
    assert np.array_equal(res["t"].view(np.uint32), small[k + "t"].view(np.uint32))
    assert np.array_equal(res["normal"], small[k + "normal"])
    lod = port.raymarch_batch(words, o, d, float(small[k + "coarse_scale"]), threads=2, t_sentinel=float(T_MISS))
    assert np.array_equal(lod["hit"] > 0, small[k + "lod_hit"] > 0)
    assert np.array_equal(lod["t"].view(np.uint32), small[k + "lod_t"].view(np.uint32))


@pytest.mark.parametrize("ci", range(len(CAMERAS)))
def test_port_full_resolution_hashes(port, dragon_words, pins, ci):
    words, center = dragon_words
    cam = pins["cameras"][ci]
    model, view = np.array(cam["model"], np.float32), np.array(cam["view"], np.float32)
    m2, v2 = port.orbit_camera(*cam["pitch_yaw_radius"])
    assert np.array_equal(m2.view(np.uint32), model.view(np.uint32))
    assert np.array_equal(v2.view(np.uint32), view.view(np.uint32))
    for strips in (16, 8):
        f = port.frame_constants(model, view, center, 1280, 720, strips)
        rgba, depth, cc, cf = port.render_frame(words, f, want_depth=True)
        g = cam[f"frame_1280x720x{strips}"]
        assert sha(rgba) == g["rgba_sha256"]
        assert sha(depth) == g["depth_sha256"]
    f = port.frame_constants(model, view, center, 1280, 720, 16)
    o, d = pixel_rays(f)
    b = cam["batch_1280x720"]
    assert sha(o) == b["rays_o_sha256"] and sha(d) == b["rays_d_sha256"]
    res = port.raymarch_batch(words, o, d, 0.0, t_sentinel=float(T_MISS))
    hit = (res["hit"] > 0).astype(np.uint8)
    assert int(hit.sum()) == b["hits"]
    assert sha(hit) == b["hit_sha256"] and sha(res["t"]) == b["t_sha256"] and sha(res["normal"]) == b["normal_sha256"]
    lod = port.raymarch_batch(words, o, d, f.coarse_scale, t_sentinel=float(T_MISS))
    bl = cam["batch_lod_1280x720"]
    assert sha((lod["hit"] > 0).astype(np.uint8)) == bl["hit_sha256"] and sha(lod["t"]) == bl["t_sha256"]


def test_survey_known_answers(port, dragon_words, pins):
    """default camera raw cast and 16-strip frame."""
    b = pins["cameras"][0]["batch_1280x720"]
    assert b["hits"] == 323762
    assert abs(b["sum_t"] - 303634.524310) < 1e-3
    assert b["xor_normals"] == 0xE59DE086
    fr = pins["cameras"][0]["frame_1280x720x16"]
    assert (fr["lit"], fr["written"], fr["coarse_hits"]) == (322364, 376504, 6793)
    words, center = dragon_words
    m, v = port.orbit_camera(0, 0, 1.0)
    f = port.frame_constants(m, v, center, 1280, 720, 16)
    _, _, cc, cf = port.render_frame(words, f)
    assert (cc.rays, cf.rays) == (18032, 376504)
    assert cc.lod_exits == 6793
    # node traffic per ray, frame mode: 66.7 B over coarse + fine
    total = 4.0 * (cc.words + cf.words) / (cc.rays + cf.rays)
    assert abs(total - 66.7) < 0.1


def test_port_matches_reference_object_code(port, ref, dragon_words):
    """Live cross-check against oracle/_ref on cameras and sizes that are NOT in the golden set."""
    words, center = dragon_words
    h = ref.tree_from_words(words, center)
    rng = np.random.default_rng(7)
    for _ in range(3):
        cam = (float(rng.uniform(-80, 80)), float(rng.uniform(0, 360)), float(rng.uniform(0.2, 2.0)))
        model, view = ref.orbit_camera(*cam)
        W, H, S = int(rng.integers(40, 300)), int(rng.integers(30, 200)), int(rng.integers(1, 7))
        f = port.frame_constants(model, view, center, W, H, S)
        rgba_r, depth_r, _ = ref.render_frames(h, W, H, S, model, view, threads=2, want_depth=True)
        rgba_p, depth_p, _, _ = port.render_frame(words, f, threads=2, want_depth=True)
        assert np.array_equal(rgba_r, rgba_p), (cam, W, H, S)
        assert np.array_equal(depth_r.view(np.uint32), depth_p.view(np.uint32))
    # random rays through the cube, including axis-parallel and tiny direction components
    n = 20000
    o = rng.uniform(0.0, 3.0, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[::7, 0] = 0.0
    d[::11, 1] = np.float32(-5e-5)
    d[::13, 2] = np.float32(1e-4)
    for rs in (0.0, 0.01, 0.2):
        hit, t, nrm, _ = ref.raymarch_batch(h, o, d, rs, normal_sentinel=5, t_sentinel=-3.0)
        res = port.raymarch_batch(words, o, d, rs, threads=2, normal_sentinel=5, t_sentinel=-3.0)
        assert np.array_equal(res["hit"] > 0, hit > 0)
        assert np.array_equal(res["t"].view(np.uint32), t.view(np.uint32))
        assert np.array_equal(res["normal"], nrm)
    ref.tree_destroy(h)


def test_port_shading_helpers_match_reference(port, ref):
    rng = np.random.default_rng(3)
    for w in rng.integers(0, 2**32, 200, dtype=np.uint64):
        w = int(w) & ~0x60000000 | (int(rng.integers(0, 3)) << 29)   # face 3 is not a valid encoding
        n_r, s_r = ref.decompress_material(w)
        n_p, s_p = port.decompress_material(w)
        assert np.array_equal(n_r.view(np.uint32), n_p.view(np.uint32)) and s_r == s_p
    for x in rng.uniform(1e-6, 1e6, 200):
        assert ref.inv_sqrt(x) == port.inv_sqrt(x)


# ---- row f2: the builder restatement ------------------------------------------------------------

def _volume(seed, w, h, d, p):
    rng = np.random.default_rng(seed)
    vox = ((rng.random((d, h, w)) < p) * rng.integers(1, 2**32, (d, h, w), dtype=np.uint64)).astype(np.uint32)
    vox[0, 0, 0] = 1
    return vox


BUILD_CASES = [(1, 8, 8, 8, 0.3), (2, 20, 12, 6, 0.2), (3, 33, 17, 10, 0.1), (4, 7, 9, 5, 0.5), (5, 64, 64, 64, 1.0),
               (6, 100, 60, 30, 0.03), (7, 2, 2, 2, 1.0)]


def test_builder_port_equals_golden(port):
    """svo_oracle_build_octree against hashes minted from the reference builder (make_build_golden.py)."""
    build_pins = json.loads((GOLDEN / "build_pins.json").read_text())
    assert len(build_pins["builder"]) == len(BUILD_CASES)
    for case, pin in zip(BUILD_CASES, build_pins["builder"]):
        assert list(case) == pin["case"]
        words, center = port.build_octree(_volume(*case))
        assert words.size == pin["n_words"] and sha(words) == pin["words_sha256"], case
        assert [float(x) for x in center] == pin["center"]


def test_builder_port_equals_reference_object_code(port, ref, tmp_path):
    """VoxelData(path, mem) + VoxelOctree(VoxelData*) (VoxelOctree.cpp:125-205) on raw .voxel files, with
    the whole volume in one cache block and with 16^3 cache blocks (even depths: see svo_build.cu)."""
    import struct
    for case in BUILD_CASES + [(8, 48, 40, 36, 0.15)]:
        vox = _volume(*case)
        d, h, w = vox.shape
        path = tmp_path / "v.voxel"
        with open(path, "wb") as fp:
            fp.write(struct.pack("<iii", w, h, d))
            fp.write(vox.tobytes())
        got, center = port.build_octree(vox)
        for mem in ((1 << 30,) if d % 2 else (1 << 30, 1 << 17)):
            hnd = ref.tree_build_voxel_file(path, mem)
            assert np.array_equal(ref.tree_words_view(hnd), got), (case, mem)
            assert np.array_equal(ref.tree_center(hnd), center)
            ref.tree_destroy(hnd)


def test_port_preview_stride_equals_reference(port, ref, dragon_words):
    """renderTile with stride 3 (renderHalfSize, Main.cpp:101-106, 161): restatement == reference object code."""
    words, center = dragon_words
    h = ref.tree_from_words(words, center)
    for cam in [(0.0, 0.0, 1.0), (20.0, 135.0, 0.5)]:
        for (W, H, S) in [(320, 180, 4), (333, 187, 5)]:
            model, view = ref.orbit_camera(*cam)
            theirs, _, _ = ref.render_frames(h, W, H, S, model, view, half_size=True)
            f = port.frame_constants(model, view, center, W, H, S)
            ours, _, _, cf = port.render_frame(words, f, pixel_stride=3)
            full, _, _, cf_full = port.render_frame(words, f)
            assert np.array_equal(ours, theirs)
            assert cf.rays < cf_full.rays / 6 and not np.array_equal(ours, full)
    ref.tree_destroy(h)


# ---- row f3: the voxeliser restatement (oracle/svo_oracle_ply.c) ----------------------------------

def _ico_ply(path, freq):
    from tools import make_scenes
    assert make_scenes.gen_lib().svo_scene_icosphere_ply(str(path).encode(), freq, make_scenes.SEED) > 0
    return path


def voxelise_cases(directory):
    """(name, ply path, resolution) of the meshes the f3 pins are minted on (tests/golden/make_voxelise_golden.py)."""
    from ply_meshes import write_variants
    cases = [("ico6", _ico_ply(directory / "ico6.ply", 6), 64), ("ico20", _ico_ply(directory / "ico20.ply", 20), 128)]
    cases += [(name, p, res) for name, p in write_variants(directory) if not name.startswith("be_") for res in (48, 128)]
    return cases


def test_voxeliser_port_equals_golden(port, tmp_path):
    """The same pin without a reference build: the node arrays the reference produced for these meshes when the vectors
    were minted (tests/golden/voxelise_pins.json, with the thread-pool size of that run), against the restatement run
    with that pool size."""
    import hashlib
    pins = json.loads((GOLDEN / "voxelise_pins.json").read_text())
    by_key = {(c["name"], c["resolution"]): c for c in pins["cases"]}
    assert len(by_key) >= 6
    for name, ply, res in voxelise_cases(tmp_path):
        pin = by_key[(name, res)]
        assert hashlib.sha256(ply.read_bytes()).hexdigest() == pin["ply_sha256"], name       # same mesh file as minted
        vol, _ = port.voxelize_ply(ply, res, pins["pool_threads"])
        words, center = port.build_octree(vol)
        assert words.size == pin["words"], (name, res)
        assert hashlib.sha256(np.ascontiguousarray(words).tobytes()).hexdigest() == pin["sha256"], (name, res)
        assert np.array_equal(center, np.array(pin["center"], np.float32)), (name, res)


def test_block_lists_never_alias_another_sub_block(port, tmp_path):
    """The one reference behaviour the GPU voxeliser does not mirror (csrc/svo_voxelize.cu header): a triangle listed
    for a sub-block whose x or y index lies outside the sub-block grid lands, through the flat index
    x + gridW*(y + gridH*z) (PlyLoader.cpp:263), in a block of the next row as well. Counted in the restatement's
    iterateOverlappingBlocks over the benchmark mesh, the container variants and nine triangles that span the whole
    volume and touch its faces, for one and several cache blocks and pool sizes 1 .. 64: nothing is ever listed there,
    and the index ranges never even reach past the grid -- the loader rescales every mesh into [0, 1]^3, pointToGrid
    maps 1 to sideLength - 3 (the member _sideLength is already sideLength - 2, and pointToGrid subtracts 2 again,
    PlyLoader.cpp:228-233, :415), and the grid covers the volume of sideLength cells."""
    from ply_meshes import write_lowpoly, write_variants
    meshes = [_ico_ply(tmp_path / "ico20.ply", 20), write_lowpoly(tmp_path / "lowpoly.ply")]
    meshes += [p for name, p in write_variants(tmp_path) if name.startswith("le_")]
    reached = 0
    for ply in meshes:
        for res, edges in ((48, (0, 16)), (64, (0, 32)), (200, (0, 64)), (256, (0, 128, 32)), (1000, (0, 256)),
                           (2048, (0, 512)), (8192, (512,))):
            for edge in edges:
                for threads in (1, 2, 3, 16, 64):
                    listed, candidates, grid, real = port.block_list_aliases(ply, res, edge, threads)
                    assert listed == 0, (ply.name, res, edge, threads, grid, real)
                    assert all(g >= r for g, r in zip(grid, real)), (grid, real)
                    reached += candidates
    assert reached == 0


def test_voxeliser_port_equals_reference(port, ref, tmp_path):
    """PlyLoader + VoxelData(loader, res, mem) + VoxelOctree (reference src/Main.cpp:320-325): the volume the
    restatement voxelises must hold exactly the voxels of the tree the reference builds from the same PLY (same
    thread-pool size: the result depends on it), and the builder restatement must turn it into the same node
    array -- for the benchmark's icosphere and for ASCII / big-endian / normals / colours / polygon variants."""
    from ply_meshes import write_variants
    threads = ref.hardware_threads()
    cases = [(_ico_ply(tmp_path / "ico6.ply", 6), 64), (_ico_ply(tmp_path / "ico20.ply", 20), 128)]
    # (not the big-endian variant: the reference's plyfile moves raw big-endian floats through a double, which
    # quiets the byte-swapped patterns that happen to be signalling NaNs and so alters a few coordinates)
    cases += [(p, res) for name, p in write_variants(tmp_path) if not name.startswith("be_") for res in (48, 128)]
    for ply, res in cases:
        h = ref.tree_build_ply(ply, res, 1 << 30)
        words, center = ref.tree_words(h), ref.tree_center(h)
        ref.tree_destroy(h)
        vol, ntri = port.voxelize_ply(ply, res, threads)
        d, hh, w = vol.shape
        side = 1
        while side < max(w, hh, d):
            side *= 2
        assert np.array_equal(vol, port.tree_to_volume(words, side, (w, hh, d))), (ply.name, res)
        built, bcenter = port.build_octree(vol)
        assert np.array_equal(built, words) and np.array_equal(bcenter, center), (ply.name, res)
        assert ntri > 0 and int(np.count_nonzero(vol)) > 100
