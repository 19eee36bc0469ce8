"""GPU: parity at BASELINE.json's FULL sizes, one test per config.

    configs[1]  2048^3 SDF scene, 1920x1080, 16 strips             test_c2_*
    configs[2]  8192^3 displaced icosphere, 3840x2160, 16 strips   test_c3_*  (incl. the fly-through's nearest / farthest camera)
    configs[3]  16-spp ambient-occlusion rays on the 2048^3 scene  test_c4_*  (2 M-ray sample of the full set)
    configs[4]  fly-through sweep on the 8192^3 tree, 4K           test_c5_*  (10 of the 100 frames)

The checker is the plain-C oracle (oracle/svo_oracle.c, pinned to the reference's object code by
tests/test_oracle_pins.py) on the same node array; the product is libsvo_b200.so through its C ABI.
VALIDATION flavour: every pixel word, coarse depth, hit code, voxel id and t identical. FAST: >= 99.99 % of
the pixels identical (BASELINE.json north_star). The scenes ride in the snapshot (scenes/_cache: sdf2048.oct,
ico8192.words.xz); a scene that is absent AND cannot be rebuilt skips its tests, which GPUTEST then shows.
"""
import numpy as np
import pytest

from oracle.pyoracle import pixel_rays

pytestmark = pytest.mark.gpu
T_MISS = np.float32(1e10)
STRIPS = 16


def _load(pysvo, name):
    from tools import make_scenes
    if not make_scenes.scene_available(name):
        pytest.skip(f"scenes/_cache/{name} not in the snapshot")
    words, center = make_scenes.load_scene(name)
    tree = pysvo.VoxelOctree(words=words, center=center)
    return tree, words, center


@pytest.fixture(scope="module")
def ico8192(pysvo):
    tree, words, center = _load(pysvo, "ico8192")
    assert tree.depth == 13 and words.size == 400874482
    yield tree, words, center
    tree.close()


@pytest.fixture(scope="module")
def sdf2048(pysvo):
    tree, words, center = _load(pysvo, "sdf2048")
    assert tree.depth == 11
    yield tree, words, center
    tree.close()


def flythrough_camera(j, n=100):
    """bench.py's c5 path: radius geometric 2.0 -> 0.05, yaw 0 -> 180, pitch 20."""
    return (20.0, 180.0 * j / (n - 1), 2.0 * (0.05 / 2.0) ** (j / (n - 1)))


def _check_frame(pysvo, port, scene, cam_args, W, H):
    tree, words, center = scene
    cam = pysvo.orbit_camera(*cam_args)
    f = port.frame_constants(np.array(cam.model[:], np.float32), np.array(cam.view[:], np.float32), center, W, H, STRIPS)
    want, wdepth, cc, cf = port.render_frame(words, f, want_depth=True)
    got, depth, st = tree.render_frame(cam, W, H, strips=STRIPS, flavour=pysvo.FLAVOUR_VALIDATION, want_depth=True)
    assert np.array_equal(depth.view(np.uint32), wdepth.view(np.uint32)), f"{cam_args}: coarse depth differs"
    bad = int((got != want).sum())
    assert bad == 0, f"{cam_args}: {bad} of {W * H} pixels differ in the validation flavour"
    assert (st.coarse_rays, st.fine_rays) == (cc.rays, cf.rays)
    fast, _, _ = tree.render_frame(cam, W, H, strips=STRIPS, flavour=pysvo.FLAVOUR_FAST)
    same = float((fast == want).mean())
    assert same >= 0.9999, f"{cam_args}: only {same:.6f} of the pixels identical in the fast flavour"
    return cf


@pytest.mark.parametrize("cam_args", [(20.0, 40.0, 0.9), (20.0, 130.0, 0.9), flythrough_camera(0), flythrough_camera(99)],
                         ids=["orbit0", "orbit25", "fly_far_r2.0", "fly_near_r0.05"])
def test_c3_ico8192_4k_frames(pysvo, port, ico8192, cam_args):
    cf = _check_frame(pysvo, port, ico8192, cam_args, 3840, 2160)
    assert cf.rays > 1_000_000


@pytest.mark.parametrize("j", [0, 11, 22, 33, 44, 55, 66, 77, 88, 99])
def test_c5_flythrough_4k_frames(pysvo, port, ico8192, j):
    _check_frame(pysvo, port, ico8192, flythrough_camera(j), 3840, 2160)


@pytest.mark.parametrize("cam_args", [(20.0, 40.0, 0.9), (20.0, 220.0, 0.9)], ids=["orbit0", "orbit50"])
def test_c2_sdf2048_1080p_frames(pysvo, port, sdf2048, cam_args):
    cf = _check_frame(pysvo, port, sdf2048, cam_args, 1920, 1080)
    assert cf.far_fetches > 0.1 * cf.desc_fetches     # the far-word path is exercised


def test_c4_ao_sdf2048_sample(pysvo, port, sdf2048):
    """bench.py's c4 ray set (16 hemisphere directions per primary hit of the 1920x1080 frame); a 2 M-ray sample
    spread over the whole set, submission order and direction-binned order, both flavours."""
    from tools.ao_rays import ao_rays
    tree, words, center = sdf2048
    cam = pysvo.orbit_camera(20.0, 40.0, 0.9)
    f = port.frame_constants(np.array(cam.model[:], np.float32), np.array(cam.view[:], np.float32), center, 1920, 1080, STRIPS)
    o, d = pixel_rays(f)
    prim = tree.raymarch_batch(o, d, 0.0, pysvo.FLAVOUR_VALIDATION)
    want_prim = port.raymarch_batch(words, o, d, 0.0, t_sentinel=float(T_MISS))
    assert np.array_equal(prim["hit"], want_prim["hit"])
    assert np.array_equal(prim["t"].view(np.uint32), want_prim["t"].view(np.uint32))
    phit = want_prim["hit"] > 0
    for key in ("normal", "voxel"):
        assert np.array_equal(prim[key][phit], want_prim[key][phit]), key
    ao_o, ao_d, _ = ao_rays(o, d, prim["t"], prim["normal"], prim["hit"] == 1, spp=16)
    n = ao_o.shape[0]
    assert n > 10_000_000
    # 32 contiguous runs of 65,536 rays spread over the set: warps of the sample hold the same rays as in the full batch
    starts = np.linspace(0, n - 65536, 32).astype(np.int64) // 32 * 32
    idx = (starts[:, None] + np.arange(65536)[None, :]).reshape(-1)
    so, sd = np.ascontiguousarray(ao_o[idx]), np.ascontiguousarray(ao_d[idx])
    want = port.raymarch_batch(words, so, sd, 0.0, t_sentinel=float(T_MISS))
    hit = want["hit"] > 0
    assert 0.05 < hit.mean() < 0.95
    for flavour in (pysvo.FLAVOUR_VALIDATION, pysvo.FLAVOUR_VALIDATION | pysvo.BATCH_COHERENCE_ORDER,
                    pysvo.FLAVOUR_VALIDATION | pysvo.BATCH_COHERENCE_ORDER | pysvo.BATCH_LANE_REFILL):
        got = tree.raymarch_batch(so, sd, 0.0, flavour)
        assert np.array_equal(got["hit"], want["hit"])
        assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
        assert np.array_equal(got["voxel"][hit], want["voxel"][hit])
        assert np.array_equal(got["normal"][hit], want["normal"][hit])
    fast = tree.raymarch_batch(so, sd, 0.0, pysvo.FLAVOUR_FAST | pysvo.BATCH_COHERENCE_ORDER | pysvo.BATCH_LANE_REFILL)
    same = (fast["hit"] == want["hit"]) & (fast["voxel"] == np.where(hit, want["voxel"], pysvo.VOXEL_NONE))
    assert same.mean() >= 0.9999
