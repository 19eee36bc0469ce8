"""GPU: the multi-GPU handle (svo_multi_*: several devices, one process; reference Main.cpp:351-367 strip threads and
:217-219 frame barrier). On a one-GPU lease the replicas share device 0 -- the same code path (worker threads, tile
interleave, cross-stream / cross-"device" events, per-device host leg); with two or more GPUs the real peer path runs
as well. Frames must equal the single-device frames word for word."""
import numpy as np
import pytest

from conftest import DRAGON

pytestmark = pytest.mark.gpu


def _device_lists(pysvo):
    lists = [(0,), (0, 0), (0, 0, 0)]
    n = pysvo.device_count()
    if n >= 2:
        lists.append((0, 1))
    if n >= 4:
        lists.append((0, 1, 2, 3))
    return lists


@pytest.fixture(scope="module")
def cams(pysvo):
    return [pysvo.orbit_camera(10.0 + k, 30.0 * k, 1.0 - 0.05 * k) for k in range(9)]


@pytest.fixture(scope="module")
def single_frames(pysvo, gpu_dragon, cams):
    W, H, S = 648, 360, 8
    return W, H, S, [gpu_dragon.render_frame(c, W, H, strips=S, flavour=pysvo.FLAVOUR_VALIDATION)[0] for c in cams]


def test_multi_frames_equal_single_device(pysvo, cams, single_frames):
    W, H, S, want = single_frames
    for devices in _device_lists(pysvo):
        m = pysvo.MultiOctree(DRAGON, devices=devices)
        assert m.n_devices == len(devices)
        # one frame, pageable host memory
        got, st = m.render_frame(cams[0], W, H, strips=S, flavour=pysvo.FLAVOUR_VALIDATION)
        assert np.array_equal(got, want[0]), f"devices {devices}: {int((got != want[0]).sum())} pixels differ"
        assert st.fine_rays == int((want[0] != 0).sum()) and st.coarse_rays == pysvo.frame_layout(W, H, S).corners
        # a sequence into a ring of page-locked host frames, every frame seen by the callback
        ring = [pysvo.PinnedArray((H, W), np.uint32) for _ in range(3)]
        seen = {}
        stats = m.render_sequence(cams, W, H, strips=S, flavour=pysvo.FLAVOUR_VALIDATION, output=pysvo.OUTPUT_HOST,
                                  host_frames=[r.array for r in ring], on_frame=lambda k, a: seen.__setitem__(k, a.copy()))
        assert sorted(seen) == list(range(len(cams))) and stats.frames == len(cams)
        for k in range(len(cams)):
            assert np.array_equal(seen[k], want[k]), f"devices {devices}, host frame {k}"
        assert stats.fine_rays == sum(int((w != 0).sum()) for w in want)
        assert stats.lanes == 3 and stats.kernel_launches >= 3 * len(cams) * len(devices)
        # the same sequence gathered in devices[0]'s HBM (fine passes store into it; peer stores with real devices)
        stats = m.render_sequence(cams, W, H, strips=S, flavour=pysvo.FLAVOUR_VALIDATION, output=pysvo.OUTPUT_DEVICE)
        assert stats.device_ms > 0 and stats.fine_rays == sum(int((w != 0).sum()) for w in want)
        for back in range(4):
            assert np.array_equal(m.device_frame(W, H, back), want[len(cams) - 1 - back]), f"devices {devices}, back {back}"
        # the same host sequence as (grey, alpha) byte pairs: half the bytes, expanded on the host to the same words
        ring16 = [pysvo.PinnedArray((H, W), np.uint16) for _ in range(4)]
        seen16 = {}
        stats = m.render_sequence(cams, W, H, strips=S, flavour=pysvo.FLAVOUR_VALIDATION, output=pysvo.OUTPUT_HOST,
                                  host_frames=[r.array for r in ring16], on_frame=lambda k, a: seen16.__setitem__(k, a.copy()),
                                  pixel_format=pysvo.PIXELS_GREY8A8)
        for k in range(len(cams)):
            assert np.array_equal(pysvo.expand_grey8a(seen16[k]), want[k]), f"devices {devices}, packed host frame {k}"
        one16, _ = m.render_frame(cams[2], W, H, strips=S, flavour=pysvo.FLAVOUR_VALIDATION, pixel_format=pysvo.PIXELS_GREY8A8)
        assert one16.dtype == np.uint16 and np.array_equal(pysvo.expand_grey8a(one16), want[2])
        with pytest.raises(pysvo.SvoError):     # frames that stay on the GPU are RGBA words
            m.render_sequence(cams[:2], W, H, strips=S, output=pysvo.OUTPUT_DEVICE, pixel_format=pysvo.PIXELS_GREY8A8)
        # a second sequence on the same handle, other size (buffers are re-used / re-allocated)
        one, _ = m.render_frame(cams[3], 200, 120, strips=4, flavour=pysvo.FLAVOUR_VALIDATION)
        t = m.tree(0)
        ref_small, _, _ = t.render_frame(cams[3], 200, 120, strips=4, flavour=pysvo.FLAVOUR_VALIDATION)
        assert np.array_equal(one, ref_small)
        m.close()


def test_two_handles_render_concurrently(pysvo, dragon_words, cams, single_frames):
    """The stripe width belongs to the sequence, not to the process: a handle gathering frames in HBM (stripes of 4 tile
    columns) and one shipping host frames (here 27 columns: svo_sequence_stats.tile_run) render at the same time from two
    threads, and a caller-set default (svo_frame_set_tile_run) is neither read nor reset by them."""
    import threading
    W, H, S, want = single_frames
    words, center = dragon_words
    a = pysvo.MultiOctree(words=words, center=center, devices=(0, 0))
    b = pysvo.MultiOctree(words=words, center=center, devices=(0, 0, 0))
    ring = [pysvo.PinnedArray((H, W), np.uint32) for _ in range(3)]
    errors, runs = [], {}
    pysvo.frame_set_tile_run(7)
    try:
        def device_side():
            try:
                for _ in range(6):
                    st = a.render_sequence(cams, W, H, strips=S, flavour=pysvo.FLAVOUR_VALIDATION, output=pysvo.OUTPUT_DEVICE)
                    runs["device"] = st.tile_run
                    for back in range(3):
                        if not np.array_equal(a.device_frame(W, H, back), want[len(cams) - 1 - back]):
                            errors.append(f"device frame {back} back differs")
            except Exception as e:      # noqa: BLE001 -- reported by the main thread
                errors.append(repr(e))

        def host_side():
            try:
                for _ in range(6):
                    seen = {}
                    st = b.render_sequence(cams, W, H, strips=S, flavour=pysvo.FLAVOUR_VALIDATION, output=pysvo.OUTPUT_HOST,
                                           host_frames=[r.array for r in ring],
                                           on_frame=lambda k, arr: seen.__setitem__(k, arr.copy()))
                    runs["host"] = st.tile_run
                    for k in range(len(cams)):
                        if not np.array_equal(seen[k], want[k]):
                            errors.append(f"host frame {k} differs")
            except Exception as e:      # noqa: BLE001
                errors.append(repr(e))

        threads = [threading.Thread(target=device_side), threading.Thread(target=host_side)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors[:5]
        assert runs["device"] == 4 and runs["host"] == 27        # 81 tile columns over three devices
        # the single-device interleave still follows the caller's setting
        lay = pysvo.frame_layout(W, H, S)
        assert [pysvo.tile_owner(W, H, S, t, 2) for t in (0, 6, 7, 13, 14)] == [0, 0, 1, 1, 0] and lay.tile_cols == 81
    finally:
        pysvo.frame_set_tile_run(0)
        a.close()
        b.close()


def test_multi_batch_shards_rays(pysvo, port, dragon_words, cams):
    from oracle.pyoracle import pixel_rays
    words, center = dragon_words
    f = port.frame_constants(np.array(cams[1].model[:], np.float32), np.array(cams[1].view[:], np.float32), center, 320, 200, 4)
    o, d = pixel_rays(f)
    want = port.raymarch_batch(words, o, d, 0.0, t_sentinel=1e10)
    devices = (0, 1) if pysvo.device_count() >= 2 else (0, 0, 0)
    m = pysvo.MultiOctree(words=words, center=center, devices=devices)
    got = m.raymarch_batch(o, d, 0.0, pysvo.FLAVOUR_VALIDATION)
    m.close()
    assert np.array_equal(got["hit"], want["hit"])
    assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    hit = want["hit"] > 0
    assert np.array_equal(got["voxel"][hit], want["voxel"][hit]) and np.array_equal(got["normal"][hit], want["normal"][hit])


def test_multi_errors(pysvo, dragon_words):
    words, center = dragon_words
    with pytest.raises(pysvo.SvoError) as e:
        pysvo.MultiOctree(words=words, center=center, devices=(0, 99))
    assert e.value.status == 1
    m = pysvo.MultiOctree(words=words, center=center, devices=(0,))
    with pytest.raises(pysvo.SvoError):
        m.render_sequence([pysvo.orbit_camera(0, 0, 1)], 64, 64, output=pysvo.OUTPUT_HOST, host_frames=[])
    m.close()
