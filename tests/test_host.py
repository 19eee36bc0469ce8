"""CPU: the C-ABI library loads, exports what include/svo_b200.h declares, and its host-only parts
(.oct I/O, camera constants, argument checking) behave. No device compute is attempted here."""
import os
import re
import struct

import numpy as np
import pytest

from conftest import CAMERAS, DRAGON, ROOT


def test_library_exports_every_declared_symbol(pysvo):
    header = (ROOT / "include" / "svo_b200.h").read_text()
    declared = set(re.findall(r"SVO_API\s+[\w\s\*]+?\b(svo_\w+)\s*\(", header))
    assert len(declared) >= 28
    L = pysvo.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in svo_b200.h but not exported"
    assert declared == set(L._svo_symbols), "pysvo binds a different symbol set than the header declares"
    assert L.svo_abi_version() == 2


def test_no_device_fails_loudly(pysvo, dragon_words):
    if pysvo.device_count() > 0:
        pytest.skip("a CUDA device is present")
    words, center = dragon_words
    with pytest.raises(pysvo.SvoError) as e:
        pysvo.VoxelOctree(words=words, center=center)
    assert e.value.status == 6 and "no CPU fallback" in str(e.value)


def test_oct_read_dragon(pysvo, dragon_words):
    words, center = dragon_words
    assert words.dtype == np.uint32 and words.size == 119887
    assert list(center) == [0.5, 0.2265625, 0.33203125]
    assert int(words[0]) == 0x00052323


@pytest.mark.parametrize("compress", [True, False])
def test_oct_roundtrip_ours(pysvo, dragon_words, tmp_path, compress):
    words, center = dragon_words
    p = tmp_path / "a.oct"
    pysvo.oct_write(p, words, center, compress=compress)
    w2, c2 = pysvo.oct_read(p)
    assert np.array_equal(words, w2) and np.array_equal(center, c2)
    raw = p.read_bytes()
    assert struct.unpack_from("<3fQ", raw) == (*[float(x) for x in center], words.size)
    if compress:
        assert len(raw) < words.nbytes + 20


def test_oct_interop_with_reference_loader_and_saver(pysvo, ref, dragon_words, tmp_path):
    words, center = dragon_words
    for compress in (True, False):
        p = tmp_path / f"ours_{int(compress)}.oct"
        pysvo.oct_write(p, words, center, compress=compress)
        h = ref.tree_load(p)   # LZ4_decompress_fast_continue, VoxelOctree.cpp:57-90
        assert np.array_equal(ref.tree_words(h), words) and np.array_equal(ref.tree_center(h), center)
        ref.tree_destroy(h)
    h = ref.tree_from_words(words, center)
    q = tmp_path / "theirs.oct"
    ref.tree_save(h, q)        # LZ4_compress_continue, VoxelOctree.cpp:92-123
    ref.tree_destroy(h)
    w2, c2 = pysvo.oct_read(q)
    assert np.array_equal(w2, words) and np.array_equal(c2, center)


def _synthetic_words(rng, n):
    """Compressible, octree-like word soup: runs, repeats at assorted distances, noise."""
    base = rng.integers(0, 2**32, n // 8 + 1, dtype=np.uint64).astype(np.uint32)
    out = np.repeat(base, 8)[:n].copy()
    idx = rng.integers(0, n, n // 5)
    out[idx] = rng.integers(0, 2**32, idx.size, dtype=np.uint64).astype(np.uint32)
    out[0] = (1 << 18) | 0x0100   # plausible root: one leaf child at offset 1
    return out


@pytest.mark.parametrize("n_words", [2, 3, 4, 5, 17, 1000, (64 << 20) // 4 - 1, (64 << 20) // 4, (64 << 20) // 4 + 3,
                                     2 * (64 << 20) // 4 + 12345])
def test_oct_roundtrip_slice_boundaries(pysvo, tmp_path, n_words):
    """Slices are 64 MiB (VoxelOctree.cpp:55); matches may reach into the previous slice."""
    rng = np.random.default_rng(n_words)
    words = _synthetic_words(rng, n_words)
    center = np.array([0.5, 0.25, 0.125], np.float32)
    p = tmp_path / "s.oct"
    pysvo.oct_write(p, words, center, compress=True)
    w2, c2 = pysvo.oct_read(p)
    assert np.array_equal(words, w2) and np.array_equal(center, c2)


def test_oct_multi_slice_interop_with_reference(pysvo, ref, tmp_path):
    n_words = (64 << 20) // 4 + 4099
    rng = np.random.default_rng(5)
    words = _synthetic_words(rng, n_words)
    # make the start of slice 1 repeat the end of slice 0 so that our encoder emits cross-slice matches
    words[(64 << 20) // 4:(64 << 20) // 4 + 2000] = words[(64 << 20) // 4 - 2000:(64 << 20) // 4]
    center = np.array([0.5, 0.5, 0.5], np.float32)
    p = tmp_path / "m.oct"
    pysvo.oct_write(p, words, center, compress=True)
    h = ref.tree_load(p)
    assert np.array_equal(ref.tree_words_view(h), words)
    q = tmp_path / "r.oct"
    ref.tree_save(h, q)
    ref.tree_destroy(h)
    w2, _ = pysvo.oct_read(q)
    assert np.array_equal(w2, words)


def test_oct_parallel_decode(pysvo, ref, tmp_path, monkeypatch):
    """Row f1: slices decode concurrently. Files this library writes have independent slices (workers never
    block); files the reference writes chain through LZ4's streaming window (a worker blocks at its first
    cross-slice match until the predecessor is done). Both must give the same words for any thread count,
    and a broken slice must surface as a format error, not a hang."""
    n_words = 3 * (64 << 20) // 4 + 7777
    rng = np.random.default_rng(11)
    words = _synthetic_words(rng, n_words)
    for k in (1, 2, 3):   # the start of every later slice repeats the end of its predecessor
        b = k * (64 << 20) // 4
        words[b:b + 3000] = words[b - 3000:b]
    center = np.array([0.5, 0.5, 0.25], np.float32)
    ours = tmp_path / "ours.oct"
    monkeypatch.setenv("SVO_IO_THREADS", "4")
    pysvo.oct_write(ours, words, center, compress=True)
    h = ref.tree_load(ours)                       # the reference reads what the threaded writer wrote
    assert np.array_equal(ref.tree_words_view(h), words)
    theirs = tmp_path / "theirs.oct"
    ref.tree_save(h, theirs)                      # ... and writes a file with cross-slice matches
    ref.tree_destroy(h)
    for threads in ("1", "2", "4", "7"):
        monkeypatch.setenv("SVO_IO_THREADS", threads)
        for path in (ours, theirs):
            w2, c2 = pysvo.oct_read(path)
            assert np.array_equal(w2, words) and np.array_equal(c2, center), (threads, path.name)
    # single-threaded writer produces the same bytes as the threaded one
    monkeypatch.setenv("SVO_IO_THREADS", "1")
    again = tmp_path / "again.oct"
    pysvo.oct_write(again, words, center, compress=True)
    assert again.read_bytes() == ours.read_bytes()
    # damage the second slice of each file: every thread count reports it
    for path in (ours, theirs):
        raw = bytearray(path.read_bytes())
        first = struct.unpack_from("<Q", raw, 20)[0]
        second_payload = 28 + first + 8
        raw[second_payload + 100:second_payload + 100 + 64] = bytes(64)   # zero offsets are invalid
        bad = tmp_path / ("bad_" + path.name)
        bad.write_bytes(bytes(raw))
        for threads in ("1", "4"):
            monkeypatch.setenv("SVO_IO_THREADS", threads)
            with pytest.raises(pysvo.SvoError) as e:
                pysvo.oct_read(bad)
            assert e.value.status == 3
        trunc = tmp_path / ("trunc_" + path.name)
        trunc.write_bytes(bytes(path.read_bytes()[:second_payload + 1000]))
        with pytest.raises(pysvo.SvoError) as e:
            pysvo.oct_read(trunc)
        assert e.value.status == 3


def test_oct_errors_are_reported(pysvo, tmp_path, dragon_words):
    with pytest.raises(pysvo.SvoError) as e:
        pysvo.oct_read(tmp_path / "missing.oct")
    assert e.value.status == 2
    with pytest.raises(pysvo.SvoError) as e:
        pysvo.oct_write(tmp_path / "no_such_dir" / "x.oct", np.zeros(4, np.uint32), [0, 0, 0])
    assert e.value.status == 2
    raw = DRAGON.read_bytes()
    for name, blob in [("trunc_header.oct", raw[:15]), ("trunc_payload.oct", raw[:len(raw) // 2]),
                       ("garbage_count.oct", raw[:12] + struct.pack("<Q", 1 << 60) + raw[20:]),
                       ("trailing.oct", raw[:20] + struct.pack("<Q", len(raw) - 28 + 4) + raw[28:] + b"\0\0\0\0")]:
        p = tmp_path / name
        p.write_bytes(blob)
        with pytest.raises(pysvo.SvoError) as e:
            pysvo.oct_read(p)
        assert e.value.status == 3, name
    # corrupt a match offset region: must fail or decode, never crash
    blob = bytearray(raw)
    rng = np.random.default_rng(1)
    for _ in range(20):
        b2 = bytearray(blob)
        for pos in rng.integers(28, len(raw), 8):
            b2[pos] ^= 0xFF
        p = tmp_path / "fuzz.oct"
        p.write_bytes(bytes(b2))
        try:
            pysvo.oct_read(p)
        except pysvo.SvoError as err:
            assert err.status == 3


@pytest.mark.parametrize("cam", CAMERAS)
def test_camera_constants_match_oracle(pysvo, port, dragon_words, cam):
    _, center = dragon_words
    c = pysvo.orbit_camera(*cam)
    m, v = port.orbit_camera(*cam)
    assert np.array_equal(np.array(c.model[:], np.float32).view(np.uint32), m.view(np.uint32))
    assert np.array_equal(np.array(c.view[:], np.float32).view(np.uint32), v.view(np.uint32))
    for (W, H, S) in [(1280, 720, 16), (1920, 1080, 16), (3840, 2160, 16), (333, 77, 5), (8, 8, 1)]:
        a = pysvo.frame_constants(c, center, W, H, S)
        b = port.frame_constants(m, v, center, W, H, S)
        assert (a.width, a.height, a.strips, a.tile_size) == (b.width, b.height, b.strips, b.tile_size)
        assert np.array_equal(a.as_array().view(np.uint32), b.as_array().view(np.uint32))


def test_camera_matches_reference_matrix_stack(pysvo, ref):
    for cam in CAMERAS:
        c = pysvo.orbit_camera(*cam)
        m, v = ref.orbit_camera(*cam)
        assert np.array_equal(np.array(c.model[:], np.float32).view(np.uint32), m.view(np.uint32))
        assert np.array_equal(np.array(c.view[:], np.float32).view(np.uint32), v.view(np.uint32))


def test_strip_layout_matches_reference_formula(pysvo):
    # Main.cpp:351-362 for the reference's own configuration
    lay = pysvo.strip_layout(1280, 720, 16)
    assert len(lay) == 16 and lay[0] == (0, 45, 161, 7) and lay[-1] == (675, 720, 161, 7)
    assert pysvo.coarse_cells(1280, 720, 16) == 18032
    assert pysvo.coarse_cells(3840, 2160, 16) == 138528


def test_oct_beyond_two_gib(pysvo, tmp_path):
    """Trees of 2 GiB and more: the reference's loader truncates the remaining byte count to int
    (VoxelOctree.cpp:79) and cannot read them; this reader / writer count in 64 bits.
    33 slices, the last one 12 bytes long."""
    n_words = (2 << 30) // 4 + 3
    words = np.zeros(n_words, np.uint32)
    words[0] = (1 << 18) | 0x0100
    marks = np.arange(0, n_words, 9973, dtype=np.int64)
    words[marks] = (marks * 2654435761 & 0xFFFFFFFF).astype(np.uint32)
    words[-3:] = [0xAAAAAAAA, 0xBBBBBBBB, 0xCCCCCCCC]
    center = np.array([0.5, 0.5, 0.5], np.float32)
    p = tmp_path / "big.oct"
    pysvo.oct_write(p, words, center, compress=True)
    assert p.stat().st_size < n_words * 4 // 20                 # really compressed
    w2, c2 = pysvo.oct_read(p)
    assert w2.size == n_words and np.array_equal(c2, center)
    assert np.array_equal(w2[marks], words[marks]) and np.array_equal(w2[-3:], words[-3:])
    assert int(np.count_nonzero(w2)) == int(np.count_nonzero(words))


def test_ply_reader_matches_oracle(pysvo, port, tmp_path):
    """svo_ply_read_triangles (the product's PLY reader + PlyLoader's vertex rescaling, fans and face normals,
    PlyLoader.cpp:64-226) against the restatement the voxeliser oracle is built on, bit for bit, for every
    container variant; malformed files are reported."""
    from ply_meshes import write_variants
    from tools import make_scenes
    ico = tmp_path / "ico5.ply"
    assert make_scenes.gen_lib().svo_scene_icosphere_ply(str(ico).encode(), 5, 7) > 0
    big = tmp_path / "ico100.ply"       # 200,000 triangles: the reader's multi-threaded paths
    assert make_scenes.gen_lib().svo_scene_icosphere_ply(str(big).encode(), 100, 7) > 0
    # a quad and a two-index face among 70,000 triangles: the face data is exactly as long as if every face were
    # a triangle, so the reader's all-triangles fast path has to notice the counts and fall back
    import struct
    rng = np.random.default_rng(3)
    v = rng.random((500, 3)).astype(np.float32)
    faces = [list(map(int, rng.choice(500, 3, replace=False))) for _ in range(70000)]
    faces[40000:40002] = [[5, 9, 13, 17], [2, 3]]
    mixed = tmp_path / "mixed_counts.ply"
    with open(mixed, "wb") as fp:
        fp.write(("ply\nformat binary_little_endian 1.0\nelement vertex 500\nproperty float x\nproperty float y\n"
                  "property float z\nelement face %d\nproperty list uchar int vertex_indices\nend_header\n" % len(faces)).encode())
        fp.write(v.tobytes())
        for f in faces:
            fp.write(struct.pack("<B%di" % len(f), len(f), *f))
    # the 200,000-triangle mesh as ASCII: vertices and faces go through the token-indexed parallel paths (every face a
    # triangle), with an ignored vertex property in the middle and a scalar-only element behind the faces
    raw_big = big.read_bytes()
    h = raw_big.index(b"end_header\n") + 11
    header = raw_big[:h].decode()
    nv, nf = int(header.split("element vertex ")[1].split()[0]), int(header.split("element face ")[1].split()[0])
    bv = np.frombuffer(raw_big[h:h + nv * 12], np.float32).reshape(nv, 3)
    bf = np.frombuffer(raw_big[h + nv * 12:], np.uint8).reshape(nf, 13)[:, 1:].copy().view(np.int32).reshape(nf, 3)
    ascii_big = tmp_path / "ico100_ascii.ply"
    with open(ascii_big, "w") as fp:
        fp.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float quality\nproperty float y\n"
                 "property float z\nelement face %d\nproperty list uchar int vertex_indices\nelement edge 2\nproperty int a\n"
                 "property int b\nend_header\n" % (nv, nf))
        np.savetxt(fp, np.column_stack([bv[:, 0], np.full(nv, 0.25), bv[:, 1], bv[:, 2]]), fmt="%.9g")
        np.savetxt(fp, np.hstack([np.full((nf, 1), 3), bf]), fmt="%d")
        fp.write("0 1\n1 2\n")
    a, lo, hi = pysvo.ply_read_triangles(ascii_big)
    b, _, _ = pysvo.ply_read_triangles(big)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))            # same mesh as the binary file, bit for bit
    # ... and with one quad among the triangles: the token-indexed face path has to notice and fall back
    ascii_quad = tmp_path / "ico100_ascii_quad.ply"
    lines = ascii_big.read_text().split("\n")
    first_face = lines.index("end_header") + 1 + nv
    lines[first_face + 1000] = "4 0 1 2 3"
    ascii_quad.write_text("\n".join(lines))
    for name, p in [("ico5", ico), ("ico100", big), ("mixed_counts", mixed), ("ico100_ascii", ascii_big),
                    ("ico100_ascii_quad", ascii_quad)] + write_variants(tmp_path):
        a, lo, hi = pysvo.ply_read_triangles(p)
        b, lo2, hi2 = port.ply_triangles(p)
        assert a.shape == b.shape and a.shape[0] > 100, name
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), name
        assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2), name
        assert (a[:, :9].min() >= 0.0) and (a[:, :9].max() <= 1.0)          # rescaled to the unit box
    with pytest.raises(pysvo.SvoError) as e:
        pysvo.ply_read_triangles(tmp_path / "missing.ply")
    assert e.value.status == 2
    raw = ico.read_bytes()
    bigraw = bytearray(big.read_bytes())
    bigraw[-4:] = struct.pack("<i", 10**8)          # the last face's last index points past the vertices
    for name, blob in [("notply.ply", b"plx\n" + raw[4:]), ("short.ply", raw[:len(raw) // 2]),
                       ("nofaces.ply", raw.replace(b"element face", b"element fac_")), ("badindex.ply", bytes(bigraw)),
                       ("ascii_junk.ply", ascii_big.read_bytes().replace(b"\n3 ", b"\n3 x", 1)),
                       ("ascii_short.ply", ascii_big.read_bytes()[:-60000]),
                       ("ascii_badindex.ply", ascii_big.read_bytes().replace(b"\n0 1\n1 2\n", b"").rsplit(b" ", 1)[0] + b" 99999999\n0 1\n1 2\n")]:
        q = tmp_path / name
        q.write_bytes(blob)
        with pytest.raises(pysvo.SvoError) as e:
            pysvo.ply_read_triangles(q)
        assert e.value.status == 3, name


def _viewer_script(rng, n):
    """Random window events: motion with and without buttons held, presses and releases of both buttons (and of
    buttons / keys the viewer ignores, which still wake it up)."""
    ev = []
    for _ in range(n):
        r = rng.random()
        if r < 0.55:
            ev.append((4, 0, int(rng.integers(-40, 41)), int(rng.integers(-60, 61))))
        elif r < 0.75:
            ev.append((5, int(rng.choice([1, 3, 2, 4])), 0, 0))
        elif r < 0.95:
            ev.append((6, int(rng.choice([1, 3])), 0, 0))
        else:
            ev.append((2, int(rng.choice([32, 97])), 0, 0))
    return ev


def test_viewer_camera_control_matches_reference_viewer(pysvo, ref):
    """Row f4: svo_viewer_feed (Events.cpp's mouse state + renderLoop's event handling, Main.cpp:229-252) against
    the reference's own `-viewer` main loop run headless on a scripted SDL (oracle/_ref, svoref_viewer_run): the
    same number of presented frames, each after the same number of consumed events, with bit-identical MODEL /
    VIEW matrices and the same renderHalfSize -- including the reference's quirk that motion received while no
    button is held is applied by the next button press, pitch wrapping, the yaw direction flip beyond 90 degrees of
    pitch and the zoom clamps."""
    from conftest import DRAGON
    rng = np.random.default_rng(11)
    scripts = [_viewer_script(rng, 150) for _ in range(3)]
    scripts.append([(5, 3, 0, 0)] + [(4, 0, 0, -70)] * 12 + [(4, 0, 0, 90)] * 6 + [(6, 3, 0, 0)])      # zoom out to the 25 cap, in by the 0.5 clamp
    scripts.append([(5, 1, 0, 0)] + [(4, 0, 7, -50)] * 9 + [(4, 0, -300, 0), (6, 1, 0, 0), (4, 0, 5, 5), (2, 27, 0, 0), (4, 0, 1, 1)])
    for ev in scripts:
        want = ref.viewer_run(DRAGON, 16, 16, 2, ev, want_pixels=False)
        st = pysvo.viewer_init()
        got = [(np.array(st.camera.model[:], np.float32), np.array(st.camera.view[:], np.float32), st.preview, 0)]
        for i, e in enumerate(ev):
            action = pysvo.viewer_feed(st, *e)
            if action == pysvo.VIEWER_FRAME:
                got.append((np.array(st.camera.model[:], np.float32), np.array(st.camera.view[:], np.float32), st.preview, i + 1))
            elif action == pysvo.VIEWER_QUIT:
                break
        assert len(got) == len(want["half"])
        for k, (m, v, half, taken) in enumerate(got):
            assert np.array_equal(m.view(np.uint32), want["model"][k].view(np.uint32)), k
            assert np.array_equal(v.view(np.uint32), want["view"][k].view(np.uint32)), k
            assert half == want["half"][k] and taken == want["events_taken"][k], k
    assert st.quit == 1 and pysvo.viewer_feed(st, 4, 0, 1, 1) == pysvo.VIEWER_QUIT


def test_viewer_camera_control_matches_golden_sessions(pysvo, port, dragon_words):
    """The same check against the committed vectors (tests/golden/viewer_sessions.npz, minted from the reference's own
    viewer by tests/golden/make_viewer_golden.py), which needs no reference build; and the oracle restatement's frames
    for the recorded matrices equal the frames the reference's window showed (full-resolution and stride-3 previews)."""
    from conftest import GOLDEN
    g = np.load(GOLDEN / "viewer_sessions.npz")
    n_sessions = sum(1 for k in g.files if k.startswith("events_"))
    assert n_sessions >= 5
    for i in range(n_sessions):
        st = pysvo.viewer_init()
        got = [(np.array(st.camera.model[:], np.float32), np.array(st.camera.view[:], np.float32), st.preview, 0)]
        for j, e in enumerate(g[f"events_{i}"]):
            action = pysvo.viewer_feed(st, *[int(v) for v in e])
            if action == pysvo.VIEWER_FRAME:
                got.append((np.array(st.camera.model[:], np.float32), np.array(st.camera.view[:], np.float32), st.preview, j + 1))
            elif action == pysvo.VIEWER_QUIT:
                break
        assert len(got) == len(g[f"half_{i}"]), i
        for k, (m, v, half, taken) in enumerate(got):
            assert np.array_equal(m.view(np.uint32), g[f"model_{i}"][k].view(np.uint32)), (i, k)
            assert np.array_equal(v.view(np.uint32), g[f"view_{i}"][k].view(np.uint32)), (i, k)
            assert half == g[f"half_{i}"][k] and taken == g[f"taken_{i}"][k], (i, k)
    words, center = dragon_words
    W, H, S = (int(v) for v in g["frame_shape"])
    for k in range(len(g["half_0"])):
        f = port.frame_constants(g["model_0"][k], g["view_0"][k], center, W, H, S)
        img = port.render_frame(words, f, pixel_stride=3 if g["half_0"][k] else 1)[0]
        assert np.array_equal(img, g["rgba_0"][k]), k


def test_headless_driver_reports_errors_without_a_gpu(pysvo, tmp_path):
    """svo_headless: a malformed or missing event script fails before anything touches the GPU; without a device the
    driver says so (no CPU fallback) instead of rendering anything."""
    import subprocess
    from conftest import DRAGON, ROOT
    pkg = ROOT / "sparse-voxel-octrees_b200"
    subprocess.check_call(["make", "-C", str(pkg), "headless"], stdout=subprocess.DEVNULL)
    exe = str(pkg / "svo_headless")
    bad = tmp_path / "bad.events"
    bad.write_text("# session\nmotion 1 2\nwiggle 3\n")
    out = subprocess.run([exe, str(DRAGON), "--events", str(bad)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "bad.events:3: cannot parse event" in out.stderr
    out = subprocess.run([exe, str(DRAGON), "--events", str(tmp_path / "none.events")], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "cannot read" in out.stderr
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 2 and "--events script" in out.stderr and "-builder" in out.stderr
    out = subprocess.run([exe, str(DRAGON), "--bogus"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 2 and "unknown option" in out.stderr
    # --check: the file's node array walked on the host, no GPU involved
    out = subprocess.run([exe, str(DRAGON), "--check"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "119887 words = 29156 descriptors + 24 far words + 90707 leaf words, depth 8" in out.stdout
    damaged = tmp_path / "damaged.oct"
    words, center = pysvo.oct_read(DRAGON)
    words = words.copy()
    words[1] &= np.uint32(0x1FFFF)          # the root's first child loses its (far) pointer: a node that is its own child
    pysvo.oct_write(damaged, words, center, compress=True)
    out = subprocess.run([exe, str(damaged), "--check"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "zero child offset at 1" in out.stderr
    if pysvo.device_count() < 1:
        good = tmp_path / "ok.events"
        good.write_text("down left\nmotion 4 -3\nup left\n")
        out = subprocess.run([exe, str(DRAGON), "--events", str(good)], capture_output=True, text=True, timeout=60)
        assert out.returncode == 1 and "no CPU fallback" in out.stderr


def test_readers_survive_mutated_files(pysvo, tmp_path):
    """.oct and PLY readers on 150 mutated files each (tests/fuzz_readers.py, in a child process): every file decodes
    or is rejected with a status -- no crash, no hang."""
    import subprocess
    import sys
    from conftest import ROOT
    out = subprocess.run([sys.executable, str(ROOT / "tests" / "fuzz_readers.py"), "5", "150", str(tmp_path)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.returncode, out.stderr[-2000:])
    assert "fuzz done" in out.stdout and "'ply rejected': 0" not in out.stdout and "'oct rejected': 0" not in out.stdout


def test_headless_png_writer(tmp_path):
    """host/png_write.hpp (svo_headless --png): signature, chunk CRCs, IHDR, and the pixels after inflating the IDAT
    stream, for a ragged and a larger-than-one-deflate-block frame."""
    import struct
    import subprocess
    import zlib
    from conftest import ROOT
    exe = tmp_path / "png_test"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", str(ROOT / "sparse-voxel-octrees_b200" / "host"),
                           str(ROOT / "tests" / "cpp" / "png_test.cpp"), "-o", str(exe)])
    for w, h in ((333, 187), (5, 3), (640, 360)):
        path = tmp_path / f"f_{w}x{h}.png"
        subprocess.check_call([str(exe), str(path), str(w), str(h)])
        b = path.read_bytes()
        assert b[:8] == b"\x89PNG\r\n\x1a\n"
        pos, chunks = 8, []
        while pos < len(b):
            n, = struct.unpack(">I", b[pos:pos + 4])
            kind, data = b[pos + 4:pos + 8], b[pos + 8:pos + 8 + n]
            assert zlib.crc32(kind + data) == struct.unpack(">I", b[pos + 8 + n:pos + 12 + n])[0], kind
            chunks.append((kind, data))
            pos += 12 + n
        assert [k for k, _ in chunks] == [b"IHDR", b"IDAT", b"IEND"]
        assert struct.unpack(">IIBBBBB", chunks[0][1]) == (w, h, 8, 2, 0, 0, 0)
        rows = np.frombuffer(zlib.decompress(chunks[1][1]), np.uint8).reshape(h, 1 + 3 * w)
        assert (rows[:, 0] == 0).all()
        px = rows[:, 1:].reshape(h, w, 3)
        y, x = np.mgrid[0:h, 0:w]
        assert np.array_equal(px[..., 0], ((x * 7 + y * 3) & 255).astype(np.uint8))
        assert np.array_equal(px[..., 1], ((x ^ y) & 255).astype(np.uint8))
        assert np.array_equal(px[..., 2], ((x * y) & 255).astype(np.uint8))


def test_words_validate(pysvo, dragon_words, monkeypatch):
    """svo_words_validate: the full host-side walk of a node array. The sample tree gives the survey's counts and
    accounts for every word; damaged arrays are rejected with the first violation named (and never crash the walk);
    the tree constructors run it by default (SVO_VALIDATE_TREES=0 opts out) before a tree is handed out."""
    words, center = dragon_words
    rep = pysvo.words_validate(words)
    assert (rep.descriptors, rep.leaves, rep.far_words, rep.depth) == (29156, 90707, 24, 8)
    assert rep.min_leaf_depth == rep.max_leaf_depth == 8
    assert rep.descriptors + rep.leaves + rep.far_words == words.size          # depth-first by block: no unused word

    def rejected(w, needle):
        with pytest.raises(pysvo.SvoError) as e:
            pysvo.words_validate(w)
        assert e.value.status == 3 and needle in str(e.value), str(e.value)

    # a descriptor deep in the tree whose children are leaves, found by walking first children from the root
    p, chain = 0, []
    while words[p] & 0xFF:
        chain.append(p)
        off = int(words[p]) >> 18
        if words[p] & 0x20000:
            off = (off << 32) | int(words[p + 1])
        p += off
    chain.append(p)
    leaf_parent, inner = chain[-1], chain[-3]
    w = words.copy(); w[inner] = (w[inner] & 0x3FFFF) | (0x3FFF << 18)          # child block far behind the array's end?
    w[inner] &= ~np.uint32(0x20000)
    if inner + 0x3FFF < words.size:                                             # (the Dragon is big enough: make it a far pointer instead)
        w[inner] |= np.uint32(0x20000); w[inner + 1] = np.uint32(words.size + 5)
    rejected(w, "past the end")
    w = words.copy(); w[inner] &= np.uint32(0x3FFFF)                            # zero offset: a node that is its own child
    rejected(w, "zero child offset")
    # a leaf parent OFF the first-child chain (the chain is what the depth is measured along): second child of the
    # lowest ancestor that has two, then first children down to the leaves' parent
    def child_block(q):
        off = int(words[q]) >> 18
        if words[q] & 0x20000:
            off = (off << 32) | int(words[q + 1])
        return q + off, (2 if words[q] & 0x10000 else 1)
    other = None
    for anc in reversed(chain[:-1]):
        if bin(int(words[anc]) & 0xFF).count("1") >= 2:
            base, stride = child_block(anc)
            other = base + stride
            break
    assert other is not None
    while words[other] & 0xFF:
        other = child_block(other)[0]
    w = words.copy(); w[other] |= np.uint32((int(w[other]) >> 8) & 0xFF)        # its leaves turned into nodes: one level too deep
    rejected(w, "deeper than 8 levels")
    w = words.copy(); w[0] = (w[0] & ~np.uint32(0xFF00)) | np.uint32(0)         # root without children
    rejected(w, "no children")
    rng = np.random.default_rng(9)
    outcomes = {True: 0, False: 0}
    for _ in range(300):
        w = words.copy()
        for _ in range(int(rng.integers(1, 6))):
            w[int(rng.integers(0, w.size))] = np.uint32(rng.integers(0, 2**32))
        try:
            pysvo.words_validate(w)
            outcomes[True] += 1
        except pysvo.SvoError as e:
            assert e.status == 3
            outcomes[False] += 1
    assert outcomes[True] > 0 and outcomes[False] > 0                           # damaged leaf words pass, damaged descriptors mostly do not
    if pysvo.device_count() < 1:
        # the constructors validate by default (SVO_VALIDATE_TREES=0 opts out), and a format error wins over the
        # missing device. The damage is off the first-child chain, so nothing but the full walk can see it.
        bad = words.copy(); bad[other] |= np.uint32((int(bad[other]) >> 8) & 0xFF)
        monkeypatch.delenv("SVO_VALIDATE_TREES", raising=False)
        with pytest.raises(pysvo.SvoError) as e:
            pysvo.VoxelOctree(words=bad, center=center)
        assert e.value.status == 3
        with pytest.raises(pysvo.SvoError) as e:
            pysvo.VoxelOctree(words=words, center=center)
        assert e.value.status == 6
        monkeypatch.setenv("SVO_VALIDATE_TREES", "0")
        with pytest.raises(pysvo.SvoError) as e:
            pysvo.VoxelOctree(words=bad, center=center)
        assert e.value.status == 6


def test_pixels_expand_grey8a(pysvo):
    """svo_pixels_expand_grey8a: (grey, alpha) byte pairs -> the reference's pixel words (Main.cpp:128-132, :165)."""
    rng = np.random.default_rng(2)
    g = rng.integers(0, 256, 10007, dtype=np.uint32)
    a = rng.choice(np.array([0, 255], np.uint32), 10007)
    g[a == 0] = 0
    rgba = (a << 24) | (g << 16) | (g << 8) | g
    packed = (g | (a << 8)).astype(np.uint16)
    assert np.array_equal(pysvo.expand_grey8a(packed), rgba)
    assert pysvo.expand_grey8a(np.zeros(0, np.uint16)).size == 0


def test_plain_c_example_builds_against_the_abi(pysvo, tmp_path):
    """host/example_render.c: the boundary is usable from pedantic C99 (no C++ in the header), links against the
    library, and without a device fails with the library's message instead of doing anything else."""
    import subprocess
    from conftest import DRAGON, ROOT
    pkg = ROOT / "sparse-voxel-octrees_b200"
    exe = tmp_path / "example_render"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", str(ROOT / "include"),
                           str(pkg / "host" / "example_render.c"), "-L", str(pkg), "-lsvo_b200", f"-Wl,-rpath,{pkg}", "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 2 and "usage" in out.stderr
    if pysvo.device_count() < 1:
        out = subprocess.run([str(exe), str(DRAGON), str(tmp_path / "x.ppm")], capture_output=True, text=True, timeout=60)
        assert out.returncode == 1 and "no CPU fallback" in out.stderr
