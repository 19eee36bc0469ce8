"""ctypes bindings for the parity oracles. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product
(sparse-voxel-octrees_b200/) never does.

Two libraries:

* ``oracle/_ref/libsvo_ref.so``  -- the reference's own object code behind a
  C-ABI harness (oracle/ref_harness.cpp, built by oracle/build_ref.sh from
  /root/reference/src). ``Ref`` below.
* ``oracle/liboracle.so``        -- the plain-C restatement (oracle/svo_oracle.c),
  instrumented with fetch counters. ``Port`` below.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_SO = HERE / "_ref" / "libsvo_ref.so"
PORT_SO = HERE / "liboracle.so"
REFERENCE_ROOT = Path(os.environ.get("SVO_REFERENCE", "/root/reference"))

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


def build_port(force: bool = False) -> Path:
    """gcc the C restatement (seconds). -ffp-contract=off is load-bearing."""
    srcs = [HERE / "svo_oracle.c", HERE / "svo_oracle_ply.c"]
    if force or not PORT_SO.exists() or PORT_SO.stat().st_mtime < max(
            [s.stat().st_mtime for s in srcs] + [(HERE / "svo_oracle.h").stat().st_mtime]):
        subprocess.check_call([
            "gcc", "-std=c11", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-pthread",
            "-Wall", "-Wextra", "-o", str(PORT_SO)] + [str(s) for s in srcs] + ["-lm"])
    return PORT_SO


def build_ref(force: bool = False) -> Path | None:
    """Build oracle/_ref from /root/reference when it is present; else keep the prebuilt file."""
    if (REFERENCE_ROOT / "src").is_dir():
        deps = [HERE / "ref_harness.cpp", HERE / "build_ref.sh", HERE / "ref_shim" / "SDL.h"]
        if force or not REF_SO.exists() or REF_SO.stat().st_mtime < max(d.stat().st_mtime for d in deps):
            subprocess.check_call(["bash", str(HERE / "build_ref.sh")], stdout=subprocess.DEVNULL)
    return REF_SO if REF_SO.exists() else None


def _null_or(arr):
    return None if arr is None else arr.ctypes.data_as(C.c_void_p)


class Ref:
    """The reference's own code (VoxelOctree.cpp / Main.cpp), via oracle/_ref."""

    def __init__(self):
        if build_ref() is None:
            raise FileNotFoundError(f"{REF_SO} missing and {REFERENCE_ROOT} not available to build it")
        # The reference prints progress on std::cout; keep it, it is harmless.
        L = self.lib = C.CDLL(str(REF_SO))
        L.svoref_tree_load.restype = C.c_void_p
        L.svoref_tree_load.argtypes = [C.c_char_p]
        L.svoref_tree_from_words.restype = C.c_void_p
        L.svoref_tree_from_words.argtypes = [_u32p, C.c_uint64, _f32p]
        L.svoref_tree_build_voxel_file.restype = C.c_void_p
        L.svoref_tree_build_voxel_file.argtypes = [C.c_char_p, C.c_uint64]
        L.svoref_tree_build_ply.restype = C.c_void_p
        L.svoref_tree_build_ply.argtypes = [C.c_char_p, C.c_int, C.c_uint64]
        L.svoref_ply_to_voxel_file.restype = C.c_int
        L.svoref_ply_to_voxel_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_uint64]
        L.svoref_tree_save.restype = None
        L.svoref_tree_save.argtypes = [C.c_void_p, C.c_char_p]
        L.svoref_tree_word_count.restype = C.c_uint64
        L.svoref_tree_word_count.argtypes = [C.c_void_p]
        L.svoref_tree_words.restype = C.POINTER(C.c_uint32)
        L.svoref_tree_words.argtypes = [C.c_void_p]
        L.svoref_tree_center.restype = None
        L.svoref_tree_center.argtypes = [C.c_void_p, _f32p]
        L.svoref_tree_destroy.restype = None
        L.svoref_tree_destroy.argtypes = [C.c_void_p]
        L.svoref_compress_material.restype = C.c_uint32
        L.svoref_compress_material.argtypes = [_f32p, C.c_float]
        L.svoref_decompress_material.restype = None
        L.svoref_decompress_material.argtypes = [C.c_uint32, _f32p, C.POINTER(C.c_float)]
        L.svoref_inv_sqrt.restype = C.c_float
        L.svoref_inv_sqrt.argtypes = [C.c_float]
        L.svoref_raymarch.restype = C.c_int
        L.svoref_raymarch.argtypes = [C.c_void_p, _f32p, _f32p, C.c_float, C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        L.svoref_raymarch_batch.restype = C.c_double
        L.svoref_raymarch_batch.argtypes = [C.c_void_p, C.c_uint64, _f32p, _f32p, C.c_float,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.svoref_orbit_camera.restype = None
        L.svoref_orbit_camera.argtypes = [C.c_float, C.c_float, C.c_float, _f32p, _f32p]
        L.svoref_inv_modelview.restype = None
        L.svoref_inv_modelview.argtypes = [_f32p, _f32p, _f32p]
        L.svoref_render_frames.restype = C.c_int
        L.svoref_render_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int,
                                           _u32p, C.c_void_p, C.c_void_p]
        L.svoref_render_frames_subset.restype = C.c_int
        L.svoref_render_frames_subset.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p,
                                                  C.c_int, _u32p, C.c_void_p, C.c_void_p]
        L.svoref_hardware_threads.restype = C.c_int
        L.svoref_viewer_run.restype = C.c_int
        L.svoref_viewer_run.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.svoref_set_half_size.restype = None
        L.svoref_set_half_size.argtypes = [C.c_int]

    # -- trees
    def tree_load(self, path):
        h = self.lib.svoref_tree_load(str(path).encode())
        if not h:
            raise FileNotFoundError(path)
        return h

    def tree_from_words(self, words, center):
        words = np.ascontiguousarray(words, np.uint32)
        return self.lib.svoref_tree_from_words(words, words.size, np.ascontiguousarray(center, np.float32))

    def tree_build_voxel_file(self, path, mem=1 << 30):
        h = self.lib.svoref_tree_build_voxel_file(str(path).encode(), int(mem))
        if not h:
            raise FileNotFoundError(path)
        return h

    def ply_to_voxel_file(self, ply, out, resolution, mem=1 << 30):
        """PlyLoader(ply).convertToVolume(out, resolution, mem) (PlyLoader.cpp:505-534)."""
        if self.lib.svoref_ply_to_voxel_file(str(ply).encode(), str(out).encode(), int(resolution), int(mem)) != 0:
            raise FileNotFoundError(ply)

    def tree_build_ply(self, path, resolution, mem=1 << 30):
        h = self.lib.svoref_tree_build_ply(str(path).encode(), int(resolution), int(mem))
        if not h:
            raise FileNotFoundError(path)
        return h

    def tree_save(self, h, path):
        self.lib.svoref_tree_save(h, str(path).encode())

    def tree_words(self, h):
        n = self.lib.svoref_tree_word_count(h)
        return np.ctypeslib.as_array(self.lib.svoref_tree_words(h), shape=(n,)).copy()

    def tree_words_view(self, h):
        n = self.lib.svoref_tree_word_count(h)
        return np.ctypeslib.as_array(self.lib.svoref_tree_words(h), shape=(n,))

    def tree_center(self, h):
        c = np.zeros(3, np.float32)
        self.lib.svoref_tree_center(h, c)
        return c

    def tree_destroy(self, h):
        self.lib.svoref_tree_destroy(h)

    # -- small helpers
    def compress_material(self, n, shade):
        return int(self.lib.svoref_compress_material(np.ascontiguousarray(n, np.float32), float(shade)))

    def decompress_material(self, word):
        n = np.zeros(3, np.float32)
        s = C.c_float()
        self.lib.svoref_decompress_material(int(word), n, C.byref(s))
        return n, np.float32(s.value)

    def inv_sqrt(self, x):
        return np.float32(self.lib.svoref_inv_sqrt(float(x)))

    def orbit_camera(self, pitch_deg, yaw_deg, radius):
        m = np.zeros(16, np.float32)
        v = np.zeros(16, np.float32)
        self.lib.svoref_orbit_camera(float(pitch_deg), float(yaw_deg), float(radius), m, v)
        return m, v

    def inv_modelview(self, model, view):
        out = np.zeros(16, np.float32)
        self.lib.svoref_inv_modelview(np.ascontiguousarray(model, np.float32), np.ascontiguousarray(view, np.float32), out)
        return out

    # -- traversal
    def raymarch(self, h, o, d, ray_scale=0.0, normal_sentinel=0xDEADBEEF, t_sentinel=-1.0):
        n = C.c_uint32(normal_sentinel)
        t = C.c_float(t_sentinel)
        hit = self.lib.svoref_raymarch(h, np.ascontiguousarray(o, np.float32), np.ascontiguousarray(d, np.float32),
                                       float(ray_scale), C.byref(n), C.byref(t))
        return bool(hit), np.float32(t.value), int(n.value)

    def raymarch_batch(self, h, o, d, ray_scale=0.0, threads=1, normal_sentinel=0, t_sentinel=0.0):
        """Returns (hit u8[n], t f32[n], normal u32[n], seconds). t/normal keep the sentinel where the
        reference leaves them untouched (misses; normal on LOD exits)."""
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = o.shape[0]
        hit = np.zeros(n, np.uint8)
        t = np.full(n, t_sentinel, np.float32)
        normal = np.full(n, normal_sentinel, np.uint32)
        secs = self.lib.svoref_raymarch_batch(h, n, o, d, float(ray_scale), _null_or(hit), _null_or(t),
                                              _null_or(normal), int(threads))
        return hit, t, normal, secs

    # -- frame loop
    def render_frames(self, h, W, H, strips, models, views, threads=None, want_depth=False, strip_modulo=1,
                      half_size=False):
        """Renders len(models) frames; returns (rgba u32[H,W] of the last frame, depth or None, seconds[f]).
        strip_modulo > 1 renders only every strip_modulo-th strip (bounded timing sample); half_size sets
        the reference's renderHalfSize (stride-3 preview, Main.cpp:161)."""
        self.lib.svoref_set_half_size(1 if half_size else 0)
        models = np.ascontiguousarray(models, np.float32).reshape(-1, 16)
        views = np.ascontiguousarray(views, np.float32).reshape(-1, 16)
        nf = models.shape[0]
        if threads is None:
            threads = min(strips, os.cpu_count() or 1)
        rgba = np.zeros((H, W), np.uint32)
        secs = np.zeros(nf, np.float64)
        depth = None
        if want_depth:
            depth = np.zeros(coarse_cells(W, H, strips), np.float32)
        rc = self.lib.svoref_render_frames_subset(h, W, H, strips, int(strip_modulo), nf, models, views, int(threads),
                                                  rgba.reshape(-1), _null_or(depth), _null_or(secs))
        self.lib.svoref_set_half_size(0)
        if rc != 0:
            raise ValueError("svoref_render_frames: bad arguments")
        return rgba, depth, secs

    def hardware_threads(self):
        return int(self.lib.svoref_hardware_threads())

    # -- the interactive viewer (row f4)
    def viewer_run(self, oct_path, W, H, strips, events, max_frames=None, want_pixels=True):
        """The reference's own `main -viewer oct_path` (Main.cpp:332-376, renderLoop :204-258, Events.cpp,
        ThreadBarrier.cpp) run against a scripted SDL: `events` is a list of (type, code, xrel, yrel) in SDL 1.2's
        numbering (2 key down, 3 key up, 4 mouse motion, 5 button down, 6 button up; buttons 1 left, 3 right);
        Escape is pressed when the script runs out. Returns a dict of per-presented-frame arrays: rgba
        [n, H, W], model [n, 16], view [n, 16], half [n] (renderHalfSize the frame was rendered with) and
        events_taken [n] (script events consumed when the frame was shown)."""
        ev = np.ascontiguousarray(np.asarray(events, np.int32).reshape(-1, 4))
        mf = int(max_frames or ev.shape[0] + 2)
        rgba = np.zeros((mf, H, W), np.uint32) if want_pixels else None
        models, views = np.zeros((mf, 16), np.float32), np.zeros((mf, 16), np.float32)
        half, taken = np.zeros(mf, np.int32), np.zeros(mf, np.int32)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        n = self.lib.svoref_viewer_run(str(oct_path).encode(), W, H, strips, ev.shape[0], p(ev), mf, p(rgba), p(models),
                                       p(views), p(half), p(taken))
        if n < 0:
            raise ValueError("svoref_viewer_run: bad arguments")
        n = min(n, mf)
        return {"rgba": None if rgba is None else rgba[:n], "model": models[:n], "view": views[:n], "half": half[:n],
                "events_taken": taken[:n]}


def strip_layout(W, H, strips, tile=8):
    """Main.cpp:351-362: (y0, y1, tilesX, tilesY) for every strip that owns at least one row."""
    stride = (H - 1) // strips + 1
    out = []
    for i in range(strips):
        y0 = i * stride
        if y0 >= H:
            break
        y1 = min((i + 1) * stride, H)
        out.append((y0, y1, (W - 1) // tile + 2, (y1 - y0 - 1) // tile + 2))
    return out


def coarse_cells(W, H, strips, tile=8):
    return sum(tx * ty for (_, _, tx, ty) in strip_layout(W, H, strips, tile))


# --------------------------------------------------------------------------------------
# The plain-C restatement (oracle/svo_oracle.c)
# --------------------------------------------------------------------------------------

class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "rays", "iterations", "desc_fetches", "far_fetches", "leaf_fetches", "pushes", "pops",
        "max_iterations", "hits", "lod_exits")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}

    @property
    def words(self):
        return self.desc_fetches + self.far_fetches + self.leaf_fetches

    def node_bytes_per_ray(self):
        """B_ray (node part) = 4 B x (descriptor + far + leaf fetches) / rays."""
        return 4.0 * self.words / max(self.rays, 1)


class Frame(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("strips", C.c_int32), ("tile_size", C.c_int32),
                ("pos", C.c_float * 3),
                ("a11", C.c_float), ("a12", C.c_float), ("a21", C.c_float), ("a22", C.c_float),
                ("a31", C.c_float), ("a32", C.c_float),
                ("zx", C.c_float), ("zy", C.c_float), ("zz", C.c_float),
                ("scale", C.c_float), ("tile_scale", C.c_float), ("coarse_scale", C.c_float), ("aspect", C.c_float),
                ("light", C.c_float * 3), ("beam_bias", C.c_float)]

    def as_array(self):
        """The 24 floats after the four ints, for bit-level comparisons."""
        return np.frombuffer(bytes(self), dtype=np.float32, offset=16).copy()


class TreeStats(C.Structure):
    _fields_ = [("descriptors", C.c_uint64), ("leaves", C.c_uint64), ("far_words", C.c_uint64),
                ("far_blocks", C.c_uint64), ("depth", C.c_uint32), ("per_level", C.c_uint64 * 24),
                ("max_index", C.c_uint64)]


class Port:
    """The C restatement with fetch counters. `kind: port` in bench.py's cpu_baseline."""

    MISS, LEAF, LOD = 0, 1, 2

    def __init__(self):
        build_port()
        L = self.lib = C.CDLL(str(PORT_SO))
        L.svo_oracle_raymarch.restype = C.c_int
        L.svo_oracle_raymarch.argtypes = [_u32p, _f32p, _f32p, C.c_float, C.POINTER(C.c_uint32), C.POINTER(C.c_float),
                                          C.POINTER(C.c_uint64), C.POINTER(Counters)]
        L.svo_oracle_raymarch_batch.restype = None
        L.svo_oracle_raymarch_batch.argtypes = [_u32p, C.c_uint64, _f32p, _f32p, C.c_float, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.POINTER(Counters), C.c_int]
        L.svo_oracle_frame_constants.restype = None
        L.svo_oracle_frame_constants.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.POINTER(Frame)]
        L.svo_oracle_orbit_camera.restype = None
        L.svo_oracle_orbit_camera.argtypes = [C.c_float, C.c_float, C.c_float, _f32p, _f32p]
        L.svo_oracle_shade.restype = C.c_float
        L.svo_oracle_shade.argtypes = [C.c_uint32, _f32p, _f32p]
        L.svo_oracle_pack.restype = C.c_uint32
        L.svo_oracle_pack.argtypes = [C.c_float]
        L.svo_oracle_decompress_material.restype = None
        L.svo_oracle_decompress_material.argtypes = [C.c_uint32, _f32p, C.POINTER(C.c_float)]
        L.svo_oracle_inv_sqrt.restype = C.c_float
        L.svo_oracle_inv_sqrt.argtypes = [C.c_float]
        L.svo_oracle_render_frame.restype = C.c_int
        L.svo_oracle_render_frame.argtypes = [_u32p, C.POINTER(Frame), _u32p, C.c_void_p, C.POINTER(Counters),
                                              C.POINTER(Counters), C.c_int]
        L.svo_oracle_tree_walk.restype = C.c_int
        L.svo_oracle_tree_walk.argtypes = [_u32p, C.c_uint64, C.POINTER(TreeStats)]
        L.svo_oracle_render_frame_strided.restype = C.c_int
        L.svo_oracle_render_frame_strided.argtypes = [_u32p, C.POINTER(Frame), C.c_int, _u32p, C.c_void_p,
                                                      C.POINTER(Counters), C.POINTER(Counters), C.c_int]
        L.svo_oracle_build_octree.restype = C.POINTER(C.c_uint32)
        L.svo_oracle_build_octree.argtypes = [_u32p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), _f32p]
        L.svo_oracle_free.restype = None
        L.svo_oracle_free.argtypes = [C.c_void_p]
        L.svo_oracle_voxelize_ply.restype = C.POINTER(C.c_uint32)
        L.svo_oracle_voxelize_ply.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int * 3), C.POINTER(C.c_uint64)]
        L.svo_oracle_block_list_aliases.restype = C.c_int64
        L.svo_oracle_block_list_aliases.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64),
                                                    C.POINTER(C.c_int * 3), C.POINTER(C.c_int * 3)]
        L.svo_oracle_ply_triangles.restype = C.POINTER(C.c_float)
        L.svo_oracle_ply_triangles.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), _f32p, _f32p]
        L.svo_oracle_tree_to_volume.restype = None
        L.svo_oracle_tree_to_volume.argtypes = [_u32p, C.c_int, _u32p, C.c_int, C.c_int, C.c_int]

    def raymarch(self, words, o, d, ray_scale=0.0, normal_sentinel=0xDEADBEEF, t_sentinel=-1.0):
        n = C.c_uint32(normal_sentinel)
        t = C.c_float(t_sentinel)
        v = C.c_uint64(0)
        c = Counters()
        r = self.lib.svo_oracle_raymarch(words, np.ascontiguousarray(o, np.float32), np.ascontiguousarray(d, np.float32),
                                         float(ray_scale), C.byref(n), C.byref(t), C.byref(v), C.byref(c))
        return int(r), np.float32(t.value), int(n.value), int(v.value), c

    def raymarch_batch(self, words, o, d, ray_scale=0.0, threads=None, normal_sentinel=0, t_sentinel=0.0):
        """Returns dict(hit u8 codes, t, normal, voxel, counters)."""
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = o.shape[0]
        hit = np.zeros(n, np.uint8)
        t = np.full(n, t_sentinel, np.float32)
        normal = np.full(n, normal_sentinel, np.uint32)
        voxel = np.zeros(n, np.uint64)
        c = Counters()
        if threads is None:
            threads = os.cpu_count() or 1
        self.lib.svo_oracle_raymarch_batch(words, n, o, d, float(ray_scale), _null_or(hit), _null_or(t),
                                           _null_or(normal), _null_or(voxel), C.byref(c), int(threads))
        return dict(hit=hit, t=t, normal=normal, voxel=voxel, counters=c)

    def frame_constants(self, model, view, center, W, H, strips):
        f = Frame()
        self.lib.svo_oracle_frame_constants(np.ascontiguousarray(model, np.float32),
                                            np.ascontiguousarray(view, np.float32),
                                            np.ascontiguousarray(center, np.float32), W, H, strips, C.byref(f))
        return f

    def orbit_camera(self, pitch_deg, yaw_deg, radius):
        m = np.zeros(16, np.float32)
        v = np.zeros(16, np.float32)
        self.lib.svo_oracle_orbit_camera(float(pitch_deg), float(yaw_deg), float(radius), m, v)
        return m, v

    def render_frame(self, words, frame, threads=None, want_depth=False, pixel_stride=1):
        """Returns (rgba u32[H,W], depth or None, coarse Counters, fine Counters). pixel_stride 3 = the
        reference's renderHalfSize preview (Main.cpp:101-106, 161)."""
        W, H = frame.width, frame.height
        rgba = np.zeros((H, W), np.uint32)
        depth = np.zeros(coarse_cells(W, H, frame.strips), np.float32) if want_depth else None
        cc, cf = Counters(), Counters()
        if threads is None:
            threads = os.cpu_count() or 1
        rc = self.lib.svo_oracle_render_frame_strided(words, C.byref(frame), int(pixel_stride), rgba.reshape(-1),
                                                      _null_or(depth), C.byref(cc), C.byref(cf), int(threads))
        if rc != 0:
            raise ValueError("svo_oracle_render_frame: bad arguments")
        return rgba, depth, cc, cf

    def shade(self, material, ray, light):
        return np.float32(self.lib.svo_oracle_shade(int(material), np.ascontiguousarray(ray, np.float32),
                                                    np.ascontiguousarray(light, np.float32)))

    def pack(self, v):
        return int(self.lib.svo_oracle_pack(float(v)))

    def decompress_material(self, word):
        n = np.zeros(3, np.float32)
        s = C.c_float()
        self.lib.svo_oracle_decompress_material(int(word), n, C.byref(s))
        return n, np.float32(s.value)

    def inv_sqrt(self, x):
        return np.float32(self.lib.svo_oracle_inv_sqrt(float(x)))

    def build_octree(self, voxels):
        """Row f2. voxels: uint32[D, H, W] (x fastest, 0 = empty) -> (words, center), VoxelOctree.cpp:125-205."""
        voxels = np.ascontiguousarray(voxels, np.uint32)
        d, h, w = voxels.shape
        n = C.c_uint64(0)
        center = np.zeros(3, np.float32)
        ptr = self.lib.svo_oracle_build_octree(voxels.reshape(-1), w, h, d, C.byref(n), center)
        try:
            return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy(), center
        finally:
            self.lib.svo_oracle_free(ptr)

    def voxelize_ply(self, path, resolution, threads):
        """Row f3 (oracle only so far): PlyLoader(path) + VoxelData(loader, resolution, mem) with the volume in one
        cache block (PlyLoader.cpp:64-474). -> (voxels uint32[D, H, W], triangle count)."""
        dims = (C.c_int * 3)()
        ntri = C.c_uint64(0)
        ptr = self.lib.svo_oracle_voxelize_ply(str(path).encode(), int(resolution), int(threads), C.byref(dims), C.byref(ntri))
        if not ptr:
            raise ValueError(f"svo_oracle_voxelize_ply: cannot read {path}")
        try:
            w, h, d = dims[0], dims[1], dims[2]
            return np.ctypeslib.as_array(ptr, shape=(d, h, w)).copy(), int(ntri.value)
        finally:
            self.lib.svo_oracle_free(ptr)

    def block_list_aliases(self, path, resolution, block_edge, threads):
        """PlyLoader's block lists for a cubic cache block of `block_edge` cells (0: one block) and a pool of `threads`:
        -> (listings outside the sub-block grid -- flat indices that alias another block, PlyLoader.cpp:263 --,
        rejected candidates outside it, (gridW, gridH, gridD), sub-blocks that exist per axis)."""
        cand = C.c_uint64(0)
        grid, real = (C.c_int * 3)(), (C.c_int * 3)()
        n = self.lib.svo_oracle_block_list_aliases(str(path).encode(), int(resolution), int(block_edge), int(threads),
                                                   C.byref(cand), C.byref(grid), C.byref(real))
        if n < 0:
            raise ValueError(f"svo_oracle_block_list_aliases: cannot read {path}")
        return int(n), int(cand.value), tuple(grid), tuple(real)

    def ply_triangles(self, path):
        """-> (float32[n, 33], lower[3], upper[3]): PlyLoader's triangle list after its constructor."""
        n = C.c_uint64(0)
        lo, hi = np.zeros(3, np.float32), np.zeros(3, np.float32)
        ptr = self.lib.svo_oracle_ply_triangles(str(path).encode(), C.byref(n), lo, hi)
        if not ptr:
            raise ValueError(f"svo_oracle_ply_triangles: cannot read {path}")
        try:
            return np.ctypeslib.as_array(ptr, shape=(n.value, 33)).copy(), lo, hi
        finally:
            self.lib.svo_oracle_free(ptr)

    def tree_to_volume(self, words, side, dims):
        """Material words of a node array spanning side^3 voxels as a dense (d, h, w) volume."""
        w, h, d = dims
        vol = np.zeros((d, h, w), np.uint32)
        self.lib.svo_oracle_tree_to_volume(np.ascontiguousarray(words, np.uint32), int(side), vol.reshape(-1), w, h, d)
        return vol

    def tree_walk(self, words):
        st = TreeStats()
        rc = self.lib.svo_oracle_tree_walk(words, words.size, C.byref(st))
        if rc != 0:
            raise ValueError("malformed node array")
        return st


def pixel_rays(frame: "Frame"):
    """One primary ray per pixel, no beam pass (Main.cpp:97-113 with tile origin (0,0) replaced by
    direct per-pixel running sums from the image origin): origin = pos, dir = normalised.
    Used for batch-mode (K1) parity; not a frame-loop restatement."""
    W, H = frame.width, frame.height
    f32 = np.float32
    scale = f32(frame.scale)
    dx = np.empty(W, f32)
    acc = f32(-1.0)
    for x in range(W):
        dx[x] = acc
        acc = f32(acc + scale)
    dy = np.empty(H, f32)
    acc = f32(frame.aspect)
    for y in range(H):
        dy[y] = acc
        acc = f32(acc - scale)
    DX, DY = np.meshgrid(dx, dy)
    comps = []
    for (a1, a2, z) in ((frame.a11, frame.a12, frame.zx), (frame.a21, frame.a22, frame.zy),
                        (frame.a31, frame.a32, frame.zz)):
        comps.append((DX * f32(a1) + DY * f32(a2)) + f32(z))
    d = np.stack(comps, -1).astype(f32)
    len2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    i = (np.uint32(0x5f3759df) - (len2.view(np.uint32) >> np.uint32(1))).astype(np.uint32)
    y = i.view(f32)
    y = y * (f32(1.5) - (len2 * f32(0.5)) * y * y)
    d = (d * y[..., None]).astype(f32).reshape(-1, 3)
    o = np.broadcast_to(np.array(list(frame.pos), f32), d.shape).copy()
    return o, d
