"""pysvo -- thin ctypes binding of libsvo_b200.so (C ABI: include/svo_b200.h).

This is plumbing for the tests, the benchmark and Python users; the product is
the CUDA library. There is no fallback of any kind: if the library is not
built, or there is no CUDA device, calls raise.

The class surface mirrors the reference's ``VoxelOctree`` (reference
src/VoxelOctree.hpp:48-57): ``VoxelOctree(path)``, ``save(path)``,
``raymarch(o, d, rayScale)``, ``center()``, plus the batched and per-frame
entry points the GPU needs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent.parent
LIB_PATH = PKG_DIR / "libsvo_b200.so"
if os.environ.get("PYSVO_LIB"):      # A/B experiments only: another build of the same library (same ABI)
    LIB_PATH = Path(os.environ["PYSVO_LIB"]).resolve()

FLAVOUR_VALIDATION = 0
FLAVOUR_FAST = 1
BATCH_COHERENCE_ORDER = 0x100   # OR into the flavour of raymarch_batch[_device] for incoherent rays
BATCH_LANE_REFILL = 0x200       # ... persistent warps, a lane that finishes its ray takes the next one
MISS, HIT_LEAF, HIT_LOD = 0, 1, 2
T_MISS = np.float32(1e10)
VOXEL_NONE = np.uint64(0xFFFFFFFFFFFFFFFF)
IPC_HANDLE_BYTES = 64

_STATUS = {0: "SVO_OK", 1: "SVO_ERR_INVALID_ARGUMENT", 2: "SVO_ERR_IO", 3: "SVO_ERR_FORMAT",
           4: "SVO_ERR_OUT_OF_MEMORY", 5: "SVO_ERR_CUDA", 6: "SVO_ERR_NO_DEVICE", 7: "SVO_ERR_UNSUPPORTED"}


class SvoError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{_STATUS.get(status, status)}: {message}")
        self.status = status


class Camera(C.Structure):
    _fields_ = [("model", C.c_float * 16), ("view", C.c_float * 16)]

    @classmethod
    def from_matrices(cls, model, view):
        cam = cls()
        cam.model[:] = [float(x) for x in np.asarray(model, np.float32).reshape(16)]
        cam.view[:] = [float(x) for x in np.asarray(view, np.float32).reshape(16)]
        return cam


class ViewerEvent(C.Structure):
    _fields_ = [("type", C.c_int32), ("code", C.c_int32), ("dx", C.c_int32), ("dy", C.c_int32)]


class ViewerState(C.Structure):
    """svo_viewer_state: the reference viewer's camera control (Events.cpp + Main.cpp:229-252)."""
    _fields_ = [("radius", C.c_float), ("pitch", C.c_float), ("yaw", C.c_float),
                ("mouse_down", C.c_int32 * 2), ("mouse_dx", C.c_int32), ("mouse_dy", C.c_int32),
                ("escape_down", C.c_int32), ("preview", C.c_int32), ("quit", C.c_int32), ("camera", Camera)]


EVENT_KEY_DOWN, EVENT_KEY_UP, EVENT_MOUSE_MOTION, EVENT_BUTTON_DOWN, EVENT_BUTTON_UP = 2, 3, 4, 5, 6
BUTTON_LEFT, BUTTON_RIGHT, KEY_ESCAPE = 1, 3, 27
VIEWER_WAIT, VIEWER_FRAME, VIEWER_QUIT = 0, 1, 2


class FrameConstants(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("strips", C.c_int32), ("tile_size", C.c_int32),
                ("pos", C.c_float * 3),
                ("a11", C.c_float), ("a12", C.c_float), ("a21", C.c_float), ("a22", C.c_float),
                ("a31", C.c_float), ("a32", C.c_float),
                ("zx", C.c_float), ("zy", C.c_float), ("zz", C.c_float),
                ("scale", C.c_float), ("tile_scale", C.c_float), ("coarse_scale", C.c_float), ("aspect", C.c_float),
                ("light", C.c_float * 3), ("beam_bias", C.c_float)]

    def as_array(self):
        return np.frombuffer(bytes(self), dtype=np.float32, offset=16).copy()


class WordsReport(C.Structure):
    _fields_ = [("descriptors", C.c_uint64), ("leaves", C.c_uint64), ("far_words", C.c_uint64),
                ("min_leaf_depth", C.c_uint32), ("max_leaf_depth", C.c_uint32), ("depth", C.c_uint32), ("reserved", C.c_uint32)]


class FrameDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("strips", C.c_int32), ("flavour", C.c_int32),
                ("tile_rank", C.c_int32), ("tile_world", C.c_int32), ("pixel_stride", C.c_int32), ("pixel_format", C.c_int32)]


class FrameStats(C.Structure):
    _fields_ = [("coarse_rays", C.c_uint64), ("fine_rays", C.c_uint64), ("tiles_rendered", C.c_uint64),
                ("tiles_total", C.c_uint64), ("kernel_launches", C.c_uint32), ("reserved", C.c_uint32),
                ("coarse_ms", C.c_float), ("fine_ms", C.c_float)]

    @property
    def rays(self):
        return int(self.coarse_rays + self.fine_rays)


class SequenceStats(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("coarse_rays", C.c_uint64), ("fine_rays", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("device_ms", C.c_float), ("wall_ms", C.c_float),
                ("lanes", C.c_int32), ("tile_run", C.c_int32)]

    @property
    def rays(self):
        return int(self.coarse_rays + self.fine_rays)


FRAME_CALLBACK = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_void_p)
OUTPUT_DEVICE, OUTPUT_HOST = 0, 1
PIXELS_RGBA8, PIXELS_GREY8A8 = 0, 1     # svo_pixel_format of host frames (MultiOctree.render_frame / render_sequence)


class FrameLayout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_strips", "strip_rows", "tiles_x", "tiles_y_full", "tiles_y_last",
                                         "tile_cols", "tiles", "corners")]


class TreeInfo(C.Structure):
    _fields_ = [("n_words", C.c_uint64), ("center", C.c_float * 3), ("depth", C.c_uint32), ("device", C.c_int32),
                ("device_bytes", C.c_uint64)]


class BuildStats(C.Structure):
    _fields_ = [("voxels", C.c_uint64), ("nodes", C.c_uint64), ("far_blocks", C.c_uint64), ("words", C.c_uint64),
                ("gather_ms", C.c_float), ("sort_ms", C.c_float), ("levels_ms", C.c_float), ("emit_ms", C.c_float)]


class VoxelizeStats(C.Structure):
    _fields_ = [("triangles", C.c_uint64), ("cell_records", C.c_uint64), ("voxels", C.c_uint64), ("dims", C.c_int32 * 3),
                ("cache_block", C.c_int32), ("sub_block", C.c_int32 * 3), ("large_triangles", C.c_int32),
                ("overlap_ms", C.c_float), ("sort_ms", C.c_float), ("fold_ms", C.c_float), ("reserved2", C.c_float)]


def build_library(force: bool = False) -> Path:
    """make -C sparse-voxel-octrees_b200 (nvcc, sm_100a). Cross-compiles without a GPU."""
    args = ["make", "-C", str(PKG_DIR), "-j8"]
    if force:
        args.insert(1, "-B")
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """The loaded library. Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with `make -C {PKG_DIR}` (or __graft_entry__.build()). "
            "pysvo has no CPU or PyTorch fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, u64, i32, f32 = C.c_void_p, C.c_uint64, C.c_int, C.c_float
    P = C.POINTER
    sigs = {
        "svo_abi_version": (i32, []),
        "svo_last_error": (C.c_char_p, []),
        "svo_device_count": (i32, [P(i32)]),
        "svo_free": (None, [vp]),
        "svo_words_validate": (i32, [vp, u64, P(WordsReport)]),
        "svo_host_alloc": (i32, [C.c_size_t, P(vp)]),
        "svo_host_register": (i32, [i32, vp, C.c_size_t, P(vp)]),
        "svo_host_unregister": (i32, [vp]),
        "svo_frame_copy_owned_tiles": (i32, [i32, P(FrameDesc), vp, vp, vp]),
        "svo_host_free": (i32, [vp]),
        "svo_oct_read": (i32, [C.c_char_p, P(P(C.c_uint32)), P(u64), P(f32)]),
        "svo_oct_write": (i32, [C.c_char_p, vp, u64, P(f32), i32]),
        "svo_tree_create_from_words": (i32, [vp, u64, P(f32), i32, P(vp)]),
        "svo_tree_load_oct": (i32, [C.c_char_p, i32, P(vp)]),
        "svo_tree_save_oct": (i32, [vp, C.c_char_p, i32]),
        "svo_tree_get_info": (i32, [vp, P(TreeInfo)]),
        "svo_tree_download_words": (i32, [vp, vp, u64]),
        "svo_tree_destroy": (i32, [vp]),
        "svo_tree_build_from_voxels": (i32, [vp, i32, i32, i32, i32, P(vp)]),
        "svo_tree_build_from_voxel_file": (i32, [C.c_char_p, i32, P(vp)]),
        "svo_tree_build_from_sparse": (i32, [vp, vp, u64, i32, i32, i32, i32, P(vp)]),
        "svo_build_last_stats": (i32, [P(BuildStats)]),
        "svo_tree_build_from_ply": (i32, [C.c_char_p, i32, u64, i32, i32, P(vp)]),
        "svo_ply_read_triangles": (i32, [C.c_char_p, P(P(f32)), P(u64), P(f32), P(f32)]),
        "svo_voxelize_last_stats": (i32, [P(VoxelizeStats)]),
        "svo_tree_extract_voxels": (i32, [vp, vp, vp, u64, P(u64)]),
        "svo_tree_rebuild": (i32, [vp, i32, i32, i32, P(vp)]),
        "svo_raymarch_batch": (i32, [vp, u64, vp, vp, f32, i32, vp, vp, vp, vp]),
        "svo_raymarch_batch_device": (i32, [vp, u64, vp, vp, f32, i32, vp, vp, vp, vp, vp]),
        "svo_raymarch": (i32, [vp, P(f32), P(f32), f32, P(C.c_uint32), P(f32), P(i32)]),
        "svo_shade_batch": (i32, [vp, u64, vp, vp, vp, P(f32), vp]),
        "svo_orbit_camera": (None, [f32, f32, f32, P(Camera)]),
        "svo_viewer_init": (None, [P(ViewerState)]),
        "svo_viewer_feed": (i32, [P(ViewerState), P(ViewerEvent)]),
        "svo_frame_constants_from_camera": (i32, [P(Camera), P(f32), i32, i32, i32, P(FrameConstants)]),
        "svo_frame_get_layout": (i32, [i32, i32, i32, P(FrameLayout)]),
        "svo_frame_tile_rect": (i32, [i32, i32, i32, i32, P(C.c_int32)]),
        "svo_frame_tile_owner": (i32, [i32, i32, i32, i32, i32]),
        "svo_frame_set_tile_run": (i32, [i32]),
        "svo_render_frame": (i32, [vp, P(Camera), P(FrameDesc), vp, vp, P(FrameStats)]),
        "svo_render_frame_device": (i32, [vp, P(Camera), P(FrameDesc), vp, vp, vp, P(FrameStats), i32]),
        "svo_render_frame_async": (i32, [vp, P(Camera), P(FrameDesc), vp, vp, i32, P(i32)]),
        "svo_frame_wait": (i32, [vp, P(FrameDesc), i32, P(FrameStats)]),
        "svo_device_alloc": (i32, [i32, C.c_size_t, P(vp)]),
        "svo_device_free": (i32, [i32, vp]),
        "svo_device_memset": (i32, [i32, vp, i32, C.c_size_t]),
        "svo_device_to_host": (i32, [i32, vp, vp, C.c_size_t]),
        "svo_host_to_device": (i32, [i32, vp, vp, C.c_size_t]),
        "svo_device_to_host_async": (i32, [i32, vp, vp, C.c_size_t, vp]),
        "svo_device_synchronize": (i32, [i32]),
        "svo_ipc_export": (i32, [i32, vp, vp]),
        "svo_ipc_open": (i32, [i32, vp, P(vp)]),
        "svo_ipc_close": (i32, [i32, vp]),
        "svo_multi_create_from_words": (i32, [vp, u64, P(f32), P(i32), i32, P(vp)]),
        "svo_multi_load_oct": (i32, [C.c_char_p, P(i32), i32, P(vp)]),
        "svo_multi_destroy": (i32, [vp]),
        "svo_multi_device_count": (i32, [vp]),
        "svo_multi_tree": (vp, [vp, i32]),
        "svo_multi_render_sequence": (i32, [vp, P(Camera), i32, P(FrameDesc), i32, P(vp), i32, FRAME_CALLBACK, vp,
                                            P(SequenceStats)]),
        "svo_multi_render_frame": (i32, [vp, P(Camera), P(FrameDesc), vp, P(FrameStats)]),
        "svo_multi_device_frame": (i32, [vp, i32, P(vp)]),
        "svo_multi_raymarch_batch": (i32, [vp, u64, vp, vp, f32, i32, vp, vp, vp, vp]),
        "svo_pixels_expand_grey8a": (None, [vp, u64, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._svo_symbols = tuple(sigs)
    _lib = L
    return L


def _check(status):
    if status != 0:
        raise SvoError(status, lib().svo_last_error().decode(errors="replace"))


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in np.asarray(v, np.float32).reshape(3)])


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def device_count() -> int:
    n = C.c_int(0)
    st = lib().svo_device_count(C.byref(n))
    return int(n.value) if st == 0 else 0


# ---- .oct files (host only) -------------------------------------------------------------

def oct_read(path):
    """-> (words uint32[n], center float32[3]). Replaces VoxelOctree(const char*), VoxelOctree.cpp:57-90."""
    words = C.POINTER(C.c_uint32)()
    n = C.c_uint64(0)
    center = (C.c_float * 3)()
    _check(lib().svo_oct_read(str(path).encode(), C.byref(words), C.byref(n), center))
    try:
        arr = np.ctypeslib.as_array(words, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint32)
    finally:
        lib().svo_free(words)
    return arr, np.array(list(center), np.float32)


def ply_read_triangles(path):
    """-> (float32[n, 33], lower[3], upper[3]). Replaces PlyLoader(path) + tris(), PlyLoader.cpp:64-226."""
    tris = C.POINTER(C.c_float)()
    n = C.c_uint64(0)
    lo, hi = (C.c_float * 3)(), (C.c_float * 3)()
    _check(lib().svo_ply_read_triangles(str(path).encode(), C.byref(tris), C.byref(n), lo, hi))
    try:
        arr = np.ctypeslib.as_array(tris, shape=(n.value, 33)).copy()
    finally:
        lib().svo_free(tris)
    return arr, np.array(list(lo), np.float32), np.array(list(hi), np.float32)


def oct_write(path, words, center, compress=True):
    """Replaces VoxelOctree::save, VoxelOctree.cpp:92-123."""
    words = np.ascontiguousarray(words, np.uint32)
    _check(lib().svo_oct_write(str(path).encode(), _ptr(words), words.size, _f3(center), 1 if compress else 0))


# ---- camera ---------------------------------------------------------------------------------

def orbit_camera(pitch_deg, yaw_deg, radius) -> Camera:
    cam = Camera()
    lib().svo_orbit_camera(float(pitch_deg), float(yaw_deg), float(radius), C.byref(cam))
    return cam


def viewer_init() -> ViewerState:
    st = ViewerState()
    lib().svo_viewer_init(C.byref(st))
    return st


def viewer_feed(state: ViewerState, type_, code=0, dx=0, dy=0) -> int:
    """-> VIEWER_WAIT / VIEWER_FRAME / VIEWER_QUIT; `state` is updated in place."""
    ev = ViewerEvent(int(type_), int(code), int(dx), int(dy))
    r = lib().svo_viewer_feed(C.byref(state), C.byref(ev))
    if r < 0:
        _check(-r)
    return r


def frame_constants(cam: Camera, center, width, height, strips) -> FrameConstants:
    out = FrameConstants()
    _check(lib().svo_frame_constants_from_camera(C.byref(cam), _f3(center), width, height, strips, C.byref(out)))
    return out


# ---- pinned host memory ------------------------------------------------------------------------

class PinnedArray:
    """numpy view over page-locked host memory from svo_host_alloc."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(np.atleast_1d(shape))
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        _check(lib().svo_host_alloc(nbytes, C.byref(p)))
        self._ptr = p
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._ptr is not None:
            self.array = None
            lib().svo_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceBuffer:
    """Raw device allocation on `device` (svo_device_alloc); `.ptr` is an int."""

    def __init__(self, device, nbytes):
        self.device = int(device)
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        _check(lib().svo_device_alloc(self.device, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def zero(self):
        _check(lib().svo_device_memset(self.device, C.c_void_p(self.ptr), 0, self.nbytes))

    def to_host(self, dtype, count=None):
        dtype = np.dtype(dtype)
        count = self.nbytes // dtype.itemsize if count is None else count
        out = np.empty(count, dtype)
        _check(lib().svo_device_to_host(self.device, _ptr(out), C.c_void_p(self.ptr), count * dtype.itemsize))
        return out

    def from_host(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        _check(lib().svo_host_to_device(self.device, C.c_void_p(self.ptr), _ptr(arr), arr.nbytes))

    def ipc_export(self) -> bytes:
        h = (C.c_uint8 * IPC_HANDLE_BYTES)()
        _check(lib().svo_ipc_export(self.device, C.c_void_p(self.ptr), h))
        return bytes(h)

    def free(self):
        if self.ptr:
            lib().svo_device_free(self.device, C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def ipc_open(device, handle: bytes) -> int:
    h = (C.c_uint8 * IPC_HANDLE_BYTES).from_buffer_copy(handle)
    p = C.c_void_p()
    _check(lib().svo_ipc_open(int(device), h, C.byref(p)))
    return p.value


def ipc_close(device, ptr: int):
    _check(lib().svo_ipc_close(int(device), C.c_void_p(ptr)))


def device_to_host_async(device, host_array, device_ptr, nbytes, stream=0):
    _check(lib().svo_device_to_host_async(int(device), _ptr(host_array), C.c_void_p(device_ptr), int(nbytes),
                                          C.c_void_p(stream or None)))


def words_validate(words) -> WordsReport:
    """Full host-side check of a node array (svo_words_validate); raises SvoError(status 3) naming the first violation."""
    words = np.ascontiguousarray(words, np.uint32)
    rep = WordsReport()
    _check(lib().svo_words_validate(_ptr(words), words.size, C.byref(rep)))
    return rep


def frame_set_tile_run(run):
    """Width of the ranks' vertical stripes in tile columns (process-wide; <= 0 restores the default of 4)."""
    _check(lib().svo_frame_set_tile_run(int(run)))


def host_register(device, array) -> int:
    """Page-locks and maps an existing host array (e.g. a shared-memory segment) for `device`; returns the device
    address kernels and copies use. Undo with host_unregister(array) before the array goes away."""
    out = C.c_void_p()
    _check(lib().svo_host_register(int(device), _ptr(array), int(array.nbytes), C.byref(out)))
    return int(out.value)


def host_unregister(array):
    _check(lib().svo_host_unregister(_ptr(array)))


def frame_copy_owned_tiles(device, width, height, strips, tile_rank, tile_world, src_ptr, dst_ptr, stream=0):
    """The pixels of the tiles (tile_rank of tile_world) owns, from framebuffer src_ptr to framebuffer dst_ptr
    (device addresses: HBM, a peer mapping, or host memory mapped with host_register). Asynchronous on `stream`."""
    desc = FrameDesc(width, height, strips, FLAVOUR_VALIDATION, tile_rank, tile_world, 1)
    _check(lib().svo_frame_copy_owned_tiles(int(device), C.byref(desc), C.c_void_p(src_ptr), C.c_void_p(dst_ptr),
                                            C.c_void_p(stream or None)))


def device_synchronize(device=0):
    _check(lib().svo_device_synchronize(int(device)))


# ---- the tree -----------------------------------------------------------------------------------

class VoxelOctree:
    """GPU-resident octree with the reference's VoxelOctree surface (VoxelOctree.hpp:48-57)."""

    def __init__(self, path=None, *, words=None, center=None, device=0, _handle=None):
        h = C.c_void_p()
        if _handle is not None:
            h = _handle
        elif path is not None:
            _check(lib().svo_tree_load_oct(str(path).encode(), int(device), C.byref(h)))
        else:
            if words is None or center is None:
                raise ValueError("VoxelOctree needs a path, or words and center")
            words = np.ascontiguousarray(words, np.uint32)
            _check(lib().svo_tree_create_from_words(_ptr(words), words.size, _f3(center), int(device), C.byref(h)))
        self._h = h
        self.info = TreeInfo()
        _check(lib().svo_tree_get_info(self._h, C.byref(self.info)))
        self.device = int(self.info.device)

    # construction on the GPU: VoxelOctree(VoxelData*), VoxelOctree.cpp:125-205
    @classmethod
    def build_from_voxels(cls, voxels, device=0):
        """voxels: uint32[D, H, W] (x fastest, 0 = empty), host memory."""
        voxels = np.ascontiguousarray(voxels, np.uint32)
        d, hh, w = voxels.shape
        h = C.c_void_p()
        _check(lib().svo_tree_build_from_voxels(_ptr(voxels), w, hh, d, int(device), C.byref(h)))
        return cls(_handle=h)

    @classmethod
    def build_from_voxel_file(cls, path, device=0):
        """Raw .voxel file (VoxelData.cpp:36-48)."""
        h = C.c_void_p()
        _check(lib().svo_tree_build_from_voxel_file(str(path).encode(), int(device), C.byref(h)))
        return cls(_handle=h)

    @classmethod
    def build_from_ply(cls, path, resolution=256, mem_budget=0, threads=0, device=0):
        """PlyLoader + VoxelData(loader, resolution, mem) + VoxelOctree(VoxelData*), Main.cpp:320-325; threads =
        size of the reference's thread pool the result is to match (0 = this host's hardware threads)."""
        h = C.c_void_p()
        _check(lib().svo_tree_build_from_ply(str(path).encode(), int(resolution), int(mem_budget), int(threads),
                                             int(device), C.byref(h)))
        return cls(_handle=h)

    @staticmethod
    def last_voxelize_stats():
        st = VoxelizeStats()
        _check(lib().svo_voxelize_last_stats(C.byref(st)))
        return st

    @classmethod
    def build_from_sparse(cls, xyz, values, dims, device=0):
        """xyz: uint32[n, 3] (x, y, z); values: uint32[n]; dims = (w, h, d)."""
        xyz = np.ascontiguousarray(xyz, np.uint32).reshape(-1, 3)
        values = np.ascontiguousarray(values, np.uint32).reshape(-1)
        if xyz.shape[0] != values.shape[0]:
            raise ValueError("xyz and values differ in length")
        h = C.c_void_p()
        _check(lib().svo_tree_build_from_sparse(_ptr(xyz), _ptr(values), values.size, int(dims[0]), int(dims[1]),
                                                int(dims[2]), int(device), C.byref(h)))
        return cls(_handle=h)

    def extract_voxels(self):
        """-> (xyz uint32[n, 3], values uint32[n]) in Morton order."""
        n = C.c_uint64(0)
        _check(lib().svo_tree_extract_voxels(self._h, None, None, 0, C.byref(n)))
        xyz = np.empty((n.value, 3), np.uint32)
        values = np.empty(n.value, np.uint32)
        _check(lib().svo_tree_extract_voxels(self._h, _ptr(xyz), _ptr(values), n.value, C.byref(n)))
        return xyz, values

    def rebuild(self, dims=None):
        """build(extract(self)) in HBM; dims default to the full 2^depth cube."""
        side = 1 << self.depth
        w, hh, d = dims if dims is not None else (side, side, side)
        h = C.c_void_p()
        _check(lib().svo_tree_rebuild(self._h, int(w), int(hh), int(d), C.byref(h)))
        return VoxelOctree(_handle=h)

    @staticmethod
    def last_build_stats():
        st = BuildStats()
        _check(lib().svo_build_last_stats(C.byref(st)))
        return st

    # reference surface
    def save(self, path, compress=True):
        _check(lib().svo_tree_save_oct(self._h, str(path).encode(), 1 if compress else 0))

    def center(self):
        return np.array(list(self.info.center), np.float32)

    def raymarch(self, o, d, ray_scale=0.0, normal=0, t=0.0):
        """Single ray, reference semantics: returns (hit, normal, t) with normal / t passed through
        unchanged where the reference leaves them untouched."""
        n = C.c_uint32(int(normal))
        tt = C.c_float(float(t))
        hit = C.c_int(0)
        _check(lib().svo_raymarch(self._h, _f3(o), _f3(d), float(ray_scale), C.byref(n), C.byref(tt), C.byref(hit)))
        return bool(hit.value), int(n.value), np.float32(tt.value)

    # batched / per-frame surface
    @property
    def n_words(self):
        return int(self.info.n_words)

    @property
    def depth(self):
        return int(self.info.depth)

    def words(self):
        out = np.empty(self.n_words, np.uint32)
        _check(lib().svo_tree_download_words(self._h, _ptr(out), out.size))
        return out

    def raymarch_batch(self, o, d, ray_scale=0.0, flavour=FLAVOUR_VALIDATION, want_voxel=True, out=None):
        """HOST arrays in, HOST arrays out (copies inside). Returns dict(hit, t, normal, voxel)."""
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = o.shape[0]
        if out is None:
            out = dict(hit=np.empty(n, np.uint8), t=np.empty(n, np.float32), normal=np.empty(n, np.uint32),
                       voxel=np.empty(n, np.uint64) if want_voxel else None)
        _check(lib().svo_raymarch_batch(self._h, n, _ptr(o), _ptr(d), float(ray_scale), int(flavour),
                                        _ptr(out.get("hit")), _ptr(out.get("t")), _ptr(out.get("normal")),
                                        _ptr(out.get("voxel"))))
        return out

    def raymarch_batch_device(self, n, d_o, d_d, ray_scale, flavour, d_hit=0, d_t=0, d_normal=0, d_voxel=0, stream=0):
        """DEVICE pointers (ints); asynchronous on `stream`."""
        vp = C.c_void_p
        _check(lib().svo_raymarch_batch_device(self._h, int(n), vp(d_o), vp(d_d), float(ray_scale), int(flavour),
                                               vp(d_hit or None), vp(d_t or None), vp(d_normal or None),
                                               vp(d_voxel or None), vp(stream or None)))

    def shade_batch(self, normal, d, light, hit=None):
        """shade + pixel pack per ray (Main.cpp:81-90, 128-132) -> rgba uint32[n]."""
        normal = np.ascontiguousarray(normal, np.uint32).reshape(-1)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        if hit is not None:
            hit = np.ascontiguousarray(hit, np.uint8).reshape(-1)
        rgba = np.empty(normal.size, np.uint32)
        _check(lib().svo_shade_batch(self._h, normal.size, _ptr(hit), _ptr(normal), _ptr(d), _f3(light), _ptr(rgba)))
        return rgba

    def render_frame(self, cam: Camera, width, height, strips=16, flavour=FLAVOUR_VALIDATION, tile_rank=0,
                     tile_world=1, rgba=None, want_depth=False, want_stats=True, pixel_stride=1):
        """HOST buffers (copies inside, synchronous). Returns (rgba uint32[H,W], depth|None, FrameStats|None).
        pixel_stride=3 is the reference's renderHalfSize preview (Main.cpp:101-106, 161)."""
        desc = FrameDesc(width, height, strips, flavour, tile_rank, tile_world, pixel_stride)
        if rgba is None:
            rgba = np.empty((height, width), np.uint32)
        depth = np.empty(coarse_cells(width, height, strips), np.float32) if want_depth else None
        stats = FrameStats() if want_stats else None
        _check(lib().svo_render_frame(self._h, C.byref(cam), C.byref(desc), _ptr(rgba), _ptr(depth),
                                      C.byref(stats) if stats is not None else None))
        return rgba, depth, stats

    def render_frame_async(self, cam: Camera, width, height, rgba, strips=16, flavour=FLAVOUR_FAST, depth=None,
                           want_stats=False):
        """Pipelined host-buffer variant: returns a (desc, ticket) pair for frame_wait. Up to four in flight."""
        desc = FrameDesc(width, height, strips, flavour, 0, 1)
        ticket = C.c_int(0)
        _check(lib().svo_render_frame_async(self._h, C.byref(cam), C.byref(desc), _ptr(rgba), _ptr(depth),
                                            1 if want_stats else 0, C.byref(ticket)))
        return desc, ticket.value

    def frame_wait(self, pending, want_stats=False):
        desc, ticket = pending
        stats = FrameStats() if want_stats else None
        _check(lib().svo_frame_wait(self._h, C.byref(desc), ticket, C.byref(stats) if stats is not None else None))
        return stats

    def render_frame_device(self, cam: Camera, width, height, d_rgba, strips=16, flavour=FLAVOUR_FAST, tile_rank=0,
                            tile_world=1, d_depth=0, stream=0, want_stats=False, pixel_stride=1):
        """DEVICE framebuffer pointer (int; may be a peer mapping); asynchronous unless want_stats."""
        desc = FrameDesc(width, height, strips, flavour, tile_rank, tile_world, pixel_stride)
        stats = FrameStats() if want_stats else None
        vp = C.c_void_p
        _check(lib().svo_render_frame_device(self._h, C.byref(cam), C.byref(desc), vp(d_rgba), vp(d_depth or None),
                                             vp(stream or None), C.byref(stats) if stats is not None else None,
                                             1 if want_stats else 0))
        return stats

    def close(self):
        if getattr(self, "_h", None):
            if not getattr(self, "_borrowed", False):     # a MultiOctree's replica belongs to that handle
                lib().svo_tree_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiOctree:
    """svo_multi: the node array replicated on several GPUs of this node, driven from this one process (the
    reference's strip threads + frame barrier, Main.cpp:351-367, :217-219). A device may be listed twice."""

    def __init__(self, path=None, *, words=None, center=None, devices=(0,)):
        self._h = None
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        if path is not None:
            _check(lib().svo_multi_load_oct(str(path).encode(), devs, len(devices), C.byref(h)))
        else:
            words = np.ascontiguousarray(words, np.uint32)
            _check(lib().svo_multi_create_from_words(_ptr(words), words.size, _f3(center), devs, len(devices), C.byref(h)))
        self._h = h
        self.devices = tuple(int(d) for d in devices)

    @property
    def n_devices(self):
        return int(lib().svo_multi_device_count(self._h))

    def tree(self, index=0) -> "VoxelOctree":
        """The replica on devices[index] as a VoxelOctree that does not own its handle."""
        h = lib().svo_multi_tree(self._h, int(index))
        if not h:
            _check(1)
        t = VoxelOctree(_handle=C.c_void_p(h))
        t._borrowed = True
        return t

    def _desc(self, width, height, strips, flavour, pixel_stride=0, pixel_format=PIXELS_RGBA8):
        return FrameDesc(int(width), int(height), int(strips), int(flavour), 0, 1, int(pixel_stride), int(pixel_format))

    def render_frame(self, cam: Camera, width, height, strips=16, flavour=FLAVOUR_VALIDATION, rgba=None, pixel_stride=0,
                     pixel_format=PIXELS_RGBA8):
        """-> (rgba uint32[H, W] -- or (grey, alpha) byte pairs as uint16[H, W] for PIXELS_GREY8A8 --, FrameStats):
        renderBatch over all strips, all devices, into host memory."""
        if rgba is None:
            rgba = np.zeros((height, width), np.uint16 if pixel_format == PIXELS_GREY8A8 else np.uint32)
        st = FrameStats()
        d = self._desc(width, height, strips, flavour, pixel_stride, pixel_format)
        _check(lib().svo_multi_render_frame(self._h, C.byref(cam), C.byref(d), _ptr(rgba), C.byref(st)))
        return rgba, st

    def render_sequence(self, cams, width, height, strips=16, flavour=FLAVOUR_FAST, output=OUTPUT_DEVICE, host_frames=None,
                        on_frame=None, pixel_format=PIXELS_RGBA8):
        """Renders the camera path back to back (up to four frames in flight). OUTPUT_HOST: frame k lands in
        host_frames[k % len(host_frames)] (page-locked numpy arrays, e.g. PinnedArray.array); on_frame(k, array)
        is called for every finished frame. -> SequenceStats."""
        n = len(cams)
        arr = (Camera * n)(*cams)
        d = self._desc(width, height, strips, flavour, 0, pixel_format)
        st = SequenceStats()
        frames, nh = None, 0
        if output == OUTPUT_HOST:
            nh = len(host_frames)
            frames = (C.c_void_p * nh)(*[f.ctypes.data for f in host_frames])
        if on_frame is not None:
            def _cb(user, k, ptr, _frames=host_frames, _nh=nh):
                on_frame(int(k), _frames[int(k) % _nh])
            cb = FRAME_CALLBACK(_cb)
        else:
            cb = C.cast(None, FRAME_CALLBACK)
        _check(lib().svo_multi_render_sequence(self._h, arr, n, C.byref(d), int(output), frames, nh, cb, None, C.byref(st)))
        return st

    def device_frame(self, width, height, back=0):
        """Frame (last - back) of the last OUTPUT_DEVICE sequence, copied from devices[0]'s HBM -> uint32[H, W]."""
        p = C.c_void_p()
        _check(lib().svo_multi_device_frame(self._h, int(back), C.byref(p)))
        out = np.empty((height, width), np.uint32)
        _check(lib().svo_device_to_host(self.devices[0], _ptr(out), p, out.nbytes))
        return out

    def raymarch_batch(self, o, d, ray_scale=0.0, flavour=FLAVOUR_VALIDATION, want_voxel=True):
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = o.shape[0]
        out = dict(hit=np.zeros(n, np.uint8), t=np.zeros(n, np.float32), normal=np.zeros(n, np.uint32),
                   voxel=np.zeros(n, np.uint64) if want_voxel else None)
        _check(lib().svo_multi_raymarch_batch(self._h, n, _ptr(o), _ptr(d), float(ray_scale), int(flavour), _ptr(out["hit"]),
                                              _ptr(out["t"]), _ptr(out["normal"]), _ptr(out["voxel"])))
        return out

    def close(self):
        if self._h is not None:
            lib().svo_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def expand_grey8a(packed):
    """(grey, alpha) byte pairs (uint16 array, PIXELS_GREY8A8) -> the reference's RGBA words, svo_pixels_expand_grey8a."""
    packed = np.ascontiguousarray(packed, np.uint16)
    out = np.empty(packed.shape, np.uint32)
    lib().svo_pixels_expand_grey8a(_ptr(packed), packed.size, _ptr(out))
    return out


def frame_layout(width, height, strips) -> FrameLayout:
    out = FrameLayout()
    _check(lib().svo_frame_get_layout(width, height, strips, C.byref(out)))
    return out


def tile_rect(width, height, strips, tile):
    r = (C.c_int32 * 4)()
    _check(lib().svo_frame_tile_rect(width, height, strips, int(tile), r))
    return tuple(r)


def tile_owner(width, height, strips, tile, world):
    r = lib().svo_frame_tile_owner(width, height, strips, int(tile), int(world))
    if r < 0:
        raise SvoError(1, lib().svo_last_error().decode(errors="replace"))
    return r


def strip_layout(width, height, strips, tile=8):
    """Rows and corner-grid sizes of the strips that own at least one row (Main.cpp:351-362)."""
    stride = (height - 1) // strips + 1
    out = []
    for i in range(strips):
        y0 = i * stride
        if y0 >= height:
            break
        y1 = min(y0 + stride, height)
        out.append((y0, y1, (width - 1) // tile + 2, (y1 - y0 - 1) // tile + 2))
    return out


def coarse_cells(width, height, strips, tile=8):
    return sum(tx * ty for (_, _, tx, ty) in strip_layout(width, height, strips, tile))
