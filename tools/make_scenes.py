#!/usr/bin/env python
"""Builds the synthetic benchmark scenes with the REFERENCE'S OWN BUILDER and caches them as .oct.

BENCH TOOLING (not product): tools/scene_gen.c writes the builder's inputs (a sparse raw `.voxel`
volume for the SDF scene, a binary PLY for the displaced icosphere); the reference's
VoxelData / PlyLoader / VoxelOctree builder (reference src/Main.cpp:316-325 call patterns), run
through oracle/_ref/libsvo_ref.so, turns them into the node array; the reference's own
VoxelOctree::save writes the .oct into scenes/_cache/ (git-ignored, travels with gpurun).

    python tools/make_scenes.py sdf2048            # BASELINE.json configs[1]
    python tools/make_scenes.py ico8192            # BASELINE.json configs[2] (minutes, ~48 GB RAM budget optional)
    python tools/make_scenes.py sdf256 ico512      # small variants for tests
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
CACHE = ROOT / "scenes" / "_cache"
GEN_SO = ROOT / "tools" / "libscene_gen.so"
SEED = 1234

# name -> (kind, resolution, extra)
SCENES = {
    "sdf128": ("sdf", 128, None), "sdf256": ("sdf", 256, None), "sdf512": ("sdf", 512, None),
    "sdf1024": ("sdf", 1024, None), "sdf2048": ("sdf", 2048, None),
    "ico256": ("ico", 256, 22), "ico512": ("ico", 512, 44), "ico2048": ("ico", 2048, 177),
    "ico4096": ("ico", 4096, 354),
    "ico8192": ("ico", 8192, 707),   # 20*707^2 = 9,996,980 triangles
}


def gen_lib():
    src = ROOT / "tools" / "scene_gen.c"
    if not GEN_SO.exists() or GEN_SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", str(GEN_SO), str(src), "-lm"])
    L = C.CDLL(str(GEN_SO))
    L.svo_scene_sdf_voxel_file.restype = C.c_int64
    L.svo_scene_sdf_voxel_file.argtypes = [C.c_char_p, C.c_int, C.c_uint32]
    L.svo_scene_icosphere_ply.restype = C.c_int64
    L.svo_scene_icosphere_ply.argtypes = [C.c_char_p, C.c_int, C.c_uint32]
    L.svo_scene_compress_material.restype = C.c_uint32
    L.svo_scene_compress_material.argtypes = [C.POINTER(C.c_float), C.c_float]
    L.svo_scene_fbm.restype = C.c_double
    L.svo_scene_fbm.argtypes = [C.c_double, C.c_double, C.c_double, C.c_uint32]
    L.svo_scene_sdf.restype = C.c_double
    L.svo_scene_sdf.argtypes = [C.c_double, C.c_double, C.c_double, C.c_uint32]
    return L


def scene_path(name: str) -> Path:
    return CACHE / f"{name}.oct"


def builder_memory_budget() -> int:
    """VoxelData's `mem` (reference src/Main.cpp:269 uses 1 GiB). Larger => larger cache blocks => faster
    (>= 5.14 GiB buys 1024^3 blocks on the PLY path). Use a quarter of free RAM, capped."""
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 8 << 30
    return int(max(1 << 30, min(avail // 4, 12 << 30)))


def make_scene(name: str, scratch: str | None = None, verbose: bool = True) -> Path:
    from oracle.pyoracle import Ref
    kind, res, freq = SCENES[name]
    out = scene_path(name)
    CACHE.mkdir(parents=True, exist_ok=True)
    L = gen_lib()
    ref = Ref()
    mem = builder_memory_budget()
    t0 = time.time()
    with tempfile.TemporaryDirectory(prefix="svo_scene_", dir=scratch) as tmp:
        if kind == "sdf":
            raw = os.path.join(tmp, f"{name}.voxel")
            filled = L.svo_scene_sdf_voxel_file(raw.encode(), res, SEED)
            if filled < 0:
                raise RuntimeError("scene_gen: could not write the raw voxel file")
            t1 = time.time()
            h = ref.tree_build_voxel_file(raw, mem)
            meta = {"kind": "sdf sphere + fBm", "resolution": res, "filled_voxels": int(filled), "seed": SEED}
        else:
            ply = os.path.join(tmp, f"{name}.ply")
            tris = L.svo_scene_icosphere_ply(ply.encode(), freq, SEED)
            if tris < 0:
                raise RuntimeError("scene_gen: could not write the PLY")
            t1 = time.time()
            h = ref.tree_build_ply(ply, res, mem)
            meta = {"kind": "displaced icosphere PLY", "resolution": res, "triangles": int(tris), "seed": SEED}
        t2 = time.time()
        tmp_out = str(out) + ".tmp"
        ref.tree_save(h, tmp_out)          # the reference's own VoxelOctree::save
        meta.update(n_words=int(ref.lib.svoref_tree_word_count(h)), builder_mem=mem, builder_threads=ref.hardware_threads(),
                    seconds={"generate_input": round(t1 - t0, 2), "reference_builder": round(t2 - t1, 2),
                             "save": round(time.time() - t2, 2)})
        ref.tree_destroy(h)
        os.replace(tmp_out, out)
    out.with_suffix(".json").write_text(json.dumps(meta, indent=1) + "\n")
    if verbose:
        print(f"{name}: {meta}", file=sys.stderr)
    return out


# ---- transport: gpurun snapshots are capped at 512 MiB and the 8192^3 .oct is 749 MB (LZ4 only halves
# it), so big scenes also get a sidecar with the same words in independently xz-compressed chunks
# (~0.22 of raw). Pure transport tooling: the words are the reference builder's, bit for bit.
XZ_CHUNK_WORDS = 16 << 20


def packed_path(name: str) -> Path:
    return CACHE / f"{name}.words.xz"


def pack_scene(name: str, workers: int | None = None) -> Path:
    """scenes/_cache/<name>.oct -> <name>.words.xz (header JSON line, then length-prefixed xz chunks)."""
    import lzma
    import struct
    from concurrent.futures import ThreadPoolExecutor
    import numpy as np
    sys.path.insert(0, str(ROOT / "sparse-voxel-octrees_b200"))
    import pysvo
    words, center = pysvo.oct_read(scene_path(name))
    chunks = [words[i:i + XZ_CHUNK_WORDS] for i in range(0, words.size, XZ_CHUNK_WORDS)]
    with ThreadPoolExecutor(workers or os.cpu_count() or 1) as ex:
        blobs = list(ex.map(lambda c: lzma.compress(c.tobytes(), preset=1), chunks))
    out = packed_path(name)
    with open(str(out) + ".tmp", "wb") as fp:
        header = json.dumps({"n_words": int(words.size), "center": [float(x) for x in center],
                             "chunk_words": XZ_CHUNK_WORDS, "chunks": len(blobs)}).encode()
        fp.write(struct.pack("<I", len(header)) + header)
        for b in blobs:
            fp.write(struct.pack("<Q", len(b)) + b)
    os.replace(str(out) + ".tmp", out)
    return out


def unpack_scene(name: str, workers: int | None = None):
    """-> (words uint32[n], center float32[3]) from the xz sidecar."""
    import lzma
    import struct
    from concurrent.futures import ThreadPoolExecutor
    import numpy as np
    with open(packed_path(name), "rb") as fp:
        (hl,) = struct.unpack("<I", fp.read(4))
        header = json.loads(fp.read(hl))
        blobs = []
        for _ in range(header["chunks"]):
            (n,) = struct.unpack("<Q", fp.read(8))
            blobs.append(fp.read(n))
    words = np.empty(header["n_words"], np.uint32)
    cw = header["chunk_words"]

    def work(i):
        raw = lzma.decompress(blobs[i])
        words[i * cw:i * cw + len(raw) // 4] = np.frombuffer(raw, np.uint32)

    with ThreadPoolExecutor(workers or os.cpu_count() or 1) as ex:
        list(ex.map(work, range(len(blobs))))
    return words, np.array(header["center"], np.float32)


def scene_available(name: str) -> bool:
    return scene_path(name).exists() or packed_path(name).exists()


def load_scene(name: str, scratch: str | None = None):
    """-> (words, center): from the .oct, else from the xz sidecar, else built now with the reference builder."""
    sys.path.insert(0, str(ROOT / "sparse-voxel-octrees_b200"))
    import pysvo
    if scene_path(name).exists():
        return pysvo.oct_read(scene_path(name))
    if packed_path(name).exists():
        return unpack_scene(name)
    return pysvo.oct_read(make_scene(name, scratch, verbose=False))


def ensure_scene(name: str, scratch: str | None = None) -> Path:
    p = scene_path(name)
    if p.exists():
        return p
    return make_scene(name, scratch)


if __name__ == "__main__":
    args = sys.argv[1:] or ["sdf256"]
    if args[0] == "pack":
        for n in args[1:]:
            t0 = time.time()
            p = pack_scene(n)
            print(p, f"{p.stat().st_size / 1e6:.0f} MB in {time.time() - t0:.0f} s")
    else:
        for n in args:
            print(make_scene(n))
