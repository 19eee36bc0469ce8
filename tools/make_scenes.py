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
        meta.update(n_words=int(ref.lib.svoref_tree_word_count(h)), builder_mem=mem,
                    seconds={"generate_input": round(t1 - t0, 2), "reference_builder": round(t2 - t1, 2),
                             "save": round(time.time() - t2, 2)})
        ref.tree_destroy(h)
        os.replace(tmp_out, out)
    out.with_suffix(".json").write_text(json.dumps(meta, indent=1) + "\n")
    if verbose:
        print(f"{name}: {meta}", file=sys.stderr)
    return out


def ensure_scene(name: str, scratch: str | None = None) -> Path:
    p = scene_path(name)
    if p.exists():
        return p
    return make_scene(name, scratch)


if __name__ == "__main__":
    names = sys.argv[1:] or ["sdf256"]
    for n in names:
        print(make_scene(n))
