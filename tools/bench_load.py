#!/usr/bin/env python
"""Row f1: start-up time of a tree, .oct file -> node array resident in HBM.

    python tools/bench_load.py [scene] [--threads 1,4,16]

Times svo_oct_read (decode only) and svo_tree_load_oct (decode pipelined with the host->device copy)
for two files holding the same words: one written by the reference's VoxelOctree::save (slices chained
through LZ4's streaming window: decode is serial, only reads / page faults / uploads overlap) and one
written by this library (independent slices: decode is parallel). Prints one JSON object.
The reference's own loader (oracle/_ref, single-threaded, reference src/VoxelOctree.cpp:57-90) is timed
beside them as the CPU baseline of this row.
"""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "sparse-voxel-octrees_b200"))

import numpy as np  # noqa: E402

import pysvo  # noqa: E402
from tools import make_scenes  # noqa: E402


def best(fn, repeat=2):
    times = []
    for _ in range(repeat):
        t = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t)
    return min(times)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene", nargs="?", default="ico8192")
    ap.add_argument("--threads", default="1,4,16")
    ap.add_argument("--dir", default=None)
    a = ap.parse_args()
    path = make_scenes.scene_path(a.scene)
    if Path(path).exists():
        words, center = pysvo.oct_read(path)
    else:
        words, center = make_scenes.unpack_scene(a.scene)
    out = {"scene": a.scene, "words": int(words.size), "bytes": int(words.nbytes), "host_cores": os.cpu_count()}
    tmp = Path(a.dir or tempfile.mkdtemp(prefix="svo_load_"))
    ours, theirs = tmp / "ours.oct", tmp / "theirs.oct"
    t = time.perf_counter()
    pysvo.oct_write(ours, words, center, compress=True)
    out["write_ours_s"] = round(time.perf_counter() - t, 3)
    out["ours_bytes"] = ours.stat().st_size
    ref = None
    try:
        from oracle.pyoracle import Ref
        ref = Ref()
        h = ref.tree_from_words(words, center)
        t = time.perf_counter()
        ref.tree_save(h, theirs)
        out["write_reference_s"] = round(time.perf_counter() - t, 3)
        ref.tree_destroy(h)
        out["theirs_bytes"] = theirs.stat().st_size
    except (FileNotFoundError, OSError) as e:
        out["reference"] = f"unavailable: {e}"
    xor = int(np.bitwise_xor.reduce(words))
    del words
    have_gpu = pysvo.device_count() > 0
    if have_gpu:
        pysvo.VoxelOctree(ROOT / "tests" / "golden" / "XYZRGB-Dragon.oct").close()   # context creation is not load time
    rows = []
    for name, p in (("reference-written", theirs), ("library-written", ours)):
        if not p.exists():
            continue
        for th in a.threads.split(","):
            os.environ["SVO_IO_THREADS"] = th
            row = {"file": name, "threads": int(th)}

            def read():
                w, _ = pysvo.oct_read(p)
                assert int(np.bitwise_xor.reduce(w)) == xor

            row["oct_read_s"] = round(best(read), 3)
            if have_gpu:
                def load():
                    pysvo.VoxelOctree(p).close()
                row["tree_load_oct_s"] = round(best(load), 3)
                row["GB_per_s_into_hbm"] = round(out["bytes"] / row["tree_load_oct_s"] / 1e9, 2)
            rows.append(row)
    out["loads"] = rows
    if ref is not None and theirs.exists():
        def ref_load():
            ref.tree_destroy(ref.tree_load(theirs))
        out["reference_loader_s"] = round(best(ref_load), 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
