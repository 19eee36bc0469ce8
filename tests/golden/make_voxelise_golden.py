#!/usr/bin/env python
"""Golden vectors for row f3 (triangle voxelisation), minted from the REFERENCE'S OWN OBJECT CODE.

    python tests/golden/make_voxelise_golden.py        (needs /root/reference to build oracle/_ref)

Every mesh of tests/test_oracle_pins.py::voxelise_cases goes through the reference's PlyLoader + VoxelData(loader,
resolution, 1 GiB) + VoxelOctree (reference src/Main.cpp:320-325, the in-memory -builder path) on this machine's thread
pool; word count, centre, SHA-256 of the node array and the POOL SIZE the run had (the voxels depend on it:
PlyLoader.cpp:381-440) go to tests/golden/voxelise_pins.json (committed). The meshes themselves are regenerated from
seeds by the test (tools/scene_gen.c icospheres, tests/ply_meshes.py variants)."""
import hashlib
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle.pyoracle import Ref  # noqa: E402
from test_oracle_pins import voxelise_cases  # noqa: E402


def main():
    ref = Ref()
    out = {"source": "reference object code (oracle/_ref/libsvo_ref.so): PlyLoader + VoxelData(loader, res, 1 GiB) + VoxelOctree",
           "pool_threads": ref.hardware_threads(), "cases": []}
    with tempfile.TemporaryDirectory() as tmp:
        for name, ply, res in voxelise_cases(Path(tmp)):
            h = ref.tree_build_ply(ply, res, 1 << 30)
            words, center = ref.tree_words(h), ref.tree_center(h)
            ref.tree_destroy(h)
            out["cases"].append({"name": name, "resolution": res, "words": int(words.size),
                                 "center": [float(c) for c in center],
                                 "sha256": hashlib.sha256(np.ascontiguousarray(words).tobytes()).hexdigest(),
                                 "ply_sha256": hashlib.sha256(Path(ply).read_bytes()).hexdigest()})
            print(name, res, words.size)
    (ROOT / "tests" / "golden" / "voxelise_pins.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
