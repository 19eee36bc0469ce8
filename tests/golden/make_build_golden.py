#!/usr/bin/env python
"""Golden vectors for row f2 (octree construction), minted from the REFERENCE'S OWN OBJECT CODE.

    python tests/golden/make_build_golden.py        (needs /root/reference to build oracle/_ref)

For every case of tests/test_oracle_pins.py::BUILD_CASES the seeded volume is written as a raw .voxel
file and built by the reference's VoxelData(path, mem) + VoxelOctree(VoxelData*) (reference
src/Main.cpp:318-319, src/VoxelOctree.cpp:125-205); length, centre and SHA-256 of the node array go to
tests/golden/build_pins.json (committed).
"""
import hashlib
import json
import struct
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle.pyoracle import Ref  # noqa: E402
from test_oracle_pins import BUILD_CASES, _volume  # noqa: E402


def main():
    ref = Ref()
    out = {"source": "reference object code (oracle/_ref/libsvo_ref.so): VoxelData(path, 1 GiB) + VoxelOctree(VoxelData*)",
           "builder": []}
    with tempfile.TemporaryDirectory() as tmp:
        for case in BUILD_CASES:
            vox = _volume(*case)
            d, h, w = vox.shape
            path = Path(tmp) / "v.voxel"
            with open(path, "wb") as fp:
                fp.write(struct.pack("<iii", w, h, d))
                fp.write(vox.tobytes())
            hnd = ref.tree_build_voxel_file(path, 1 << 30)
            words = ref.tree_words(hnd)
            center = ref.tree_center(hnd)
            ref.tree_destroy(hnd)
            out["builder"].append({"case": list(case), "n_words": int(words.size),
                                   "words_sha256": hashlib.sha256(words.tobytes()).hexdigest(),
                                   "center": [float(x) for x in center]})
    (ROOT / "tests" / "golden" / "build_pins.json").write_text(json.dumps(out, indent=1) + "\n")
    print("wrote tests/golden/build_pins.json")


if __name__ == "__main__":
    main()
