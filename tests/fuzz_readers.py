"""Mutation fuzzing of the two file readers of the product (.oct and PLY), run in a child process by
tests/test_host.py::test_readers_survive_mutated_files so that a crash shows up as a non-zero exit code.

    python tests/fuzz_readers.py <seed> <iterations> <work dir>

Truncations, byte flips, corrupted headers and counts, inserted garbage: every file must either decode or be rejected
with an svo_status (the reference does neither: no error channel, VoxelOctree.cpp:60, and plyfile aborts)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "sparse-voxel-octrees_b200"))
sys.path.insert(0, str(ROOT / "tests"))
import pysvo  # noqa: E402


def mutate(rng, src, header_bytes):
    b = bytearray(src)
    mode = int(rng.integers(0, 5))
    if mode == 0:
        b = b[:int(rng.integers(0, len(b)))]
    elif mode == 1:
        for _ in range(int(rng.integers(1, 12))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
    elif mode == 2:
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, header_bytes))] = int(rng.integers(0, 256))
    elif mode == 3:
        pos = int(rng.integers(0, len(b)))
        b[pos:pos] = bytes(rng.integers(0, 256, int(rng.integers(1, 50)), dtype=np.uint8))
    else:
        b = b.replace(b"element face ", b"element face 9", 1).replace(b"element vertex ", b"element vertex 9", 1)
    return bytes(b)


def main():
    seed, iterations, work = int(sys.argv[1]), int(sys.argv[2]), Path(sys.argv[3])
    rng = np.random.default_rng(seed)
    from ply_meshes import write_variants
    from tools import make_scenes
    plys = [p.read_bytes() for _, p in write_variants(work)]
    ico = work / "ico.ply"
    make_scenes.gen_lib().svo_scene_icosphere_ply(str(ico).encode(), 30, 7)
    plys.append(ico.read_bytes())
    dragon = ROOT / "tests" / "golden" / "XYZRGB-Dragon.oct"
    words, center = pysvo.oct_read(dragon)
    pysvo.oct_write(work / "ours.oct", words, center, compress=True)
    pysvo.oct_write(work / "literal.oct", words[:5000], center, compress=False)
    octs = [dragon.read_bytes(), (work / "ours.oct").read_bytes(), (work / "literal.oct").read_bytes()]
    counts = {"ply ok": 0, "ply rejected": 0, "oct ok": 0, "oct rejected": 0}
    for it in range(iterations):
        src = plys[it % len(plys)]
        path = work / "mutant.ply"
        path.write_bytes(mutate(rng, src, max(src.find(b"end_header"), 1)))
        try:
            pysvo.ply_read_triangles(path)
            counts["ply ok"] += 1
        except (pysvo.SvoError, MemoryError):
            counts["ply rejected"] += 1
        path = work / "mutant.oct"
        path.write_bytes(mutate(rng, octs[it % len(octs)], 40))
        try:
            pysvo.oct_read(path)
            counts["oct ok"] += 1
        except (pysvo.SvoError, MemoryError):
            counts["oct rejected"] += 1
    print("fuzz done", counts)


if __name__ == "__main__":
    main()
