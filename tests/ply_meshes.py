"""Small PLY meshes in the container variants the reference's loader can read (row f3): ASCII, binary little and
big endian, with and without vertex normals and uchar colours, triangles / quads / a pentagon (fans), an
ignored vertex property and an ignored extra element. (The product's reader also takes the newer type names
and 8-byte big-endian values, which the reference's plyfile crashes on.)"""
import struct

import numpy as np


def _uv_sphere(nu, nv, seed):
    rng = np.random.default_rng(seed)
    verts, normals, colors = [], [], []
    for i in range(nv + 1):
        th = np.pi * i / nv
        for j in range(nu):
            ph = 2 * np.pi * j / nu
            n = np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)])
            r = 1.0 + 0.15 * np.sin(3 * ph) * np.sin(2 * th) + 0.02 * rng.standard_normal()
            verts.append(n * r * np.array([1.0, 0.8, 0.6]) + np.array([3.0, -2.0, 0.5]))
            normals.append(n)
            colors.append(rng.integers(0, 256, 3))
    faces = []
    for i in range(nv):
        for j in range(nu):
            a, b = i * nu + j, i * nu + (j + 1) % nu
            c, d = (i + 1) * nu + (j + 1) % nu, (i + 1) * nu + j
            faces.append([a, b, c, d] if (i + j) % 3 else [a, b, c])
            if not (i + j) % 3:
                faces.append([a, c, d])
    faces.append([0, 1, 2, 3, 4])          # a pentagon on the degenerate pole row: fan of three
    return np.array(verts), np.array(normals), np.array(colors), faces


def write_variants(directory, seed=5):
    """-> list of (name, path)."""
    v, n, c, faces = _uv_sphere(24, 16, seed)
    out = []

    p = directory / "ascii_normals_colors.ply"
    with open(p, "w") as fp:
        fp.write("ply\nformat ascii 1.0\ncomment variants\nelement vertex %d\n" % len(v))
        fp.write("property float x\nproperty float y\nproperty float z\nproperty float confidence\n")
        fp.write("property float nx\nproperty float ny\nproperty float nz\n")
        fp.write("property uchar red\nproperty uchar green\nproperty uchar blue\n")
        fp.write("element face %d\nproperty list uchar int vertex_indices\nelement edge 2\nproperty int a\nproperty int b\nend_header\n" % len(faces))
        for i in range(len(v)):
            fp.write("%.9g %.9g %.9g 0.5 %.9g %.9g %.9g %d %d %d\n" % (*np.float32(v[i]), *np.float32(n[i]), *c[i]))
        for f in faces:
            fp.write("%d %s\n" % (len(f), " ".join(map(str, f))))
        fp.write("0 1\n1 2\n")
    out.append(("ascii_normals_colors", p))

    # (the reference's plyfile only byte-swaps what PlyLoader swaps afterwards: 4-byte values)
    p = directory / "be_float_colors_no_normals.ply"
    with open(p, "wb") as fp:
        fp.write(("ply\nformat binary_big_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                  "element face %d\nproperty list uchar int vertex_indices\nend_header\n" % (len(v), len(faces))).encode())
        for i in range(len(v)):
            fp.write(struct.pack(">fff", *v[i]))
        for f in faces:
            fp.write(struct.pack(">B%di" % len(f), len(f), *f))
    out.append(("be_float_colors_no_normals", p))

    p = directory / "le_float_normals.ply"
    with open(p, "wb") as fp:
        fp.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                  "property float nx\nproperty float ny\nproperty float nz\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n"
                  "element face %d\nproperty list uchar uint vertex_indices\nend_header\n" % (len(v), len(faces))).encode())
        for i in range(len(v)):
            fp.write(struct.pack("<ffffffBBB", *v[i], *n[i], *[int(x) for x in c[i]]))
        for f in faces:
            fp.write(struct.pack("<B%dI" % len(f), len(f), *f))
    out.append(("le_float_normals", p))
    return out


def write_lowpoly(path):
    """Nine triangles that are large at any resolution: a skewed, slightly rotated octahedron and one thin sliver across
    the whole volume (binary little-endian, positions only: face normals come from the loader)."""
    c, s_ = np.cos(0.3), np.sin(0.3)
    rot = np.array([[c, -s_, 0.0], [s_, c, 0.0], [0.0, 0.0, 1.0]]) @ np.array([[1.0, 0.0, 0.0], [0.0, np.cos(0.2), -np.sin(0.2)], [0.0, np.sin(0.2), np.cos(0.2)]])
    octa = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float64) * [1.0, 0.7, 0.85]
    verts = np.vstack([octa @ rot.T, [[-0.9, -0.6, -0.8], [0.95, 0.62, 0.7], [0.9, 0.66, 0.74]]]).astype(np.float32)
    faces = [(0, 2, 4), (2, 1, 4), (1, 3, 4), (3, 0, 4), (2, 0, 5), (1, 2, 5), (3, 1, 5), (0, 3, 5), (6, 7, 8)]
    with open(path, "wb") as fp:
        fp.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                  "element face %d\nproperty list uchar int vertex_indices\nend_header\n" % (len(verts), len(faces))).encode())
        for v in verts:
            fp.write(struct.pack("<fff", *v))
        for f in faces:
            fp.write(struct.pack("<Biii", 3, *f))
    return path
