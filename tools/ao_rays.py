"""Deterministic ambient-occlusion ray generator (BASELINE.json configs[3]).

For every primary hit: hit point p = o + d*t, normal n from the leaf's material word
(reference src/Util.hpp:86-100 decode), `spp` directions per pixel from a counter-based integer hash
(seed 1234, pixel, sample) mapped to the hemisphere around n with +, *, /, sqrt only, origin p + n*2^-11.
numpy float32; the rays are INPUTS to both the GPU path and the oracle, so only determinism matters here.
"""
import numpy as np

SEED = 1234


def _hash(a):
    a = a.astype(np.uint32)
    a ^= a >> np.uint32(16)
    a *= np.uint32(0x7FEB352D)
    a ^= a >> np.uint32(15)
    a *= np.uint32(0x846CA68B)
    a ^= a >> np.uint32(16)
    return a


def decode_normals(words):
    words = words.astype(np.uint32)
    sign = np.where(words >> np.uint32(31), np.float32(-1), np.float32(1)).astype(np.float32)
    face = ((words >> np.uint32(29)) & np.uint32(3)).astype(np.int64) % 3
    u = ((words >> np.uint32(18)) & np.uint32(0x7FF)).astype(np.float32) * np.float32(4.8852e-4) * np.float32(2) - np.float32(1)
    v = ((words >> np.uint32(7)) & np.uint32(0x7FF)).astype(np.float32) * np.float32(4.8852e-4) * np.float32(2) - np.float32(1)
    n = np.zeros((words.size, 3), np.float32)
    idx = np.arange(words.size)
    n[idx, face] = sign
    n[idx, (face + 1) % 3] = u
    n[idx, (face + 2) % 3] = v
    n /= np.sqrt((n * n).sum(1, keepdims=True)).astype(np.float32)
    return n


def ao_rays(o, d, t, normal_words, hit_mask, spp=16):
    """o, d: (n,3) primary rays; t: (n,) hit distances; normal_words: (n,) material words; hit_mask: (n,) bool.
    Returns (origins (m*spp,3) f32, directions (m*spp,3) f32, pixel index (m*spp,) i64)."""
    pix = np.nonzero(hit_mask)[0]
    p = (o[pix] + d[pix] * t[pix, None]).astype(np.float32)
    n = decode_normals(normal_words[pix])
    # orient the normal against the viewing direction
    flip = (n * d[pix]).sum(1) > 0
    n[flip] = -n[flip]
    org = (p + n * np.float32(2.0 ** -11)).astype(np.float32)
    # tangent frame with +,*,/,sqrt only
    a = np.where(np.abs(n[:, :1]) > 0.5, np.array([[0, 1, 0]], np.float32), np.array([[1, 0, 0]], np.float32))
    tx = np.cross(n, a).astype(np.float32)
    tx /= np.sqrt((tx * tx).sum(1, keepdims=True)).astype(np.float32)
    ty = np.cross(n, tx).astype(np.float32)
    outs_o, outs_d, outs_p = [], [], []
    for s in range(spp):
        salt = np.uint32((SEED + s * 0x9E3779B9) & 0xFFFFFFFF)
        h1 = _hash(pix.astype(np.uint32) * np.uint32(2654435761) + salt)
        h2 = _hash(h1 + np.uint32(0x85EBCA6B))
        u1 = (h1 >> np.uint32(8)).astype(np.float32) / np.float32(1 << 24)
        u2 = (h2 >> np.uint32(8)).astype(np.float32) / np.float32(1 << 24)
        # uniform point in the unit disc by rejection-free square -> disc (concentric, algebraic form)
        x = np.float32(2) * u1 - np.float32(1)
        y = np.float32(2) * u2 - np.float32(1)
        r2 = np.minimum(x * x + y * y, np.float32(0.999))
        # cosine-weighted: lift the (clamped) point onto the hemisphere
        scale = np.where(x * x + y * y > 0.999, np.sqrt(np.float32(0.999) / np.maximum(x * x + y * y, np.float32(1e-12))), np.float32(1)).astype(np.float32)
        x, y = x * scale, y * scale
        z = np.sqrt(np.float32(1) - r2).astype(np.float32)
        dirs = (tx * x[:, None] + ty * y[:, None] + n * z[:, None]).astype(np.float32)
        outs_o.append(org)
        outs_d.append(dirs)
        outs_p.append(pix)
    return (np.ascontiguousarray(np.stack(outs_o, 1).reshape(-1, 3)),
            np.ascontiguousarray(np.stack(outs_d, 1).reshape(-1, 3)),
            np.stack(outs_p, 1).reshape(-1))
