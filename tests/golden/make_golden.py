#!/usr/bin/env python
"""Generates the committed golden vectors from the REFERENCE'S OWN OBJECT CODE.

Run in the build container (needs /root/reference to build oracle/_ref/libsvo_ref.so):

    python tests/golden/make_golden.py

Outputs (committed):
    tests/golden/dragon_pins.json   header / tree-walk / full-resolution hash pins
    tests/golden/dragon_small.npz   full per-pixel outputs at 160x90 (3 cameras)

The reference has no tests or golden vectors of its own; these are minted by
running its unmodified raymarch (src/VoxelOctree.cpp:207-346) and renderBatch (src/Main.cpp:139-202)
on its sample tree models/XYZRGB-Dragon.oct (copied to tests/golden/ as a data fixture).
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import Port, Ref, pixel_rays  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
CAMERAS = [(0.0, 0.0, 1.0), (20.0, 135.0, 0.5), (0.0, 0.0, 2.5), (-35.0, 250.0, 0.8), (89.0, 10.0, 0.3)]
T_MISS = np.float32(1e10)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = Ref()
    port = Port()  # only for the frame-constant struct that feeds pixel_rays (pinned to ref by the frame hashes)
    h = ref.tree_load(GOLDEN / "XYZRGB-Dragon.oct")
    words = ref.tree_words(h)
    center = ref.tree_center(h)

    pins = {
        "source": "reference object code (oracle/_ref/libsvo_ref.so), -O3 -DNDEBUG -ffp-contract=off",
        "n_words": int(words.size),
        "center": [float(x) for x in center],
        "root_word": int(words[0]),
        "words_sha256": sha(words),
        "cameras": [],
    }

    small = {}
    for ci, cam in enumerate(CAMERAS):
        model, view = ref.orbit_camera(*cam)
        entry = {"pitch_yaw_radius": list(cam), "model": [float(x) for x in model], "view": [float(x) for x in view]}

        # full-resolution frame, the reference's default configuration (Main.cpp:57-63)
        rgba, depth, _ = ref.render_frames(h, 1280, 720, 16, model, view, threads=8, want_depth=True)
        entry["frame_1280x720x16"] = {
            "rgba_sha256": sha(rgba), "depth_sha256": sha(depth),
            "lit": int(((rgba & 0xFFFFFF) != 0).sum()), "written": int((rgba != 0).sum()),
            "coarse_hits": int((depth < 1e9).sum()),
        }
        # a second strip count: the image depends on it
        rgba8, depth8, _ = ref.render_frames(h, 1280, 720, 8, model, view, threads=8, want_depth=True)
        entry["frame_1280x720x8"] = {"rgba_sha256": sha(rgba8), "depth_sha256": sha(depth8)}

        # full-resolution raw cast (one ray per pixel from the eye, no beam pass)
        f = port.frame_constants(model, view, center, 1280, 720, 16)
        o, d = pixel_rays(f)
        hit, t, normal, _ = ref.raymarch_batch(h, o, d, 0.0, threads=8, normal_sentinel=0, t_sentinel=float(T_MISS))
        entry["batch_1280x720"] = {
            "rays_o_sha256": sha(o), "rays_d_sha256": sha(d),
            "hits": int(hit.sum()), "sum_t": float(t[hit > 0].astype(np.float64).sum()),
            "xor_normals": int(np.bitwise_xor.reduce(normal[hit > 0])) if hit.any() else 0,
            "hit_sha256": sha(hit), "t_sha256": sha(t), "normal_sha256": sha(normal),
        }
        hitL, tL, normalL, _ = ref.raymarch_batch(h, o, d, f.coarse_scale, threads=8, normal_sentinel=0,
                                                 t_sentinel=float(T_MISS))
        # normal is left untouched on LOD exits: force 0 there is NOT possible from outside, so pin hit/t only
        entry["batch_lod_1280x720"] = {"hits": int(hitL.sum()), "hit_sha256": sha(hitL), "t_sha256": sha(tL)}
        pins["cameras"].append(entry)

        if ci in (0, 1, 3):
            W, H, S = 160, 90, 4
            rgbaS, depthS, _ = ref.render_frames(h, W, H, S, model, view, threads=4, want_depth=True)
            fS = port.frame_constants(model, view, center, W, H, S)
            oS, dS = pixel_rays(fS)
            hitS, tS, nS, _ = ref.raymarch_batch(h, oS, dS, 0.0, threads=4, normal_sentinel=0, t_sentinel=float(T_MISS))
            hitSL, tSL, _, _ = ref.raymarch_batch(h, oS, dS, fS.coarse_scale, threads=4, normal_sentinel=0,
                                                  t_sentinel=float(T_MISS))
            k = f"cam{ci}_"
            small[k + "model"] = model
            small[k + "view"] = view
            small[k + "rgba"] = rgbaS
            small[k + "depth"] = depthS
            small[k + "o"] = oS
            small[k + "d"] = dS
            small[k + "hit"] = hitS
            small[k + "t"] = tS
            small[k + "normal"] = nS
            small[k + "lod_hit"] = hitSL
            small[k + "lod_t"] = tSL
            small[k + "coarse_scale"] = np.float32(fS.coarse_scale)

    st = port.tree_walk(words)
    pins["tree_walk"] = {"descriptors": int(st.descriptors), "leaves": int(st.leaves), "far_words": int(st.far_words),
                         "far_blocks": int(st.far_blocks), "depth": int(st.depth),
                         "per_level": [int(x) for x in list(st.per_level)[:st.depth]]}

    (GOLDEN / "dragon_pins.json").write_text(json.dumps(pins, indent=1) + "\n")
    np.savez_compressed(GOLDEN / "dragon_small.npz", **small)
    print("wrote", GOLDEN / "dragon_pins.json", GOLDEN / "dragon_small.npz")


if __name__ == "__main__":
    main()
