#!/usr/bin/env python
"""Golden vectors for row f4 (the viewer's camera control), minted from the REFERENCE'S OWN OBJECT CODE.

    python tests/golden/make_viewer_golden.py        (needs /root/reference to build oracle/_ref)

Scripted window sessions are played to the reference's own `-viewer` main loop (reference src/Main.cpp:332-376 with
renderLoop :204-258, src/Events.cpp, src/ThreadBarrier.cpp) behind the scripted SDL of oracle/ref_shim/SDL.h
(svoref_viewer_run); for every frame the viewer presents, the MODEL / VIEW matrices it was rendered with, renderHalfSize
and the number of script events consumed go to tests/golden/viewer_sessions.npz (committed), together with the first
session's 96 x 64 frames. tests/test_host.py checks svo_viewer_feed against them; the GPU tests the frames.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle.pyoracle import Ref  # noqa: E402

DRAGON = ROOT / "tests" / "golden" / "XYZRGB-Dragon.oct"
FRAME = (96, 64, 2)            # W, H, strips of the recorded frames


def sessions():
    """Event lists (type, code, xrel, yrel) in SDL 1.2's numbering; see test_host._viewer_script for the random ones."""
    from test_host import _viewer_script
    rng = np.random.default_rng(2024)
    out = [[(4, 0, 3, 2), (5, 1, 0, 0), (4, 0, 25, -10), (4, 0, 40, -35), (6, 1, 0, 0), (4, 0, 9, 9), (5, 3, 0, 0),
            (4, 0, 0, 30), (4, 0, 2, -45), (6, 3, 0, 0), (5, 1, 0, 0), (4, 0, -15, 120), (6, 1, 0, 0)]]
    out += [_viewer_script(rng, 120) for _ in range(2)]
    out.append([(5, 3, 0, 0)] + [(4, 0, 0, -70)] * 12 + [(4, 0, 0, 90)] * 6 + [(6, 3, 0, 0)])
    out.append([(5, 1, 0, 0)] + [(4, 0, 7, -50)] * 9 + [(4, 0, -300, 0), (6, 1, 0, 0), (4, 0, 5, 5), (2, 27, 0, 0), (4, 0, 1, 1)])
    return out


def main():
    ref = Ref()
    data = {"frame_shape": np.array(FRAME, np.int32)}
    for i, ev in enumerate(sessions()):
        W, H, S = FRAME if i == 0 else (16, 16, 2)
        run = ref.viewer_run(DRAGON, W, H, S, ev, want_pixels=i == 0)
        data[f"events_{i}"] = np.asarray(ev, np.int32)
        data[f"model_{i}"] = run["model"]
        data[f"view_{i}"] = run["view"]
        data[f"half_{i}"] = run["half"]
        data[f"taken_{i}"] = run["events_taken"]
        if i == 0:
            data["rgba_0"] = run["rgba"]
        print(f"session {i}: {len(ev)} events -> {len(run['half'])} frames, {int(run['half'].sum())} of them previews")
    np.savez_compressed(ROOT / "tests" / "golden" / "viewer_sessions.npz", **data)


if __name__ == "__main__":
    main()
