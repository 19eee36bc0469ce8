"""Kernel design tool (no GPU): how many leading pushes the rays of a fine-pass warp share (oracle traces)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from tools.simulate_divergence import trace, P
for scene in ("sdf2048",):
    ops, counts = trace(scene, 3840, 2160, stride=16)
    nW = ops.shape[0]
    lead = []
    trips = []
    for w in range(nW):
        n = counts[w]
        if n.max() == 0: continue
        o = ops[w]
        # common leading pushes over live lanes
        live = n > 0
        k = 0
        while k < n[live].min() and (o[live, k] == P).all(): k += 1
        lead.append(k); trips.append(int(n.max()))
    lead = np.array(lead); trips = np.array(trips)
    print(scene, "warps", len(lead), "mean common leading pushes", lead.mean(), "mean warp trips", trips.mean(), "share", lead.sum()/trips.sum())
    print(np.bincount(lead))
