"""Model: block-level compaction. Two warps per tile run until the tile's live rays fit one warp (<= 32, checked every
`period` trips), then warp 1's rays move into warp 0's free lanes (cost `move` warp instructions once) and one warp
finishes the tile."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from tools.simulate_divergence import trace, scheme_single_loop, P, A, Q, Lf, X
now = dict(H=13, V=8, P=26, A=14, Q=31, L=105, pro=150)   # current SASS section costs (loop 102: header 13, push 34 = V+P, advance 14, pop 31)
ops, counts = trace(sys.argv[1] if len(sys.argv) > 1 else "sdf2048", 3840, 2160, 16)
nt = ops.shape[0]//2
rays = int((counts > 0).sum())
t0, u0, _ = scheme_single_loop(ops, counts, now)
print(f"now: {t0/rays:.1f} warp-instr/ray, eff {u0/t0:.3f}")
def trip_cost(cur, running, c):
    isP, isL = running & (cur == P), running & (cur == Lf)
    isA = running & ((cur == A) | (cur == Q) | (cur == X)); isQ = running & ((cur == Q) | (cur == X))
    return c['H'] + (c['V'] if (isP | isL).any() else 0) + (c['P'] if isP.any() else 0) + (c['L'] if isL.any() else 0) + (c['A'] if isA.any() else 0) + (c['Q'] if isQ.any() else 0)
for period, move in ((1, 150), (4, 150), (8, 200), (8, 400)):
    total = 0.0
    for t in range(nt):
        o = ops[2*t:2*t+2]; n = counts[2*t:2*t+2].astype(np.int64)
        if n.max() == 0: continue
        total += 2*now['pro'] if (n[0].max() > 0 and n[1].max() > 0) else now['pro']
        T = int(n.max()); merged = False
        for k in range(T):
            live0, live1 = n[0] > k, n[1] > k
            if not merged and k % period == 0 and k > 0 and live0.any() and live1.any() and live0.sum() + live1.sum() <= 32:
                merged = True; total += move
            if merged:
                cur = np.concatenate([o[0][:, k], o[1][:, k]]); run = np.concatenate([live0, live1])
                if run.any(): total += trip_cost(cur, run, now)
            else:
                if live0.any(): total += trip_cost(o[0][:, k], live0, now)
                if live1.any(): total += trip_cost(o[1][:, k], live1, now)
    print(f"compaction check every {period} trips, move cost {move}: {total/rays:.1f} warp-instr/ray ({t0/total:.3f}x)")
