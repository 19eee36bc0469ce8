#!/usr/bin/env python
"""Kernel design tool: replays per-ray loop traces from the instrumented oracle through warp-level cost
models of different loop organisations (no GPU needed). Costs are SASS instruction counts per section
read off the ncu source page of the current kernel.

    python tools/simulate_divergence.py [scene] [W H] [tile stride]
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "sparse-voxel-octrees_b200"))
import pysvo  # noqa: E402
from oracle.pyoracle import Port, Frame  # noqa: E402


def trace(scene="sdf2048", W=3840, H=2160, stride=64, cam=(20.0, 40.0, 0.9), max_ops=512):
    port = Port()
    path = ROOT / "scenes" / "_cache" / f"{scene}.oct" if scene != "dragon" else ROOT / "tests/golden/XYZRGB-Dragon.oct"
    words, center = pysvo.oct_read(path)
    m, v = port.orbit_camera(*cam)
    f = port.frame_constants(m, v, center, W, H, 16)
    L = port.lib
    L.svo_oracle_trace_fine_warps.restype = C.c_int64
    L.svo_oracle_trace_fine_warps.argtypes = [np.ctypeslib.ndpointer(np.uint32), C.POINTER(Frame), C.c_int,
                                              np.ctypeslib.ndpointer(np.uint8), C.c_uint32,
                                              np.ctypeslib.ndpointer(np.uint32), C.c_int64]
    max_warps = 20000
    ops = np.zeros((max_warps * 32, max_ops), np.uint8)
    counts = np.zeros(max_warps * 32, np.uint32)
    n = L.svo_oracle_trace_fine_warps(words, C.byref(f), stride, ops, max_ops, counts, max_warps)
    return ops[:n * 32].reshape(n, 32, max_ops), counts[:n * 32].reshape(n, 32)


P, A, Q, Lf, X = ord('P'), ord('A'), ord('Q'), ord('L'), ord('X')


def scheme_single_loop(ops, counts, c):
    """Current kernel: every trip all live lanes run the header, then the union of the paths they need."""
    total = 0.0
    lanes_useful = 0.0
    trips = 0
    for w in range(ops.shape[0]):
        n = counts[w]
        T = int(n.max())
        if T == 0:
            continue
        o = ops[w][:, :T]
        live = np.arange(T)[None, :] < n[:, None]
        isP = (o == P) & live
        isL = (o == Lf) & live
        isA = ((o == A) | (o == Q) | (o == X)) & live
        isQ = ((o == Q) | (o == X)) & live
        anyP, anyL, anyA, anyQ = isP.any(0), isL.any(0), isA.any(0), isQ.any(0)
        total += T * c['H'] + (anyP | anyL).sum() * c['V'] + anyP.sum() * c['P'] + anyL.sum() * c['L'] \
            + anyA.sum() * c['A'] + anyQ.sum() * c['Q'] + c['pro']
        trips += T
        lanes_useful += live.sum() * c['H'] + (isP | isL).sum() * c['V'] + isP.sum() * c['P'] + isL.sum() * c['L'] \
            + isA.sum() * c['A'] + isQ.sum() * c['Q'] + (n > 0).sum() * c['pro']
    return total, lanes_useful / 32.0, trips


def scheme_while_while(ops, counts, c):
    """Outer trip = descend phase (lanes whose next op is P/L loop together) then step phase (A/Q lanes loop
    until their next op is a P/L)."""
    total = 0.0
    for w in range(ops.shape[0]):
        n = counts[w].astype(np.int64)
        if n.max() == 0:
            continue
        pos = np.zeros(32, np.int64)
        o = ops[w]
        total += c['pro']
        while True:
            live = pos < n
            if not live.any():
                break
            # descend phase
            while True:
                cur = np.where(live, o[np.arange(32), np.minimum(pos, n - 1 + (n == 0))], 0)
                want = live & ((cur == P) | (cur == Lf))
                if not want.any():
                    break
                total += c['H'] + c['V'] + (c['P'] if (cur[want] == P).any() else 0) + (c['L'] if (cur[want] == Lf).any() else 0)
                pos = pos + want
                live = pos < n
            # step phase
            while True:
                cur = np.where(live, o[np.arange(32), np.minimum(pos, n - 1 + (n == 0))], 0)
                want = live & ((cur == A) | (cur == Q) | (cur == X))
                if not want.any():
                    break
                total += c['H'] + c['A'] + (c['Q'] if ((cur[want] == Q) | (cur[want] == X)).any() else 0)
                pos = pos + want
                live = pos < n
    return total


def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "sdf2048"
    W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3840, 2160)
    stride = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    ops, counts = trace(scene, W, H, stride)
    nw = ops.shape[0]
    rays = int((counts > 0).sum())
    print(f"{nw} warps, {rays} rays, trips/ray {counts[counts > 0].mean():.2f}, max {counts.max()}")
    flat = ops[np.arange(ops.shape[2])[None, None, :] < counts[:, :, None]]
    for ch in 'PAQLX':
        print(f"  {ch}: {(flat == ord(ch)).sum() / rays:.2f} per ray")
    for k in range(3):
        w, l = nw // 2 + k, 5
        print("  sample:", bytes(ops[w, l, :counts[w, l]]).decode())
    cur = dict(H=18, V=9, P=44, A=13, Q=39, L=110, pro=264)
    tight = dict(H=17, V=8, P=32, A=13, Q=26, L=105, pro=120)
    for name, c in (("current SASS", cur), ("tightened", tight)):
        t0, useful, trips = scheme_single_loop(ops, counts, c)
        t1 = scheme_while_while(ops, counts, c)
        print(f"[{name}] single loop: {t0 / rays:.1f} warp-instr/ray, simt eff {useful / t0:.3f}, warp trips/warp {trips / nw:.1f}"
              f" | while-while: {t1 / rays:.1f} warp-instr/ray ({t0 / t1:.2f}x)")


if __name__ == "__main__":
    main()


def scheme_vote(ops, counts, c, mode="majority"):
    """Each round the warp runs ONE body (descend or step), picked by vote; other lanes wait with their
    header result cached."""
    total = 0.0
    for w in range(ops.shape[0]):
        n = counts[w].astype(np.int64)
        if n.max() == 0:
            continue
        pos = np.zeros(32, np.int64)
        o = ops[w]
        total += c['pro']
        ar = np.arange(32)
        fresh = n > 0          # lanes that need a header evaluation
        while True:
            live = pos < n
            if not live.any():
                break
            cur = np.where(live, o[ar, np.minimum(pos, np.maximum(n - 1, 0))], 0)
            wantP = live & ((cur == P) | (cur == Lf))
            wantA = live & ((cur == A) | (cur == Q) | (cur == X))
            if fresh.any():
                total += c['H']
            nP, nA = wantP.sum(), wantA.sum()
            if mode == "majority":
                pickP = nP >= nA
            else:  # prefer stepping lanes so that they catch up into descents
                pickP = nA == 0
            if pickP:
                total += c['V'] + (c['P'] if (cur[wantP] == P).any() else 0) + (c['L'] if (cur[wantP] == Lf).any() else 0)
                pos = pos + wantP
                fresh = wantP.copy()
            else:
                total += c['A'] + (c['Q'] if ((cur[wantA] == Q) | (cur[wantA] == X)).any() else 0)
                pos = pos + wantA
                fresh = wantA.copy()
    return total


if __name__ == "__main__" and len(sys.argv) > 5 and sys.argv[5] == "vote":
    ops, counts = trace(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    rays = int((counts > 0).sum())
    now = dict(H=18, V=9, P=28, A=13, Q=33, L=105, pro=150)
    t0, useful, trips = scheme_single_loop(ops, counts, now)
    print(f"single loop {t0 / rays:.1f}  eff {useful / t0:.3f}")
    for mode in ("majority", "step-first"):
        print(mode, f"{scheme_vote(ops, counts, now, mode) / rays:.1f}")


def scheme_deferred_pop(ops, counts, c, threshold):
    """Single loop, but lanes whose advance needs a pop wait (idle) until at least `threshold` lanes of the
    warp are waiting for one -- or nothing else can run -- and then all pop together."""
    total = 0.0
    for w in range(ops.shape[0]):
        n = counts[w].astype(np.int64)
        if n.max() == 0:
            continue
        pos = np.zeros(32, np.int64)
        o = ops[w]
        ar = np.arange(32)
        pend = np.zeros(32, bool)       # advanced, pop still owed
        total += c['pro']
        while True:
            live = pos < n
            if not (live | pend).any():
                break
            cur = np.where(live & ~pend, o[ar, np.minimum(pos, np.maximum(n - 1, 0))], 0)
            run = live & ~pend
            isP, isL = run & (cur == P), run & (cur == Lf)
            isA = run & ((cur == A) | (cur == Q) | (cur == X))
            isQ = run & ((cur == Q) | (cur == X))
            if run.any():
                total += c['H'] + ((c['V']) if (isP | isL).any() else 0) + (c['P'] if isP.any() else 0) \
                    + (c['L'] if isL.any() else 0) + (c['A'] if isA.any() else 0)
                pos = pos + (isP | isL | (isA & ~isQ))
                pend = pend | isQ
            if pend.sum() >= threshold or (pend.any() and not ((pos < n) & ~pend).any()):
                total += c['Q']
                pos = pos + pend
                pend[:] = False
    return total


if __name__ == "__main__" and len(sys.argv) > 5 and sys.argv[5] == "defer":
    ops, counts = trace(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    rays = int((counts > 0).sum())
    now = dict(H=15, V=8, P=27, A=15, Q=37, L=105, pro=150)
    t0, useful, trips = scheme_single_loop(ops, counts, now)
    print(f"single loop {t0 / rays:.1f}  eff {useful / t0:.3f}")
    for th in (1, 4, 8, 12, 16, 24):
        print("defer pops until", th, f"lanes: {scheme_deferred_pop(ops, counts, now, th) / rays:.1f}")


if __name__ == "__main__" and len(sys.argv) > 5 and sys.argv[5] == "regroup":
    ops, counts = trace(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    rays = int((counts > 0).sum())
    now = dict(H=15, V=8, P=27, A=15, Q=37, L=105, pro=150)
    t0, useful, trips = scheme_single_loop(ops, counts, now)
    print(f"8x4 halves: {t0 / rays:.1f} warp-instr/ray, eff {useful / t0:.3f}")
    # oracle-knowledge regrouping of each tile's 64 rays into two warps (upper bounds on any predictor)
    nt = ops.shape[0] // 2
    o64 = ops[:2 * nt].reshape(nt, 64, -1)
    c64 = counts[:2 * nt].reshape(nt, 64)
    for name, key in (("by trip count", lambda o, c: c),
                      ("by first divergence position", None),
                      ("4x4 quads interleaved (checkerboard)", "quad")):
        if key == "quad":
            lane = np.arange(64)
            x, y = lane % 8, lane // 8
            order = np.argsort(((x // 4 + y // 4) % 2) * 64 + lane, kind="stable")
            oo = o64[:, order]
            cc = c64[:, order]
        elif key is None:
            # sort by the op string itself (lexicographic on the first 48 trips)
            keys = [np.lexsort(o64[t, :, :48].T[::-1]) for t in range(nt)]
            oo = np.stack([o64[t, keys[t]] for t in range(nt)])
            cc = np.stack([c64[t, keys[t]] for t in range(nt)])
        else:
            idx = np.argsort(c64, axis=1, kind="stable")
            oo = np.take_along_axis(o64, idx[:, :, None], axis=1)
            cc = np.take_along_axis(c64, idx, axis=1)
        t1, u1, _ = scheme_single_loop(oo.reshape(2 * nt, 32, -1), cc.reshape(2 * nt, 32), now)
        print(f"{name}: {t1 / rays:.1f} warp-instr/ray ({t0 / t1:.3f}x), eff {u1 / t1:.3f}")


def scheme_refill(ops, counts, c, batch):
    """One warp per tile (64 rays): lanes whose ray ended wait until `batch` lanes are free (or no ray is
    running), then set up the next rays of the tile together (cost pro) -- bounded lane re-fill."""
    total = 0.0
    nt = ops.shape[0] // 2
    for t in range(nt):
        o = ops[2 * t:2 * t + 2].reshape(64, -1)
        n = counts[2 * t:2 * t + 2].reshape(64).astype(np.int64)
        if n.max() == 0:
            continue
        queue = list(np.nonzero(n > 0)[0])
        lane_ray = -np.ones(32, np.int64)
        pos = np.zeros(32, np.int64)
        while True:
            free = np.nonzero(lane_ray < 0)[0]
            running = lane_ray >= 0
            if queue and (len(free) >= batch or not running.any()):
                total += c['pro']
                for ln in free:
                    if not queue:
                        break
                    lane_ray[ln] = queue.pop(0)
                    pos[ln] = 0
                running = lane_ray >= 0
            if not running.any():
                break
            lr = np.where(running, lane_ray, 0)
            cur = np.where(running, o[lr, np.minimum(pos, n[lr] - 1)], 0)
            isP, isL = running & (cur == P), running & (cur == Lf)
            isA = running & ((cur == A) | (cur == Q) | (cur == X))
            isQ = running & ((cur == Q) | (cur == X))
            total += c['H'] + (c['V'] if (isP | isL).any() else 0) + (c['P'] if isP.any() else 0) \
                + (c['L'] if isL.any() else 0) + (c['A'] if isA.any() else 0) + (c['Q'] if isQ.any() else 0)
            pos = pos + running
            done = running & (pos >= n[lr])
            lane_ray[done] = -1
    return total


if __name__ == "__main__" and len(sys.argv) > 5 and sys.argv[5] == "refill":
    ops, counts = trace(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    rays = int((counts > 0).sum())
    now = dict(H=15, V=8, P=27, A=15, Q=37, L=105, pro=150)
    t0, useful, trips = scheme_single_loop(ops, counts, now)
    print(f"two warps per tile, no refill: {t0 / rays:.1f} warp-instr/ray")
    for b in (1, 4, 8, 16, 32):
        print(f"one warp per tile, refill when {b} lanes free: {scheme_refill(ops, counts, now, b) / rays:.1f}")
