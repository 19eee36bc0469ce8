#!/usr/bin/env python
"""bench.py -- Mrays/s of the ESVO ray-casting hot path on N B200s, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl own|reference] [--workload NAME]

A "step" is one frame: the reference's renderBatch over all strips (beam pass + fine pass + shading,
reference src/Main.cpp:139-202) for one camera of an orbit. A "ray" is one raymarch invocation
(reference src/VoxelOctree.cpp:207), coarse and fine alike, as the reference would issue them for
that frame. `value` = rays of the K timed frames / device time, inputs (the octree) resident in HBM;
`e2e` = the same through the host-buffer C-ABI call, the frame copied back to pinned host memory
every step. With N > 1 the octree is replicated, the 8x8 tiles are interleaved over the ranks and
every rank's fine-pass kernel stores its finished tiles straight into rank 0's framebuffer over
NVLink (CUDA IPC peer mapping) -- strong scaling of one frame.

--impl reference times the reference's own CPU renderer (its unmodified object code in
oracle/_ref/libsvo_ref.so) on all host cores on the same workload; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "sparse-voxel-octrees_b200"))

# name -> (scene file stem, width, height, camera: pitch, yaw0, yaw step, radius), BASELINE.json configs
WORKLOADS = {
    "c3_ico8192_4k": dict(scene="ico8192", width=3840, height=2160, pitch=20.0, yaw0=40.0, yaw_step=3.6, radius=0.9,
                          text="configs[2]: ~10M-triangle displaced icosphere -> 8192^3 via reference PlyLoader, 3840x2160 primary rays, orbit"),
    "c2_sdf2048_1080p": dict(scene="sdf2048", width=1920, height=1080, pitch=20.0, yaw0=40.0, yaw_step=3.6, radius=0.9,
                             text="configs[1]: SDF sphere + fBm voxelised to 2048^3 by the reference VoxelData builder, 1920x1080 primary rays, orbit"),
    "c2_sdf2048_4k": dict(scene="sdf2048", width=3840, height=2160, pitch=20.0, yaw0=40.0, yaw_step=3.6, radius=0.9,
                          text="configs[1] scene at 3840x2160"),
    "c1_dragon_720p": dict(scene=None, width=1280, height=720, pitch=0.0, yaw0=0.0, yaw_step=3.6, radius=1.0,
                           text="configs[0]: models/XYZRGB-Dragon.oct (256^3), 1280x720 primary rays, orbit radius 1"),
    "c5_flythrough_ico8192": dict(scene="ico8192", width=3840, height=2160, camera_path="flythrough", pitch=20.0, yaw0=0.0,
                                  yaw_step=0.0, radius=2.0,
                                  text="configs[4]: camera fly-through, 100 frames, radius geometric 2.0 -> 0.05, yaw 0 -> 180 "
                                       "degrees, pitch 20, on the 8192^3 octree at 3840x2160"),
    "c5_flythrough_sdf2048": dict(scene="sdf2048", width=3840, height=2160, camera_path="flythrough", pitch=20.0, yaw0=0.0,
                                  yaw_step=0.0, radius=2.0,
                                  text="configs[4] camera path (100 frames, radius 2.0 -> 0.05, yaw 0 -> 180, pitch 20) on "
                                       "the 2048^3 SDF scene at 3840x2160"),
    "c4_ao_sdf2048": dict(scene="sdf2048", kind="ao", width=1920, height=1080, spp=16, pitch=20.0, yaw0=40.0, yaw_step=3.6,
                          radius=0.9,
                          text="configs[3]: incoherent ambient-occlusion rays, 16 random hemisphere directions per primary "
                               "hit of the 1920x1080 frame on the 2048^3 SDF scene, through the batch raymarch API"),
    "fallback_ico2048_4k": dict(scene="ico2048", width=3840, height=2160, pitch=20.0, yaw0=40.0, yaw_step=3.6, radius=0.9,
                                text="FALLBACK (8192^3 scene file absent): displaced icosphere -> 2048^3 via reference PlyLoader, 3840x2160"),
}
STRIPS = 16            # the reference's NumThreads (Main.cpp:57): part of the image definition
ORBIT = 100            # distinct cameras (3.6 degrees apart)
LANES = 4              # frames in flight per rank (SVO_BENCH_LANES overrides): hides the fine pass's long-ray tail and,
                       # with several GPUs, the frame barrier. Measured 2 / 4 / 6 lanes, 2048^3 @ 4K: N=8 48.6 / 61.6 / 63.0
                       # Grays/s, N=4 31.2 / 34.6 / -, N=1 9.30 / 9.37 / - (profiles/experiments/r01_lanes.md)
DRAGON = ROOT / "tests" / "golden" / "XYZRGB-Dragon.oct"


def pick_workload(name):
    from tools import make_scenes
    if name:
        w = dict(WORKLOADS[name], name=name)
    else:
        for cand in ("c3_ico8192_4k", "c2_sdf2048_1080p"):
            if make_scenes.scene_available(WORKLOADS[cand]["scene"]):
                w = dict(WORKLOADS[cand], name=cand)
                break
        else:
            w = dict(WORKLOADS["fallback_ico2048_4k"], name="fallback_ico2048_4k")
    w["path"] = DRAGON if w["scene"] is None else make_scenes.scene_path(w["scene"])
    return w


def ensure_scene(w, rank, barrier):
    """Rank 0 builds a missing scene with the reference builder (oracle/_ref); others wait."""
    if not Path(w["path"]).exists():
        if rank == 0:
            from tools import make_scenes
            if make_scenes.packed_path(w["scene"]).exists():
                # transport sidecar (snapshot size cap): same words, xz chunks -> literal-only .oct on local disk
                import pysvo
                words, center = make_scenes.unpack_scene(w["scene"])
                pysvo.oct_write(str(w["path"]) + ".tmp", words, center, compress=False)
                os.replace(str(w["path"]) + ".tmp", w["path"])
                del words
            else:
                make_scenes.make_scene(w["scene"], verbose=False)
    barrier()


def cameras(pysvo_or_none, w, count):
    """(pitch, yaw, radius) of step k; both paths repeat after ORBIT = 100 distinct cameras."""
    out = []
    for k in range(count):
        j = k % ORBIT
        if w.get("camera_path") == "flythrough":
            out.append((w["pitch"], 180.0 * j / (ORBIT - 1), 2.0 * (0.05 / 2.0) ** (j / (ORBIT - 1))))
        else:
            out.append((w["pitch"], w["yaw0"] + w["yaw_step"] * j, w["radius"]))
    return out


class ClockSampler:
    """Samples SM clock + throttle reasons every ~20 ms through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples = []
        self._stop = threading.Event()
        self._thread = None
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.ok:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable: " + getattr(self, "err", "")}
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonApplicationsClocksSetting", 0x2): "applications_clocks_setting",
        }
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        window = "timed region"
        if not inside:
            inside = self.samples
            window = "warm-up + timed + e2e (timed region shorter than the sampling period)"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no samples"}
        mhz = sorted(s[1] for s in inside)
        bits = 0
        for s in inside:
            bits |= s[2]
        reasons = sorted(n for b, n in names.items() if b and (bits & b))
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(inside), "window": window}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def profile_traffic(workload):
    """dram bytes per fine-pass launch of this workload from the committed ncu --set full capture, if any."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(workload)
        except Exception:  # noqa: BLE001
            return None
    return None


# ------------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------------

def reference_sample(w, steps, warmup, budget_s, quiet=False):
    """Times the reference's own renderer (oracle/_ref) on all host cores. Returns dict(value, ...)."""
    from oracle.pyoracle import Ref, strip_layout
    import pysvo  # host-only .oct reader (64-bit safe; the reference loader truncates >= 2 GiB, App. E.1)
    ref = Ref()
    cores = os.cpu_count() or 1
    H = w["height"]
    strips = max(1, min(cores, H // 8))       # strips = OS threads = nproc
    words, center = pysvo.oct_read(w["path"])
    h = ref.tree_from_words(words, center)
    del words
    cams = cameras(None, w, warmup + steps)
    mv = [ref.orbit_camera(*c) for c in cams]
    models = np.stack([m for m, _ in mv])
    views = np.stack([v for _, v in mv])
    # size the per-step sample: probe one frame on every 8th strip, then pick the modulo
    modulo = 1
    t0 = time.perf_counter()
    ref.render_frames(h, w["width"], H, strips, models[:1], views[:1], threads=cores, strip_modulo=8)
    probe = (time.perf_counter() - t0) * 8.0
    while modulo < 64 and probe * (steps + warmup) / modulo > budget_s and strips // (modulo * 2) >= 1:
        modulo *= 2
    rgba, _, secs = ref.render_frames(h, w["width"], H, strips, models[:warmup] if warmup else models[:1],
                                      views[:warmup] if warmup else views[:1], threads=cores, strip_modulo=modulo)
    total_rays = 0
    total_s = 0.0
    lay = strip_layout(w["width"], H, strips)
    for k in range(steps):
        rgba, _, secs = ref.render_frames(h, w["width"], H, strips, models[warmup + k:warmup + k + 1],
                                          views[warmup + k:warmup + k + 1], threads=cores, strip_modulo=modulo)
        rays = 0
        for s, (y0, y1, tx, ty) in enumerate(lay):
            if s % modulo == 0:
                rays += tx * ty + int((rgba[y0:y1] != 0).sum())
        total_rays += rays
        total_s += float(secs[0])
    ref.tree_destroy(h)
    sample = (f"{steps} frame(s) after {warmup} warm-up, every strip" if modulo == 1 else
              f"{steps} frame(s) after {warmup} warm-up, every {modulo}th of {strips} strips per frame")
    return dict(value=total_rays / total_s / 1e6, seconds=total_s, rays=total_rays, cores=cores, strips=strips,
                sample=sample, ms_per_step=total_s / steps * 1e3, modulo=modulo)


def ao_workload_rays(w, tree_or_none, words, center):
    """Primary hits of the frame (oracle-independent: traced by whoever asks) -> AO ray set."""
    from oracle.pyoracle import Port, pixel_rays
    from tools.ao_rays import ao_rays
    port = Port()
    m, v = port.orbit_camera(w["pitch"], w["yaw0"], w["radius"])
    f = port.frame_constants(m, v, center, w["width"], w["height"], STRIPS)
    o, d = pixel_rays(f)
    if tree_or_none is not None:
        prim = tree_or_none.raymarch_batch(o, d, 0.0, 0)
    else:
        prim = port.raymarch_batch(words, o, d, 0.0, t_sentinel=1e10)
    return ao_rays(o, d, prim["t"], prim["normal"], prim["hit"] == 1, spp=w["spp"])[:2]


def run_reference_ao(args, w):
    from oracle.pyoracle import Ref
    import pysvo
    ref = Ref()
    cores = os.cpu_count() or 1
    words, center = pysvo.oct_read(w["path"])
    ao_o, ao_d = ao_workload_rays(w, None, words, center)
    h = ref.tree_from_words(words, center)
    steps = args.steps if args.steps is not None else 2
    warmup = args.warmup if args.warmup is not None else 1
    n = ao_o.shape[0]
    sample = min(n, 4_000_000)                  # bounded sample: the first `sample` rays of the set
    for _ in range(warmup):
        ref.raymarch_batch(h, ao_o[:sample], ao_d[:sample], 0.0, threads=cores)
    secs = 0.0
    for _ in range(steps):
        secs += ref.raymarch_batch(h, ao_o[:sample], ao_d[:sample], 0.0, threads=cores)[3]
    value = sample * steps / secs / 1e6
    line = {"impl": "reference", "metric": "Mrays/s ESVO traversal (batch raymarch calls / time)", "value": value,
            "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": secs / steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "description": w["text"], "rays": n, "host_threads": cores},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "reference",
                             "sample": f"first {sample} of {n} AO rays per step, {steps} step(s) after {warmup} warm-up"},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def run_own_ao(args, w):
    """configs[3]: the batch API on incoherent rays. Rays sharded contiguously over the ranks, no exchange."""
    import torch
    import pysvo
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    steps = args.steps if args.steps is not None else 20
    warmup = max(args.warmup if args.warmup is not None else 3, 3)
    words, center = pysvo.oct_read(w["path"])
    tree = pysvo.VoxelOctree(words=words, center=center, device=local_rank)
    flavour = pysvo.FLAVOUR_VALIDATION if args.validation else pysvo.FLAVOUR_FAST
    if not os.environ.get("SVO_BENCH_NO_COHERENCE_ORDER"):     # (experiment switch, never set by default)
        flavour |= pysvo.BATCH_COHERENCE_ORDER               # direction-binned thread order, inside the timed call
    ao_o, ao_d = ao_workload_rays(w, tree, words, center)
    n_total = ao_o.shape[0]
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    ao_o, ao_d = ao_o[lo:hi], ao_d[lo:hi]
    n = ao_o.shape[0]
    d_o, d_d = torch.from_numpy(ao_o).cuda(), torch.from_numpy(ao_d).cuda()
    d_hit = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_t = torch.empty(n, dtype=torch.float32, device="cuda")
    d_n = torch.empty(n, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        tree.raymarch_batch_device(n, d_o.data_ptr(), d_d.data_ptr(), 0.0, flavour, d_hit.data_ptr(), d_t.data_ptr(),
                                   d_n.data_ptr(), 0, stream)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t_end = time.perf_counter()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if dist:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    value = n_total * steps / (total_ms * 1e-3) / 1e6

    # end to end: host rays in, host results out, every step (pinned buffers, chunked double-buffered copies)
    ho, hd = pysvo.PinnedArray((n, 3), np.float32), pysvo.PinnedArray((n, 3), np.float32)
    ho.array[:], hd.array[:] = ao_o, ao_d
    pinned_out = [pysvo.PinnedArray(n, np.uint8), pysvo.PinnedArray(n, np.float32), pysvo.PinnedArray(n, np.uint32)]
    out = dict(hit=pinned_out[0].array, t=pinned_out[1].array, normal=pinned_out[2].array, voxel=None)
    e2e_steps = min(steps, 5)
    tree.raymarch_batch(ho.array, hd.array, 0.0, flavour, out=out)
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        tree.raymarch_batch(ho.array, hd.array, 0.0, flavour, out=out)
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if dist:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n_total * e2e_steps / float(e2e_s.item()) / 1e6
    if sampler:
        sampler.stop()
    if rank != 0:
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return

    # parity + algorithmic bytes on a bounded sample, roofline of the batch kernel
    from oracle.pyoracle import Port
    port = Port()
    sample = min(n, 2_000_000)
    want = port.raymarch_batch(words, ao_o[:sample], ao_d[:sample], 0.0, t_sentinel=1e10)
    got_hit = d_hit[:sample].cpu().numpy()
    got_n = d_n[:sample].cpu().numpy().view(np.uint32)
    leaf = want["hit"] == 1
    identical = float(((got_hit == want["hit"]) & np.where(leaf, got_n == want["normal"], True)).mean())
    c = want["counters"]
    b_ray = 4.0 * c.words / c.rays + 24.0 + 9.0      # node words + 24 B ray read + hit/t/normal written
    peak, peak_src = measured_peak()
    achieved = b_ray * n_total * steps / (total_ms * 1e-3) / 1e9
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle.pyoracle import Ref
        ref = Ref()
        h = ref.tree_from_words(words, center)
        cores = os.cpu_count() or 1
        ref.raymarch_batch(h, ao_o[:sample], ao_d[:sample], 0.0, threads=cores)
        secs = ref.raymarch_batch(h, ao_o[:sample], ao_d[:sample], 0.0, threads=cores)[3]
        cpu = {"value": sample / secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference",
               "sample": f"first {sample} of {n} AO rays, one pass after one warm-up pass"}
    line = {"metric": "Mrays/s ESVO traversal (batch raymarch calls / time)", "value": value, "unit": "Mrays/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "description": w["text"], "rays": n_total, "spp": w["spp"],
                       "flavour": "validation" if args.validation else "fast",
                       "thread_order": "direction-binned (SVO_BATCH_COHERENCE_ORDER: key pass + one 6-bit radix pass "
                                       "inside every timed call)" if flavour & pysvo.BATCH_COHERENCE_ORDER else "submission order",
                       "l2": f"no flush: ray + result arrays {n_total * 41 / 1e6:.0f} MB per step exceed the 126 MB L2",
                       "parallelism": "rays sharded contiguously over ranks, octree replicated" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": n_total * 24, "d2h_bytes_per_step": n_total * 9,
                    "steps": e2e_steps, "note": "svo_raymarch_batch with pinned host arrays: 2 Mi-ray chunks alternate between two streams"},
            "gpu_launches": steps * world * (1 + (4 if flavour & pysvo.BATCH_COHERENCE_ORDER else 0)),
            "clocks": sampler.summary(t_begin, t_end),
            "parity": {"identical_rays_fast_vs_oracle": identical, "sample": sample},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "raymarchBatchKernel<FAST>", "peak_source": peak_src,
                         "bytes_per_ray": b_ray}}
    if cpu:
        line["cpu_baseline"] = cpu
    emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = pick_workload(args.workload)
    ensure_scene(w, 0, lambda: None)
    if w.get("kind") == "ao":
        return run_reference_ao(args, w)
    steps = args.steps if args.steps is not None else 3
    warmup = args.warmup if args.warmup is not None else 1
    r = reference_sample(w, steps, warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": "Mrays/s ESVO traversal (coarse + fine raymarch calls per frame / time)",
        "value": r["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "description": w["text"], "width": w["width"], "height": w["height"],
                   "strips": r["strips"], "host_threads": r["cores"]},
        "cpu_baseline": {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": "reference",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------

def run_own(args):
    import torch
    import pysvo

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not pysvo.LIB_PATH.exists():
        raise SystemExit(f"{pysvo.LIB_PATH} missing: run __graft_entry__.build() first (no fallback exists)")
    if pysvo.device_count() < 1:
        raise SystemExit("no CUDA device: this benchmark has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    steps = args.steps if args.steps is not None else 300
    warmup = args.warmup if args.warmup is not None else 10
    warmup = max(warmup, 3)

    w = pick_workload(args.workload)
    ensure_scene(w, rank, barrier)
    if w.get("kind") == "ao":
        if dist is not None:
            dist.destroy_process_group()
        return run_own_ao(args, w)
    W, H = w["width"], w["height"]
    tree = pysvo.VoxelOctree(w["path"], device=local_rank)
    flavour = pysvo.FLAVOUR_VALIDATION if args.validation else pysvo.FLAVOUR_FAST
    stream = torch.cuda.current_stream().cuda_stream
    cams = [pysvo.orbit_camera(*c) for c in cameras(pysvo, w, ORBIT)]

    # ---- framebuffer: rank 0 owns it; other ranks map it and store their tiles into it over NVLink
    nbytes = W * H * 4
    n_lanes = min(int(os.environ.get("SVO_BENCH_LANES", "0")) or LANES, 8)    # the library's frame ring is 8 deep
    fbs = []
    fb_ptrs = [0] * n_lanes
    if rank == 0:
        for i in range(n_lanes):    # one framebuffer per frame in flight: frame k is consumed while k+1.. are rendered
            fbs.append(pysvo.DeviceBuffer(local_rank, nbytes))
            fbs[i].zero()
            fb_ptrs[i] = fbs[i].ptr
    if world > 1:
        handle = [[f.ipc_export() for f in fbs] if rank == 0 else None]
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            fb_ptrs = [pysvo.ipc_open(local_rank, hdl) for hdl in handle[0]]
    fb_ptr = fb_ptrs[0]
    flag = torch.zeros(1, device="cuda", dtype=torch.int32) if world > 1 else None

    def frame(k, want_stats=False, fb=None, sync=True):
        st = tree.render_frame_device(cams[k % ORBIT], W, H, fb or fb_ptr, strips=STRIPS, flavour=flavour,
                                      tile_rank=rank, tile_world=world, stream=stream, want_stats=want_stats)
        if world > 1 and sync:
            dist.all_reduce(flag)      # frame complete on every rank before rank 0 may use it
        return st

    # ---- rays per camera (untimed): coarse once per frame, fine summed over ranks
    n_distinct = min(ORBIT, steps + warmup)
    fine = torch.zeros(ORBIT, dtype=torch.int64, device="cuda")
    coarse = int(pysvo.frame_layout(W, H, STRIPS).corners)   # beam rays of the frame as the reference issues them
    kernel_ms = []
    for k in range(n_distinct):
        st = frame(k, want_stats=True)
        fine[k] = int(st.fine_rays)
        kernel_ms.append((st.coarse_ms, st.fine_ms))
    if os.environ.get("SVO_BENCH_DEBUG"):
        km = np.array(kernel_ms)
        print(f"[rank {rank}] beam pass {km[:, 0].mean():.4f} ms, classify+fine {km[:, 1].mean():.4f} ms "
              f"(min {km[:, 1].min():.4f}, max {km[:, 1].max():.4f}) over {len(km)} cameras", file=sys.stderr, flush=True)
    if world > 1:
        dist.all_reduce(fine)
    fine = fine.cpu().numpy()
    rays_of = lambda k: coarse + int(fine[k % ORBIT])  # noqa: E731

    # ---- timed region: W warm-up frames, then exactly K frames between device events.
    # Double-buffered like a swap chain: frame k renders into framebuffer k & 1 on stream k & 1, so the
    # long-ray tail of one fine pass overlaps the start of the next frame (each frame is still complete,
    # in rank 0's memory, when its stream reaches the frame barrier).
    main = torch.cuda.current_stream()
    lanes = [torch.cuda.Stream() for _ in range(n_lanes)]
    comm = torch.cuda.Stream() if world > 1 else None
    gate = [None] * n_lanes

    def pipelined_frame(k):
        slot = k % n_lanes
        with torch.cuda.stream(lanes[slot]):
            tree.render_frame_device(cams[k % ORBIT], W, H, fb_ptrs[slot], strips=STRIPS, flavour=flavour,
                                     tile_rank=rank, tile_world=world, stream=lanes[slot].cuda_stream)
            if world > 1 and not os.environ.get("SVO_BENCH_NO_BARRIER"):   # (experiment switch, never set by default)
                done = torch.cuda.Event()
                done.record(lanes[slot])
                comm.wait_event(done)
                with torch.cuda.stream(comm):
                    dist.all_reduce(flag)          # frame barrier: every rank's tiles are in rank 0's framebuffer
                    gate[slot] = torch.cuda.Event()
                    gate[slot].record(comm)
                lanes[slot].wait_event(gate[slot])  # the slot is reused (frame k+2) only after the barrier

    def run_frames(first, count):
        start = torch.cuda.Event(enable_timing=True)
        stop = torch.cuda.Event(enable_timing=True)
        start.record(main)
        for ln in lanes:
            ln.wait_event(start)
        for k in range(first, first + count):
            pipelined_frame(k)
        for ln in lanes:
            main.wait_stream(ln)
        if comm is not None:
            main.wait_stream(comm)
        stop.record(main)
        return start, stop

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    run_frames(0, warmup)
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    t_begin = time.perf_counter()
    e0, e1 = run_frames(warmup, steps)
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    total_rays = sum(rays_of(k) for k in range(warmup, warmup + steps))
    value = total_rays / (total_ms * 1e-3) / 1e6

    # ---- end to end: host-visible frame every step (pinned host memory), same cameras
    e2e_steps = min(steps, 300)
    host_bufs = [pysvo.PinnedArray((H, W), np.uint32) for _ in range(4)] if rank == 0 else None
    host = host_bufs[0] if rank == 0 else None
    hosts = [h.array for h in host_bufs] if rank == 0 else None
    copy_stream = torch.cuda.Stream()
    if world == 1:
        # untimed: the library creates its staging framebuffers on first use
        for k in range(len(hosts)):
            tree.frame_wait(tree.render_frame_async(cams[k % ORBIT], W, H, hosts[k], strips=STRIPS, flavour=flavour))
    e2e_mode = "single"
    e2e_lanes = min(n_lanes, 4)
    frame_done = [None] * e2e_lanes
    shared_host = None
    per_rank = [world > 1 and not os.environ.get("SVO_BENCH_E2E_VIA_RANK0")]
    if world > 1:
        if rank == 0 and per_rank[0]:
            try:        # the shared host frames live in /dev/shm (like NCCL's own shared-memory segments)
                vfs = os.statvfs("/dev/shm")
                per_rank[0] = vfs.f_bavail * vfs.f_frsize > 2 * e2e_lanes * nbytes
            except OSError:
                per_rank[0] = False
        dist.broadcast_object_list(per_rank, src=0)
    per_rank = bool(per_rank[0])
    if per_rank:
        # untimed set-up of the per-rank path: the shared host frames (one per frame in flight) and local framebuffers
        name = [f"/dev/shm/svo_bench_{os.getpid()}.frames" if rank == 0 else None]
        dist.broadcast_object_list(name, src=0)
        ok, registered = True, False
        try:
            if rank == 0:
                shared_host = np.memmap(name[0], dtype=np.uint32, mode="w+", shape=(e2e_lanes, H, W))
                shared_host[:] = 0
        except Exception as e:  # noqa: BLE001
            ok = False
            print(f"[rank {rank}] shared host frame: {e!r}", file=sys.stderr, flush=True)
        barrier()
        try:
            if rank != 0:
                shared_host = np.memmap(name[0], dtype=np.uint32, mode="r+", shape=(e2e_lanes, H, W))
            shared_dev = pysvo.host_register(local_rank, shared_host)
            registered = True
        except Exception as e:  # noqa: BLE001
            ok = False
            print(f"[rank {rank}] mapping the shared host frame: {e!r}", file=sys.stderr, flush=True)
        all_ok = torch.tensor([1 if ok else 0], device="cuda", dtype=torch.int32)
        dist.all_reduce(all_ok, op=dist.ReduceOp.MIN)
        if not int(all_ok.item()):       # some rank could not map it: every rank takes the rank-0 path instead
            if registered:
                pysvo.host_unregister(shared_host)
            shared_host = None
            per_rank = False
            barrier()
            if rank == 0 and os.path.exists(name[0]):
                os.unlink(name[0])
    if per_rank:
        # wider stripes for this leg: a rank's rows of pixels are what one PCIe write burst carries (measured: 128-byte
        # runs 28 GB/s, whole rows 50 GB/s); the widest run <= 16 tile columns that still deals every rank the same
        # number of columns, else the default of 4. The image does not depend on it.
        tile_cols = (W - 1) // 8 + 1
        e2e_run = int(os.environ.get("SVO_BENCH_E2E_RUN", "0")) or next(
            (r for r in range(16, 4, -1) if tile_cols % (world * r) == 0), 4)
        pysvo.frame_set_tile_run(e2e_run)
        if rank == 0:
            local_fbs = fb_ptrs[:e2e_lanes]
        else:
            own = [pysvo.DeviceBuffer(local_rank, nbytes) for _ in range(e2e_lanes)]
            for b in own:
                b.zero()
            local_fbs = [b.ptr for b in own]
        for k in range(e2e_lanes):      # first use of the copy kernel / the mapping, untimed
            pysvo.frame_copy_owned_tiles(local_rank, W, H, STRIPS, rank, world, local_fbs[k], shared_dev + k * nbytes, stream)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    if world == 1:
        # pipelined host-buffer API: four frames in flight, each lands in its own pinned host buffer
        pending = [None] * len(hosts)
        dbg = [0.0, 0.0]
        for k in range(warmup, warmup + e2e_steps):
            slot = k % len(hosts)
            ta = time.perf_counter()
            if pending[slot] is not None:
                tree.frame_wait(pending[slot])          # frame k-4 is in host memory; its buffer is free again
            tb = time.perf_counter()
            pending[slot] = tree.render_frame_async(cams[k % ORBIT], W, H, hosts[slot], strips=STRIPS, flavour=flavour)
            dbg[0] += tb - ta
            dbg[1] += time.perf_counter() - tb
        for pnd in pending:
            if pnd is not None:
                tree.frame_wait(pnd)
        if os.environ.get("SVO_BENCH_DEBUG"):
            print(f"[e2e] wait {dbg[0] / e2e_steps * 1e3:.3f} ms/frame, issue {dbg[1] / e2e_steps * 1e3:.3f} ms/frame",
                  file=sys.stderr, flush=True)
    elif not per_rank:
        e2e_mode = "rank0"
        # every rank stores its tiles into rank 0's framebuffer k & 1 over NVLink; after the frame barrier
        # rank 0 copies it to pinned host memory on a side stream while frame k+1 is rendered into the other one
        copied = [None, None]
        for k in range(warmup, warmup + e2e_steps):
            slot = k & 1
            frame(k, fb=fb_ptrs[slot], sync=False)
            if rank == 0 and copied[slot ^ 1] is not None:
                main.wait_event(copied[slot ^ 1])   # framebuffer slot^1 is rewritten by frame k+1 after this barrier
            dist.all_reduce(flag)
            if rank == 0:
                if copied[slot] is not None:
                    copied[slot].synchronize()      # frame k-2 is in host memory: its pinned buffer is free again
                copy_stream.wait_stream(main)
                pysvo.device_to_host_async(local_rank, hosts[slot], fb_ptrs[slot], nbytes, copy_stream.cuda_stream)
                copied[slot] = torch.cuda.Event()
                copied[slot].record(copy_stream)
        if rank == 0:
            for ev in copied:
                if ev is not None:
                    ev.synchronize()
    else:
        # every rank renders its tiles into its OWN framebuffer and ships them itself (svo_frame_copy_owned_tiles)
        # into one page-locked host frame that all ranks have mapped (a shared-memory segment): 1 / world of the
        # frame per PCIe link instead of all of it over rank 0's. Four frames in flight, like the N = 1 path; the
        # frame barrier (all-reduce enqueued behind each rank's copy) tells rank 0 that a host frame is complete.
        e2e_mode = "per-rank"
        for k in range(warmup, warmup + e2e_steps):
            slot = k % e2e_lanes
            if frame_done[slot] is not None:
                frame_done[slot].synchronize()      # frame k-4 is complete in host memory: its slot is free again
            with torch.cuda.stream(lanes[slot]):
                tree.render_frame_device(cams[k % ORBIT], W, H, local_fbs[slot], strips=STRIPS, flavour=flavour,
                                         tile_rank=rank, tile_world=world, stream=lanes[slot].cuda_stream)
                pysvo.frame_copy_owned_tiles(local_rank, W, H, STRIPS, rank, world, local_fbs[slot],
                                             shared_dev + slot * nbytes, lanes[slot].cuda_stream)
                shipped = torch.cuda.Event()
                shipped.record(lanes[slot])
            comm.wait_event(shipped)
            with torch.cuda.stream(comm):
                dist.all_reduce(flag)
                frame_done[slot] = torch.cuda.Event()
                frame_done[slot].record(comm)
        for ev in frame_done:
            if ev is not None:
                ev.synchronize()
    torch.cuda.synchronize()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_rays = sum(rays_of(k) for k in range(warmup, warmup + e2e_steps))
    e2e_value = e2e_rays / float(e2e_s.item()) / 1e6
    e2e_frame_identical = None
    if shared_host is not None:
        if rank == 0:
            # the last host frame, assembled by all ranks, against the same camera rendered by rank 0 alone
            k_last = warmup + e2e_steps - 1
            tree.render_frame_device(cams[k_last % ORBIT], W, H, fb_ptrs[0], strips=STRIPS, flavour=flavour, stream=stream)
            torch.cuda.synchronize()
            alone = fbs[0].to_host(np.uint32).reshape(H, W)
            e2e_frame_identical = bool(np.array_equal(alone, shared_host[k_last % e2e_lanes]))
        barrier()
        pysvo.frame_set_tile_run(0)
        pysvo.host_unregister(shared_host)
        del shared_host
        if rank == 0:
            os.unlink(name[0])
    if sampler:
        sampler.stop()

    if rank != 0:
        if world > 1:
            barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fine pass): algorithmic bytes from the instrumented oracle
    from oracle.pyoracle import Port
    port = Port()
    words, center = pysvo.oct_read(w["path"])
    roof_cams = [0, min(25, n_distinct - 1)] if world == 1 else [0]
    alg_bytes, fine_ms_sum, px_sum, fine_rays_sum, node_bytes_sum = 0.0, 0.0, 0, 0, 0
    parity = {}
    for k in roof_cams:
        cam = cams[k]
        f = port.frame_constants(np.array(cam.model[:], np.float32), np.array(cam.view[:], np.float32), center, W, H, STRIPS)
        want, _, cc, cf = port.render_frame(words, f)
        node_bytes = 4 * cf.words
        if world == 1:
            alg_bytes += node_bytes + 4 * cf.rays
            fine_ms_sum += kernel_ms[k][1]
            px_sum += cf.rays
            fine_rays_sum += cf.rays
            node_bytes_sum += node_bytes
        if k == roof_cams[0]:
            got, _, _ = tree.render_frame(cam, W, H, strips=STRIPS, flavour=pysvo.FLAVOUR_FAST, rgba=host.array,
                                          want_stats=False) if world == 1 else (None, None, None)
            if got is not None:
                parity["fast_identical_pixels"] = float((got == want).mean())
                val, _, _ = tree.render_frame(cam, W, H, strips=STRIPS, flavour=pysvo.FLAVOUR_VALIDATION,
                                              rgba=host.array, want_stats=False)
                parity["validation_identical_pixels"] = float((val == want).mean())
            parity["oracle_rays"] = int(cc.rays + cf.rays)
            parity["gpu_rays"] = rays_of(k)
            coarse_b_ray = 4.0 * cc.words / max(cc.rays, 1)
            fine_b_ray = 4.0 * cf.words / max(cf.rays, 1)
    peak, peak_src = measured_peak()
    roofline = None
    if world == 1 and fine_ms_sum > 0:
        achieved = alg_bytes / (fine_ms_sum * 1e-3) / 1e9
        traffic = profile_traffic(w["name"])
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (traffic or {}).get("dram_bytes_per_launch"),
                    "kernel": "finePassKernel<FAST>", "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes / len(roof_cams),
                    "bytes_per_fine_ray": {"node_words": node_bytes_sum / max(fine_rays_sum, 1), "pixel_store": 4.0},
                    "launch_ms": fine_ms_sum / len(roof_cams),
                    "kernel_share_of_step": float(np.mean([f_ / (c_ + f_) for c_, f_ in kernel_ms])),
                    "ncu": {k: (traffic or {}).get(k) for k in ("l1_hit_pct", "l2_hit_pct", "issue_active_pct",
                                                               "threads_per_instruction", "source")},
                    "note": "instruction-issue bound pointer chasing with SIMT divergence, not HBM-bound: node "
                            "fetches hit L1/L2 (see profiles/ and DESIGN.md section 4)"}
        # the bound that does bind: one warp instruction per SM sub-partition per cycle (148 SMs x 4)
        if traffic and traffic.get("warp_instructions") and traffic.get("sm_cycles_elapsed"):
            slots = 148 * 4 * float(traffic["sm_cycles_elapsed"])
            roofline["issue_roofline"] = {
                "warp_instructions_per_launch": traffic["warp_instructions"],
                "issue_slots_per_launch": slots, "frac": traffic["warp_instructions"] / slots,
                "threads_per_instruction": traffic.get("threads_per_instruction"),
                "source": traffic.get("source"),
                "note": "fraction of the SMs' issue slots (148 SMs x 4 schedulers x elapsed cycles of the ncu capture) "
                        "that issued an instruction of this kernel: the limit this pass actually runs against"}

    # ---- CPU baseline: the reference's own renderer on this box's host cores (bounded sample)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = reference_sample(w, steps=2, warmup=1, budget_s=25.0)
            cpu = {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": "reference", "sample": r["sample"]}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": f"unavailable: {e!r}"}

    clocks = sampler.summary(t_begin, t_end)
    line = {
        "metric": "Mrays/s ESVO traversal (coarse + fine raymarch calls per frame / time)",
        "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "description": w["text"], "width": W, "height": H, "strips": STRIPS,
                   "flavour": "validation" if args.validation else "fast",
                   "octree_words": tree.n_words, "octree_depth": tree.depth,
                   "parallelism": "replicated octree, interleaved 8x8 tiles, fine-pass stores into rank 0's framebuffer over NVLink" if world > 1 else "single GPU",
                   "l2": f"no flush: octree {tree.n_words * 4 / 1e6:.0f} MB vs 126 MB L2, camera moves every step",
                   "pipelining": f"{n_lanes} frames in flight (framebuffer k % {n_lanes} on stream k % {n_lanes}); beam passes run ahead on internal streams",
                   "rays_per_frame_mean": total_rays / steps, "ms_per_frame": total_ms / steps},
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 128, "d2h_bytes_per_step": nbytes,
                "steps": e2e_steps, "ms_per_step": float(e2e_s.item()) / e2e_steps * 1e3,
                "note": "per-step input is the 128 B camera (kernel parameters); the octree stays resident; " + {
                    "single": "svo_render_frame_async, four frames in flight, every frame copied to pinned host memory",
                    "per-rank": "every rank ships the tiles it rendered into ONE page-locked host frame shared by all ranks "
                                "(svo_frame_copy_owned_tiles: 1 / world of the frame per PCIe link, stripes "
                                f"{e2e_run if world > 1 and per_rank else 4} tile columns wide), four frames in flight, "
                                "frame barrier behind the copies",
                    "rank0": "tiles gathered in rank 0's HBM over NVLink, rank 0 copies every frame to pinned host memory",
                }[e2e_mode]},
        "gpu_launches": 3 * steps * world,       # beam pass + tile classifier + fine pass per frame and rank
        "clocks": clocks,
        "parity": dict(parity, **({"e2e_host_frame_identical_to_single_rank": e2e_frame_identical}
                                  if e2e_frame_identical is not None else {})),
        "bytes_per_ray": {"coarse_node": coarse_b_ray, "fine_node": fine_b_ray},
    }
    if roofline is not None:
        line["roofline"] = roofline
    if cpu is not None:
        line["cpu_baseline"] = cpu
    emit(line)
    if world > 1:
        barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The one JSON line, on the process's real stdout."""
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, text.encode())
    else:
        sys.stdout.write(text)
        sys.stdout.flush()


def quarantine_stdout():
    """The reference's C++ code prints progress on std::cout (VoxelOctree.cpp:87, PlyLoader.cpp:465, ...), and
    its buffer may flush at exit, after our line. Point fd 1 at stderr for the whole run and keep the real
    stdout for the single JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    quarantine_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--validation", action="store_true", help="time the bit-exact flavour instead of FAST")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
