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

def reference_tree(ref, w):
    """The workload's node array as a tree handle of the reference's own code (oracle/_ref) -- WITHOUT loading
    libsvo_b200.so: the reference arm must not touch the product. The xz transport sidecar is read with
    lzma + numpy (tools/make_scenes.unpack_scene) and adopted through the reference's constructor; a plain
    .oct goes through the reference's own loader (VoxelOctree.cpp:57-90; every scene here is below the 2 GiB
    where its int truncation starts, App. E.1). Returns (handle, words or None, center)."""
    from tools import make_scenes
    path = Path(w["path"])
    if w["scene"] is not None and not path.exists() and make_scenes.packed_path(w["scene"]).exists():
        words, center = make_scenes.unpack_scene(w["scene"])
        return ref.tree_from_words(words, center), words, np.asarray(center, np.float32)
    if not path.exists():
        make_scenes.make_scene(w["scene"], verbose=False)        # the reference builder, also without the product
    h = ref.tree_load(path)
    return h, None, ref.tree_center(h)


def reference_sample(w, steps, warmup, budget_s, strips=None, ref=None, tree=None):
    """Times the reference's own renderer (oracle/_ref) on the host cores: `strips` strips = OS threads (the
    reference's NumThreads, Main.cpp:57,351-367); default STRIPS, the GPU arm's configuration, so that both arms
    render the same image. Returns dict(value, ...)."""
    from oracle.pyoracle import Ref, strip_layout
    own_tree = tree is None
    if ref is None:
        ref = Ref()
    cores = os.cpu_count() or 1
    H = w["height"]
    strips = STRIPS if strips is None else max(1, min(strips, H // 8))
    threads = min(cores, strips)
    h = reference_tree(ref, w)[0] if own_tree else tree
    cams = cameras(None, w, warmup + steps)
    mv = [ref.orbit_camera(*c) for c in cams]
    models = np.stack([m for m, _ in mv])
    views = np.stack([v for _, v in mv])
    # size the sample: one whole frame as a probe (also the first warm-up), then as many of the requested steps as fit
    # the budget; strips are only sub-sampled (every modulo-th strip, which also idles threads) when a single frame
    # does not fit
    modulo = 1
    t0 = time.perf_counter()
    ref.render_frames(h, w["width"], H, strips, models[:1], views[:1], threads=threads)
    probe = time.perf_counter() - t0
    while modulo < 64 and probe * (2 + warmup) / modulo > budget_s and strips // (modulo * 2) >= 1:
        modulo *= 2
    steps = max(2, min(steps, int(budget_s * modulo / max(probe, 1e-6)) - warmup))
    rgba, _, secs = ref.render_frames(h, w["width"], H, strips, models[:warmup] if warmup else models[:1],
                                      views[:warmup] if warmup else views[:1], threads=threads, strip_modulo=modulo)
    total_rays = 0
    total_s = 0.0
    lay = strip_layout(w["width"], H, strips)
    for k in range(steps):
        rgba, _, secs = ref.render_frames(h, w["width"], H, strips, models[warmup + k:warmup + k + 1],
                                          views[warmup + k:warmup + k + 1], threads=threads, strip_modulo=modulo)
        rays = 0
        for s, (y0, y1, tx, ty) in enumerate(lay):
            if s % modulo == 0:
                rays += tx * ty + int((rgba[y0:y1] != 0).sum())
        total_rays += rays
        total_s += float(secs[0])
    if own_tree:
        ref.tree_destroy(h)
    sample = (f"{steps} frame(s) after {warmup} warm-up, every strip" if modulo == 1 else
              f"{steps} frame(s) after {warmup} warm-up, every {modulo}th of {strips} strips per frame")
    return dict(value=total_rays / total_s / 1e6, seconds=total_s, rays=total_rays, cores=threads, host_cores=cores,
                strips=strips, sample=sample, ms_per_step=total_s / steps * 1e3, modulo=modulo)


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
    ref = Ref()
    cores = os.cpu_count() or 1
    h, words, center = reference_tree(ref, w)
    if words is None:
        words = ref.tree_words(h)
    ao_o, ao_d = ao_workload_rays(w, None, words, center)
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
    ensure_scene(w, rank, dist.barrier if dist else (lambda: None))
    steps = args.steps if args.steps is not None else 20
    warmup = max(args.warmup if args.warmup is not None else 3, 3)
    words, center = pysvo.oct_read(w["path"])
    tree = pysvo.VoxelOctree(words=words, center=center, device=local_rank)
    flavour = pysvo.FLAVOUR_VALIDATION if args.validation else pysvo.FLAVOUR_FAST
    if not os.environ.get("SVO_BENCH_NO_COHERENCE_ORDER"):     # (experiment switch, never set by default)
        flavour |= pysvo.BATCH_COHERENCE_ORDER               # direction-binned thread order, inside the timed call
    if not os.environ.get("SVO_BENCH_NO_LANE_REFILL"):        # (experiment switch, never set by default)
        flavour |= pysvo.BATCH_LANE_REFILL                   # persistent warps, finished lanes take the next rays
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
                       "lane_refill": bool(flavour & pysvo.BATCH_LANE_REFILL),
                       "l2": f"no flush: ray + result arrays {n_total * 41 / 1e6:.0f} MB per step exceed the 126 MB L2",
                       "parallelism": "rays sharded contiguously over ranks, octree replicated" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": n_total * 24, "d2h_bytes_per_step": n_total * 9,
                    "steps": e2e_steps, "note": "svo_raymarch_batch with pinned host arrays: 2 Mi-ray chunks alternate between two streams"},
            "gpu_launches": steps * world * (1 + (3 if flavour & pysvo.BATCH_COHERENCE_ORDER else 0)),
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
    """The reference's own CPU renderer on this box's host cores; never loads libsvo_b200.so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = pick_workload(args.workload)
    if w.get("kind") == "ao":
        return run_reference_ao(args, w)
    from oracle.pyoracle import Ref
    ref = Ref()
    steps = args.steps if args.steps is not None else 3
    warmup = args.warmup if args.warmup is not None else 1
    h = reference_tree(ref, w)[0]
    # the ratio's arm: the GPU arm's configuration (16 strips = the reference's NumThreads, Main.cpp:57 -- the image
    # depends on it), its 16 render threads on this box's cores
    r = reference_sample(w, steps, warmup, budget_s=120.0, strips=STRIPS, ref=ref, tree=h)
    # beside it: every host core busy (strips = threads = nproc) -- a slightly different image
    cores = os.cpu_count() or 1
    allc = None
    if cores != STRIPS:
        a = reference_sample(w, steps, warmup, budget_s=60.0, strips=cores, ref=ref, tree=h)
        allc = {"value": a["value"], "unit": "Mrays/s", "strips": a["strips"], "threads": a["cores"], "sample": a["sample"],
                "note": "strips = threads = nproc: not the GPU arm's image (tile grids are anchored per strip)"}
    ref.tree_destroy(h)
    line = {
        "impl": "reference", "metric": "Mrays/s ESVO traversal (coarse + fine raymarch calls per frame / time)",
        "value": r["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "description": w["text"], "width": w["width"], "height": w["height"],
                   "strips": r["strips"], "host_threads": r["cores"], "host_cores": r["host_cores"]},
        "cpu_baseline": {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": "reference",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if allc is not None:
        line["all_cores"] = allc
    emit(line)


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------

def kernel_source_hash():
    """sha256 over the kernel sources: ties an ncu capture under profiles/ to the code it was taken from."""
    import hashlib
    h = hashlib.sha256()
    for name in ("svo_traverse.cuh", "svo_kernels.cu", "svo_kernels.cuh"):
        h.update((ROOT / "sparse-voxel-octrees_b200" / "csrc" / name).read_bytes())
    return h.hexdigest()[:16]


def git_head():
    try:
        import subprocess
        return subprocess.run(["git", "-C", str(ROOT), "rev-parse", "--short", "HEAD"], capture_output=True, text=True,
                              timeout=10).stdout.strip() or None
    except Exception:  # noqa: BLE001
        return None


def run_own(args):
    """Frame workloads. ONE process drives all N GPUs through the library's multi-GPU handle (svo_multi_*: replicated
    octree, interleaved tile columns, worker thread per device, CUDA-event frame barrier, fine passes storing into
    device 0's framebuffer over NVLink). Under torchrun the other ranks initialise NCCL, meet rank 0 at the barriers
    and take part in the max-over-ranks reduction of the timing, but issue no work: their GPUs are driven by rank 0."""
    import torch
    import pysvo

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not pysvo.LIB_PATH.exists():
        raise SystemExit(f"{pysvo.LIB_PATH} missing: run __graft_entry__.build() first (no fallback exists)")
    if pysvo.device_count() < 1:
        raise SystemExit("no CUDA device: this benchmark has no CPU fallback (use --impl reference for the CPU arm)")
    w = pick_workload(args.workload)
    if w.get("kind") == "ao":
        return run_own_ao(args, w)
    torch.cuda.set_device(local_rank)
    dist = idle = None
    if world > 1:
        import datetime
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(minutes=30))
        probe = torch.ones(1, device="cuda")
        dist.all_reduce(probe)                       # NCCL up on all N ranks / GPUs
        assert int(probe.item()) == world
        # the long waits (rank 0 renders, the others idle) go through a CPU group: a pending NCCL barrier would be a
        # spinning kernel on every idle rank's GPU -- the GPUs rank 0 is timing
        idle = dist.new_group(backend="gloo", timeout=datetime.timedelta(minutes=30))

    def barrier():
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier(group=idle)

    def max_over_ranks(x):
        if dist is None:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=idle)
        return float(t.item())

    # GPUs rendering every frame: the torchrun world size, or --gpus when launched as a plain process (no NCCL then:
    # the frame barrier is the library's, CUDA events across devices)
    n_dev = world if world > 1 else max(1, min(args.gpus, pysvo.device_count()))
    steps = args.steps if args.steps is not None else 300
    warmup = max(args.warmup if args.warmup is not None else 10, 3)
    if rank != 0:
        barrier()                 # scene ready, octree replicated, warm-up done
        barrier()                 # timed region over
        max_over_ranks(0.0)
        barrier()                 # e2e over
        max_over_ranks(0.0)
        barrier()
        dist.destroy_process_group()
        return

    ensure_scene(w, 0, lambda: None)
    W, H = w["width"], w["height"]
    flavour = pysvo.FLAVOUR_VALIDATION if args.validation else pysvo.FLAVOUR_FAST
    devices = tuple(range(n_dev))
    if os.environ.get("SVO_BENCH_DEVICES"):      # (experiment switch, never set by default: e.g. "0,0" = two replicas on one GPU)
        devices = tuple(int(x) for x in os.environ["SVO_BENCH_DEVICES"].split(","))
        n_dev = len(devices)
    multi = pysvo.MultiOctree(w["path"], devices=devices)
    tree = multi.tree(0)
    cams = [pysvo.orbit_camera(*c) for c in cameras(pysvo, w, ORBIT)]
    path = lambda first, count: [cams[k % ORBIT] for k in range(first, first + count)]  # noqa: E731
    nbytes = W * H * 4

    # ---- device-timed: W warm-up frames, then rounds of exactly K frames (one svo_multi_render_sequence call each:
    # four frames in flight, CUDA events on device 0 around the round). Rounds repeat until the timed region is
    # >= 0.6 s so that the clock sampler sees it; the MEDIAN round is reported.
    sampler = ClockSampler(local_rank)
    sampler.start()
    multi.render_sequence(path(0, warmup), W, H, strips=STRIPS, flavour=flavour, output=pysvo.OUTPUT_DEVICE)
    barrier()
    t_begin = time.perf_counter()
    rounds = []
    while (not rounds or sum(r.device_ms for r in rounds) < 600.0) and len(rounds) < 400:
        rounds.append(multi.render_sequence(path(warmup, steps), W, H, strips=STRIPS, flavour=flavour,
                                            output=pysvo.OUTPUT_DEVICE))
    t_end = time.perf_counter()
    barrier()
    rounds.sort(key=lambda r: r.device_ms)
    med = rounds[len(rounds) // 2]
    total_ms = max_over_ranks(med.device_ms)
    total_rays = med.rays
    value = total_rays / (total_ms * 1e-3) / 1e6
    device_launches = int(med.kernel_launches)
    n_lanes, dev_run = int(med.lanes), int(med.tile_run)

    # ---- end to end: every frame of the same camera path lands in page-locked host memory (ring of four frames);
    # with N > 1 every GPU ships the stripes it rendered over its own PCIe link. Host clock around the call.
    e2e_steps = min(steps, 300)
    ring = [pysvo.PinnedArray((H, W), np.uint32) for _ in range(int(os.environ.get("SVO_MULTI_LANES", "0") or 0) or 6)]
    hosts = [r.array for r in ring]
    multi.render_sequence(path(0, 8), W, H, strips=STRIPS, flavour=flavour, output=pysvo.OUTPUT_HOST, host_frames=hosts)
    e2e_rounds = []
    while (not e2e_rounds or sum(r.wall_ms for r in e2e_rounds) < 600.0) and len(e2e_rounds) < 100:
        e2e_rounds.append(multi.render_sequence(path(warmup, e2e_steps), W, H, strips=STRIPS, flavour=flavour,
                                                output=pysvo.OUTPUT_HOST, host_frames=hosts))
    barrier()
    e2e_rounds.sort(key=lambda r: r.wall_ms)
    e2e_med = e2e_rounds[len(e2e_rounds) // 2]
    e2e_ms = max_over_ranks(e2e_med.wall_ms)
    e2e_value = e2e_med.rays / (e2e_ms * 1e-3) / 1e6
    k_last = warmup + e2e_steps - 1
    last_host = hosts[(e2e_steps - 1) % len(hosts)].copy()       # the last e2e frame, as assembled by all devices
    # ---- beside it (not the headline: the reference's framebuffer is RGBA words): the same frames shipped as (grey, alpha)
    # byte pairs, SVO_PIXELS_GREY8A8 -- half the bytes on host links that are the limit at 4 and 8 GPUs and for small frames
    ring16 = [pysvo.PinnedArray((H, W), np.uint16) for _ in range(len(ring))]
    hosts16 = [r.array for r in ring16]
    multi.render_sequence(path(0, 8), W, H, strips=STRIPS, flavour=flavour, output=pysvo.OUTPUT_HOST, host_frames=hosts16,
                          pixel_format=pysvo.PIXELS_GREY8A8)
    g_rounds = []
    while (not g_rounds or sum(r.wall_ms for r in g_rounds) < 400.0) and len(g_rounds) < 100:
        g_rounds.append(multi.render_sequence(path(warmup, e2e_steps), W, H, strips=STRIPS, flavour=flavour, output=pysvo.OUTPUT_HOST,
                                              host_frames=hosts16, pixel_format=pysvo.PIXELS_GREY8A8))
    g_rounds.sort(key=lambda r: r.wall_ms)
    g_med = g_rounds[len(g_rounds) // 2]
    grey_identical = bool(np.array_equal(pysvo.expand_grey8a(hosts16[(e2e_steps - 1) % len(hosts16)]), last_host))
    sampler.stop()

    # ---- parity on the benchmark's own scene and size: the frame all N devices assemble against the oracle
    from oracle.pyoracle import Port
    port = Port()
    words, center = pysvo.oct_read(w["path"])
    roof_cams = [0, 25] if n_dev == 1 else [0]
    parity = {}
    cam = cams[k_last % ORBIT]
    f = port.frame_constants(np.array(cam.model[:], np.float32), np.array(cam.view[:], np.float32), center, W, H, STRIPS)
    want, _, cc, cf = port.render_frame(words, f)
    parity["e2e_last_frame_identical_pixels"] = float((last_host == want).mean())
    val, vst = multi.render_frame(cam, W, H, strips=STRIPS, flavour=pysvo.FLAVOUR_VALIDATION, rgba=hosts[0])
    parity["validation_identical_pixels"] = float((val == want).mean())
    fast, _ = multi.render_frame(cam, W, H, strips=STRIPS, flavour=pysvo.FLAVOUR_FAST, rgba=hosts[1])
    parity["fast_identical_pixels"] = float((fast == want).mean())
    parity["oracle_rays"] = int(cc.rays + cf.rays)
    parity["gpu_rays"] = int(vst.coarse_rays + vst.fine_rays)
    if n_dev > 1:
        alone, _, _ = tree.render_frame(cam, W, H, strips=STRIPS, flavour=flavour)
        parity["e2e_host_frame_identical_to_single_rank"] = bool(np.array_equal(alone, last_host))
    coarse_b_ray = 4.0 * cc.words / max(cc.rays, 1)
    fine_b_ray = 4.0 * cf.words / max(cf.rays, 1)

    # ---- roofline of the dominant kernel (fine pass), N = 1: algorithmic bytes from the instrumented oracle on the
    # same cameras / the kernel's own duration (CUDA events around classifier + fine pass on the launching stream,
    # svo_frame_stats.fine_ms), both measured in this run
    peak, peak_src = measured_peak()
    roofline = None
    if n_dev == 1:
        fb = pysvo.DeviceBuffer(local_rank, nbytes)
        stream = torch.cuda.current_stream().cuda_stream
        alg_bytes, fine_ms_sum, fine_rays_sum, node_bytes_sum, shares = 0.0, 0.0, 0, 0, []
        for k in roof_cams:
            ck = cams[k]
            fk = port.frame_constants(np.array(ck.model[:], np.float32), np.array(ck.view[:], np.float32), center, W, H, STRIPS)
            _, _, _, cfk = port.render_frame(words, fk)
            samples = []
            for _ in range(5):
                st = tree.render_frame_device(ck, W, H, fb.ptr, strips=STRIPS, flavour=flavour, stream=stream, want_stats=True)
                samples.append((st.coarse_ms, st.fine_ms))
            samples.sort(key=lambda x: x[1])
            c_ms, f_ms = samples[len(samples) // 2]
            alg_bytes += 4 * cfk.words + 4 * cfk.rays
            node_bytes_sum += 4 * cfk.words
            fine_rays_sum += cfk.rays
            fine_ms_sum += f_ms
            shares.append(f_ms / (c_ms + f_ms))
        achieved = alg_bytes / (fine_ms_sum * 1e-3) / 1e9
        traffic = profile_traffic(w["name"]) or {}
        src_hash = kernel_source_hash()
        # the capture is quoted only if it profiled THIS machine code: the hash of the fine-pass kernel's SASS in the library
        # being run (cuobjdump), or -- where cuobjdump is missing -- the hash of the kernel sources
        from tools.ncu_summary import kernel_sass_hash
        sass_hash = kernel_sass_hash(pysvo.LIB_PATH)
        if sass_hash and traffic.get("kernel_sass_hash"):
            fresh = traffic.get("kernel_sass_hash") == sass_hash
        else:
            fresh = traffic.get("kernel_source_hash") == src_hash
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic.get("dram_bytes_per_launch") if fresh else None,
                    "kernel": "finePassKernel<FAST>", "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes / len(roof_cams),
                    "bytes_per_fine_ray": {"node_words": node_bytes_sum / max(fine_rays_sum, 1), "pixel_store": 4.0},
                    "launch_ms": fine_ms_sum / len(roof_cams),
                    "kernel_share_of_step": float(np.mean(shares)),
                    "kernel_source_hash": src_hash, "kernel_sass_hash": sass_hash, "source_commit": git_head(),
                    "ncu_capture": {"fresh": fresh, "capture_kernel_source_hash": traffic.get("kernel_source_hash"),
                                    "capture_kernel_sass_hash": traffic.get("kernel_sass_hash"),
                                    "capture_commit": traffic.get("commit"), "source": traffic.get("source")},
                    "note": "instruction-issue bound pointer chasing with SIMT divergence, not HBM-bound: node "
                            "fetches hit L1/L2 (see profiles/ and DESIGN.md section 4)"}
        if fresh:
            roofline["ncu"] = {k: traffic.get(k) for k in ("l1_hit_pct", "l2_hit_pct", "issue_active_pct", "threads_per_instruction",
                                                         "l2_throughput_pct", "l1_throughput_pct")}
            if traffic.get("warp_instructions") and traffic.get("sm_cycles_elapsed"):
                slots = 148 * 4 * float(traffic["sm_cycles_elapsed"])
                roofline["issue_roofline"] = {
                    "warp_instructions_per_launch": traffic["warp_instructions"], "issue_slots_per_launch": slots,
                    "frac": traffic["warp_instructions"] / slots,
                    "note": "fraction of the SMs' issue slots (148 SMs x 4 schedulers x elapsed cycles of the ncu capture) "
                            "that issued an instruction of this kernel: the limit this pass actually runs against"}
        else:
            roofline["ncu_capture"]["note"] = ("profiles/traffic.json was captured from other kernel code (or is absent): "
                                               "traffic / ncu counters are withheld rather than reported stale")
        fb.free()

    # ---- CPU baseline: the reference's own renderer on this box's host cores (bounded sample, ~10-30 s)
    cpu = None
    if n_dev == 1 and not args.no_cpu_baseline:
        try:
            r = reference_sample(w, steps=20, warmup=2, budget_s=25.0, strips=STRIPS)
            cpu = {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": "reference", "sample": r["sample"],
                   "strips": r["strips"], "host_cores": r["host_cores"]}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": f"unavailable: {e!r}"}

    clocks = sampler.summary(t_begin, t_end)
    e2e_launches = int(e2e_med.kernel_launches)
    line = {
        "metric": "Mrays/s ESVO traversal (coarse + fine raymarch calls per frame / time)",
        "value": value, "unit": "Mrays/s", "n_gpus": n_dev, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "description": w["text"], "width": W, "height": H, "strips": STRIPS,
                   "flavour": "validation" if args.validation else "fast",
                   "octree_words": tree.n_words, "octree_depth": tree.depth,
                   "parallelism": ("one process, svo_multi_*: replicated octree, tile columns dealt in stripes of "
                                   f"{dev_run} to {n_dev} GPUs, fine passes store into GPU 0's framebuffer over NVLink, "
                                   "CUDA-event frame barrier (no NCCL on the data path)") if n_dev > 1 else "single GPU",
                   "l2": f"no flush: octree {tree.n_words * 4 / 1e6:.0f} MB vs 126 MB L2, camera moves every step",
                   "pipelining": f"{n_lanes} frames in flight; beam passes run ahead on internal streams",
                   "timed_region": {"rounds": len(rounds), "reported": "median round", "steps_per_round": steps,
                                    "seconds": (t_end - t_begin),
                                    "round_ms_min_med_max": [rounds[0].device_ms, med.device_ms, rounds[-1].device_ms]},
                   "rays_per_frame_mean": total_rays / steps, "ms_per_frame": total_ms / steps},
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 128, "d2h_bytes_per_step": nbytes,
                "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "rounds": len(e2e_rounds),
                "note": "svo_multi_render_sequence(SVO_OUTPUT_HOST): per-step input is the 128 B camera (kernel parameters), "
                        "the octree stays resident; every frame is copied to page-locked host memory, " + (
                            f"copy engine, {int(e2e_med.lanes)} frames in flight" if n_dev == 1 else
                            f"every GPU ships the stripes it rendered ({int(e2e_med.tile_run)} tile columns wide) into the ONE "
                            f"host frame itself (1 / N of the frame per PCIe link, copy engine), {int(e2e_med.lanes)} frames in flight")},
        "e2e_grey8a8": {"value": g_med.rays / (g_med.wall_ms * 1e-3) / 1e6, "unit": "Mrays/s", "d2h_bytes_per_step": nbytes // 2,
                        "ms_per_step": g_med.wall_ms / e2e_steps, "rounds": len(g_rounds),
                        "expands_to_the_rgba_frame": grey_identical,
                        "note": "NOT the headline e2e: the same sequence with desc.pixel_format = SVO_PIXELS_GREY8A8 (grey + coverage, "
                                "two bytes per pixel, packed on the GPU; svo_pixels_expand_grey8a gives back the reference's words)"},
        "gpu_launches": device_launches,
        "gpu_launches_note": "kernels of the reported device-timed round, all GPUs (beam pass + tile classifier + fine pass "
                             f"per frame and GPU); the e2e round launched {e2e_launches}",
        "clocks": clocks,
        "parity": parity,
        "bytes_per_ray": {"coarse_node": coarse_b_ray, "fine_node": fine_b_ray},
    }
    if roofline is not None:
        line["roofline"] = roofline
    if cpu is not None:
        line["cpu_baseline"] = cpu
    emit(line)
    multi.close()
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
