// Several GPUs of one node in one process (include/svo_b200.h, svo_multi_*).
//
// What it replaces in the reference: the viewer cuts the frame into NumThreads strips, spawns one render thread per
// strip (src/Main.cpp:351-367) and brackets every frame with a two-phase barrier (src/Main.cpp:217-219,
// src/ThreadBarrier.cpp:41-59). Here the unit is the GPU: the node array is replicated, tile columns are dealt to the
// devices in stripes (svo_frame_desc.tile_rank / tile_world of the single-device path), one host thread per device
// enqueues that device's share of every frame, and the frame barrier is CUDA events waited for on streams -- across
// devices too. The host never waits for the GPU inside a sequence except to hand a finished host frame to the caller.
#include "svo_capi_internal.hpp"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <thread>

#include "oct_io.hpp"

namespace {

constexpr int kLanes = 8;       // most frames in flight (streams, framebuffers, events per device)
// frames in flight: hides the fine pass's long-ray tail and the frame barrier. Measured on one 8-GPU box (c3, device-
// timed, N = 1 on the same box 10.00 Grays/s): 4 / 6 / 8 lanes -> 72.6 / 78.0 / 78.3 Grays/s at N = 8 (0.91 / 0.975 / 0.98 of
// linear); at N = 1 the lanes beyond four change nothing (10.01 / 10.01 / 10.01).
constexpr int kDefaultLanes = 6;

// SVO_MULTI_LANES (experiment switch): frames in flight, 1 .. 8
int lanesInFlight() {
    static const int lanes = [] {
        const char *e = getenv("SVO_MULTI_LANES");
        const int v = e ? atoi(e) : 0;
        return v >= 1 && v <= kLanes ? v : kDefaultLanes;
    }();
    return lanes;
}

struct Job {
    const svo_camera *cams = nullptr;
    int nFrames = 0;
    svo_frame_desc desc{};
    int output = SVO_OUTPUT_DEVICE;
    uint32_t *const *hostFrames = nullptr;
    int nHost = 0;
    int lanes = kDefaultLanes;
    int tileRun = 1;                    // stripe width of this sequence, in tile columns (per job: handles do not share it)
};

struct Replica {
    int index = 0, device = 0;
    svo_tree *tree = nullptr;
    cudaStream_t lane[kLanes] = {};
    cudaStream_t copy = nullptr;
    uint32_t *fb[kLanes] = {};          // local framebuffers (host output)
    size_t fbBytes = 0;
    uint16_t *fb16[kLanes] = {};        // ... packed to (grey, alpha) pairs for SVO_PIXELS_GREY8A8 host frames
    size_t fb16Bytes = 0;
    cudaEvent_t done[kLanes] = {};      // the device's share of the frame in lane l is rendered
    cudaEvent_t copied[kLanes] = {};    // ... and has left fb[l] for host memory
    bool copiedRecorded[kLanes] = {};
    std::thread worker;
    std::atomic<int64_t> issued{0};     // frames of the current sequence this device has enqueued
    std::atomic<bool> failed{false};
    int status = SVO_OK;
    std::string error;
    uint64_t launches = 0;
};

} // namespace

struct svo_multi {
    std::vector<std::unique_ptr<Replica>> reps;
    bool peer = true;                   // every device can store into devices[0]'s memory
    uint32_t *gather[kLanes] = {};      // devices[0]: the frames of SVO_OUTPUT_DEVICE sequences
    size_t gatherBytes = 0;
    cudaStream_t gatherStream = nullptr;
    cudaEvent_t slotFree[kLanes] = {};  // gather[l] may be overwritten (recorded on gatherStream)
    bool slotFreeRecorded[kLanes] = {};
    cudaEvent_t seqStart = nullptr, seqStop = nullptr;
    int lastFrames = 0, lastLanes = 4;

    std::mutex callMutex;               // one sequence at a time
    std::mutex m;
    std::condition_variable cvJob, cvDone;
    Job job;
    std::atomic<uint64_t> generation{0};   // bumped (under m) when `job` is new; workers that are still spinning see it without the lock
    int finished = 0;
    bool quit = false;
    std::atomic<int> ready{0};              // workers that have taken the current job and prepared their buffers
    std::atomic<uint64_t> go{0};            // == generation once the sequence's start event is recorded: workers may enqueue
    std::atomic<int64_t> consumed{0};   // frames of the current sequence whose slot has been released
    std::atomic<bool> abort{false};
};

namespace {

using svo_detail::checkDesc;

void relax(unsigned &spins) {
    if (++spins < 2000) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    } else {
        std::this_thread::yield();
    }
}

// Host leg with several devices: strided copies on the copy engine (default) or a kernel storing into the mapped host
// frame (SVO_MULTI_HOST_COPY=kernel; measured against each other in profiles/experiments/r02_host_leg.md).
bool hostCopyByKernel() {
    static const bool byKernel = [] {
        const char *e = getenv("SVO_MULTI_HOST_COPY");
        return e && !strcmp(e, "kernel");
    }();
    return byKernel;
}

// One device's share of a sequence. Runs on the device's worker thread with the device current.
int runSequence(svo_multi *M, Replica &rep, const Job &job, uint64_t generation) {
    const int N = int(M->reps.size());
    svo_tree *tree = rep.tree;
    svo_frame_desc desc = job.desc;
    desc.tile_rank = rep.index;
    desc.tile_world = N;
    const svo::TileShare share = svo::tileShare(rep.index, N, job.tileRun);
    const size_t frameBytes = size_t(desc.width)*size_t(desc.height)*sizeof(uint32_t);
    FramePlan *plan = nullptr;
    {
        std::lock_guard<std::mutex> lock(tree->mutex);
        int st = svo_detail::getPlan(tree, desc.width, desc.height, desc.strips, &plan);
        if (st != SVO_OK) return st;
    }
    if (job.output == SVO_OUTPUT_HOST) {
        if (rep.fbBytes < frameBytes) {
            SVO_CUDA(cudaDeviceSynchronize());
            for (int l = 0; l < kLanes; ++l) {
                if (rep.fb[l]) cudaFree(rep.fb[l]);
                rep.fb[l] = nullptr;
                rep.copiedRecorded[l] = false;
            }
            rep.fbBytes = 0;
            for (int l = 0; l < kLanes; ++l) {
                SVO_CUDA(cudaMalloc(&rep.fb[l], frameBytes));
                SVO_CUDA(cudaMemset(rep.fb[l], 0, frameBytes));
            }
            rep.fbBytes = frameBytes;
        }
        if (desc.pixel_format == SVO_PIXELS_GREY8A8 && rep.fb16Bytes < frameBytes/2) {
            SVO_CUDA(cudaDeviceSynchronize());
            for (int l = 0; l < kLanes; ++l) {
                if (rep.fb16[l]) cudaFree(rep.fb16[l]);
                rep.fb16[l] = nullptr;
            }
            rep.fb16Bytes = 0;
            for (int l = 0; l < kLanes; ++l) SVO_CUDA(cudaMalloc(&rep.fb16[l], frameBytes/2));
            rep.fb16Bytes = frameBytes/2;
        }
    }
    // ready; the caller records the sequence's start event once every device is, then opens the gate
    M->ready.fetch_add(1, std::memory_order_acq_rel);
    {
        unsigned spins = 0;
        while (M->go.load(std::memory_order_acquire) != generation) {
            if (M->abort.load(std::memory_order_relaxed)) return fail(SVO_ERR_CUDA, "sequence aborted (another device failed)");
            relax(spins);
        }
    }
    SVO_CUDA(cudaMemsetAsync(plan->dFineTotal, 0, sizeof(unsigned long long), rep.lane[0]));
    cudaEvent_t zeroed = rep.done[0];   // the other lanes' classifiers add to the total, too: order them behind the memset
    SVO_CUDA(cudaEventRecord(zeroed, rep.lane[0]));
    for (int l = 1; l < job.lanes; ++l) SVO_CUDA(cudaStreamWaitEvent(rep.lane[l], zeroed, 0));

    for (int k = 0; k < job.nFrames; ++k) {
        const int l = k % job.lanes;
        unsigned spins = 0;
        while (M->consumed.load(std::memory_order_acquire) < int64_t(k) - job.lanes + 1) {   // the lane's previous frame has been let go
            if (M->abort.load(std::memory_order_relaxed)) return fail(SVO_ERR_CUDA, "sequence aborted (another device failed)");
            relax(spins);
        }
        cudaStream_t s = rep.lane[l];
        uint32_t *target;
        if (job.output == SVO_OUTPUT_DEVICE) {
            target = M->gather[l];
            if (M->slotFreeRecorded[l]) SVO_CUDA(cudaStreamWaitEvent(s, M->slotFree[l], 0));
        } else {
            target = rep.fb[l];
            if (rep.copiedRecorded[l]) SVO_CUDA(cudaStreamWaitEvent(s, rep.copied[l], 0));
        }
        uint32_t launches = 0;
        {
            std::lock_guard<std::mutex> lock(tree->mutex);
            int st = svo_detail::enqueueFrame(tree, plan, &job.cams[k], &desc, share, target, nullptr, s, false, &launches, nullptr);
            if (st != SVO_OK) return st;
        }
        rep.launches += launches;
        SVO_CUDA(cudaEventRecord(rep.done[l], s));
        if (job.output == SVO_OUTPUT_HOST) {
            uint32_t *host = job.hostFrames[k % job.nHost];
            SVO_CUDA(cudaStreamWaitEvent(rep.copy, rep.done[l], 0));
            const bool packed = desc.pixel_format == SVO_PIXELS_GREY8A8;
            const size_t pixelBytes = packed ? sizeof(uint16_t) : sizeof(uint32_t);
            const unsigned char *srcBytes = reinterpret_cast<const unsigned char *>(rep.fb[l]);
            bool shipped = false;
            if (packed && N > 1) {
                // two bytes per pixel, several devices: every device packs its stripes to (grey, alpha) pairs and stores them
                // straight into the mapped host frame, 256 contiguous bytes per warp. (Strided copies of such narrow rows are
                // bound by the copy engine's per-row cost, not by bytes: at N = 8, 480-byte rows took 0.286 ms per 4K frame
                // where the RGBA frame's 960-byte rows take 0.308.)
                void *mapped = nullptr;
                SVO_CUDA(cudaHostGetDevicePointer(&mapped, host, 0));
                SVO_CUDA(svo::launchPackGrey8a8(plan->dev, desc.width, desc.height, rep.fb[l], static_cast<uint16_t *>(mapped), share, rep.copy));
                ++rep.launches;
                shipped = true;
            } else if (packed) {
                // one device: packed on the GPU, then one contiguous copy on the copy engine
                SVO_CUDA(svo::launchPackGrey8a8(plan->dev, desc.width, desc.height, rep.fb[l], rep.fb16[l], share, rep.copy));
                ++rep.launches;
                srcBytes = reinterpret_cast<const unsigned char *>(rep.fb16[l]);
            }
            unsigned char *hostBytes = reinterpret_cast<unsigned char *>(host);
            if (shipped) {
                // nothing left to copy
            } else if (N == 1) {
                SVO_CUDA(cudaMemcpyAsync(hostBytes, srcBytes, frameBytes/sizeof(uint32_t)*pixelBytes, cudaMemcpyDeviceToHost, rep.copy));   // copy engine
            } else if (hostCopyByKernel() && !packed) {
                // this device's stripes only, stored by a kernel straight into the (mapped, page-locked) host frame
                void *mapped = nullptr;
                SVO_CUDA(cudaHostGetDevicePointer(&mapped, host, 0));
                SVO_CUDA(svo::launchCopyOwnedColumns(plan->dev, desc.width, desc.height, rep.fb[l], static_cast<uint32_t *>(mapped),
                                                     share, rep.copy));
                ++rep.launches;
            } else {
                // this device's stripes only, one strided copy per stripe on the copy engine (no SM involved)
                const int run = share.run;
                const size_t pitch = size_t(desc.width)*pixelBytes;
                for (int tx0 = rep.index*run; tx0 < plan->dev.tileCols; tx0 += N*run) {
                    const int x0 = tx0*8, x1 = std::min((tx0 + run)*8, desc.width);
                    SVO_CUDA(cudaMemcpy2DAsync(hostBytes + size_t(x0)*pixelBytes, pitch, srcBytes + size_t(x0)*pixelBytes, pitch,
                                               size_t(x1 - x0)*pixelBytes, size_t(desc.height), cudaMemcpyDeviceToHost, rep.copy));
                }
            }
            SVO_CUDA(cudaEventRecord(rep.copied[l], rep.copy));
            rep.copiedRecorded[l] = true;
        }
        rep.issued.store(k + 1, std::memory_order_release);
    }
    return SVO_OK;
}

void workerMain(svo_multi *M, Replica *rep) {
    cudaSetDevice(rep->device);
    uint64_t seen = 0;
    for (;;) {
        Job job;
        {
            // sequences usually come back to back (a benchmark's rounds, a player's chunks): stay hot for 2 ms before sleeping
            const auto spinUntil = std::chrono::steady_clock::now() + std::chrono::milliseconds(2);
            while (M->generation.load(std::memory_order_acquire) == seen && std::chrono::steady_clock::now() < spinUntil) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            std::unique_lock<std::mutex> lock(M->m);
            M->cvJob.wait(lock, [&] { return M->quit || M->generation.load(std::memory_order_relaxed) != seen; });
            if (M->quit) return;
            seen = M->generation.load(std::memory_order_relaxed);
            job = M->job;
        }
        rep->launches = 0;
        rep->status = runSequence(M, *rep, job, seen);
        if (rep->status != SVO_OK) {
            rep->error = g_lastError;
            rep->failed.store(true, std::memory_order_release);
            M->abort.store(true, std::memory_order_release);
        }
        {
            std::lock_guard<std::mutex> lock(M->m);
            ++M->finished;
        }
        M->cvDone.notify_all();
    }
}

int destroyMulti(svo_multi *M) {
    if (!M) return SVO_OK;
    {
        std::lock_guard<std::mutex> lock(M->m);
        M->quit = true;
    }
    M->cvJob.notify_all();
    for (auto &r : M->reps) if (r->worker.joinable()) r->worker.join();
    for (auto &r : M->reps) {
        if (!r->tree) continue;         // never came up (e.g. a device index out of range): nothing of it exists
        DeviceScope scope(r->device);
        cudaDeviceSynchronize();
        for (int l = 0; l < kLanes; ++l) {
            if (r->lane[l]) cudaStreamDestroy(r->lane[l]);
            if (r->fb[l]) cudaFree(r->fb[l]);
            if (r->fb16[l]) cudaFree(r->fb16[l]);
            if (r->done[l]) cudaEventDestroy(r->done[l]);
            if (r->copied[l]) cudaEventDestroy(r->copied[l]);
        }
        if (r->copy) cudaStreamDestroy(r->copy);
    }
    if (!M->reps.empty() && M->reps[0]->tree) {
        DeviceScope scope(M->reps[0]->device);
        for (int l = 0; l < kLanes; ++l) {
            if (M->gather[l]) cudaFree(M->gather[l]);
            if (M->slotFree[l]) cudaEventDestroy(M->slotFree[l]);
        }
        if (M->gatherStream) cudaStreamDestroy(M->gatherStream);
        if (M->seqStart) cudaEventDestroy(M->seqStart);
        if (M->seqStop) cudaEventDestroy(M->seqStop);
    }
    for (auto &r : M->reps) if (r->tree) svo_tree_destroy(r->tree);
    delete M;
    cudaGetLastError();                 // leave no stale error behind for the caller's next launch check
    return SVO_OK;
}

int createMulti(const uint32_t *words, uint64_t nWords, const float center[3], const int *devices, int nDevices,
                svo_multi **out) {
    if (!words || !center || !devices || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_multi_create: null argument");
    *out = nullptr;
    if (nDevices < 1 || nDevices > 64) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_multi_create: %d devices", nDevices);
    std::unique_ptr<svo_multi, int (*)(svo_multi *)> M(new svo_multi, destroyMulti);
    // replicas are uploaded in parallel; the node array is walked once (by the first)
    std::vector<svo_tree *> trees(size_t(nDevices), nullptr);
    std::vector<int> status(size_t(nDevices), SVO_OK);
    std::vector<std::string> errors(static_cast<size_t>(nDevices));
    {
        std::vector<std::thread> pool;
        for (int i = 0; i < nDevices; ++i)
            pool.emplace_back([&, i] {
                status[size_t(i)] = svo_detail::createTreeOnDevice(words, nWords, center, devices[i], i == 0, &trees[size_t(i)]);
                if (status[size_t(i)] != SVO_OK) errors[size_t(i)] = g_lastError;
            });
        for (auto &t : pool) t.join();
    }
    for (int i = 0; i < nDevices; ++i) {
        std::unique_ptr<Replica> r(new Replica);
        r->index = i;
        r->device = devices[i];
        r->tree = trees[size_t(i)];
        M->reps.push_back(std::move(r));
    }
    for (int i = 0; i < nDevices; ++i)
        if (status[size_t(i)] != SVO_OK) return fail(status[size_t(i)], "device %d: %s", devices[i], errors[size_t(i)].c_str());

    const int dev0 = devices[0];
    for (int i = 0; i < nDevices; ++i) {
        Replica &r = *M->reps[size_t(i)];
        SVO_DEVICE(r.device);
        for (int l = 0; l < kLanes; ++l) {
            SVO_CUDA(cudaStreamCreateWithFlags(&r.lane[l], cudaStreamNonBlocking));
            SVO_CUDA(cudaEventCreateWithFlags(&r.done[l], cudaEventDisableTiming));
            SVO_CUDA(cudaEventCreateWithFlags(&r.copied[l], cudaEventDisableTiming));
        }
        // high priority: the packing kernel of SVO_PIXELS_GREY8A8 frames must not queue behind the pending blocks of the
        // next frames' fine passes (the copies themselves run on the copy engines)
        int prioLow = 0, prioHigh = 0;
        SVO_CUDA(cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh));
        SVO_CUDA(cudaStreamCreateWithPriority(&r.copy, cudaStreamNonBlocking, prioHigh));
        if (r.device != dev0) {
            int can = 0;
            SVO_CUDA(cudaDeviceCanAccessPeer(&can, r.device, dev0));
            if (can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(dev0, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
                if (e != cudaSuccess) return failCuda(e, "cudaDeviceEnablePeerAccess");
            } else {
                M->peer = false;
            }
        }
    }
    {
        SVO_DEVICE(dev0);
        SVO_CUDA(cudaStreamCreateWithFlags(&M->gatherStream, cudaStreamNonBlocking));
        for (int l = 0; l < kLanes; ++l) SVO_CUDA(cudaEventCreateWithFlags(&M->slotFree[l], cudaEventDisableTiming));
        SVO_CUDA(cudaEventCreate(&M->seqStart));
        SVO_CUDA(cudaEventCreate(&M->seqStop));
    }
    for (auto &r : M->reps) r->worker = std::thread(workerMain, M.get(), r.get());
    *out = M.release();
    return SVO_OK;
}

// widest stripe <= 32 tile columns that still deals every device the same number of columns: a stripe's row is what one
// DMA burst of the host leg carries (measured at N = 2, 4K frames, copy engine: 15 / 30 / 60 columns wide -> 60.9 / 67.9 /
// 66.7 GB/s into host memory; wider stripes also unbalance the devices' shares of the rays)
int hostStripeRun(int width, int nDevices) {
    const int tileCols = (width - 1)/8 + 1;
    if (const char *e = getenv("SVO_MULTI_HOST_RUN")) {      // experiment switch
        const int v = atoi(e);
        if (v > 0) return v;
    }
    for (int r = 32; r > 4; --r)
        if (tileCols % (nDevices*r) == 0) return r;
    return 4;
}

} // namespace

extern "C" {

int svo_multi_create_from_words(const uint32_t *words, uint64_t n_words, const float center[3], const int *devices,
                                int n_devices, svo_multi **out) {
    return createMulti(words, n_words, center, devices, n_devices, out);
}

int svo_multi_load_oct(const char *path, const int *devices, int n_devices, svo_multi **out) {
    if (!path || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_multi_load_oct: null argument");
    *out = nullptr;
    svo::OctFile f;
    std::string err;
    int status = 0;
    if (!svo::readOctFile(path, f, err, status)) return fail(status, "%s", err.c_str());
    int st = createMulti(f.words, f.nWords, f.center, devices, n_devices, out);
    free(f.words);
    return st;
}

int svo_multi_destroy(svo_multi *m) { return destroyMulti(m); }

int svo_multi_device_count(const svo_multi *m) { return m ? int(m->reps.size()) : 0; }

svo_tree *svo_multi_tree(svo_multi *m, int index) {
    if (!m || index < 0 || index >= int(m->reps.size())) {
        fail(SVO_ERR_INVALID_ARGUMENT, "svo_multi_tree: index %d out of range", index);
        return nullptr;
    }
    return m->reps[size_t(index)]->tree;
}

int svo_multi_render_sequence(svo_multi *M, const svo_camera *cams, int n_frames, const svo_frame_desc *desc_in, int output,
                              uint32_t *const *host_frames, int n_host_frames, svo_frame_callback on_frame, void *user,
                              svo_sequence_stats *stats) {
    if (!M || !cams || !desc_in) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_multi_render_sequence: null argument");
    if (n_frames < 1) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_multi_render_sequence: %d frames", n_frames);
    if (output != SVO_OUTPUT_DEVICE && output != SVO_OUTPUT_HOST) return fail(SVO_ERR_INVALID_ARGUMENT, "unknown output mode %d", output);
    svo_frame_desc desc = *desc_in;
    desc.tile_rank = 0;
    desc.tile_world = 1;
    int st = checkDesc(&desc);
    if (st != SVO_OK) return st;
    const int N = int(M->reps.size());
    if (output == SVO_OUTPUT_HOST) {
        if (!host_frames || n_host_frames < 1) return fail(SVO_ERR_INVALID_ARGUMENT, "SVO_OUTPUT_HOST needs at least one host frame");
        for (int i = 0; i < n_host_frames; ++i)
            if (!host_frames[i]) return fail(SVO_ERR_INVALID_ARGUMENT, "host frame %d is null", i);
    } else if (!M->peer) {
        return fail(SVO_ERR_UNSUPPORTED, "SVO_OUTPUT_DEVICE needs peer access from every device to devices[0]; use SVO_OUTPUT_HOST");
    }
    if (output == SVO_OUTPUT_DEVICE && desc.pixel_format != SVO_PIXELS_RGBA8)
        return fail(SVO_ERR_UNSUPPORTED, "SVO_OUTPUT_DEVICE frames are SVO_PIXELS_RGBA8 (the packed format is for host frames)");
    std::lock_guard<std::mutex> call(M->callMutex);
    const auto wallStart = std::chrono::steady_clock::now();
    const size_t frameBytes = size_t(desc.width)*size_t(desc.height)*sizeof(uint32_t);
    const int dev0 = M->reps[0]->device;
    Job job;
    job.cams = cams;
    job.nFrames = n_frames;
    job.desc = desc;
    job.output = output;
    job.hostFrames = host_frames;
    job.nHost = n_host_frames;
    job.lanes = output == SVO_OUTPUT_HOST ? std::min(lanesInFlight(), n_host_frames) : lanesInFlight();
    const int run = N > 1 ? (output == SVO_OUTPUT_HOST ? hostStripeRun(desc.width, N) : 4) : 1;
    job.tileRun = run;

    if (output == SVO_OUTPUT_DEVICE) {
        SVO_DEVICE(dev0);
        if (M->gatherBytes < frameBytes) {
            SVO_CUDA(cudaDeviceSynchronize());
            for (int l = 0; l < kLanes; ++l) {
                if (M->gather[l]) cudaFree(M->gather[l]);
                M->gather[l] = nullptr;
                M->slotFreeRecorded[l] = false;
            }
            M->gatherBytes = 0;
            for (int l = 0; l < kLanes; ++l) {
                SVO_CUDA(cudaMalloc(&M->gather[l], frameBytes));
                SVO_CUDA(cudaMemset(M->gather[l], 0, frameBytes));
            }
            M->gatherBytes = frameBytes;
        }
    }

    M->consumed.store(0, std::memory_order_relaxed);
    M->abort.store(false, std::memory_order_relaxed);
    M->ready.store(0, std::memory_order_relaxed);
    for (auto &r : M->reps) {
        r->issued.store(0, std::memory_order_relaxed);
        r->failed.store(false, std::memory_order_relaxed);
    }
    uint64_t generation = 0;
    {
        std::lock_guard<std::mutex> lock(M->m);
        M->job = job;
        M->finished = 0;
        generation = M->generation.load(std::memory_order_relaxed) + 1;
        M->generation.store(generation, std::memory_order_release);
    }
    M->cvJob.notify_all();
    // every worker has taken the job and has its buffers: only now does the sequence's clock start (the workers' wake-up is
    // not rendering time), and nothing of the sequence can start before it: the first frames wait for the lanes' events
    {
        unsigned spins = 0;
        while (M->ready.load(std::memory_order_acquire) < N) {
            bool dead = false;
            for (auto &r : M->reps) dead = dead || r->failed.load(std::memory_order_acquire);
            if (dead) break;
            relax(spins);
        }
    }
    if (output == SVO_OUTPUT_DEVICE) {
        SVO_DEVICE(dev0);
        cudaError_t e = cudaEventRecord(M->seqStart, M->gatherStream);
        for (int l = 0; l < kLanes && e == cudaSuccess; ++l) {
            e = cudaEventRecord(M->slotFree[l], M->gatherStream);
            M->slotFreeRecorded[l] = true;
        }
        if (e != cudaSuccess) M->abort.store(true, std::memory_order_release);
    }
    const auto gateOpened = std::chrono::steady_clock::now();
    M->go.store(generation, std::memory_order_release);

    // the frame barrier
    cudaError_t cudaErr = cudaSuccess;
    bool failed = false;
    {
        DeviceScope scope(dev0);
        for (int k = 0; k < n_frames && !failed; ++k) {
            const int l = k % job.lanes;
            for (auto &r : M->reps) {
                unsigned spins = 0;
                while (r->issued.load(std::memory_order_acquire) <= k) {
                    if (M->abort.load(std::memory_order_acquire)) { failed = true; break; }
                    relax(spins);
                }
                if (failed) break;
            }
            if (failed) break;
            if (output == SVO_OUTPUT_DEVICE) {
                for (auto &r : M->reps)
                    if (cudaErr == cudaSuccess) cudaErr = cudaStreamWaitEvent(M->gatherStream, r->done[l], 0);
                if (cudaErr == cudaSuccess) cudaErr = cudaEventRecord(M->slotFree[l], M->gatherStream);
            } else {
                for (auto &r : M->reps)
                    if (cudaErr == cudaSuccess) cudaErr = cudaEventSynchronize(r->copied[l]);
                if (cudaErr == cudaSuccess && on_frame) on_frame(user, k, host_frames[k % n_host_frames]);
            }
            if (cudaErr != cudaSuccess) {
                M->abort.store(true, std::memory_order_release);
                failed = true;
            }
            M->consumed.store(k + 1, std::memory_order_release);
        }
        if (!failed && output == SVO_OUTPUT_DEVICE) {
            cudaErr = cudaEventRecord(M->seqStop, M->gatherStream);
            if (cudaErr == cudaSuccess) cudaErr = cudaEventSynchronize(M->seqStop);
        }
    }
    {
        std::unique_lock<std::mutex> lock(M->m);
        M->cvDone.wait(lock, [&] { return M->finished == N; });
    }
    for (auto &r : M->reps)
        if (r->status != SVO_OK) return fail(r->status, "device %d: %s", r->device, r->error.c_str());
    if (cudaErr != cudaSuccess) return failCuda(cudaErr, "frame barrier");
    M->lastFrames = n_frames;
    M->lastLanes = job.lanes;
    const auto wallStop = std::chrono::steady_clock::now();

    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->frames = uint64_t(n_frames);
        svo::FramePlanDev p{};
        svo_detail::planGeometry(desc.width, desc.height, desc.strips, p);
        stats->coarse_rays = uint64_t(p.totalCorners)*uint64_t(n_frames);
        for (auto &r : M->reps) {
            SVO_DEVICE(r->device);
            FramePlan *plan = nullptr;
            {
                std::lock_guard<std::mutex> lock(r->tree->mutex);
                if ((st = svo_detail::getPlan(r->tree, desc.width, desc.height, desc.strips, &plan)) != SVO_OK) return st;
            }
            for (int l = 0; l < job.lanes; ++l) SVO_CUDA(cudaStreamSynchronize(r->lane[l]));
            unsigned long long fine = 0;
            SVO_CUDA(cudaMemcpy(&fine, plan->dFineTotal, sizeof fine, cudaMemcpyDeviceToHost));
            stats->fine_rays += fine;
            stats->kernel_launches += r->launches;
        }
        if (output == SVO_OUTPUT_DEVICE) {
            SVO_DEVICE(dev0);
            SVO_CUDA(cudaEventElapsedTime(&stats->device_ms, M->seqStart, M->seqStop));
        }
        // host clock from the moment every worker was ready to enqueue (their wake-up is not rendering time) to the last frame
        stats->wall_ms = float(std::chrono::duration<double, std::milli>(wallStop - gateOpened).count());
        (void)wallStart;
        stats->lanes = job.lanes;
        stats->tile_run = run;
    }
    return SVO_OK;
}

int svo_multi_render_frame(svo_multi *m, const svo_camera *cam, const svo_frame_desc *desc, uint32_t *rgba, svo_frame_stats *stats) {
    if (!m || !cam || !desc || !rgba) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_multi_render_frame: null argument");
    svo_sequence_stats seq;
    uint32_t *frames[1] = {rgba};
    // with several devices the host leg is a kernel storing into the host frame: it must be mapped for them
    const int N = int(m->reps.size());
    bool registered = false;
    if (N > 1) {
        cudaPointerAttributes attr{};
        cudaError_t e = cudaPointerGetAttributes(&attr, rgba);
        if (e != cudaSuccess || attr.type == cudaMemoryTypeUnregistered) {
            cudaGetLastError();
            const size_t bytes = size_t(desc->width)*size_t(desc->height)*(desc->pixel_format == SVO_PIXELS_GREY8A8 ? sizeof(uint16_t) : sizeof(uint32_t));
            e = cudaHostRegister(rgba, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
            if (e != cudaSuccess) return failCuda(e, "cudaHostRegister(host frame)");
            registered = true;
        }
    }
    int st = svo_multi_render_sequence(m, cam, 1, desc, SVO_OUTPUT_HOST, frames, 1, nullptr, nullptr, &seq);
    if (registered) cudaHostUnregister(rgba);
    if (st != SVO_OK) return st;
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->coarse_rays = seq.coarse_rays;
        stats->fine_rays = seq.fine_rays;
        stats->kernel_launches = uint32_t(seq.kernel_launches);
        svo::FramePlanDev p{};
        svo_detail::planGeometry(desc->width, desc->height, desc->strips, p);
        stats->tiles_total = uint64_t(p.totalTiles);
    }
    return SVO_OK;
}

int svo_multi_device_frame(svo_multi *m, int back, uint32_t **d_rgba) {
    if (!m || !d_rgba) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_multi_device_frame: null argument");
    if (m->lastFrames < 1 || !m->gather[0]) return fail(SVO_ERR_INVALID_ARGUMENT, "no SVO_OUTPUT_DEVICE sequence has been rendered");
    if (back < 0 || back >= m->lastLanes || back >= m->lastFrames) return fail(SVO_ERR_INVALID_ARGUMENT, "frame %d back is no longer held", back);
    *d_rgba = m->gather[(m->lastFrames - 1 - back) % m->lastLanes];
    return SVO_OK;
}

int svo_multi_raymarch_batch(svo_multi *m, uint64_t n, const float *o, const float *d, float ray_scale, int flavour,
                             uint8_t *hit, float *t, uint32_t *normal, uint64_t *voxel) {
    if (!m || (n && (!o || !d))) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_multi_raymarch_batch: null argument");
    const uint64_t N = m->reps.size();
    std::vector<int> status(size_t(N), SVO_OK);
    std::vector<std::string> errors(static_cast<size_t>(N));
    std::vector<std::thread> pool;
    for (uint64_t i = 0; i < N; ++i) {
        const uint64_t lo = n*i/N, hi = n*(i + 1)/N;
        if (hi == lo) continue;
        pool.emplace_back([=, &status, &errors] {
            status[size_t(i)] = svo_raymarch_batch(m->reps[size_t(i)]->tree, hi - lo, o + 3*lo, d + 3*lo, ray_scale, flavour,
                                                   hit ? hit + lo : nullptr, t ? t + lo : nullptr, normal ? normal + lo : nullptr,
                                                   voxel ? voxel + lo : nullptr);
            if (status[size_t(i)] != SVO_OK) errors[size_t(i)] = svo_last_error();
        });
    }
    for (auto &th : pool) th.join();
    for (uint64_t i = 0; i < N; ++i)
        if (status[size_t(i)] != SVO_OK) return fail(status[size_t(i)], "device %d: %s", m->reps[size_t(i)]->device, errors[size_t(i)].c_str());
    return SVO_OK;
}

} // extern "C"
