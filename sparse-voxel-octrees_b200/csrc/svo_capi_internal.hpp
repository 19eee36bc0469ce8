// Internals shared by the C ABI's translation units (svo_capi.cu: single-device entry points; svo_multi.cu: the
// multi-GPU handle): error channel, device scope, per-configuration frame plans, the tree handle.
#pragma once

#include "../../include/svo_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>

#include "svo_kernels.cuh"

extern thread_local std::string g_lastError;   // svo_capi.cu

inline int fail(int status, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_lastError = buf;
    return status;
}

inline int failCuda(cudaError_t e, const char *what) {
    int status = (e == cudaErrorMemoryAllocation) ? SVO_ERR_OUT_OF_MEMORY
               : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? SVO_ERR_NO_DEVICE : SVO_ERR_CUDA;
    return fail(status, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

#define SVO_CUDA(call)                                         \
    do {                                                       \
        cudaError_t e_ = (call);                               \
        if (e_ != cudaSuccess) return failCuda(e_, #call);     \
    } while (0)

// Makes `device` current for the scope, restoring the caller's device after.
struct DeviceScope {
    int previous = -1;
    cudaError_t error = cudaSuccess;
    explicit DeviceScope(int device) {
        error = cudaGetDevice(&previous);
        if (error == cudaSuccess && previous != device) error = cudaSetDevice(device);
    }
    ~DeviceScope() {
        int now = -1;
        if (previous >= 0 && cudaGetDevice(&now) == cudaSuccess && now != previous) cudaSetDevice(previous);
    }
};

#define SVO_DEVICE(device)                                     \
    DeviceScope scope_(device);                                \
    if (scope_.error != cudaSuccess) return failCuda(scope_.error, "cudaSetDevice")

// Everything one (width, height, strips) configuration needs on the device. Frames cycle through a
// ring of kRing slots (depth buffer, tile list, counters) so that the beam passes of the next frames
// can run ahead on the tree's two high-priority internal streams while earlier fine passes are
// still busy -- the beam pass is latency-bound (its longest ray), and with several GPUs sharing a
// frame it would otherwise be the critical path. Host-buffer frames additionally alternate between
// two staging framebuffers so the device->host copy of frame i overlaps the rendering of frame i+1.
constexpr int kRing = 8;
constexpr int kHostLanes = 4;   // host-buffer frames in flight (staging framebuffers, render streams, tickets)

struct FramePlan {
    svo::FramePlanDev dev{};
    float *dTables = nullptr;              // dxCoarse | dyCoarse | dxFine | dyFine
    float *dDepth[kRing] = {};             // totalCorners floats each
    svo::TileRecord *dTiles[kRing] = {};   // totalTiles records each (worst case: every tile rendered)
    svo::FrameCounters *dCounters[kRing] = {};
    uint32_t *dPrefix[kRing] = {};         // per tile: the traversal state its corner rays share (FAST fine pass)
    svo::FrameCounters *hCounters = nullptr;            // pinned, kRing entries
    unsigned long long *dFineTotal = nullptr;           // fine rays of every frame since it was last zeroed (frame sequences)
    cudaEvent_t coarseDone[kRing] = {};    // beam pass of the slot finished (internal stream)
    cudaEvent_t fineDone[kRing] = {};      // last fine pass that read the slot's depth / tile list
    cudaEvent_t laneReady[kRing] = {};     // what the caller had ordered on its stream before the frame (framebuffer free)
    cudaEvent_t prepDone[kRing] = {};      // tile list (+ prefix records) of the slot written
    cudaEvent_t timing[kRing][4] = {};     // coarse start/end, fine start/end (stats only)
    bool fineRecorded[kRing] = {};
    uint64_t frameNumber = 0;

    // host-buffer entry points only
    uint32_t *dRgba[kHostLanes] = {};                   // staging framebuffers (lazy)
    cudaEvent_t copyDone[kHostLanes] = {};              // device->host copy out of dRgba[i] finished
    bool copyRecorded[kHostLanes] = {};
    bool pending[kHostLanes] = {};                      // svo_render_frame_async issued, not yet waited for
    bool pendingStats[kHostLanes] = {};
    int pendingRing[kHostLanes] = {};
    uint32_t pendingLaunches[kHostLanes] = {};
    svo::TileShare pendingShare[kHostLanes] = {};      // the frame's part of the image, as it was enqueued
    uint64_t hostFrameNumber = 0;

    void destroy() {
        if (dTables) cudaFree(dTables);
        if (hCounters) cudaFreeHost(hCounters);
        if (dFineTotal) cudaFree(dFineTotal);
        for (int b = 0; b < kRing; ++b) {
            if (dDepth[b]) cudaFree(dDepth[b]);
            if (dTiles[b]) cudaFree(dTiles[b]);
            if (dCounters[b]) cudaFree(dCounters[b]);
            if (dPrefix[b]) cudaFree(dPrefix[b]);
            if (coarseDone[b]) cudaEventDestroy(coarseDone[b]);
            if (fineDone[b]) cudaEventDestroy(fineDone[b]);
            if (laneReady[b]) cudaEventDestroy(laneReady[b]);
            if (prepDone[b]) cudaEventDestroy(prepDone[b]);
            for (int k = 0; k < 4; ++k) if (timing[b][k]) cudaEventDestroy(timing[b][k]);
        }
        for (int b = 0; b < kHostLanes; ++b) {
            if (dRgba[b]) cudaFree(dRgba[b]);
            if (copyDone[b]) cudaEventDestroy(copyDone[b]);
        }
        *this = FramePlan();
    }
};

struct GrowBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t want) {
        if (want <= bytes) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
};

struct svo_tree {
    int device = 0;
    uint32_t *dWords = nullptr;
    uint64_t nWords = 0;
    float center[3] = {0, 0, 0};
    uint32_t depth = 0;
    cudaStream_t stream = nullptr;          // batches; classifier + fine pass of even host-buffer frames
    cudaStream_t stream2 = nullptr;         // ... of host-buffer frames 1 mod 4 (so that consecutive fine passes overlap)
    cudaStream_t stream34[2] = {nullptr, nullptr};      // ... 2 and 3 mod 4
    cudaStream_t coarseStream[2] = {nullptr, nullptr};  // beam passes, high priority, alternating per frame
    cudaStream_t prepStream[2] = {nullptr, nullptr};    // tile classifier + shared-prefix pass, high priority
    cudaStream_t copyStream = nullptr;      // device->host frame copies
    std::mutex mutex;
    std::map<std::tuple<int, int, int>, FramePlan> plans;
    GrowBuffer batchIn, batchOut;
    GrowBuffer orderWorkspace;               // svo_raymarch_batch_device with SVO_BATCH_COHERENCE_ORDER
    GrowBuffer refillCursors;                // SVO_BATCH_LANE_REFILL: ring of eight ray cursors
    unsigned refillCursorNext = 0;
    std::vector<cudaStream_t> l2WindowStreams;   // streams that already carry the access-policy window (experiment)

    svo::TreeDev dev() const { return svo::TreeDev{dWords, nWords, depth}; }
};


namespace svo_detail {

// svo_capi.cu. Callers hold tree->mutex and have made the tree's device current.
int checkDesc(const svo_frame_desc *desc);
int getPlan(svo_tree *tree, int width, int height, int strips, FramePlan **out);
int enqueueFrame(svo_tree *tree, FramePlan *plan, const svo_camera *cam, const svo_frame_desc *desc,
                 const svo::TileShare &share, uint32_t *dRgba, float *userDepth, cudaStream_t stream, bool wantStats,
                 uint32_t *launches, int *slotOut);
void planGeometry(int width, int height, int strips, svo::FramePlanDev &p);
// Replica of a node array on `device`; validate == false skips the host-side walk (the caller has done it once).
int createTreeOnDevice(const uint32_t *words, uint64_t nWords, const float center[3], int device, bool validate, svo_tree **out);

} // namespace svo_detail
