// C ABI implementation (include/svo_b200.h): handles, upload, per-configuration
// frame plans, kernel sequencing, host<->device copies. No CPU fallback: every
// device entry point needs a CUDA device and says so when there is none.
#include "../../include/svo_b200.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <atomic>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>

#include "camera.hpp"
#include "oct_io.hpp"
#include "ply_io.hpp"
#include "svo_build.cuh"
#include "svo_kernels.cuh"
#include "svo_capi_internal.hpp"
#include "svo_voxelize.cuh"

#include <chrono>
#include <thread>

thread_local std::string g_lastError;

namespace {

// Depth = number of descriptor levels, found by following first children from
// the root: the builder puts every leaf at the same depth (reference
// src/VoxelOctree.cpp:139-205 recurses until halfSize == 1).
int measureDepth(const uint32_t *words, uint64_t nWords, uint32_t &depthOut) {
    uint64_t p = 0;
    uint32_t depth = 0;
    for (;;) {
        if (p >= nWords) return fail(SVO_ERR_FORMAT, "node array: child pointer %llu past the end (%llu words)",
                                     (unsigned long long)p, (unsigned long long)nWords);
        uint32_t desc = words[p];
        ++depth;
        if (depth > 23) return fail(SVO_ERR_FORMAT, "node array: deeper than 23 levels");
        if (((desc >> 8) & 0xFFu) == 0) return fail(SVO_ERR_FORMAT, "node array: descriptor %llu has no children", (unsigned long long)p);
        uint64_t offset = desc >> 18;
        if (desc & 0x20000u) {
            if (p + 1 >= nWords) return fail(SVO_ERR_FORMAT, "node array: far word past the end");
            offset = (offset << 32) | words[p + 1];
        }
        if ((desc & 0xFFu) == 0) {
            if (p + offset >= nWords) return fail(SVO_ERR_FORMAT, "node array: leaf pointer past the end");
            break;
        }
        if (offset == 0) return fail(SVO_ERR_FORMAT, "node array: zero child offset at %llu", (unsigned long long)p);
        p += offset;
    }
    depthOut = depth;
    return SVO_OK;
}

// Device, node-array allocation (+1 zeroed padding word: the traversal reads words[p + 1] next to
// every descriptor) and streams; the node array itself is filled by the caller.
int allocTree(uint64_t nWords, const float center[3], int device, std::unique_ptr<svo_tree> &treeOut,
              uint32_t *adoptWords = nullptr) {
    if (nWords < 2) return fail(SVO_ERR_FORMAT, "node array too small (%llu words)", (unsigned long long)nWords);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(SVO_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(SVO_ERR_INVALID_ARGUMENT, "device %d out of range [0, %d)", device, count);

    SVO_DEVICE(device);
    std::unique_ptr<svo_tree> tree(new (std::nothrow) svo_tree);
    if (!tree) return fail(SVO_ERR_OUT_OF_MEMORY, "out of host memory");
    tree->device = device;
    tree->nWords = nWords;
    memcpy(tree->center, center, sizeof(float)*3);

    size_t bytes = size_t(nWords + 1)*sizeof(uint32_t);
    auto cleanup = [&](cudaError_t err, const char *what) {
        if (tree->dWords) cudaFree(tree->dWords);
        if (tree->stream) cudaStreamDestroy(tree->stream);
        if (tree->stream2) cudaStreamDestroy(tree->stream2);
        for (int i = 0; i < 2; ++i) if (tree->stream34[i]) cudaStreamDestroy(tree->stream34[i]);
        for (int i = 0; i < 2; ++i) if (tree->coarseStream[i]) cudaStreamDestroy(tree->coarseStream[i]);
        for (int i = 0; i < 2; ++i) if (tree->prepStream[i]) cudaStreamDestroy(tree->prepStream[i]);
        if (tree->copyStream) cudaStreamDestroy(tree->copyStream);
        return failCuda(err, what);
    };
    if (adoptWords) {
        tree->dWords = adoptWords;   // built in place on this device, padding word included (svo_build.cu)
    } else {
        if ((e = cudaMalloc(&tree->dWords, bytes)) != cudaSuccess) return cleanup(e, "cudaMalloc(node array)");
        if ((e = cudaMemset(tree->dWords + nWords, 0, sizeof(uint32_t))) != cudaSuccess) return cleanup(e, "cudaMemset(padding)");
    }
    int prioLow = 0, prioHigh = 0;
    if ((e = cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh)) != cudaSuccess) return cleanup(e, "cudaDeviceGetStreamPriorityRange");
    if ((e = cudaStreamCreateWithFlags(&tree->stream, cudaStreamNonBlocking)) != cudaSuccess) return cleanup(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&tree->stream2, cudaStreamNonBlocking)) != cudaSuccess) return cleanup(e, "cudaStreamCreate");
    for (int i = 0; i < 2; ++i)
        if ((e = cudaStreamCreateWithFlags(&tree->stream34[i], cudaStreamNonBlocking)) != cudaSuccess) return cleanup(e, "cudaStreamCreate");
    for (int i = 0; i < 2; ++i)
        if ((e = cudaStreamCreateWithPriority(&tree->coarseStream[i], cudaStreamNonBlocking, prioHigh)) != cudaSuccess)
            return cleanup(e, "cudaStreamCreate(coarse)");
    for (int i = 0; i < 2; ++i)
        if ((e = cudaStreamCreateWithPriority(&tree->prepStream[i], cudaStreamNonBlocking, prioHigh)) != cudaSuccess)
            return cleanup(e, "cudaStreamCreate(prep)");
    if ((e = cudaStreamCreateWithFlags(&tree->copyStream, cudaStreamNonBlocking)) != cudaSuccess) return cleanup(e, "cudaStreamCreate(copy)");
    treeOut = std::move(tree);
    return SVO_OK;
}

// Full walk of a node array on the host: every descriptor, far word and leaf word a ray can be led to lies inside
// the array, child blocks lie strictly behind their parent (no cycles), no node is reachable more often than the array
// has words (not a tree), and no branch is deeper than `maxDepth` levels -- the traversal kernels size their stack
// from the depth of the first-child chain (measureDepth) and read words[parent + offset] unchecked, exactly like the
// reference (VoxelOctree.cpp:253-293), so a damaged array would otherwise mean an illegal address on the device.
// Child addressing as in SURVEY.md App. A.1: child bit b of descriptor p sits at p + offset + rank*(1 + far16) with
// rank = popcount(non-leaf mask below b) when it is a node, at p + offset + popcount(valid mask below b) when it is a
// leaf word.
struct WalkItem { uint64_t p; uint32_t depth; };

// One depth-first walk from `stack`'s items; stops expanding once `splitAt` items are pending (the caller hands them
// to threads) when splitAt != 0. Violations go to `err` (first one wins); `visited` is shared by all walkers.
bool walkWords(const uint32_t *words, uint64_t n, uint32_t maxDepth, std::vector<WalkItem> &stack, size_t splitAt,
               std::atomic<uint64_t> &visited, svo_words_report &r, std::string &err) {
    char msg[160];
    auto bad = [&](const char *fmt, unsigned long long a, unsigned long long b) {
        snprintf(msg, sizeof msg, fmt, a, b);
        err = msg;
        return false;
    };
    uint64_t local = 0;
    size_t head = 0;            // splitAt != 0: breadth first (FIFO), so that the pending items are subtrees of similar size
    while (head < stack.size() && (splitAt == 0 || stack.size() - head < splitAt)) {
        WalkItem it;
        if (splitAt) { it = stack[head++]; } else { it = stack.back(); stack.pop_back(); }
        const uint64_t p = it.p;
        if (p >= n) return bad("node array: descriptor index %llu past the end (%llu words)", p, n);
        ++r.descriptors;
        if (++local == 4096) {
            if (visited.fetch_add(local, std::memory_order_relaxed) + local > n) return bad("node array: not a tree (more reachable descriptors than its %llu words)%.0llu", n, 0);
            local = 0;
        }
        if (it.depth > maxDepth) return bad("node array: a branch at word %llu is deeper than %llu levels", p, maxDepth);
        const uint32_t d = words[p];
        const uint32_t valid = (d >> 8) & 0xFFu, nonLeaf = d & 0xFFu;
        if (valid == 0) continue;                                  // nothing below: rays step over it
        uint64_t offset = d >> 18;
        if (d & 0x20000u) {
            if (p + 1 >= n) return bad("node array: far word of descriptor %llu past the end (%llu words)", p, n);
            offset = (offset << 32) | words[p + 1];
            ++r.far_words;
        }
        if (offset == 0) return bad("node array: zero child offset at %llu%.0llu", p, 0);
        const uint64_t base = p + offset, stride = (d & 0x10000u) ? 2 : 1;
        for (uint32_t b = 0; b < 8; ++b) {
            if (!((valid >> b) & 1u)) continue;
            const uint32_t below = (1u << b) - 1u;
            if ((nonLeaf >> b) & 1u) {
                stack.push_back({base + uint64_t(__builtin_popcount(nonLeaf & below))*stride, it.depth + 1});
            } else {
                const uint64_t leaf = base + uint64_t(__builtin_popcount(valid & below));
                if (leaf >= n) return bad("node array: leaf word %llu of descriptor %llu past the end", leaf, p);
                ++r.leaves;
                if (it.depth < r.min_leaf_depth) r.min_leaf_depth = it.depth;
                if (it.depth > r.max_leaf_depth) r.max_leaf_depth = it.depth;
            }
        }
    }
    if (head) stack.erase(stack.begin(), stack.begin() + long(head));
    if (visited.fetch_add(local, std::memory_order_relaxed) + local > n) return bad("node array: not a tree (more reachable descriptors than its %llu words)%.0llu", n, 0);
    return true;
}

int validateWords(const uint32_t *words, uint64_t n, uint32_t maxDepth, svo_words_report *report) {
    std::atomic<uint64_t> visited(0);
    svo_words_report total = {};
    total.min_leaf_depth = ~0u;
    std::string err;
    // the top of the tree on this thread until there is enough pending work to share, then subtrees on all cores
    int threads = int(std::thread::hardware_concurrency());
    threads = threads < 1 ? 1 : threads > 16 ? 16 : threads;
    if (n < (uint64_t(1) << 20)) threads = 1;
    std::vector<WalkItem> top;
    top.push_back({0, 1});
    if (!walkWords(words, n, maxDepth, top, threads > 1 ? size_t(threads)*64 : 0, visited, total, err)) return fail(SVO_ERR_FORMAT, "%s", err.c_str());
    if (!top.empty()) {
        std::vector<svo_words_report> parts((size_t(threads)), svo_words_report{});
        const size_t nThreads = size_t(threads);
        std::vector<std::string> errs(nThreads);
        std::atomic<size_t> next(0);
        std::vector<std::thread> pool;
        auto work = [&](int t) {
            parts[size_t(t)].min_leaf_depth = ~0u;
            for (;;) {
                const size_t k = next.fetch_add(1);
                if (k >= top.size() || !errs[size_t(t)].empty()) return;
                std::vector<WalkItem> mine(1, top[k]);
                if (!walkWords(words, n, maxDepth, mine, 0, visited, parts[size_t(t)], errs[size_t(t)])) return;
            }
        };
        for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto &th : pool) th.join();
        for (int t = 0; t < threads; ++t) {
            if (!errs[size_t(t)].empty()) return fail(SVO_ERR_FORMAT, "%s", errs[size_t(t)].c_str());
            total.descriptors += parts[size_t(t)].descriptors;
            total.leaves += parts[size_t(t)].leaves;
            total.far_words += parts[size_t(t)].far_words;
            if (parts[size_t(t)].min_leaf_depth < total.min_leaf_depth) total.min_leaf_depth = parts[size_t(t)].min_leaf_depth;
            if (parts[size_t(t)].max_leaf_depth > total.max_leaf_depth) total.max_leaf_depth = parts[size_t(t)].max_leaf_depth;
        }
    }
    if (report) *report = total;
    return SVO_OK;
}

// On unless SVO_VALIDATE_TREES=0: a node array from a file or a caller is walked in full before a kernel may follow
// its child pointers (the kernels size their shared-memory stack from the tree's depth and read
// words[parent + offset] unchecked, like the reference).
bool validationRequested() {
    const char *e = getenv("SVO_VALIDATE_TREES");
    return !(e && *e == '0');
}

int createTree(const uint32_t *words, uint64_t nWords, const float center[3], int device, svo_tree **out, bool validate = true) {
    if (!words || !center || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_tree_create: null argument");
    if (nWords < 2) return fail(SVO_ERR_FORMAT, "node array too small (%llu words)", (unsigned long long)nWords);
    *out = nullptr;
    uint32_t depth = 0;
    int st = measureDepth(words, nWords, depth);
    if (st != SVO_OK) return st;
    // The full host-side walk (all cores) runs on this thread while a helper thread allocates and uploads the array;
    // no kernel can see the tree before both are done, and a format error wins over a device error (so a damaged
    // array is reported as such even on a box without a GPU).
    std::unique_ptr<svo_tree> tree;
    int upStatus = SVO_OK;
    std::string upError;
    std::thread uploader([&] {
        upStatus = allocTree(nWords, center, device, tree);
        if (upStatus == SVO_OK) {
            DeviceScope scope(device);
            cudaError_t e = scope.error;
            if (e == cudaSuccess) e = cudaMemcpy(tree->dWords, words, size_t(nWords)*sizeof(uint32_t), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) upStatus = failCuda(e, "cudaMemcpy(node array)");
        }
        if (upStatus != SVO_OK) upError = g_lastError;      // thread-local: carry it over to the caller's thread
    });
    if (validate && validationRequested()) st = validateWords(words, nWords, depth, nullptr);
    uploader.join();
    if (st != SVO_OK || upStatus != SVO_OK) {
        std::string keep = st != SVO_OK ? g_lastError : upError;
        if (tree) svo_tree_destroy(tree.release());
        g_lastError = keep;
        return st != SVO_OK ? st : upStatus;
    }
    tree->depth = depth;
    *out = tree.release();
    return SVO_OK;
}

// The running sums renderBatch / renderTile use for screen coordinates
// (reference src/Main.cpp:97-100, 167-170), tabulated once per configuration.
void planGeometry(int width, int height, int strips, svo::FramePlanDev &p) {
    p.width = width;
    p.height = height;
    p.stripRows = (height - 1)/strips + 1;                       // Main.cpp:351
    p.nStrips = (height + p.stripRows - 1)/p.stripRows;          // strips with y0 < height
    p.tilesX = (width - 1)/8 + 2;                                // Main.cpp:359
    int lastRows = height - (p.nStrips - 1)*p.stripRows;
    p.tilesYLast = (lastRows - 1)/8 + 2;                         // Main.cpp:360
    p.tilesYFull = p.nStrips > 1 ? (p.stripRows - 1)/8 + 2 : p.tilesYLast;
    p.tileRowsFull = p.tilesYFull - 1;
    p.tileRowsLast = p.tilesYLast - 1;
    p.tileCols = p.tilesX - 1;
    p.totalTileRows = (p.nStrips - 1)*p.tileRowsFull + p.tileRowsLast;
    p.totalTiles = p.totalTileRows*p.tileCols;
    p.totalCorners = (p.nStrips - 1)*p.tilesX*p.tilesYFull + p.tilesX*p.tilesYLast;
}

int buildPlan(svo_tree *tree, int width, int height, int strips, FramePlan &plan) {
    svo::FramePlanDev &p = plan.dev;
    planGeometry(width, height, strips, p);

    const float scale = 2.0f/width;                              // Main.cpp:156
    const float tileScale = 8*scale;                             // Main.cpp:157
    const float aspect = height/(float)width;                    // Main.cpp:62

    size_t nDxC = p.tilesX, nDyC = size_t(p.nStrips)*p.tilesYFull, nDxF = width, nDyF = height;
    std::vector<float> tables(nDxC + nDyC + nDxF + nDyF);
    float *dxC = tables.data(), *dyC = dxC + nDxC, *dxF = dyC + nDyC, *dyF = dxF + nDxF;

    float dx = -1.0f + 0*scale;                                  // x0 == 0, Main.cpp:169
    for (int x = 0; x < p.tilesX; ++x, dx += tileScale) dxC[x] = dx;
    for (int s = 0; s < p.nStrips; ++s) {
        int y0 = s*p.stripRows;
        float dy = aspect - y0*scale;                            // Main.cpp:167
        for (int y = 0; y < p.tilesYFull; ++y, dy -= tileScale) dyC[s*p.tilesYFull + y] = dy;
    }
    for (int tx0 = 0; tx0 < width; tx0 += 8) {
        float fx = -1.0f + tx0*scale;                            // Main.cpp:99
        for (int x = tx0; x < tx0 + 8 && x < width; ++x, fx += scale) dxF[x] = fx;
    }
    for (int s = 0; s < p.nStrips; ++s) {
        int y0 = s*p.stripRows, y1 = y0 + p.stripRows < height ? y0 + p.stripRows : height;
        for (int ty0 = y0; ty0 < y1; ty0 += 8) {
            float fy = aspect - ty0*scale;                       // Main.cpp:97
            for (int y = ty0; y < ty0 + 8 && y < y1; ++y, fy -= scale) dyF[y] = fy;
        }
    }

    SVO_CUDA(cudaMalloc(&plan.dTables, tables.size()*sizeof(float)));
    SVO_CUDA(cudaMemcpy(plan.dTables, tables.data(), tables.size()*sizeof(float), cudaMemcpyHostToDevice));
    SVO_CUDA(cudaMallocHost(&plan.hCounters, kRing*sizeof(svo::FrameCounters)));
    SVO_CUDA(cudaMalloc(&plan.dFineTotal, sizeof(unsigned long long)));
    SVO_CUDA(cudaMemset(plan.dFineTotal, 0, sizeof(unsigned long long)));
    for (int b = 0; b < kRing; ++b) {
        SVO_CUDA(cudaMalloc(&plan.dDepth[b], size_t(p.totalCorners)*sizeof(float)));
        SVO_CUDA(cudaMalloc(&plan.dTiles[b], size_t(p.totalTiles > 0 ? p.totalTiles : 1)*sizeof(svo::TileRecord)));
        SVO_CUDA(cudaMalloc(&plan.dCounters[b], sizeof(svo::FrameCounters)));
        if (const int words = svo::prefixRecordWords(tree->dev()))
            SVO_CUDA(cudaMalloc(&plan.dPrefix[b], size_t(p.totalTiles > 0 ? p.totalTiles : 1)*size_t(words)*sizeof(uint32_t)));
        SVO_CUDA(cudaMemset(plan.dCounters[b], 0, sizeof(svo::FrameCounters)));
        SVO_CUDA(cudaEventCreateWithFlags(&plan.coarseDone[b], cudaEventDisableTiming));
        SVO_CUDA(cudaEventCreateWithFlags(&plan.fineDone[b], cudaEventDisableTiming));
        SVO_CUDA(cudaEventCreateWithFlags(&plan.laneReady[b], cudaEventDisableTiming));
        SVO_CUDA(cudaEventCreateWithFlags(&plan.prepDone[b], cudaEventDisableTiming));
        for (int k = 0; k < 4; ++k) SVO_CUDA(cudaEventCreate(&plan.timing[b][k]));
    }
    for (int b = 0; b < kHostLanes; ++b) SVO_CUDA(cudaEventCreateWithFlags(&plan.copyDone[b], cudaEventDisableTiming));
    p.dxCoarse = plan.dTables;
    p.dyCoarse = p.dxCoarse + nDxC;
    p.dxFine = p.dyCoarse + nDyC;
    p.dyFine = p.dxFine + nDxF;
    (void)tree;
    return SVO_OK;
}

int getPlan(svo_tree *tree, int width, int height, int strips, FramePlan **out) {
    auto key = std::make_tuple(width, height, strips);
    auto it = tree->plans.find(key);
    if (it == tree->plans.end()) {
        FramePlan plan;
        int st = buildPlan(tree, width, height, strips, plan);
        if (st != SVO_OK) {
            plan.destroy();
            return st;
        }
        it = tree->plans.emplace(key, plan).first;
    }
    *out = &it->second;
    return SVO_OK;
}

int checkDesc(const svo_frame_desc *desc) {
    if (!desc) return fail(SVO_ERR_INVALID_ARGUMENT, "null frame descriptor");
    if (desc->width < 1 || desc->height < 1 || desc->width > 65536 || desc->height > 65536)
        return fail(SVO_ERR_INVALID_ARGUMENT, "bad frame size %dx%d", desc->width, desc->height);
    if (desc->strips < 1 || desc->strips > desc->height)
        return fail(SVO_ERR_INVALID_ARGUMENT, "strips must be in [1, height] (got %d)", desc->strips);
    if (desc->flavour != SVO_FLAVOUR_VALIDATION && desc->flavour != SVO_FLAVOUR_FAST)
        return fail(SVO_ERR_INVALID_ARGUMENT, "unknown flavour %d", desc->flavour);
    if (desc->tile_world < 1 || desc->tile_rank < 0 || desc->tile_rank >= desc->tile_world)
        return fail(SVO_ERR_INVALID_ARGUMENT, "bad tile interleave rank %d of %d", desc->tile_rank, desc->tile_world);
    if (desc->pixel_stride < 0 || desc->pixel_stride > 8)
        return fail(SVO_ERR_INVALID_ARGUMENT, "pixel_stride must be in [0, 8] (got %d)", desc->pixel_stride);
    if (desc->pixel_format != SVO_PIXELS_RGBA8 && desc->pixel_format != SVO_PIXELS_GREY8A8)
        return fail(SVO_ERR_INVALID_ARGUMENT, "unknown pixel format %d", desc->pixel_format);
    return SVO_OK;
}

svo::FrameConsts toDeviceConsts(const svo_frame_constants &c) {
    svo::FrameConsts f;
    f.posX = c.pos[0]; f.posY = c.pos[1]; f.posZ = c.pos[2];
    f.a11 = c.a11; f.a12 = c.a12; f.a21 = c.a21; f.a22 = c.a22; f.a31 = c.a31; f.a32 = c.a32;
    f.zx = c.zx; f.zy = c.zy; f.zz = c.zz;
    f.lightX = c.light[0]; f.lightY = c.light[1]; f.lightZ = c.light[2];
    f.coarseScale = c.coarse_scale;
    f.beamBias = c.beam_bias;
    return f;
}

// EXPERIMENT SWITCH (off unless SVO_L2_PERSIST_MB is set; measured and left off): an L2
// access-policy window over the start of the node array, persisting hits. The window is limited to
// cudaDevAttrMaxAccessPolicyWindowSize bytes and the array is in depth-first block order, so "the upper
// levels" are not a contiguous range that a window could cover.
int l2PersistMegabytes() {
    static const int mb = [] {
        const char *e = getenv("SVO_L2_PERSIST_MB");
        return e ? atoi(e) : 0;
    }();
    return mb;
}

void applyL2Window(svo_tree *tree, cudaStream_t s) {
    const int mb = l2PersistMegabytes();
    if (mb <= 0) return;
    for (cudaStream_t seen : tree->l2WindowStreams) if (seen == s) return;
    tree->l2WindowStreams.push_back(s);
    int maxWindow = 0, maxPersist = 0;
    cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, tree->device);
    cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, tree->device);
    size_t persist = std::min(size_t(mb) << 20, size_t(maxPersist));
    if (tree->l2WindowStreams.size() == 1) {
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist);
        fprintf(stderr, "[svo] L2 persisting window: %zu MB set aside (max %d MB), window <= %d MB\n", persist >> 20,
                maxPersist >> 20, maxWindow >> 20);
    }
    cudaStreamAttrValue attr = {};
    attr.accessPolicyWindow.base_ptr = tree->dWords;
    attr.accessPolicyWindow.num_bytes = std::min(size_t(tree->nWords)*sizeof(uint32_t), size_t(maxWindow));
    attr.accessPolicyWindow.hitRatio = float(std::min(1.0, double(persist)/double(attr.accessPolicyWindow.num_bytes)));
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
}

// SVO_PREP_ON_LANE=1 (experiment switch): tile classifier and prefix pass on the caller's stream, as before round 2.
bool prepOnCallerStream() {
    static const bool on = [] {
        const char *e = getenv("SVO_PREP_ON_LANE");
        return e && *e == '1';
    }();
    return on;
}

// SVO_NO_PREFIX_RESTART=1 (experiment switch): the FAST fine pass starts every ray at the root, like VALIDATION.
bool prefixRestartEnabled() {
    static const bool on = [] {
        const char *e = getenv("SVO_NO_PREFIX_RESTART");
        return !(e && *e == '1');
    }();
    return on;
}

// Enqueues one frame. The beam pass goes to the tree's high-priority internal stream (unless the
// caller wants the depth buffer in its own memory, which must be ordered on `stream`), the tile
// classifier and the fine pass to `stream`. Returns the double-buffer slot used. Caller holds
// tree->mutex and has made the device current. `share` is this call's part of the frame (desc's tile_rank /
// tile_world plus the stripe width, which the descriptor does not carry).
int enqueueFrame(svo_tree *tree, FramePlan *plan, const svo_camera *cam, const svo_frame_desc *desc,
                 const svo::TileShare &share, uint32_t *dRgba, float *userDepth, cudaStream_t stream, bool wantStats,
                 uint32_t *launches, int *slotOut) {
    svo_frame_constants c;
    svo::frameConstants(*cam, tree->center, desc->width, desc->height, desc->strips, c);
    svo::FrameConsts f = toDeviceConsts(c);
    // a camera with a NaN / infinite entry would make every ray of the frame spin for ever (see raymarchBatchKernel)
    {
        const float *v = &f.posX;
        float sum = 0.0f;
        for (size_t i = 0; i < sizeof(svo::FrameConsts)/sizeof(float); ++i) sum += v[i] < 0 ? -v[i] : v[i];
        if (!(sum < 3.0e38f)) return fail(SVO_ERR_INVALID_ARGUMENT, "camera matrices contain a non-finite value");
    }
    const int b = int(plan->frameNumber++ % kRing);
    float *depth = userDepth ? userDepth : plan->dDepth[b];
    // Three streams per frame. The beam pass runs on a high-priority internal stream (`cs`), frames ahead of the fine
    // passes. The tile classifier and the shared-prefix pass -- small kernels between the beam pass and the fine pass --
    // run on a second high-priority stream (`ps`): on the caller's stream their blocks would queue behind the pending
    // blocks of the previous frames' fine passes and only run in such a pass's tail, so the next fine pass would start
    // two dependent launches late, once per frame (18 % of a frame when eight GPUs share it). With a caller-owned depth
    // buffer everything is ordered on the caller's stream.
    const bool split = !userDepth && !prepOnCallerStream();
    cudaStream_t cs = userDepth ? stream : tree->coarseStream[b & 1];
    cudaStream_t ps = split ? tree->prepStream[b & 1] : stream;
    applyL2Window(tree, stream);
    applyL2Window(tree, cs);
    uint32_t n = 0;

    // the slot's depth buffer, tile list and counters are free once the fine pass of kRing frames ago is done
    // (also with a caller-owned depth buffer: only the depths are the caller's, the counters and the tile list are not)
    if (plan->fineRecorded[b]) SVO_CUDA(cudaStreamWaitEvent(cs, plan->fineDone[b], 0));
    if (wantStats) SVO_CUDA(cudaEventRecord(plan->timing[b][0], cs));
    SVO_CUDA(svo::launchCoarsePass(tree->dev(), plan->dev, f, desc->flavour, depth, plan->dCounters[b], share, cs));
    ++n;
    if (wantStats) SVO_CUDA(cudaEventRecord(plan->timing[b][1], cs));
    if (!userDepth) {
        SVO_CUDA(cudaEventRecord(plan->coarseDone[b], cs));
        SVO_CUDA(cudaStreamWaitEvent(ps, plan->coarseDone[b], 0));
    }
    if (split) {
        // the classifier zero-fills skipped tiles in the framebuffer: not before what the caller ordered on `stream`
        SVO_CUDA(cudaEventRecord(plan->laneReady[b], stream));
        SVO_CUDA(cudaStreamWaitEvent(ps, plan->laneReady[b], 0));
    }
    if (wantStats && !split) SVO_CUDA(cudaEventRecord(plan->timing[b][2], stream));
    SVO_CUDA(svo::launchClassifyTiles(plan->dev, f, depth, dRgba, share, desc->pixel_stride,
                                      plan->dTiles[b], plan->dCounters[b], plan->dFineTotal, ps));
    ++n;
    uint32_t *prefix = prefixRestartEnabled() && svo::finePassUsesPrefix(tree->dev(), desc->flavour, desc->pixel_stride) ? plan->dPrefix[b] : nullptr;
    if (prefix) {
        SVO_CUDA(svo::launchTilePrefix(tree->dev(), plan->dev, f, desc->flavour, plan->dTiles[b], plan->dCounters[b],
                                       share, desc->pixel_stride, prefix, ps));
        ++n;
    }
    if (split) {
        SVO_CUDA(cudaEventRecord(plan->prepDone[b], ps));
        SVO_CUDA(cudaStreamWaitEvent(stream, plan->prepDone[b], 0));
        if (wantStats) SVO_CUDA(cudaEventRecord(plan->timing[b][2], stream));
    }
    SVO_CUDA(svo::launchFinePass(tree->dev(), plan->dev, f, desc->flavour, plan->dTiles[b], plan->dCounters[b], dRgba,
                                 share, desc->pixel_stride, prefix, stream));
    ++n;
    if (wantStats) {
        SVO_CUDA(cudaEventRecord(plan->timing[b][3], stream));
        SVO_CUDA(cudaMemcpyAsync(plan->hCounters + b, plan->dCounters[b], sizeof(svo::FrameCounters),
                                 cudaMemcpyDeviceToHost, stream));
    }
    SVO_CUDA(cudaEventRecord(plan->fineDone[b], stream));
    plan->fineRecorded[b] = true;
    if (launches) *launches = n;
    if (slotOut) *slotOut = b;
    return SVO_OK;
}

// Valid once the frame's stream work has completed (caller synchronised).
void fillStats(const FramePlan *plan, int slot, const svo::TileShare &share, uint32_t launches, svo_frame_stats *stats) {
    const svo::FramePlanDev &p = plan->dev;
    const int world = share.world, rank = share.rank, run = share.run;
    // corner columns this rank traces: those on either side of its tile-column runs
    int cols = 0;
    for (int cx = 0; cx < p.tilesX; ++cx) {
        bool right = cx < p.tileCols && (cx/run) % world == rank;        // tile column to the right of the corner
        bool left = cx > 0 && ((cx - 1)/run) % world == rank;            // ... to the left
        if (world == 1 || right || left) ++cols;
    }
    const int cornerRows = (p.nStrips - 1)*p.tilesYFull + p.tilesYLast;
    stats->coarse_rays = uint64_t(cols)*uint64_t(cornerRows);
    stats->fine_rays = plan->hCounters[slot].fineRays;
    stats->tiles_rendered = plan->hCounters[slot].tilesRendered;
    stats->tiles_total = uint64_t(svo::ownedTileColumns(p.tileCols, share))*uint64_t(p.totalTileRows);
    stats->kernel_launches = launches;
    stats->reserved = 0;
    stats->coarse_ms = stats->fine_ms = 0.0f;
    cudaEventElapsedTime(&stats->coarse_ms, plan->timing[slot][0], plan->timing[slot][1]);
    cudaEventElapsedTime(&stats->fine_ms, plan->timing[slot][2], plan->timing[slot][3]);
}

} // namespace

namespace svo_detail {   // what svo_multi.cu uses of the above
int checkDesc(const svo_frame_desc *desc) { return ::checkDesc(desc); }
int getPlan(svo_tree *tree, int width, int height, int strips, FramePlan **out) { return ::getPlan(tree, width, height, strips, out); }
int enqueueFrame(svo_tree *tree, FramePlan *plan, const svo_camera *cam, const svo_frame_desc *desc,
                 const svo::TileShare &share, uint32_t *dRgba, float *userDepth, cudaStream_t stream, bool wantStats,
                 uint32_t *launches, int *slotOut) {
    return ::enqueueFrame(tree, plan, cam, desc, share, dRgba, userDepth, stream, wantStats, launches, slotOut);
}
void planGeometry(int width, int height, int strips, svo::FramePlanDev &p) { ::planGeometry(width, height, strips, p); }
int createTreeOnDevice(const uint32_t *words, uint64_t nWords, const float center[3], int device, bool validate, svo_tree **out) {
    return ::createTree(words, nWords, center, device, out, validate);
}
} // namespace svo_detail

extern "C" {

int svo_abi_version(void) { return SVO_ABI_VERSION; }

const char *svo_last_error(void) { return g_lastError.c_str(); }

int svo_device_count(int *count) {
    if (!count) return fail(SVO_ERR_INVALID_ARGUMENT, "null count");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(SVO_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return SVO_OK;
}

void svo_free(void *p) { free(p); }

void svo_pixels_expand_grey8a(const uint16_t *src, uint64_t n, uint32_t *dst) {
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t p = src[i], g = p & 0xFFu;
        dst[i] = ((p & 0xFF00u) << 16) | (g << 16) | (g << 8) | g;
    }
}

int svo_host_alloc(size_t bytes, void **out) {
    if (!out) return fail(SVO_ERR_INVALID_ARGUMENT, "null out");
    *out = nullptr;
    SVO_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return SVO_OK;
}

int svo_host_free(void *p) {
    if (p) SVO_CUDA(cudaFreeHost(p));
    return SVO_OK;
}

int svo_host_register(int device, void *p, size_t bytes, void **device_ptr) {
    if (!p || !bytes || !device_ptr) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_host_register: null argument");
    *device_ptr = nullptr;
    SVO_DEVICE(device);
    SVO_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    cudaError_t e = cudaHostGetDevicePointer(device_ptr, p, 0);
    if (e != cudaSuccess) {
        cudaHostUnregister(p);
        return failCuda(e, "cudaHostGetDevicePointer");
    }
    return SVO_OK;
}

int svo_host_unregister(void *p) {
    if (p) SVO_CUDA(cudaHostUnregister(p));
    return SVO_OK;
}

/* ---- .oct ------------------------------------------------------------------ */

int svo_oct_read(const char *path, uint32_t **words, uint64_t *n_words, float center[3]) {
    if (!path || !words || !n_words || !center) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_oct_read: null argument");
    *words = nullptr;
    *n_words = 0;
    svo::OctFile f;
    std::string err;
    int status = 0;
    if (!svo::readOctFile(path, f, err, status)) return fail(status, "%s", err.c_str());
    *words = f.words;
    *n_words = f.nWords;
    memcpy(center, f.center, sizeof(float)*3);
    return SVO_OK;
}

int svo_oct_write(const char *path, const uint32_t *words, uint64_t n_words, const float center[3], int compress) {
    if (!path || (!words && n_words) || !center) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_oct_write: null argument");
    std::string err;
    int status = 0;
    if (!svo::writeOctFile(path, words, n_words, center, compress != 0, err, status)) return fail(status, "%s", err.c_str());
    return SVO_OK;
}

/* ---- trees ----------------------------------------------------------------- */

int svo_words_validate(const uint32_t *words, uint64_t n_words, svo_words_report *report) {
    if (!words) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_words_validate: null argument");
    if (n_words < 2) return fail(SVO_ERR_FORMAT, "node array too small (%llu words)", (unsigned long long)n_words);
    uint32_t depth = 0;
    int st = measureDepth(words, n_words, depth);
    if (st != SVO_OK) return st;
    if ((st = validateWords(words, n_words, depth, report)) != SVO_OK) return st;
    if (report) report->depth = depth;
    return SVO_OK;
}

int svo_tree_create_from_words(const uint32_t *words, uint64_t n_words, const float center[3], int device, svo_tree **out) {
    return createTree(words, n_words, center, device, out);
}

// Decode and upload are pipelined: the reader's worker threads decode the
// 64 MiB LZ4 slices while this thread copies every finished slice to the device, in file order.
int svo_tree_load_oct(const char *path, int device, svo_tree **out) {
    if (!path || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_tree_load_oct: null argument");
    *out = nullptr;
    svo::OctReader reader;
    std::string err;
    int status = 0;
    if (!reader.open(path, err, status)) return fail(status, "%s", err.c_str());
    std::unique_ptr<svo_tree> tree;
    int st = allocTree(reader.nWords, reader.center, device, tree);
    if (st != SVO_OK) return st;
    SVO_DEVICE(device);
    uint32_t *words = svo::allocNodeArray(reader.nWords);
    if (!words) {
        svo_tree_destroy(tree.release());
        return fail(SVO_ERR_OUT_OF_MEMORY, "out of host memory for the node array");
    }
    cudaError_t copyErr = cudaSuccess;
    auto upload = [&](uint64_t firstWord, uint64_t wordCount) {
        copyErr = cudaMemcpyAsync(tree->dWords + firstWord, words + firstWord, size_t(wordCount)*sizeof(uint32_t),
                                  cudaMemcpyHostToDevice, tree->copyStream);
        return copyErr == cudaSuccess;
    };
    bool ok = reader.decode(words, upload, err, status);
    if (copyErr == cudaSuccess) copyErr = cudaStreamSynchronize(tree->copyStream);
    uint32_t depth = 0;
    if (ok && copyErr == cudaSuccess) st = measureDepth(words, reader.nWords, depth);
    if (ok && copyErr == cudaSuccess && st == SVO_OK && validationRequested()) st = validateWords(words, reader.nWords, depth, nullptr);
    free(words);
    if (copyErr != cudaSuccess) {
        svo_tree_destroy(tree.release());
        return failCuda(copyErr, "cudaMemcpyAsync(node array slice)");
    }
    if (!ok) {
        svo_tree_destroy(tree.release());
        return fail(status, "%s", err.c_str());
    }
    if (st != SVO_OK) {
        svo_tree_destroy(tree.release());
        return st;
    }
    tree->depth = depth;
    *out = tree.release();
    return SVO_OK;
}

/* ---- construction (row f2) ---------------------------------------------------- */

namespace {

thread_local svo_build_stats g_buildStats = {};

int requireDevice(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(SVO_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(SVO_ERR_INVALID_ARGUMENT, "device %d out of range [0, %d)", device, count);
    return SVO_OK;
}

int finishBuild(svo::OctreeBuilder &builder, int device, svo_tree **out) {
    svo::BuildResult r;
    std::string err;
    if (!builder.finish(r, err)) return fail(SVO_ERR_FORMAT, "octree construction: %s", err.c_str());
    std::unique_ptr<svo_tree> tree;
    int st = allocTree(r.nWords, r.center, device, tree, r.dWords);
    if (st != SVO_OK) {
        cudaFree(r.dWords);
        return st;
    }
    tree->depth = r.depth;
    g_buildStats.voxels = r.stats.voxels;
    g_buildStats.nodes = r.stats.nodes;
    g_buildStats.far_blocks = r.stats.farBlocks;
    g_buildStats.words = r.nWords;
    g_buildStats.gather_ms = r.stats.gatherMs;
    g_buildStats.sort_ms = r.stats.sortMs;
    g_buildStats.levels_ms = r.stats.levelsMs;
    g_buildStats.emit_ms = r.stats.emitMs;
    *out = tree.release();
    return SVO_OK;
}

constexpr uint64_t kBuildChunkVoxels = 16ull << 20;   // 64 MiB of voxels per upload

} // namespace

int svo_tree_build_from_voxels(const uint32_t *voxels, int w, int h, int d, int device, svo_tree **out) {
    if (!voxels || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_tree_build_from_voxels: null argument");
    *out = nullptr;
    int st = requireDevice(device);
    if (st != SVO_OK) return st;
    SVO_DEVICE(device);
    svo::OctreeBuilder builder;
    std::string err;
    if (!builder.begin(w, h, d, err)) return fail(SVO_ERR_INVALID_ARGUMENT, "octree construction: %s", err.c_str());
    const uint64_t total = uint64_t(w)*uint64_t(h)*uint64_t(d);
    uint32_t *dChunk = nullptr;
    SVO_CUDA(cudaMalloc(&dChunk, size_t(std::min(total, kBuildChunkVoxels))*sizeof(uint32_t)));
    for (uint64_t first = 0; first < total; first += kBuildChunkVoxels) {
        const uint64_t count = std::min(kBuildChunkVoxels, total - first);
        cudaError_t e = cudaMemcpy(dChunk, voxels + first, size_t(count)*sizeof(uint32_t), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(dChunk); return failCuda(e, "cudaMemcpy(voxel chunk)"); }
        if (!builder.addDenseChunk(dChunk, first, count, err)) { cudaFree(dChunk); return fail(SVO_ERR_CUDA, "octree construction: %s", err.c_str()); }
    }
    cudaFree(dChunk);
    return finishBuild(builder, device, out);
}

// Raw .voxel file: int32 w, h, d, then w*h*d uint32 voxels, x fastest (reference src/VoxelData.cpp:36-48,
// 183-201). Streamed through two pinned buffers: the read of chunk k+1 overlaps the upload and gather of chunk k.
int svo_tree_build_from_voxel_file(const char *path, int device, svo_tree **out) {
    if (!path || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_tree_build_from_voxel_file: null argument");
    *out = nullptr;
    int st = requireDevice(device);
    if (st != SVO_OK) return st;
    FILE *fp = fopen(path, "rb");
    if (!fp) return fail(SVO_ERR_IO, "cannot open %s", path);
    std::unique_ptr<FILE, int (*)(FILE *)> closer(fp, fclose);
    int32_t dims[3] = {0, 0, 0};
    if (fread(dims, sizeof(int32_t), 3, fp) != 3) return fail(SVO_ERR_FORMAT, "short read in the header of %s", path);
    SVO_DEVICE(device);
    svo::OctreeBuilder builder;
    std::string err;
    if (!builder.begin(dims[0], dims[1], dims[2], err)) return fail(SVO_ERR_FORMAT, "%s: %s", path, err.c_str());
    const uint64_t total = uint64_t(dims[0])*uint64_t(dims[1])*uint64_t(dims[2]);
    const uint64_t chunk = std::min(total, kBuildChunkVoxels);
    uint32_t *hBuf[2] = {nullptr, nullptr}, *dBuf[2] = {nullptr, nullptr};
    cudaEvent_t uploaded[2] = {nullptr, nullptr};
    auto release = [&]() {
        for (int i = 0; i < 2; ++i) {
            if (hBuf[i]) cudaFreeHost(hBuf[i]);
            if (dBuf[i]) cudaFree(dBuf[i]);
            if (uploaded[i]) cudaEventDestroy(uploaded[i]);
        }
    };
    for (int i = 0; i < 2; ++i) {
        cudaError_t e = cudaMallocHost(&hBuf[i], size_t(chunk)*sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&dBuf[i], size_t(chunk)*sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&uploaded[i], cudaEventDisableTiming);
        if (e != cudaSuccess) { release(); return failCuda(e, "voxel staging buffers"); }
    }
    int k = 0;
    for (uint64_t first = 0; first < total; first += chunk, ++k) {
        const int b = k & 1;
        const uint64_t count = std::min(chunk, total - first);
        cudaEventSynchronize(uploaded[b]);     // the copy out of this pinned buffer two chunks ago
        if (fread(hBuf[b], sizeof(uint32_t), size_t(count), fp) != size_t(count)) {
            release();
            return fail(SVO_ERR_FORMAT, "%s: short read (voxel %llu of %llu)", path, (unsigned long long)first, (unsigned long long)total);
        }
        // default stream: ordered after the previous chunk's gather, overlapping the next fread
        cudaError_t e = cudaMemcpyAsync(dBuf[b], hBuf[b], size_t(count)*sizeof(uint32_t), cudaMemcpyHostToDevice, 0);
        if (e == cudaSuccess) e = cudaEventRecord(uploaded[b], 0);
        if (e != cudaSuccess) { release(); return failCuda(e, "cudaMemcpyAsync(voxel chunk)"); }
        if (!builder.addDenseChunk(dBuf[b], first, count, err)) { release(); return fail(SVO_ERR_CUDA, "octree construction: %s", err.c_str()); }
    }
    release();
    return finishBuild(builder, device, out);
}

int svo_tree_build_from_sparse(const uint32_t *xyz, const uint32_t *values, uint64_t n, int w, int h, int d,
                               int device, svo_tree **out) {
    if ((n && (!xyz || !values)) || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_tree_build_from_sparse: null argument");
    *out = nullptr;
    int st = requireDevice(device);
    if (st != SVO_OK) return st;
    SVO_DEVICE(device);
    svo::OctreeBuilder builder;
    std::string err;
    if (!builder.begin(w, h, d, err)) return fail(SVO_ERR_INVALID_ARGUMENT, "octree construction: %s", err.c_str());
    const uint64_t chunk = std::min<uint64_t>(n ? n : 1, kBuildChunkVoxels);
    uint32_t *dXyz = nullptr, *dVal = nullptr;
    SVO_CUDA(cudaMalloc(&dXyz, size_t(chunk)*3*sizeof(uint32_t)));
    cudaError_t e = cudaMalloc(&dVal, size_t(chunk)*sizeof(uint32_t));
    if (e != cudaSuccess) { cudaFree(dXyz); return failCuda(e, "cudaMalloc(sparse chunk)"); }
    for (uint64_t first = 0; first < n; first += chunk) {
        const uint64_t count = std::min(chunk, n - first);
        e = cudaMemcpy(dXyz, xyz + 3*first, size_t(count)*3*sizeof(uint32_t), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dVal, values + first, size_t(count)*sizeof(uint32_t), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(dXyz); cudaFree(dVal); return failCuda(e, "cudaMemcpy(sparse chunk)"); }
        if (!builder.addSparse(dXyz, dVal, count, err)) { cudaFree(dXyz); cudaFree(dVal); return fail(SVO_ERR_CUDA, "octree construction: %s", err.c_str()); }
    }
    cudaFree(dXyz);
    cudaFree(dVal);
    return finishBuild(builder, device, out);
}

namespace {
thread_local svo_voxelize_stats g_voxelizeStats = {};
}

int svo_ply_read_triangles(const char *path, float **tris, uint64_t *n, float lower[3], float upper[3]) {
    if (!path || !tris || !n || !lower || !upper) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_ply_read_triangles: null argument");
    *tris = nullptr;
    *n = 0;
    svo::Mesh mesh;
    std::string err;
    int status = 0;
    if (!svo::readPlyMesh(path, mesh, err, status)) return fail(status, "%s", err.c_str());
    static_assert(sizeof(svo::MeshTriangle) == 33*sizeof(float), "MeshTriangle is 33 packed floats");
    float *out = static_cast<float *>(malloc(mesh.triangleCount()*sizeof(svo::MeshTriangle)));
    if (!out) return fail(SVO_ERR_OUT_OF_MEMORY, "out of host memory for the triangle list");
    svo::assembleTriangles(mesh, reinterpret_cast<svo::MeshTriangle *>(out));
    *tris = out;
    *n = mesh.triangleCount();
    memcpy(lower, mesh.lower, 12);
    memcpy(upper, mesh.upper, 12);
    return SVO_OK;
}

int svo_tree_build_from_ply(const char *path, int resolution, uint64_t mem_budget, int threads, int device, svo_tree **out) {
    if (!path || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_tree_build_from_ply: null argument");
    *out = nullptr;
    int st = requireDevice(device);
    if (st != SVO_OK) return st;
    svo::Mesh mesh;
    std::string err;
    int status = 0;
    const bool debugTiming = getenv("SVO_BUILD_DEBUG") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); };
    auto t0 = now();
    if (!svo::readPlyMesh(path, mesh, err, status)) return fail(status, "%s", err.c_str());
    if (debugTiming) fprintf(stderr, "[svo] PLY read: %.3f s (%zu vertices, %zu triangles)\n", since(t0), mesh.verts.size(), mesh.triangleCount());
    if (mem_budget == 0) mem_budget = uint64_t(1024)*1024*1024;                   // Main.cpp:269
    if (threads <= 0) threads = int(std::thread::hardware_concurrency());
    if (threads <= 0) threads = 1;
    SVO_DEVICE(device);
    svo::OctreeBuilder builder;
    svo::VoxelizeStats vs;
    t0 = now();
    if (!svo::voxelizeMesh(mesh, resolution, mem_budget, threads, builder, vs, err))
        return fail(SVO_ERR_INVALID_ARGUMENT, "voxelisation: %s", err.c_str());
    if (debugTiming) fprintf(stderr, "[svo] upload + voxelise (wall): %.3f s\n", since(t0));
    t0 = now();
    st = finishBuild(builder, device, out);
    if (debugTiming) fprintf(stderr, "[svo] octree build (wall): %.3f s\n", since(t0));
    if (st != SVO_OK) return st;
    g_voxelizeStats.triangles = vs.triangles;
    g_voxelizeStats.cell_records = vs.cellRecords;
    g_voxelizeStats.voxels = vs.voxels;
    for (int i = 0; i < 3; ++i) { g_voxelizeStats.dims[i] = vs.dims[i]; g_voxelizeStats.sub_block[i] = vs.subBlock[i]; }
    g_voxelizeStats.cache_block = vs.cacheBlock;
    g_voxelizeStats.large_triangles = int32_t(vs.largeTriangles);
    g_voxelizeStats.overlap_ms = vs.overlapMs;
    g_voxelizeStats.sort_ms = vs.sortMs;
    g_voxelizeStats.fold_ms = vs.foldMs;
    return SVO_OK;
}

int svo_voxelize_last_stats(svo_voxelize_stats *out) {
    if (!out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_voxelize_last_stats: null argument");
    *out = g_voxelizeStats;
    return SVO_OK;
}

// Two calls: with xyz_out == values_out == NULL only *n_out is written (the count); with buffers of
// `capacity` voxels the list is copied out. The extraction runs on the device either way.
int svo_tree_extract_voxels(const svo_tree *tree, uint32_t *xyz_out, uint32_t *values_out, uint64_t capacity,
                            uint64_t *n_out) {
    if (!tree || !n_out || ((xyz_out == nullptr) != (values_out == nullptr)))
        return fail(SVO_ERR_INVALID_ARGUMENT, "svo_tree_extract_voxels: null argument");
    SVO_DEVICE(tree->device);
    uint32_t *dXyz = nullptr, *dVal = nullptr;
    uint64_t n = 0;
    std::string err;
    if (!svo::extractVoxels(tree->dWords, tree->depth, &dXyz, &dVal, &n, err)) return fail(SVO_ERR_CUDA, "voxel extraction: %s", err.c_str());
    *n_out = n;
    int st = SVO_OK;
    if (xyz_out) {
        if (capacity < n) {
            st = fail(SVO_ERR_INVALID_ARGUMENT, "svo_tree_extract_voxels: capacity %llu < %llu voxels", (unsigned long long)capacity, (unsigned long long)n);
        } else {
            cudaError_t e = cudaMemcpy(xyz_out, dXyz, size_t(n)*3*sizeof(uint32_t), cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaMemcpy(values_out, dVal, size_t(n)*sizeof(uint32_t), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) st = failCuda(e, "cudaMemcpy(voxel list)");
        }
    }
    cudaFree(dXyz);
    cudaFree(dVal);
    return st;
}

// build(extract(tree)) entirely in HBM: the canonical re-layout of a tree. For a tree the reference's
// builder (or this one) produced the result is the same array -- the round-trip property the tests use
// at sizes no CPU oracle reaches.
int svo_tree_rebuild(const svo_tree *tree, int w, int h, int d, svo_tree **out) {
    if (!tree || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_tree_rebuild: null argument");
    *out = nullptr;
    SVO_DEVICE(tree->device);
    uint32_t *dXyz = nullptr, *dVal = nullptr;
    uint64_t n = 0;
    std::string err;
    if (!svo::extractVoxels(tree->dWords, tree->depth, &dXyz, &dVal, &n, err)) return fail(SVO_ERR_CUDA, "voxel extraction: %s", err.c_str());
    svo::OctreeBuilder builder;
    bool ok = builder.begin(w, h, d, err) && builder.addSparse(dXyz, dVal, n, err);
    cudaFree(dXyz);
    cudaFree(dVal);
    if (!ok) return fail(SVO_ERR_INVALID_ARGUMENT, "octree construction: %s", err.c_str());
    return finishBuild(builder, tree->device, out);
}

int svo_build_last_stats(svo_build_stats *out) {
    if (!out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_build_last_stats: null argument");
    *out = g_buildStats;
    return SVO_OK;
}

int svo_tree_download_words(const svo_tree *tree, uint32_t *words_out, uint64_t n_words) {
    if (!tree || !words_out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_tree_download_words: null argument");
    if (n_words != tree->nWords) return fail(SVO_ERR_INVALID_ARGUMENT, "word count mismatch (tree has %llu)", (unsigned long long)tree->nWords);
    SVO_DEVICE(tree->device);
    SVO_CUDA(cudaMemcpy(words_out, tree->dWords, size_t(n_words)*sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return SVO_OK;
}

int svo_tree_save_oct(const svo_tree *tree, const char *path, int compress) {
    if (!tree || !path) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_tree_save_oct: null argument");
    std::unique_ptr<uint32_t, void (*)(void *)> words(static_cast<uint32_t *>(malloc(size_t(tree->nWords)*4)), free);
    if (!words) return fail(SVO_ERR_OUT_OF_MEMORY, "out of host memory");
    int st = svo_tree_download_words(tree, words.get(), tree->nWords);
    if (st != SVO_OK) return st;
    return svo_oct_write(path, words.get(), tree->nWords, tree->center, compress);
}

int svo_tree_get_info(const svo_tree *tree, svo_tree_info *out) {
    if (!tree || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_tree_get_info: null argument");
    out->n_words = tree->nWords;
    memcpy(out->center, tree->center, sizeof(float)*3);
    out->depth = tree->depth;
    out->device = tree->device;
    out->device_bytes = (tree->nWords + 1)*sizeof(uint32_t);
    return SVO_OK;
}

int svo_tree_destroy(svo_tree *tree) {
    if (!tree) return SVO_OK;
    {
        SVO_DEVICE(tree->device);
        cudaDeviceSynchronize();
        for (auto &kv : tree->plans) kv.second.destroy();
        tree->batchIn.release();
        tree->orderWorkspace.release();
        tree->refillCursors.release();
        tree->batchOut.release();
        if (tree->dWords) cudaFree(tree->dWords);
        if (tree->stream) cudaStreamDestroy(tree->stream);
        if (tree->stream2) cudaStreamDestroy(tree->stream2);
        for (int i = 0; i < 2; ++i) if (tree->stream34[i]) cudaStreamDestroy(tree->stream34[i]);
        for (int i = 0; i < 2; ++i) if (tree->coarseStream[i]) cudaStreamDestroy(tree->coarseStream[i]);
        for (int i = 0; i < 2; ++i) if (tree->prepStream[i]) cudaStreamDestroy(tree->prepStream[i]);
        if (tree->copyStream) cudaStreamDestroy(tree->copyStream);
    }
    delete tree;
    return SVO_OK;
}

/* ---- traversal -------------------------------------------------------------- */

int svo_raymarch_batch_device(svo_tree *tree, uint64_t n, const float *d_o, const float *d_d, float ray_scale,
                              int flavour, uint8_t *d_hit, float *d_t, uint32_t *d_normal, uint64_t *d_voxel,
                              void *stream) {
    if (!tree || (n && (!d_o || !d_d))) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_raymarch_batch_device: null argument");
    const bool reorder = (flavour & SVO_BATCH_COHERENCE_ORDER) != 0;
    const bool refill = (flavour & SVO_BATCH_LANE_REFILL) != 0;
    flavour &= ~(SVO_BATCH_COHERENCE_ORDER | SVO_BATCH_LANE_REFILL);
    if (flavour != SVO_FLAVOUR_VALIDATION && flavour != SVO_FLAVOUR_FAST) return fail(SVO_ERR_INVALID_ARGUMENT, "unknown flavour %d", flavour);
    SVO_DEVICE(tree->device);
    const uint32_t *order = nullptr;
    unsigned long long *cursor = nullptr;
    if ((reorder && n >= 4096) || refill) {
        // one workspace per tree: calls with these flags on one tree are ordered by the caller's stream (the refill
        // cursors rotate through a ring of eight, so up to eight such batches may be in flight)
        std::lock_guard<std::mutex> lock(tree->mutex);
        if (reorder && n >= 4096) {
            SVO_CUDA(tree->orderWorkspace.reserve(svo::coherenceOrderBytes(n)));
            SVO_CUDA(svo::buildCoherenceOrder(n, d_d, tree->orderWorkspace.ptr, &order, static_cast<cudaStream_t>(stream)));
        }
        if (refill) {
            SVO_CUDA(tree->refillCursors.reserve(8*sizeof(unsigned long long)));
            cursor = static_cast<unsigned long long *>(tree->refillCursors.ptr) + (tree->refillCursorNext++ & 7u);
        }
    }
    SVO_CUDA(svo::launchRaymarchBatch(tree->dev(), n, d_o, d_d, ray_scale, flavour, d_hit, d_t, d_normal, d_voxel, order,
                                      cursor, static_cast<cudaStream_t>(stream)));
    return SVO_OK;
}

int svo_raymarch_batch(svo_tree *tree, uint64_t n, const float *o, const float *d, float ray_scale, int flavour,
                       uint8_t *hit, float *t, uint32_t *normal, uint64_t *voxel) {
    if (!tree || (n && (!o || !d))) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_raymarch_batch: null argument");
    const bool reorder = (flavour & SVO_BATCH_COHERENCE_ORDER) != 0;
    const bool refill = (flavour & SVO_BATCH_LANE_REFILL) != 0;
    flavour &= ~(SVO_BATCH_COHERENCE_ORDER | SVO_BATCH_LANE_REFILL);
    if (flavour != SVO_FLAVOUR_VALIDATION && flavour != SVO_FLAVOUR_FAST) return fail(SVO_ERR_INVALID_ARGUMENT, "unknown flavour %d", flavour);
    if (n == 0) return SVO_OK;
    SVO_DEVICE(tree->device);
    std::lock_guard<std::mutex> lock(tree->mutex);

    // Chunks alternate between two streams, each with its own staging area, so that the upload of
    // chunk k+1 and the download of chunk k-1 overlap the traversal of chunk k (page-locked host
    // buffers make the copies truly asynchronous; pageable ones still work).
    const uint64_t kChunk = 1ull << 21;
    const uint64_t chunk = n < kChunk ? n : kChunk;
    auto align16 = [](size_t v) { return (v + 15) & ~size_t(15); };
    const size_t inBytes = align16(size_t(chunk)*3*sizeof(float));
    const size_t offVoxel = 0, offT = align16(offVoxel + size_t(chunk)*8), offNormal = align16(offT + size_t(chunk)*4),
                 offHit = align16(offNormal + size_t(chunk)*4), outBytes = align16(offHit + size_t(chunk));
    SVO_CUDA(tree->batchIn.reserve(4*inBytes));
    SVO_CUDA(tree->batchOut.reserve(2*outBytes));
    const size_t orderBytes = (reorder && chunk >= 4096) ? ((svo::coherenceOrderBytes(chunk) + 255) & ~size_t(255)) : 0;
    if (orderBytes) SVO_CUDA(tree->orderWorkspace.reserve(2*orderBytes));
    if (refill) SVO_CUDA(tree->refillCursors.reserve(8*sizeof(unsigned long long)));

    cudaStream_t streams[2] = {tree->stream, tree->stream2};
    for (uint64_t begin = 0, k = 0; begin < n; begin += chunk, ++k) {
        const uint64_t m = n - begin < chunk ? n - begin : chunk;
        const int slot = int(k & 1);
        cudaStream_t s = streams[slot];
        float *dO = reinterpret_cast<float *>(static_cast<unsigned char *>(tree->batchIn.ptr) + size_t(slot)*2*inBytes);
        float *dD = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(dO) + inBytes);
        unsigned char *base = static_cast<unsigned char *>(tree->batchOut.ptr) + size_t(slot)*outBytes;
        uint64_t *dVoxel = voxel ? reinterpret_cast<uint64_t *>(base + offVoxel) : nullptr;
        float *dT = t ? reinterpret_cast<float *>(base + offT) : nullptr;
        uint32_t *dNormal = normal ? reinterpret_cast<uint32_t *>(base + offNormal) : nullptr;
        uint8_t *dHit = hit ? base + offHit : nullptr;
        SVO_CUDA(cudaMemcpyAsync(dO, o + 3*begin, size_t(m)*3*sizeof(float), cudaMemcpyHostToDevice, s));
        SVO_CUDA(cudaMemcpyAsync(dD, d + 3*begin, size_t(m)*3*sizeof(float), cudaMemcpyHostToDevice, s));
        const uint32_t *order = nullptr;
        if (orderBytes && m >= 4096)
            SVO_CUDA(svo::buildCoherenceOrder(m, dD, static_cast<unsigned char *>(tree->orderWorkspace.ptr) + size_t(slot)*orderBytes, &order, s));
        unsigned long long *cursor = refill ? static_cast<unsigned long long *>(tree->refillCursors.ptr) + slot : nullptr;
        SVO_CUDA(svo::launchRaymarchBatch(tree->dev(), m, dO, dD, ray_scale, flavour, dHit, dT, dNormal, dVoxel, order, cursor, s));
        if (hit) SVO_CUDA(cudaMemcpyAsync(hit + begin, dHit, size_t(m), cudaMemcpyDeviceToHost, s));
        if (t) SVO_CUDA(cudaMemcpyAsync(t + begin, dT, size_t(m)*4, cudaMemcpyDeviceToHost, s));
        if (normal) SVO_CUDA(cudaMemcpyAsync(normal + begin, dNormal, size_t(m)*4, cudaMemcpyDeviceToHost, s));
        if (voxel) SVO_CUDA(cudaMemcpyAsync(voxel + begin, dVoxel, size_t(m)*8, cudaMemcpyDeviceToHost, s));
    }
    SVO_CUDA(cudaStreamSynchronize(streams[0]));
    SVO_CUDA(cudaStreamSynchronize(streams[1]));
    return SVO_OK;
}

int svo_raymarch(svo_tree *tree, const float o[3], const float d[3], float ray_scale, uint32_t *normal, float *t, int *hit_out) {
    if (!tree || !o || !d || !normal || !t || !hit_out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_raymarch: null argument");
    uint8_t code = 0;
    float tt = 0.0f;
    uint32_t nn = 0;
    int st = svo_raymarch_batch(tree, 1, o, d, ray_scale, SVO_FLAVOUR_VALIDATION, &code, &tt, &nn, nullptr);
    if (st != SVO_OK) return st;
    *hit_out = code != SVO_MISS;
    if (code != SVO_MISS) *t = tt;             // VoxelOctree.cpp:266 / :344
    if (code == SVO_HIT_LEAF) *normal = nn;    // VoxelOctree.cpp:282
    return SVO_OK;
}

/* ---- camera ----------------------------------------------------------------- */

void svo_orbit_camera(float pitch_deg, float yaw_deg, float radius, svo_camera *out) {
    if (out) svo::orbitCamera(pitch_deg, yaw_deg, radius, *out);
}

void svo_viewer_init(svo_viewer_state *state) {
    if (state) svo::viewerInit(*state);
}

int svo_viewer_feed(svo_viewer_state *state, const svo_viewer_event *event) {
    if (!state || !event) return -fail(SVO_ERR_INVALID_ARGUMENT, "svo_viewer_feed: null argument");
    return svo::viewerFeed(*state, *event);
}

int svo_frame_constants_from_camera(const svo_camera *cam, const float center[3], int width, int height, int strips,
                                    svo_frame_constants *out) {
    if (!cam || !center || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_frame_constants_from_camera: null argument");
    if (width < 1 || height < 1 || strips < 1) return fail(SVO_ERR_INVALID_ARGUMENT, "bad frame configuration");
    svo::frameConstants(*cam, center, width, height, strips, *out);
    return SVO_OK;
}

/* ---- frames ------------------------------------------------------------------ */

int svo_frame_get_layout(int width, int height, int strips, svo_frame_layout *out) {
    if (!out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_frame_get_layout: null argument");
    svo_frame_desc d = {width, height, strips, SVO_FLAVOUR_VALIDATION, 0, 1, 0, 0};
    int st = checkDesc(&d);
    if (st != SVO_OK) return st;
    svo::FramePlanDev p{};
    planGeometry(width, height, strips, p);
    out->n_strips = p.nStrips;
    out->strip_rows = p.stripRows;
    out->tiles_x = p.tilesX;
    out->tiles_y_full = p.tilesYFull;
    out->tiles_y_last = p.tilesYLast;
    out->tile_cols = p.tileCols;
    out->tiles = p.totalTiles;
    out->corners = p.totalCorners;
    return SVO_OK;
}

int svo_frame_tile_owner(int width, int height, int strips, int tile, int tile_world) {
    svo_frame_desc d = {width, height, strips, SVO_FLAVOUR_VALIDATION, 0, tile_world, 0, 0};
    if (checkDesc(&d) != SVO_OK) return -1;
    svo::FramePlanDev p{};
    planGeometry(width, height, strips, p);
    if (tile < 0 || tile >= p.totalTiles) {
        fail(SVO_ERR_INVALID_ARGUMENT, "tile %d out of range [0, %d)", tile, p.totalTiles);
        return -1;
    }
    return ((tile % p.tileCols)/svo::tileShare(0, tile_world).run) % tile_world;
}

int svo_frame_set_tile_run(int run) {
    if (run > 4096) return fail(SVO_ERR_INVALID_ARGUMENT, "tile run %d is out of range", run);
    svo::setDefaultTileRun(run);
    return SVO_OK;
}

int svo_frame_tile_rect(int width, int height, int strips, int tile, int32_t rect[4]) {
    if (!rect) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_frame_tile_rect: null argument");
    svo_frame_desc d = {width, height, strips, SVO_FLAVOUR_VALIDATION, 0, 1, 0, 0};
    int st = checkDesc(&d);
    if (st != SVO_OK) return st;
    svo::FramePlanDev p{};
    planGeometry(width, height, strips, p);
    if (tile < 0 || tile >= p.totalTiles) return fail(SVO_ERR_INVALID_ARGUMENT, "tile %d out of range [0, %d)", tile, p.totalTiles);
    // same arithmetic as classifyTilesKernel
    int tileRow = tile/p.tileCols, tx = tile - tileRow*p.tileCols;
    int strip = tileRow/p.tileRowsFull < p.nStrips - 1 ? tileRow/p.tileRowsFull : p.nStrips - 1;
    int ty = tileRow - strip*p.tileRowsFull;
    int stripY0 = strip*p.stripRows;
    int yEnd = stripY0 + p.stripRows < height ? stripY0 + p.stripRows : height;
    rect[0] = tx*8;
    rect[1] = stripY0 + ty*8;
    rect[2] = rect[0] + 8 < width ? rect[0] + 8 : width;
    rect[3] = rect[1] + 8 < yEnd ? rect[1] + 8 : yEnd;
    return SVO_OK;
}

int svo_render_frame_device(svo_tree *tree, const svo_camera *cam, const svo_frame_desc *desc, uint32_t *d_rgba,
                            float *d_depth, void *stream, svo_frame_stats *stats, int sync_stats) {
    if (!tree || !cam || !d_rgba) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_render_frame_device: null argument");
    int st = checkDesc(desc);
    if (st != SVO_OK) return st;
    if (desc->pixel_format != SVO_PIXELS_RGBA8) return fail(SVO_ERR_UNSUPPORTED, "svo_render_frame_device writes SVO_PIXELS_RGBA8 only");
    SVO_DEVICE(tree->device);
    std::lock_guard<std::mutex> lock(tree->mutex);
    FramePlan *plan = nullptr;
    if ((st = getPlan(tree, desc->width, desc->height, desc->strips, &plan)) != SVO_OK) return st;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    bool wantStats = stats && sync_stats;
    uint32_t launches = 0;
    int slot = 0;
    const svo::TileShare share = svo::tileShare(desc->tile_rank, desc->tile_world);
    if ((st = enqueueFrame(tree, plan, cam, desc, share, d_rgba, d_depth, s, wantStats, &launches, &slot)) != SVO_OK) return st;
    if (wantStats) {
        SVO_CUDA(cudaStreamSynchronize(s));
        fillStats(plan, slot, share, launches, stats);
    }
    return SVO_OK;
}

int svo_frame_copy_owned_tiles(int device, const svo_frame_desc *desc, const uint32_t *d_src, uint32_t *d_dst, void *stream) {
    if (!d_src || !d_dst) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_frame_copy_owned_tiles: null argument");
    int st = checkDesc(desc);
    if (st != SVO_OK) return st;
    SVO_DEVICE(device);
    svo::FramePlanDev p{};
    planGeometry(desc->width, desc->height, desc->strips, p);
    SVO_CUDA(svo::launchCopyOwnedColumns(p, desc->width, desc->height, d_src, d_dst,
                                         svo::tileShare(desc->tile_rank, desc->tile_world), static_cast<cudaStream_t>(stream)));
    return SVO_OK;
}

int svo_render_frame_async(svo_tree *tree, const svo_camera *cam, const svo_frame_desc *desc, uint32_t *rgba,
                           float *depth, int want_stats, int *ticket) {
    if (!tree || !cam || !rgba || !ticket) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_render_frame_async: null argument");
    int st = checkDesc(desc);
    if (st != SVO_OK) return st;
    if (desc->pixel_format != SVO_PIXELS_RGBA8) return fail(SVO_ERR_UNSUPPORTED, "svo_render_frame[_async] delivers SVO_PIXELS_RGBA8 only (svo_multi_render_* take SVO_PIXELS_GREY8A8)");
    SVO_DEVICE(tree->device);
    std::lock_guard<std::mutex> lock(tree->mutex);
    FramePlan *plan = nullptr;
    if ((st = getPlan(tree, desc->width, desc->height, desc->strips, &plan)) != SVO_OK) return st;
    const int b = int(plan->hostFrameNumber % kHostLanes);  // staging framebuffer / ticket of this frame
    if (plan->pending[b])
        return fail(SVO_ERR_INVALID_ARGUMENT, "%d frames are already in flight for this configuration: call svo_frame_wait first", kHostLanes);
    size_t frameBytes = size_t(desc->width)*size_t(desc->height)*sizeof(uint32_t);
    if (!plan->dRgba[b]) {
        SVO_CUDA(cudaMalloc(&plan->dRgba[b], frameBytes));
        SVO_CUDA(cudaMemset(plan->dRgba[b], 0, frameBytes));
    }
    // consecutive frames render on different streams into different staging framebuffers, so the
    // long-ray tail of one fine pass overlaps the start of the next ones
    cudaStream_t s = b == 0 ? tree->stream : b == 1 ? tree->stream2 : tree->stream34[b - 2];
    // the staging framebuffer is free once its previous device->host copy has finished
    if (plan->copyRecorded[b]) SVO_CUDA(cudaStreamWaitEvent(s, plan->copyDone[b], 0));
    uint32_t launches = 0;
    int slot = 0;
    const svo::TileShare share = svo::tileShare(desc->tile_rank, desc->tile_world);
    if ((st = enqueueFrame(tree, plan, cam, desc, share, plan->dRgba[b], nullptr, s, want_stats != 0, &launches, &slot)) != SVO_OK) return st;
    // copies run on their own stream so that the next frame's kernels are not queued behind them
    SVO_CUDA(cudaStreamWaitEvent(tree->copyStream, plan->fineDone[slot], 0));
    SVO_CUDA(cudaMemcpyAsync(rgba, plan->dRgba[b], frameBytes, cudaMemcpyDeviceToHost, tree->copyStream));
    if (depth)
        SVO_CUDA(cudaMemcpyAsync(depth, plan->dDepth[slot], size_t(plan->dev.totalCorners)*sizeof(float),
                                 cudaMemcpyDeviceToHost, tree->copyStream));
    SVO_CUDA(cudaEventRecord(plan->copyDone[b], tree->copyStream));
    ++plan->hostFrameNumber;
    plan->copyRecorded[b] = true;
    plan->pending[b] = true;
    plan->pendingStats[b] = want_stats != 0;
    plan->pendingRing[b] = slot;
    plan->pendingLaunches[b] = launches;
    plan->pendingShare[b] = share;
    *ticket = b;
    return SVO_OK;
}

int svo_frame_wait(svo_tree *tree, const svo_frame_desc *desc, int ticket, svo_frame_stats *stats) {
    if (!tree || !desc) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_frame_wait: null argument");
    if (ticket < 0 || ticket >= kHostLanes) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_frame_wait: bad ticket %d", ticket);
    SVO_DEVICE(tree->device);
    std::lock_guard<std::mutex> lock(tree->mutex);
    auto it = tree->plans.find(std::make_tuple(desc->width, desc->height, desc->strips));
    if (it == tree->plans.end() || !it->second.pending[ticket])
        return fail(SVO_ERR_INVALID_ARGUMENT, "svo_frame_wait: no frame in flight for this ticket");
    FramePlan *plan = &it->second;
    SVO_CUDA(cudaEventSynchronize(plan->copyDone[ticket]));
    plan->pending[ticket] = false;
    if (stats) {
        if (!plan->pendingStats[ticket])
            return fail(SVO_ERR_INVALID_ARGUMENT, "svo_frame_wait: statistics were not requested for this frame");
        fillStats(plan, plan->pendingRing[ticket], plan->pendingShare[ticket], plan->pendingLaunches[ticket], stats);
    }
    return SVO_OK;
}

int svo_render_frame(svo_tree *tree, const svo_camera *cam, const svo_frame_desc *desc, uint32_t *rgba, float *depth,
                     svo_frame_stats *stats) {
    int ticket = 0;
    int st = svo_render_frame_async(tree, cam, desc, rgba, depth, stats != nullptr, &ticket);
    if (st != SVO_OK) return st;
    return svo_frame_wait(tree, desc, ticket, stats);
}

/* ---- shading ------------------------------------------------------------------ */

int svo_shade_batch(svo_tree *tree, uint64_t n, const uint8_t *hit, const uint32_t *normal, const float *d,
                    const float light[3], uint32_t *rgba) {
    if (!tree || !light || (n && (!normal || !d || !rgba))) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_shade_batch: null argument");
    if (n == 0) return SVO_OK;
    SVO_DEVICE(tree->device);
    std::lock_guard<std::mutex> lock(tree->mutex);
    const uint64_t chunk = 4ull << 20;
    const size_t inBytes = size_t(std::min(n, chunk))*(1 + 4 + 12), outBytes = size_t(std::min(n, chunk))*4;
    SVO_CUDA(tree->batchIn.reserve(inBytes));
    SVO_CUDA(tree->batchOut.reserve(outBytes));
    for (uint64_t first = 0; first < n; first += chunk) {
        const uint64_t m = std::min(chunk, n - first);
        float *dD = static_cast<float *>(tree->batchIn.ptr);
        uint32_t *dNormal = reinterpret_cast<uint32_t *>(dD + 3*m);
        uint8_t *dHit = reinterpret_cast<uint8_t *>(dNormal + m);
        uint32_t *dRgba = static_cast<uint32_t *>(tree->batchOut.ptr);
        cudaStream_t s = tree->stream;
        SVO_CUDA(cudaMemcpyAsync(dD, d + 3*first, size_t(m)*12, cudaMemcpyHostToDevice, s));
        SVO_CUDA(cudaMemcpyAsync(dNormal, normal + first, size_t(m)*4, cudaMemcpyHostToDevice, s));
        if (hit) SVO_CUDA(cudaMemcpyAsync(dHit, hit + first, size_t(m), cudaMemcpyHostToDevice, s));
        SVO_CUDA(svo::launchShadeBatch(m, hit ? dHit : nullptr, dNormal, dD, light, dRgba, s));
        SVO_CUDA(cudaMemcpyAsync(rgba + first, dRgba, size_t(m)*4, cudaMemcpyDeviceToHost, s));
        SVO_CUDA(cudaStreamSynchronize(s));
    }
    return SVO_OK;
}

/* ---- device memory + peer mapping -------------------------------------------- */

int svo_device_alloc(int device, size_t bytes, void **out) {
    if (!out) return fail(SVO_ERR_INVALID_ARGUMENT, "null out");
    *out = nullptr;
    SVO_DEVICE(device);
    SVO_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    return SVO_OK;
}

int svo_device_free(int device, void *p) {
    if (!p) return SVO_OK;
    SVO_DEVICE(device);
    SVO_CUDA(cudaFree(p));
    return SVO_OK;
}

int svo_device_memset(int device, void *p, int value, size_t bytes) {
    SVO_DEVICE(device);
    SVO_CUDA(cudaMemset(p, value, bytes));
    return SVO_OK;
}

int svo_device_to_host(int device, void *host_dst, const void *device_src, size_t bytes) {
    SVO_DEVICE(device);
    SVO_CUDA(cudaMemcpy(host_dst, device_src, bytes, cudaMemcpyDeviceToHost));
    return SVO_OK;
}

int svo_device_to_host_async(int device, void *host_dst, const void *device_src, size_t bytes, void *stream) {
    SVO_DEVICE(device);
    SVO_CUDA(cudaMemcpyAsync(host_dst, device_src, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    return SVO_OK;
}

int svo_host_to_device(int device, void *device_dst, const void *host_src, size_t bytes) {
    SVO_DEVICE(device);
    SVO_CUDA(cudaMemcpy(device_dst, host_src, bytes, cudaMemcpyHostToDevice));
    return SVO_OK;
}

int svo_device_synchronize(int device) {
    SVO_DEVICE(device);
    SVO_CUDA(cudaDeviceSynchronize());
    return SVO_OK;
}

int svo_ipc_export(int device, void *p, uint8_t handle[SVO_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == SVO_IPC_HANDLE_BYTES, "IPC handle size");
    if (!p || !handle) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_ipc_export: null argument");
    SVO_DEVICE(device);
    cudaIpcMemHandle_t h;
    SVO_CUDA(cudaIpcGetMemHandle(&h, p));
    memcpy(handle, &h, sizeof h);
    return SVO_OK;
}

int svo_ipc_open(int device, const uint8_t handle[SVO_IPC_HANDLE_BYTES], void **out) {
    if (!handle || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_ipc_open: null argument");
    *out = nullptr;
    SVO_DEVICE(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    SVO_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return SVO_OK;
}

int svo_ipc_close(int device, void *p) {
    if (!p) return SVO_OK;
    SVO_DEVICE(device);
    SVO_CUDA(cudaIpcCloseMemHandle(p));
    return SVO_OK;
}

} // extern "C"
