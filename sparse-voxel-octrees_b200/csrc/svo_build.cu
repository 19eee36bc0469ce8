// Octree construction on the GPU -- SURVEY.md section 8, "next" row f2.
//
// What it computes is fixed by the reference: VoxelOctree::VoxelOctree(VoxelData*) and buildOctree
// (reference src/VoxelOctree.cpp:125-205) over the occupancy VoxelData reports
// (src/VoxelData.hpp:106-138), finalised by ChunkedAllocator (src/ChunkedAllocator.hpp:73-116). The
// result must be the same uint32 array, word for word.
//
// How it computes it is not. The reference recurses depth-first on the host, appending blocks to a
// chunked array and remembering far words to splice in afterwards. The layout that produces is a
// closed form of the subtree sizes, so here it is three data-parallel sweeps over sorted arrays in HBM:
//
//   gather    every non-empty voxel becomes (Morton key, material word); x is the lowest key bit,
//             because buildOctree visits the children of a node in the order -x-y-z ... +x+y+z reversed
//             (VoxelOctree.cpp:144-146,164,177), i.e. ascending x + 2y + 4z. One radix sort.
//   bottom-up level l = depth-1 .. 0: a node is a run of child keys with equal key >> 3: one reduce-by-key
//             (OR of the children's octant bits) gives parents and valid masks, the scan of the masks'
//             popcounts where their children start. Per node: child count c, and -- with F(child) = number of words everything below that
//             child's descriptor occupies, prefix-summed over the child level --
//               G_last = 1 + sum of F over all children but the last   (VoxelOctree.cpp:176-184; G_i =
//                        c - i + sum_{j<i} F_j is the distance from child descriptor i to its own
//                        child block, insertions included, and it is largest for the last child)
//               far    = G_last > 0x3FFF                               (:182-183, all-or-nothing per block)
//               F      = c * (1 + far) + sum of F over all children    (the block, its far words, the rest)
//             Leaf parents (halfSize == 1, :166-172) have F = c and never far words.
//   top-down  the root descriptor is word 0 and its block starts at word 1 (:128-132). A node whose block
//             starts at B with c children and stride s = 1 + far puts child i's descriptor at B + i*s and
//             that child's block at B + c*s + sum_{j<i} F_j (:186-196: inline 14-bit offset, or far word +
//             high bits, plus bit 16 on the parent / bit 17 on the child). Leaf parents write their
//             voxels' material words.
//
// Occupancy follows the reference including its blind spot: buildLowLut's thread partition only looks
// at z-planes [0, (depth/2)*2) (VoxelData.cpp:160-163), so the last plane of an odd-depth volume is
// never seen. (With a cache block smaller than the volume the reference itself becomes inconsistent
// there -- top and low LUT disagree and it emits a child-less descriptor; this builder always behaves
// like the reference with the whole volume in one cache block.)
//
// cub (shipped with the CUDA toolkit) provides the radix sort and the prefix sums; everything else is
// the kernels below.
#include "svo_build.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

namespace svo {

namespace {

constexpr int kThreads = 256;

struct Dims {
    int w, h, d;
    int skipPlane;      // z of the plane buildLowLut never sees, or -1
};

// bits of v (21 used) to every third bit
__device__ __forceinline__ uint64_t spread3(uint32_t v) {
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
__device__ __forceinline__ uint64_t mortonKey(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}

// Block-aggregated append of the threads with `keep` set: one atomic on the global cursor per block and
// iteration (a single cursor takes ~1 atomic per ns; one per warp made the gather atomics-bound at
// 300 M voxels). Every thread of the block must call it the same number of times.
__device__ __forceinline__ void appendEntry(bool keep, uint64_t key, uint32_t value, uint64_t *keys, uint32_t *vals,
                                            unsigned long long *cursor) {
    __shared__ unsigned warpCount[kThreads/32];
    __shared__ unsigned long long blockBase;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warpCount[warp] = unsigned(__popc(mask));
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned total = 0;
        for (int w = 0; w < kThreads/32; ++w) {
            const unsigned c = warpCount[w];
            warpCount[w] = total;           // exclusive prefix
            total += c;
        }
        blockBase = total ? atomicAdd(cursor, (unsigned long long)total) : 0ull;
    }
    __syncthreads();
    if (keep) {
        const unsigned long long at = blockBase + warpCount[warp] + __popc(mask & ((1u << lane) - 1u));
        keys[at] = key;
        vals[at] = value;
    }
    __syncthreads();                        // warpCount / blockBase are rewritten by the next call
}

__global__ void __launch_bounds__(kThreads)
gatherDenseKernel(const uint32_t *__restrict__ voxels, uint64_t first, uint64_t count, Dims dims, uint64_t *keys,
                  uint32_t *vals, unsigned long long *cursor) {
    const uint64_t stride = uint64_t(gridDim.x)*blockDim.x;
    const uint64_t rounded = (count + (kThreads - 1)) & ~uint64_t(kThreads - 1);   // whole blocks take part in the append
    for (uint64_t i = uint64_t(blockIdx.x)*blockDim.x + threadIdx.x; i < rounded; i += stride) {
        uint32_t v = i < count ? __ldg(voxels + i) : 0u;
        uint64_t key = 0;
        bool keep = v != 0u;
        if (keep) {
            const uint64_t g = first + i;
            const uint64_t plane = uint64_t(dims.w)*uint64_t(dims.h);
            const uint32_t z = uint32_t(g/plane);
            const uint32_t r = uint32_t(g - uint64_t(z)*plane);
            const uint32_t y = r/uint32_t(dims.w), x = r - y*uint32_t(dims.w);
            keep = int(z) != dims.skipPlane;
            key = mortonKey(x, y, z);
        }
        appendEntry(keep, key, v, keys, vals, cursor);
    }
}

__global__ void __launch_bounds__(kThreads)
gatherSparseKernel(const uint32_t *__restrict__ xyz, const uint32_t *__restrict__ values, uint64_t n, Dims dims,
                   uint64_t *keys, uint32_t *vals, unsigned long long *cursor) {
    const uint64_t stride = uint64_t(gridDim.x)*blockDim.x;
    const uint64_t rounded = (n + (kThreads - 1)) & ~uint64_t(kThreads - 1);
    for (uint64_t i = uint64_t(blockIdx.x)*blockDim.x + threadIdx.x; i < rounded; i += stride) {
        bool keep = false;
        uint64_t key = 0;
        uint32_t v = 0;
        if (i < n) {
            const uint32_t x = __ldg(xyz + 3*i), y = __ldg(xyz + 3*i + 1), z = __ldg(xyz + 3*i + 2);
            v = __ldg(values + i);
            keep = v != 0u && x < uint32_t(dims.w) && y < uint32_t(dims.h) && z < uint32_t(dims.d) && int(z) != dims.skipPlane;
            key = mortonKey(x, y, z);
        }
        appendEntry(keep, key, v, keys, vals, cursor);
    }
}

// A level is one reduce-by-key over the sorted child keys: parent key = key >> 3, value = the child's
// octant bit, reduction = OR. That yields the parents' keys and valid masks in one pass; a parent's
// child count is the popcount of its mask (children are unique), and the exclusive scan of those
// counts is where each parent's children start in the child arrays.
struct ParentKey {
    __host__ __device__ __forceinline__ uint64_t operator()(uint64_t key) const { return key >> 3; }
};
struct OctantBit {
    __host__ __device__ __forceinline__ uint32_t operator()(uint64_t key) const { return 1u << uint32_t(key & 7u); }
};
struct BitOr {
    __host__ __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a | b; }
};
struct PopCount {
    __host__ __device__ __forceinline__ uint32_t operator()(uint32_t mask) const {
#ifdef __CUDA_ARCH__
        return uint32_t(__popc(mask));
#else
        uint32_t n = 0;
        for (; mask; mask &= mask - 1u) ++n;
        return n;
#endif
    }
};

// childPrefix == nullptr: leaf parents (children are voxels)
__global__ void __launch_bounds__(kThreads)
nodeStatsKernel(const uint32_t *__restrict__ childStart, uint32_t nNodes, const uint64_t *__restrict__ childPrefix,
                uint8_t *far, uint64_t *subtree, unsigned long long *farBlocks) {
    const uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
    bool isFar = false;
    if (j < nNodes) {
        const uint32_t s = childStart[j], e = childStart[j + 1];
        const uint64_t count = e - s;
        uint64_t f = count;
        if (childPrefix) {
            const uint64_t base = childPrefix[s];
            isFar = 1u + (childPrefix[e - 1] - base) > 0x3FFFu;                  // VoxelOctree.cpp:182-183
            f = count*(isFar ? 2u : 1u) + (childPrefix[e] - base);
        }
        far[j] = isFar ? 1 : 0;
        subtree[j] = f;
    }
    const unsigned votes = __ballot_sync(0xffffffffu, isFar);
    if ((threadIdx.x & 31u) == 0 && votes) atomicAdd(farBlocks, (unsigned long long)__popc(votes));
}

// One thread per node of level l: writes the descriptors (and far words) of its children, or its
// voxels' material words, and hands every child its block address.
__global__ void __launch_bounds__(kThreads)
emitLevelKernel(uint32_t nNodes, const uint32_t *__restrict__ childStart, const uint8_t *__restrict__ far,
                const uint64_t *__restrict__ blockBase, const uint64_t *__restrict__ childPrefix,
                const uint32_t *__restrict__ childMask, const uint8_t *__restrict__ childFar, bool childIsLeafParent,
                const uint32_t *__restrict__ voxelValues, uint64_t *childBlockBase, uint32_t *out) {
    const uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= nNodes) return;
    const uint32_t s = childStart[j], e = childStart[j + 1];
    const uint64_t base = blockBase[j];
    if (voxelValues) {                                                            // VoxelOctree.cpp:166-172
        for (uint32_t c = s; c < e; ++c) out[base + (c - s)] = voxelValues[c];
        return;
    }
    const uint64_t count = e - s;
    const bool isFar = far[j] != 0;
    const uint64_t stride = isFar ? 2u : 1u;
    const uint64_t first = base + count*stride, p0 = childPrefix[s];
    for (uint32_t c = s; c < e; ++c) {
        const uint64_t at = base + uint64_t(c - s)*stride;
        const uint64_t block = first + (childPrefix[c] - p0);
        const uint64_t offset = block - at;
        const uint32_t m = childMask[c];
        uint32_t word = (m << 8) | (childIsLeafParent ? 0u : m) | (childFar[c] ? 0x10000u : 0u);  // :199-201
        if (isFar) {                                                              // :188-193
            out[at + 1] = uint32_t(offset);
            word |= 0x20000u | uint32_t((offset >> 32) << 18);
        } else {
            word |= uint32_t(offset << 18);                                       // :195
        }
        out[at] = word;
        childBlockBase[c] = block;
    }
}

__global__ void emitRootKernel(const uint32_t *mask, const uint8_t *far, bool rootIsLeafParent, uint32_t *out,
                               uint64_t *blockBase) {
    const uint32_t m = mask[0];
    out[0] = (m << 8) | (rootIsLeafParent ? 0u : m) | (far[0] ? 0x10000u : 0u) | (1u << 18);   // :128-132
    blockBase[0] = 1;
}

// Scratch memory comes from a private stream-ordered pool (one per device): a build makes some sixty
// allocations of up to gigabytes, level after level, and plain cudaMalloc / cudaFree (map, unmap and a
// device synchronisation each) cost more than the kernels between them. Blocks freed by one level are
// reused by the next; the pool is trimmed when the build is over.
cudaMemPool_t scratchPool() {
    static cudaMemPool_t pools[64] = {};
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= 64) return nullptr;
    if (!pools[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&pools[device], &props) != cudaSuccess) { pools[device] = nullptr; return nullptr; }
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pools[device], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return pools[device];
}

// Trimming the pool (unmapping and freeing tens of GB) takes 0.1 - 2 s, none of which the caller of a build
// has to wait for: it runs on a background thread, which the next build (or process exit) joins first.
struct PoolTrimmer {
    std::mutex lock;
    std::thread worker;
    void wait() {
        std::lock_guard<std::mutex> g(lock);
        if (worker.joinable()) worker.join();
    }
    void start() {
        cudaMemPool_t pool = scratchPool();
        int device = 0;
        if (!pool || cudaGetDevice(&device) != cudaSuccess) return;
        std::lock_guard<std::mutex> g(lock);
        if (worker.joinable()) worker.join();
        worker = std::thread([pool, device] {
            if (cudaSetDevice(device) == cudaSuccess) cudaMemPoolTrimTo(pool, size_t(256) << 20);
        });
    }
    ~PoolTrimmer() { if (worker.joinable()) worker.join(); }
};
PoolTrimmer &poolTrimmer() {
    static PoolTrimmer t;
    return t;
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p) { o.p = nullptr; }
    DevBuf &operator=(DevBuf &&o) noexcept { release(); p = o.p; o.p = nullptr; return *this; }
    ~DevBuf() { release(); }
    cudaError_t alloc(uint64_t n) {
        release();
        const size_t bytes = size_t(n ? n : 1)*sizeof(T);
        cudaMemPool_t pool = scratchPool();
        if (!pool) return cudaMalloc(&p, bytes);
        return cudaMallocFromPoolAsync(reinterpret_cast<void **>(&p), bytes, pool, 0);
    }
    void release() { if (p) cudaFreeAsync(p, 0); p = nullptr; }   // cudaFreeAsync also takes cudaMalloc'ed blocks
    T *detach() { T *q = p; p = nullptr; return q; }
};

struct Level {
    uint32_t n = 0;
    DevBuf<uint64_t> keys;       // n; released once the parent level is built
    DevBuf<uint32_t> childStart; // n + 1
    DevBuf<uint32_t> mask;       // n (+ 1 zero behind the last, for the scan)
    DevBuf<uint8_t> far;         // n
    DevBuf<uint64_t> prefix;     // n + 1: exclusive prefix sum of F over this level
};

inline unsigned gridFor(uint64_t n) { return unsigned((n + kThreads - 1)/kThreads); }

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~Timer() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    void start() { cudaEventRecord(a, 0); }
    float stop() {
        cudaEventRecord(b, 0);
        cudaEventSynchronize(b);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, a, b);
        return ms;
    }
};

#define SVO_BUILD_CUDA(call)                                                          \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            err = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
            return false;                                                             \
        }                                                                             \
    } while (0)

} // namespace

cudaMemPool_t buildScratchPool() { return scratchPool(); }

OctreeBuilder::~OctreeBuilder() {
    if (dKeys_) cudaFree(dKeys_);
    if (dVals_) cudaFree(dVals_);
    if (dCursor_) cudaFree(dCursor_);
}

bool OctreeBuilder::begin(int w, int h, int d, std::string &err) {
    poolTrimmer().wait();      // a trim left over from the previous build would take back what this one reuses
    if (w <= 0 || h <= 0 || d <= 0) { err = "volume dimensions must be positive"; return false; }
    if (w > (1 << 21) || h > (1 << 21) || d > (1 << 21)) { err = "volume dimensions above 2^21 are not supported"; return false; }
    w_ = w; h_ = h; d_ = d;
    side_ = 1;
    levels_ = 0;
    while (side_ < w || side_ < h || side_ < d) { side_ <<= 1; ++levels_; }     // roundToPow2 / sideLength, VoxelData.cpp:204-207,293-295
    if (levels_ < 1) { err = "a 1 x 1 x 1 volume has no octree (the reference needs side >= 2)"; return false; }
    if (levels_ > 23) { err = "deeper than 23 levels"; return false; }
    SVO_BUILD_CUDA(cudaMalloc(&dCursor_, sizeof(unsigned long long)));
    SVO_BUILD_CUDA(cudaMemset(dCursor_, 0, sizeof(unsigned long long)));
    count_ = 0;
    return true;
}

bool OctreeBuilder::reserve(uint64_t entries, std::string &err) {
    if (entries <= capacity_) return true;
    uint64_t want = std::max<uint64_t>(entries, capacity_ + capacity_/2);
    uint64_t *keys = nullptr;
    uint32_t *vals = nullptr;
    SVO_BUILD_CUDA(cudaMalloc(&keys, size_t(want)*sizeof(uint64_t)));
    cudaError_t e = cudaMalloc(&vals, size_t(want)*sizeof(uint32_t));
    if (e != cudaSuccess) { cudaFree(keys); err = std::string("cudaMalloc(voxel values): ") + cudaGetErrorString(e); return false; }
    if (count_) {
        cudaMemcpy(keys, dKeys_, size_t(count_)*sizeof(uint64_t), cudaMemcpyDeviceToDevice);
        cudaMemcpy(vals, dVals_, size_t(count_)*sizeof(uint32_t), cudaMemcpyDeviceToDevice);
    }
    if (dKeys_) cudaFree(dKeys_);
    if (dVals_) cudaFree(dVals_);
    dKeys_ = keys;
    dVals_ = vals;
    capacity_ = want;
    return true;
}

bool OctreeBuilder::addDenseChunk(const uint32_t *dVoxels, uint64_t first, uint64_t count, std::string &err) {
    if (!dCursor_) { err = "OctreeBuilder::begin was not called"; return false; }
    if (count == 0) return true;
    if (first + count > uint64_t(w_)*uint64_t(h_)*uint64_t(d_)) { err = "dense chunk past the end of the volume"; return false; }
    if (!reserve(count_ + count, err)) return false;      // worst case: every voxel of the chunk is filled
    Dims dims{w_, h_, d_, (d_ & 1) ? d_ - 1 : -1};
    Timer t;
    t.start();
    const unsigned blocks = std::min<uint64_t>(gridFor(count), 148u*16u);
    gatherDenseKernel<<<blocks, kThreads>>>(dVoxels, first, count, dims, dKeys_, dVals_, dCursor_);
    SVO_BUILD_CUDA(cudaGetLastError());
    unsigned long long now = 0;
    SVO_BUILD_CUDA(cudaMemcpy(&now, dCursor_, sizeof now, cudaMemcpyDeviceToHost));
    gatherMs_ += t.stop();
    count_ = now;
    return true;
}

bool OctreeBuilder::addSparse(const uint32_t *dXyz, const uint32_t *dValues, uint64_t n, std::string &err) {
    if (!dCursor_) { err = "OctreeBuilder::begin was not called"; return false; }
    if (n == 0) return true;
    if (!reserve(count_ + n, err)) return false;
    Dims dims{w_, h_, d_, (d_ & 1) ? d_ - 1 : -1};
    Timer t;
    t.start();
    const unsigned blocks = std::min<uint64_t>(gridFor(n), 148u*16u);
    gatherSparseKernel<<<blocks, kThreads>>>(dXyz, dValues, n, dims, dKeys_, dVals_, dCursor_);
    SVO_BUILD_CUDA(cudaGetLastError());
    unsigned long long now = 0;
    SVO_BUILD_CUDA(cudaMemcpy(&now, dCursor_, sizeof now, cudaMemcpyDeviceToHost));
    gatherMs_ += t.stop();
    count_ = now;
    return true;
}

bool OctreeBuilder::finish(BuildResult &out, std::string &err) {
    if (!dCursor_) { err = "OctreeBuilder::begin was not called"; return false; }
    if (count_ == 0) { err = "the volume has no visible non-empty voxel: nothing to build"; return false; }
    if (count_ >= (1ull << 31)) { err = "more than 2^31 - 1 voxels are not supported (32-bit item counts in the scans)"; return false; }
    const uint32_t nVoxels = uint32_t(count_);
    BuildStats stats;
    stats.voxels = count_;
    stats.gatherMs = gatherMs_;
    Timer timer;
    const bool debugTiming = getenv("SVO_BUILD_DEBUG") != nullptr;
    const auto wall0 = std::chrono::steady_clock::now();
    auto wallMs = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count(); };

    // ---- sort by Morton key
    timer.start();
    DevBuf<uint64_t> keysAlt;
    DevBuf<uint32_t> valsAlt;
    SVO_BUILD_CUDA(keysAlt.alloc(nVoxels));
    SVO_BUILD_CUDA(valsAlt.alloc(nVoxels));
    cub::DoubleBuffer<uint64_t> kb(dKeys_, keysAlt.p);
    cub::DoubleBuffer<uint32_t> vb(dVals_, valsAlt.p);
    size_t tempBytes = 0;
    SVO_BUILD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, kb, vb, int(nVoxels), 0, 3*levels_));
    DevBuf<uint8_t> temp;
    SVO_BUILD_CUDA(temp.alloc(tempBytes));
    SVO_BUILD_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tempBytes, kb, vb, int(nVoxels), 0, 3*levels_));
    const uint64_t *voxelKeys = kb.Current();
    const uint32_t *voxelVals = vb.Current();
    stats.sortMs = timer.stop();

    // ---- bottom-up
    timer.start();
    DevBuf<unsigned long long> counters;     // [0] far blocks
    SVO_BUILD_CUDA(counters.alloc(1));
    SVO_BUILD_CUDA(cudaMemset(counters.p, 0, sizeof(unsigned long long)));
    DevBuf<uint32_t> numRuns;
    SVO_BUILD_CUDA(numRuns.alloc(1));
    std::vector<Level> level(static_cast<size_t>(levels_));
    DevBuf<uint8_t> cubTemp;
    size_t cubTempCap = 0;
    auto needTemp = [&](size_t bytes) -> cudaError_t {
        if (bytes <= cubTempCap) return cudaSuccess;
        cudaError_t e = cubTemp.alloc(bytes);
        cubTempCap = e == cudaSuccess ? bytes : 0;
        return e;
    };

    const uint64_t *childKeys = voxelKeys;
    uint32_t m = nVoxels;
    for (int l = levels_ - 1; l >= 0; --l) {
        Level &lv = level[size_t(l)];
        const bool leafParent = l == levels_ - 1;
        // parents' keys and masks: at most m of them
        SVO_BUILD_CUDA(lv.keys.alloc(m));
        SVO_BUILD_CUDA(lv.mask.alloc(uint64_t(m) + 1));
        thrust::transform_iterator<ParentKey, const uint64_t *> parentKeys(childKeys, ParentKey());
        thrust::transform_iterator<OctantBit, const uint64_t *> octantBits(childKeys, OctantBit());
        size_t bytes = 0;
        SVO_BUILD_CUDA(cub::DeviceReduce::ReduceByKey(nullptr, bytes, parentKeys, lv.keys.p, octantBits, lv.mask.p,
                                                      numRuns.p, BitOr(), int(m)));
        SVO_BUILD_CUDA(needTemp(bytes));
        SVO_BUILD_CUDA(cub::DeviceReduce::ReduceByKey(cubTemp.p, bytes, parentKeys, lv.keys.p, octantBits, lv.mask.p,
                                                      numRuns.p, BitOr(), int(m)));
        SVO_BUILD_CUDA(cudaMemcpy(&lv.n, numRuns.p, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        // where each parent's children start: exclusive scan of popcount(mask), n + 1 items (the last is 0)
        SVO_BUILD_CUDA(cudaMemsetAsync(lv.mask.p + lv.n, 0, sizeof(uint32_t)));
        SVO_BUILD_CUDA(lv.childStart.alloc(uint64_t(lv.n) + 1));
        thrust::transform_iterator<PopCount, const uint32_t *> childCounts(lv.mask.p, PopCount());
        bytes = 0;
        SVO_BUILD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, childCounts, lv.childStart.p, int(lv.n) + 1));
        SVO_BUILD_CUDA(needTemp(bytes));
        SVO_BUILD_CUDA(cub::DeviceScan::ExclusiveSum(cubTemp.p, bytes, childCounts, lv.childStart.p, int(lv.n) + 1));
        if (leafParent) {
            // children of one parent are unique, so the popcounts add up to m unless a coordinate was named twice
            uint32_t covered = 0;
            SVO_BUILD_CUDA(cudaMemcpy(&covered, lv.childStart.p + lv.n, sizeof(uint32_t), cudaMemcpyDeviceToHost));
            if (covered != m) {
                err = "sparse voxel list names " + std::to_string(m - covered) + " coordinate(s) more than once";
                return false;
            }
        }
        SVO_BUILD_CUDA(lv.far.alloc(lv.n));
        SVO_BUILD_CUDA(lv.prefix.alloc(uint64_t(lv.n) + 1));
        SVO_BUILD_CUDA(cudaMemsetAsync(lv.prefix.p + lv.n, 0, sizeof(uint64_t)));
        nodeStatsKernel<<<gridFor(lv.n), kThreads>>>(lv.childStart.p, lv.n, leafParent ? nullptr : level[size_t(l) + 1].prefix.p,
                                                     lv.far.p, lv.prefix.p, counters.p);
        bytes = 0;
        SVO_BUILD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, lv.prefix.p, lv.prefix.p, int(lv.n) + 1));
        SVO_BUILD_CUDA(needTemp(bytes));
        SVO_BUILD_CUDA(cub::DeviceScan::ExclusiveSum(cubTemp.p, bytes, lv.prefix.p, lv.prefix.p, int(lv.n) + 1));
        SVO_BUILD_CUDA(cudaGetLastError());
        if (!leafParent) level[size_t(l) + 1].keys.release();
        childKeys = lv.keys.p;
        m = lv.n;
        stats.nodes += lv.n;
    }
    if (level[0].n != 1) { err = "internal error: the top level has " + std::to_string(level[0].n) + " nodes"; return false; }
    uint64_t rootSubtree = 0;
    SVO_BUILD_CUDA(cudaMemcpy(&rootSubtree, level[0].prefix.p + 1, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    unsigned long long farBlocks = 0;
    SVO_BUILD_CUDA(cudaMemcpy(&farBlocks, counters.p, sizeof farBlocks, cudaMemcpyDeviceToHost));
    stats.farBlocks = farBlocks;
    stats.levelsMs = timer.stop();

    // ---- top-down
    timer.start();
    const uint64_t nWords = 1 + rootSubtree;
    uint32_t *words = nullptr;
    SVO_BUILD_CUDA(cudaMalloc(&words, size_t(nWords + 1)*sizeof(uint32_t)));
    DevBuf<uint32_t> guard;
    guard.p = words;
    SVO_BUILD_CUDA(cudaMemsetAsync(words + nWords, 0, sizeof(uint32_t)));
    DevBuf<uint64_t> base, nextBase;
    SVO_BUILD_CUDA(base.alloc(1));
    emitRootKernel<<<1, 1>>>(level[0].mask.p, level[0].far.p, levels_ == 1, words, base.p);
    for (int l = 0; l < levels_; ++l) {
        Level &lv = level[size_t(l)];
        const bool leafParent = l == levels_ - 1;
        if (leafParent) {
            emitLevelKernel<<<gridFor(lv.n), kThreads>>>(lv.n, lv.childStart.p, lv.far.p, base.p, nullptr, nullptr, nullptr,
                                                         false, voxelVals, nullptr, words);
        } else {
            Level &ch = level[size_t(l) + 1];
            SVO_BUILD_CUDA(nextBase.alloc(ch.n));
            emitLevelKernel<<<gridFor(lv.n), kThreads>>>(lv.n, lv.childStart.p, lv.far.p, base.p, ch.prefix.p, ch.mask.p,
                                                         ch.far.p, l + 1 == levels_ - 1, nullptr, nextBase.p, words);
            base = std::move(nextBase);
        }
        SVO_BUILD_CUDA(cudaGetLastError());
    }
    stats.emitMs = timer.stop();
    SVO_BUILD_CUDA(cudaDeviceSynchronize());
    const double builtAt = wallMs();
    for (Level &lv : level) lv = Level();
    base.release();
    keysAlt.release();
    valsAlt.release();
    temp.release();
    cubTemp.release();
    counters.release();
    numRuns.release();
    cudaStreamSynchronize(0);
    const double releasedAt = wallMs();
    poolTrimmer().start();
    if (debugTiming)
        fprintf(stderr, "[svo] OctreeBuilder::finish: %.1f ms to the last kernel (device %.1f ms), %.1f ms releasing scratch, %.1f ms handing the pool to the trimmer\n",
                builtAt, stats.sortMs + stats.levelsMs + stats.emitMs, releasedAt - builtAt, wallMs() - releasedAt);

    out.dWords = guard.p;
    guard.p = nullptr;
    out.nWords = nWords;
    out.depth = uint32_t(levels_);
    out.center[0] = float(w_)*0.5f/float(side_);     // VoxelData::getCenter, VoxelData.cpp:297-303
    out.center[1] = float(h_)*0.5f/float(side_);
    out.center[2] = float(d_)*0.5f/float(side_);
    out.stats = stats;
    return true;
}

// ---- the inverse: node array -> filled voxels --------------------------------------------------
// Walks the tree breadth-first, one frontier of (descriptor index, Morton prefix) per level, expanding
// every node into its children by the same rules the traversal uses (descriptor layout: SURVEY.md
// App. A.1; reference src/VoxelOctree.cpp:253-293): valid mask in bits 8-15, ascending octant order in
// the child block, stride 2 when the block is far-interleaved (bit 16), far word after the descriptor
// (bit 17). The last level emits (x, y, z, material word). With the builder above this gives a
// size-independent round-trip property: build(extract(tree)) == tree.

namespace {

__device__ __forceinline__ uint32_t compact3(uint64_t x) {
    x &= 0x1249249249249249ull;
    x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
    x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
    x = (x ^ (x >> 8)) & 0x1f0000ff0000ffull;
    x = (x ^ (x >> 16)) & 0x1f00000000ffffull;
    x = (x ^ (x >> 32)) & 0x1fffffull;
    return uint32_t(x);
}

__global__ void __launch_bounds__(kThreads)
countChildrenKernel(const uint32_t *__restrict__ words, const uint64_t *__restrict__ addr, uint32_t n, uint32_t *counts) {
    const uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= n) return;
    counts[j] = uint32_t(__popc((words[addr[j]] >> 8) & 0xFFu));
}

// leaf == false: children are descriptors -> next frontier; leaf == true: children are voxels -> output
__global__ void __launch_bounds__(kThreads)
expandKernel(const uint32_t *__restrict__ words, const uint64_t *__restrict__ addr, const uint64_t *__restrict__ key,
             const uint32_t *__restrict__ offsets, uint32_t n, bool leaf, uint64_t *nextAddr, uint64_t *nextKey,
             uint32_t *xyz, uint32_t *values) {
    const uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t at = addr[j];
    const uint32_t desc = words[at];
    uint64_t offset = desc >> 18;
    if (desc & 0x20000u) offset = (offset << 32) | words[at + 1];
    const uint64_t block = at + offset;
    const uint64_t stride = (desc & 0x10000u) ? 2u : 1u;
    uint32_t mask = (desc >> 8) & 0xFFu;
    uint32_t out = offsets[j];
    const uint64_t prefix = key[j] << 3;
    for (uint32_t i = 0; mask; ++i, ++out) {
        const uint32_t octant = uint32_t(__ffs(int(mask)) - 1);
        mask &= mask - 1u;
        const uint64_t k = prefix | octant;
        if (leaf) {
            xyz[3*uint64_t(out)] = compact3(k);
            xyz[3*uint64_t(out) + 1] = compact3(k >> 1);
            xyz[3*uint64_t(out) + 2] = compact3(k >> 2);
            values[out] = words[block + i];
        } else {
            nextAddr[out] = block + uint64_t(i)*stride;
            nextKey[out] = k;
        }
    }
}

} // namespace

bool extractVoxels(const uint32_t *dWords, uint32_t depth, uint32_t **dXyzOut, uint32_t **dValuesOut, uint64_t *nOut,
                   std::string &err) {
    *dXyzOut = nullptr;
    *dValuesOut = nullptr;
    *nOut = 0;
    if (depth < 1 || depth > 23) { err = "tree depth out of range"; return false; }
    DevBuf<uint64_t> addr, key, nextAddr, nextKey;
    DevBuf<uint32_t> counts, offsets;
    DevBuf<uint8_t> temp;
    size_t tempCap = 0;
    SVO_BUILD_CUDA(addr.alloc(1));
    SVO_BUILD_CUDA(key.alloc(1));
    SVO_BUILD_CUDA(cudaMemset(addr.p, 0, sizeof(uint64_t)));
    SVO_BUILD_CUDA(cudaMemset(key.p, 0, sizeof(uint64_t)));
    uint32_t n = 1;
    for (uint32_t l = 0; l < depth; ++l) {
        const bool leaf = l + 1 == depth;
        SVO_BUILD_CUDA(counts.alloc(n));
        SVO_BUILD_CUDA(offsets.alloc(n));
        countChildrenKernel<<<gridFor(n), kThreads>>>(dWords, addr.p, n, counts.p);
        size_t bytes = 0;
        SVO_BUILD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, counts.p, offsets.p, int(n)));
        if (bytes > tempCap) { SVO_BUILD_CUDA(temp.alloc(bytes)); tempCap = bytes; }
        SVO_BUILD_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, bytes, counts.p, offsets.p, int(n)));
        uint32_t lastOffset = 0, lastCount = 0;
        SVO_BUILD_CUDA(cudaMemcpy(&lastOffset, offsets.p + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
        SVO_BUILD_CUDA(cudaMemcpy(&lastCount, counts.p + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
        const uint64_t total = uint64_t(lastOffset) + lastCount;
        if (total == 0 || total >= (1ull << 31)) { err = "tree level with no or too many (>= 2^31) children"; return false; }
        if (leaf) {
            DevBuf<uint32_t> xyz, values;
            SVO_BUILD_CUDA(xyz.alloc(total*3));
            SVO_BUILD_CUDA(values.alloc(total));
            expandKernel<<<gridFor(n), kThreads>>>(dWords, addr.p, key.p, offsets.p, n, true, nullptr, nullptr, xyz.p, values.p);
            SVO_BUILD_CUDA(cudaGetLastError());
            SVO_BUILD_CUDA(cudaDeviceSynchronize());
            *dXyzOut = xyz.p;
            *dValuesOut = values.p;
            xyz.p = nullptr;
            values.p = nullptr;
            *nOut = total;
            return true;
        }
        SVO_BUILD_CUDA(nextAddr.alloc(total));
        SVO_BUILD_CUDA(nextKey.alloc(total));
        expandKernel<<<gridFor(n), kThreads>>>(dWords, addr.p, key.p, offsets.p, n, false, nextAddr.p, nextKey.p, nullptr, nullptr);
        SVO_BUILD_CUDA(cudaGetLastError());
        addr = std::move(nextAddr);
        key = std::move(nextKey);
        n = uint32_t(total);
    }
    err = "unreachable";
    return false;
}

} // namespace svo
