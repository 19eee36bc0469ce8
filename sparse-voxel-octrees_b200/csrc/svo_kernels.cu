// sm_100a kernels for the reference's ray-casting hot path:
//   K1 raymarchBatchKernel  VoxelOctree::raymarch in a loop     (reference src/VoxelOctree.cpp:207-346)
//   K2 coarsePassKernel     renderBatch's tile-corner beam pass (reference src/Main.cpp:167-184)
//   K3 finePassKernel       4-corner min + renderTile + shade + pack + the strip memset, fused
//                           (reference src/Main.cpp:92-137, 165, 186-197, 81-90)
//   K4 tileStatsKernel      ray / tile counts for the Mrays/s metric
// No tensor cores: this is pointer chasing over a uint32 node array, bounded by
// instruction issue and L1/L2 latency long before HBM bandwidth.
#include "svo_kernels.cuh"

namespace svo {

namespace {

constexpr int kBatchThreads = 128;
constexpr int kCoarseThreads = 64;
constexpr int kTileThreads = 64;   // one 8x8 tile per block: two warps of 8x4 pixels

template <typename IdxT>
__device__ __forceinline__ SmemStack<IdxT> makeStack(unsigned char *smem, uint32_t slots) {
    SmemStack<IdxT> s;
    s.stride = blockDim.x;
    s.parent = reinterpret_cast<IdxT *>(smem) + threadIdx.x;
    s.maxT = reinterpret_cast<float *>(smem + size_t(slots)*blockDim.x*sizeof(IdxT)) + threadIdx.x;
    return s;
}

template <typename IdxT>
size_t stackBytes(uint32_t slots, int threads) {
    return size_t(slots)*size_t(threads)*(sizeof(IdxT) + sizeof(float));
}

// ---- K1 -------------------------------------------------------------------

template <bool FAST, bool LOD, typename IdxT>
__global__ void __launch_bounds__(kBatchThreads)
raymarchBatchKernel(const uint32_t *__restrict__ octree, uint64_t n, const float *__restrict__ o,
                    const float *__restrict__ d, float rayScale, uint32_t slots, uint8_t *__restrict__ hit,
                    float *__restrict__ t, uint32_t *__restrict__ normal, uint64_t *__restrict__ voxel) {
    extern __shared__ __align__(16) unsigned char smem[];
    SmemStack<IdxT> stack = makeStack<IdxT>(smem, slots);

    uint64_t i = uint64_t(blockIdx.x)*blockDim.x + threadIdx.x;
    if (i >= n) return;

    float ox = __ldg(o + 3*i), oy = __ldg(o + 3*i + 1), oz = __ldg(o + 3*i + 2);
    float dx = __ldg(d + 3*i), dy = __ldg(d + 3*i + 1), dz = __ldg(d + 3*i + 2);

    float tHit = kTreeMiss;
    uint32_t material = 0;
    uint64_t vox = ~uint64_t(0);
    int code = raymarch<FAST, LOD, IdxT>(octree, ox, oy, oz, dx, dy, dz, rayScale, stack, tHit, material, vox);

    if (hit) hit[i] = uint8_t(code);
    if (t) t[i] = tHit;
    if (normal) normal[i] = material;
    if (voxel) voxel[i] = vox;
}

// ---- K2 -------------------------------------------------------------------

template <bool FAST, typename IdxT>
__global__ void __launch_bounds__(kCoarseThreads)
coarsePassKernel(const uint32_t *__restrict__ octree, FramePlanDev plan, FrameConsts f, uint32_t slots,
                 float *__restrict__ depth) {
    extern __shared__ __align__(16) unsigned char smem[];
    SmemStack<IdxT> stack = makeStack<IdxT>(smem, slots);

    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= plan.totalCorners) return;

    int cellsFull = plan.tilesX*plan.tilesYFull;
    int strip = min(i/cellsFull, plan.nStrips - 1);
    int rem = i - strip*cellsFull;
    int y = rem/plan.tilesX;
    int x = rem - y*plan.tilesX;

    float dx = __ldg(plan.dxCoarse + x);
    float dy = __ldg(plan.dyCoarse + strip*plan.tilesYFull + y);
    float rx, ry, rz;
    rayDirection(f, dx, dy, rx, ry, rz);

    float tHit = kTreeMiss;
    uint32_t material;
    uint64_t vox;
    raymarch<FAST, true, IdxT>(octree, f.posX, f.posY, f.posZ, rx, ry, rz, f.coarseScale, stack, tHit, material, vox);
    depth[i] = tHit;   // stays 1e10f on a miss (Main.cpp:181-184)
}

// ---- K3 -------------------------------------------------------------------

struct TileCoords {
    int strip, tx, ty;
    int x0, y0;       // pixel origin
    int yEnd;         // strip end row (exclusive)
    int cornerIdx;    // index of corner (tx+1, ty+1) in the depth buffer
};

__device__ __forceinline__ TileCoords tileCoords(const FramePlanDev &plan, int tile) {
    TileCoords c;
    int tileRow = tile/plan.tileCols;
    c.tx = tile - tileRow*plan.tileCols;
    c.strip = min(tileRow/plan.tileRowsFull, plan.nStrips - 1);
    c.ty = tileRow - c.strip*plan.tileRowsFull;
    int stripY0 = c.strip*plan.stripRows;
    c.x0 = c.tx*8;
    c.y0 = stripY0 + c.ty*8;
    c.yEnd = min(stripY0 + plan.stripRows, plan.height);
    c.cornerIdx = c.strip*plan.tilesX*plan.tilesYFull + (c.ty + 1)*plan.tilesX + c.tx + 1;
    return c;
}

// min over the tile's four corner depths in the reference's association (Main.cpp:187-189)
__device__ __forceinline__ float tileMinDepth(const FramePlanDev &plan, const float *__restrict__ depth, int idx) {
    return minStd(minStd(__ldg(depth + idx), __ldg(depth + idx - 1)),
                  minStd(__ldg(depth + idx - plan.tilesX), __ldg(depth + idx - plan.tilesX - 1)));
}

template <bool FAST, typename IdxT>
__global__ void __launch_bounds__(kTileThreads)
finePassKernel(const uint32_t *__restrict__ octree, FramePlanDev plan, FrameConsts f, uint32_t slots,
               const float *__restrict__ depth, uint32_t *__restrict__ rgba, int tileRank, int tileWorld) {
    extern __shared__ __align__(16) unsigned char smem[];
    SmemStack<IdxT> stack = makeStack<IdxT>(smem, slots);

    int tile = blockIdx.x*tileWorld + tileRank;
    if (tile >= plan.totalTiles) return;
    TileCoords c = tileCoords(plan, tile);

    int px = c.x0 + (threadIdx.x & 7);
    int py = c.y0 + (threadIdx.x >> 3);
    if (px >= plan.width || py >= c.yEnd) return;
    uint32_t *dst = rgba + size_t(py)*size_t(plan.width) + px;

    float minT = tileMinDepth(plan, depth, c.cornerIdx);
    if (minT == kTreeMiss) {       // Main.cpp:191: tile skipped, pixels keep the memset's 0 (Main.cpp:165)
        *dst = 0u;
        return;
    }
    float startT = maxStd(subRn(minT, f.beamBias), 0.0f);   // Main.cpp:197

    float rx, ry, rz;
    rayDirection(f, __ldg(plan.dxFine + px), __ldg(plan.dyFine + py), rx, ry, rz);
    float ox = addRn(f.posX, mulRn(rx, startT));            // pos + dir*minT, Main.cpp:118
    float oy = addRn(f.posY, mulRn(ry, startT));
    float oz = addRn(f.posZ, mulRn(rz, startT));

    float tHit;
    uint32_t material = 0;
    uint64_t vox;
    int code = raymarch<FAST, false, IdxT>(octree, ox, oy, oz, rx, ry, rz, 0.0f, stack, tHit, material, vox);
    uint32_t colour = 0xFF000000u;                          // Vec3() -> black, Main.cpp:117
    if (code != kMiss) colour = packGrey(shadeMaterial(material, rx, ry, rz, f.lightX, f.lightY, f.lightZ));
    *dst = colour;
}

// ---- K4 -------------------------------------------------------------------

__global__ void __launch_bounds__(256)
tileStatsKernel(FramePlanDev plan, const float *__restrict__ depth, int tileRank, int tileWorld, int ownedTiles,
                FrameCounters *counters) {
    int k = blockIdx.x*blockDim.x + threadIdx.x;
    unsigned int pixels = 0, rendered = 0;
    if (k < ownedTiles) {
        TileCoords c = tileCoords(plan, k*tileWorld + tileRank);
        if (tileMinDepth(plan, depth, c.cornerIdx) != kTreeMiss) {
            int w = min(c.x0 + 8, plan.width) - c.x0;
            int h = min(c.y0 + 8, c.yEnd) - c.y0;
            pixels = w*h;
            rendered = 1;
        }
    }
    pixels = __reduce_add_sync(0xffffffffu, pixels);
    rendered = __reduce_add_sync(0xffffffffu, rendered);
    if ((threadIdx.x & 31) == 0 && rendered) {
        atomicAdd(&counters->fineRays, (unsigned long long)pixels);
        atomicAdd(&counters->tilesRendered, (unsigned long long)rendered);
    }
}

inline uint32_t stackSlots(const TreeDev &tree) { return tree.depth > 1 ? tree.depth - 1 : 1; }

template <typename K>
cudaError_t ensureSmem(K kernel, size_t bytes) {
    if (bytes <= 48*1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
}

template <bool FAST, bool LOD, typename IdxT>
cudaError_t launchBatchT(const TreeDev &tree, uint64_t n, const float *o, const float *d, float rayScale,
                         uint8_t *hit, float *t, uint32_t *normal, uint64_t *voxel, cudaStream_t stream) {
    uint32_t slots = stackSlots(tree);
    size_t smem = stackBytes<IdxT>(slots, kBatchThreads);
    auto kernel = raymarchBatchKernel<FAST, LOD, IdxT>;
    cudaError_t e = ensureSmem(kernel, smem);
    if (e != cudaSuccess) return e;
    uint64_t blocks = (n + kBatchThreads - 1)/kBatchThreads;
    if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    kernel<<<unsigned(blocks), kBatchThreads, smem, stream>>>(tree.words, n, o, d, rayScale, slots, hit, t, normal, voxel);
    return cudaGetLastError();
}

template <bool FAST, typename IdxT>
cudaError_t launchCoarseT(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts, float *depth,
                          cudaStream_t stream) {
    uint32_t slots = stackSlots(tree);
    size_t smem = stackBytes<IdxT>(slots, kCoarseThreads);
    auto kernel = coarsePassKernel<FAST, IdxT>;
    cudaError_t e = ensureSmem(kernel, smem);
    if (e != cudaSuccess) return e;
    int blocks = (plan.totalCorners + kCoarseThreads - 1)/kCoarseThreads;
    kernel<<<blocks, kCoarseThreads, smem, stream>>>(tree.words, plan, consts, slots, depth);
    return cudaGetLastError();
}

template <bool FAST, typename IdxT>
cudaError_t launchFineT(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts, const float *depth,
                        uint32_t *rgba, int tileRank, int tileWorld, cudaStream_t stream) {
    uint32_t slots = stackSlots(tree);
    size_t smem = stackBytes<IdxT>(slots, kTileThreads);
    auto kernel = finePassKernel<FAST, IdxT>;
    cudaError_t e = ensureSmem(kernel, smem);
    if (e != cudaSuccess) return e;
    int owned = (plan.totalTiles - tileRank + tileWorld - 1)/tileWorld;
    if (owned <= 0) return cudaSuccess;
    kernel<<<owned, kTileThreads, smem, stream>>>(tree.words, plan, consts, slots, depth, rgba, tileRank, tileWorld);
    return cudaGetLastError();
}

} // namespace

cudaError_t launchRaymarchBatch(const TreeDev &tree, uint64_t n, const float *o, const float *d, float rayScale,
                                int flavour, uint8_t *hit, float *t, uint32_t *normal, uint64_t *voxel,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    bool wide = tree.nWords >= (1ull << 32);
    bool lod = rayScale != 0.0f;
    bool fast = flavour != 0;
#define SVO_BATCH(F, L) \
    (wide ? launchBatchT<F, L, uint64_t>(tree, n, o, d, rayScale, hit, t, normal, voxel, stream) \
          : launchBatchT<F, L, uint32_t>(tree, n, o, d, rayScale, hit, t, normal, voxel, stream))
    if (fast) return lod ? SVO_BATCH(true, true) : SVO_BATCH(true, false);
    return lod ? SVO_BATCH(false, true) : SVO_BATCH(false, false);
#undef SVO_BATCH
}

cudaError_t launchCoarsePass(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts, int flavour,
                             float *depth, cudaStream_t stream) {
    // The beam pass is individually rounded in BOTH flavours: its depths feed
    // `minT - 0.03` (Main.cpp:197) and the tile-skip test (Main.cpp:191), so a
    // last-bit change here moves every ray origin of a tile or drops / adds a
    // whole tile. It is < 5 % of the rays; FAST only changes the fine pass.
    (void)flavour;
    bool wide = tree.nWords >= (1ull << 32);
    return wide ? launchCoarseT<false, uint64_t>(tree, plan, consts, depth, stream)
                : launchCoarseT<false, uint32_t>(tree, plan, consts, depth, stream);
}

cudaError_t launchFinePass(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts, int flavour,
                           const float *depth, uint32_t *rgba, int tileRank, int tileWorld, cudaStream_t stream) {
    bool wide = tree.nWords >= (1ull << 32);
    if (flavour != 0)
        return wide ? launchFineT<true, uint64_t>(tree, plan, consts, depth, rgba, tileRank, tileWorld, stream)
                    : launchFineT<true, uint32_t>(tree, plan, consts, depth, rgba, tileRank, tileWorld, stream);
    return wide ? launchFineT<false, uint64_t>(tree, plan, consts, depth, rgba, tileRank, tileWorld, stream)
                : launchFineT<false, uint32_t>(tree, plan, consts, depth, rgba, tileRank, tileWorld, stream);
}

cudaError_t launchTileStats(const FramePlanDev &plan, const float *depth, int tileRank, int tileWorld,
                            FrameCounters *counters, cudaStream_t stream) {
    int owned = (plan.totalTiles - tileRank + tileWorld - 1)/tileWorld;
    if (owned <= 0) return cudaSuccess;
    tileStatsKernel<<<(owned + 255)/256, 256, 0, stream>>>(plan, depth, tileRank, tileWorld, owned, counters);
    return cudaGetLastError();
}

} // namespace svo
