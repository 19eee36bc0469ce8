// sm_100a kernels for the reference's ray-casting hot path:
//   K1 raymarchBatchKernel  VoxelOctree::raymarch in a loop     (reference src/VoxelOctree.cpp:207-346)
//   K2 coarsePassKernel     renderBatch's tile-corner beam pass (reference src/Main.cpp:167-184)
//   K2b classifyTilesKernel 4-corner min, tile skip test, the strip memset for skipped tiles,
//                           compact list of tiles to render      (reference src/Main.cpp:165, 186-197)
//   K3 finePassKernel       renderTile + shade + pack, fused     (reference src/Main.cpp:92-137, 81-90)
// No tensor cores: this is pointer chasing over a uint32 node array, bounded by
// instruction issue and SIMT divergence long before HBM bandwidth.
#include "svo_kernels.cuh"

#include <atomic>

#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>

namespace svo {

namespace {

constexpr int kBatchThreads = 128;
constexpr int kCoarseThreads = 64;
constexpr int kTileThreads = 64;   // an 8x8 tile = two warps of 8x4 pixels
#ifndef SVO_TILES_PER_BLOCK
#define SVO_TILES_PER_BLOCK 1
#endif
constexpr unsigned kTilesPerBlock = SVO_TILES_PER_BLOCK;
constexpr int kClassifyThreads = 128;

// ---- K1 -------------------------------------------------------------------

template <bool FAST, bool LOD, typename IdxT>
__global__ void __launch_bounds__(kBatchThreads, 16)
raymarchBatchKernel(const uint32_t *__restrict__ octree, uint64_t n, const float *__restrict__ o,
                    const float *__restrict__ d, float rayScale, uint8_t *__restrict__ hit,
                    float *__restrict__ t, uint32_t *__restrict__ normal, uint64_t *__restrict__ voxel,
                    const uint32_t *__restrict__ order) {
    extern __shared__ __align__(16) unsigned char smem[];
    SmemStack<IdxT, kBatchThreads> stack;
    stack.init(smem);

    uint64_t i = uint64_t(blockIdx.x)*blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (order) i = __ldg(order + i);       // coherence order: thread k traces ray order[k], results land at order[k]

    float ox = __ldg(o + 3*i), oy = __ldg(o + 3*i + 1), oz = __ldg(o + 3*i + 2);
    float dx = __ldg(d + 3*i), dy = __ldg(d + 3*i + 1), dz = __ldg(d + 3*i + 2);

    // A ray with a NaN or infinite component never leaves the loop (every comparison is false, so no axis ever
    // steps; the reference spins the same way, VoxelOctree.cpp:252-339, but on a CPU thread, not on a GPU that the
    // caller then cannot get back). Such a ray is reported as a miss: it is swapped for one that starts behind the
    // volume and leaves it in two trips, so the warp stays whole for the traversal's barrier.
    const float magnitude = fabsf(ox) + fabsf(oy) + fabsf(oz) + fabsf(dx) + fabsf(dy) + fabsf(dz);
    const bool finite = magnitude < __int_as_float(0x7f800000);
    if (!finite) { ox = oy = oz = 3.0f; dx = dy = dz = 1.0f; }

    float tHit = kTreeMiss;
    uint64_t vox;
    int code = raymarch<FAST, LOD, IdxT, kBatchThreads>(octree, ox, oy, oz, dx, dy, dz, rayScale, stack, tHit, vox);
    if (!finite) { code = kMiss; tHit = kTreeMiss; }
    if (code == kMiss) vox = ~uint64_t(0);
    uint32_t material = code == kHitLeaf ? ldNode(octree + vox) : 0u;   // VoxelOctree.cpp:282

    if (hit) hit[i] = uint8_t(code);
    if (t) t[i] = tHit;
    if (normal) normal[i] = material;
    if (voxel) voxel[i] = vox;
}

// shade + pack for arbitrary rays (Main.cpp:81-90, 128-132)
__global__ void __launch_bounds__(kBatchThreads)
shadeBatchKernel(uint64_t n, const uint8_t *__restrict__ hit, const uint32_t *__restrict__ normal,
                 const float *__restrict__ d, float lx, float ly, float lz, uint32_t *__restrict__ rgba) {
    uint64_t i = uint64_t(blockIdx.x)*blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t colour = 0xFF000000u;
    if (!hit || hit[i] != kMiss)
        colour = packGrey(shadeMaterial(__ldg(normal + i), __ldg(d + 3*i), __ldg(d + 3*i + 1), __ldg(d + 3*i + 2), lx, ly, lz));
    rgba[i] = colour;
}

// ---- K2 -------------------------------------------------------------------

// Tile (tx, ty) belongs to rank tx % world (vertical stripes one tile wide), so a rank needs only
// the corner columns cx with cx % world == rank or (cx - 1) % world == rank: 2/world of the beam
// pass for world >= 3. Threads enumerate exactly those corners, densely: column slot j of a corner
// row maps to cx = (j >> 1)*world + rank + (j & 1).
template <typename IdxT>
__global__ void __launch_bounds__(kCoarseThreads)
coarsePassKernel(const uint32_t *__restrict__ octree, FramePlanDev plan, FrameConsts f, float *__restrict__ depth,
                 FrameCounters *__restrict__ counters, int tileRank, int tileWorld, int tileRun, int colSlots,
                 int totalSlots) {
    extern __shared__ __align__(16) unsigned char smem[];
    SmemStack<IdxT, kCoarseThreads> stack;
    stack.init(smem);

    int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k == 0) {
        counters->tilesRendered = 0;
        counters->fineRays = 0;
    }
    // A warp takes an 8 x 4 block of this rank's corner grid (column slots x corner rows of one
    // strip) rather than 32 corners of one row: neighbouring beams walk the same nodes.
    const int lane = threadIdx.x & 31;
    const int warp = k >> 5;
    const int groupsX = (colSlots + 7) >> 3;
    const int groupsPerStrip = groupsX*((plan.tilesYFull + 3) >> 2);
    const int strip = warp/groupsPerStrip;
    if (strip >= plan.nStrips) return;
    const int g = warp - strip*groupsPerStrip;
    const int gy = g/groupsX;
    const int j = (g - gy*groupsX)*8 + (lane & 7);
    const int y = gy*4 + (lane >> 3);
    if (j >= colSlots || y >= (strip == plan.nStrips - 1 ? plan.tilesYLast : plan.tilesYFull)) return;
    // slot j -> corner column: run q = j / (run + 1) of this rank starts at tile column (q*world + rank)*run
    int q = j/(tileRun + 1);
    int x = tileWorld == 1 ? j : (q*tileWorld + tileRank)*tileRun + (j - q*(tileRun + 1));
    if (x >= plan.tilesX) return;
    int i = strip*plan.tilesX*plan.tilesYFull + y*plan.tilesX + x;

    float dx = __ldg(plan.dxCoarse + x);
    float dy = __ldg(plan.dyCoarse + strip*plan.tilesYFull + y);
    float rx, ry, rz;
    rayDirection(f, dx, dy, rx, ry, rz);

    float tHit = kTreeMiss;
    uint64_t vox;
    // individually rounded in both flavours, see launchCoarsePass
    raymarch<false, true, IdxT, kCoarseThreads>(octree, f.posX, f.posY, f.posZ, rx, ry, rz, f.coarseScale, stack, tHit, vox);
    depth[i] = tHit;   // stays 1e10f on a miss (Main.cpp:181-184)
}

// ---- K2b ------------------------------------------------------------------

__global__ void __launch_bounds__(kClassifyThreads)
classifyTilesKernel(FramePlanDev plan, float beamBias, const float *__restrict__ depth, uint32_t *__restrict__ rgba,
                    int tileRank, int tileWorld, int tileRun, int ownedCols, int ownedTiles, int pixelStride,
                    TileRecord *__restrict__ tiles, FrameCounters *__restrict__ counters,
                    unsigned long long *__restrict__ fineRaysTotal) {
    int k = blockIdx.x*blockDim.x + threadIdx.x;
    bool active = k < ownedTiles;
    bool rendered = false;
    int x0 = 0, y0 = 0, yEnd = 0;
    float minT = kTreeMiss;
    if (active) {
        int tileRow = k/ownedCols;                              // this rank owns tile columns (tx / run) % world == rank
        int c = k - tileRow*ownedCols;                          // c-th owned column
        int tx = ((c/tileRun)*tileWorld + tileRank)*tileRun + c%tileRun;
        int strip = min(tileRow/plan.tileRowsFull, plan.nStrips - 1);
        int ty = tileRow - strip*plan.tileRowsFull;
        int stripY0 = strip*plan.stripRows;
        x0 = tx*8;
        y0 = stripY0 + ty*8;
        yEnd = min(stripY0 + plan.stripRows, plan.height);
        // corner (tx+1, ty+1) and its three neighbours, in the reference's association (Main.cpp:187-189)
        int idx = strip*plan.tilesX*plan.tilesYFull + (ty + 1)*plan.tilesX + tx + 1;
        minT = minStd(minStd(__ldg(depth + idx), __ldg(depth + idx - 1)),
                      minStd(__ldg(depth + idx - plan.tilesX), __ldg(depth + idx - plan.tilesX - 1)));
        rendered = minT != kTreeMiss;                           // Main.cpp:191
    }

    // warp-aggregated append: one atomic per warp
    unsigned lane = threadIdx.x & 31u;
    unsigned mask = __ballot_sync(0xffffffffu, rendered);
    int w = min(x0 + 8, plan.width) - x0;
    int h = min(y0 + 8, yEnd) - y0;
    // fine rays of the tile: every pixelStride-th pixel of every pixelStride-th row (Main.cpp:101-106)
    const int raysX = (w + pixelStride - 1)/pixelStride, raysY = (h + pixelStride - 1)/pixelStride;
    unsigned pixels = __reduce_add_sync(0xffffffffu, rendered ? unsigned(raysX*raysY) : 0u);
    unsigned base = 0;
    if (lane == 0 && mask) {
        base = atomicAdd(&counters->tilesRendered, unsigned(__popc(mask)));
        atomicAdd(&counters->fineRays, (unsigned long long)pixels);
        if (fineRaysTotal) atomicAdd(fineRaysTotal, (unsigned long long)pixels);   // running total over a frame sequence
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (rendered) {
        TileRecord r;
        r.xy = uint32_t(x0) | (uint32_t(y0) << 16);
        r.yEnd = uint32_t(yEnd);
        r.startT = maxStd(subRn(minT, beamBias), 0.0f);        // Main.cpp:197
        r.pad = 0;
        tiles[base + __popc(mask & ((1u << lane) - 1u))] = r;
    } else if (active) {
        // skipped tile: its pixels keep the strip memset's zero (Main.cpp:165)
        if (w == 8 && (plan.width & 3) == 0 && (reinterpret_cast<uintptr_t>(rgba) & 15) == 0) {
            for (int r = 0; r < h; ++r) {
                uint4 *row = reinterpret_cast<uint4 *>(rgba + size_t(y0 + r)*size_t(plan.width) + x0);
                row[0] = make_uint4(0, 0, 0, 0);
                row[1] = make_uint4(0, 0, 0, 0);
            }
        } else {
            for (int r = 0; r < h; ++r)
                for (int c = 0; c < w; ++c) rgba[size_t(y0 + r)*size_t(plan.width) + x0 + c] = 0u;
        }
    }
}

// ---- K3 -------------------------------------------------------------------

template <bool FAST, typename IdxT>
__global__ void __launch_bounds__(kTileThreads, 32)
finePassKernel(const uint32_t *__restrict__ octree, FramePlanDev plan, FrameConsts f,
               const TileRecord *__restrict__ tiles, const FrameCounters *__restrict__ counters,
               uint32_t *__restrict__ rgba) {
    extern __shared__ __align__(16) unsigned char smem[];
    SmemStack<IdxT, kTileThreads> stack;
    stack.init(smem);

    // kTilesPerBlock consecutive list entries per block: each warp renders the same 8x4 half of each of
    // them in turn, which evens out the two warps' lifetimes (a block's warp slots are only released
    // when its slowest warp is done) and amortises the block launch.
    const unsigned nTiles = counters->tilesRendered;
#pragma unroll 1
    for (unsigned t = 0; t < kTilesPerBlock; ++t) {
        const unsigned tile = blockIdx.x*kTilesPerBlock + t;
        if (tile >= nTiles) return;
        const uint4 rec = __ldg(reinterpret_cast<const uint4 *>(tiles) + tile);
        const int px = int(rec.x & 0xFFFFu) + int(threadIdx.x & 7);
        const int py = int(rec.x >> 16) + int(threadIdx.x >> 3);
        if (px >= plan.width || py >= int(rec.y)) continue;
        const float startT = __uint_as_float(rec.z);

        float rx, ry, rz;
        rayDirection(f, __ldg(plan.dxFine + px), __ldg(plan.dyFine + py), rx, ry, rz);
        const float ox = addRn(f.posX, mulRn(rx, startT));      // pos + dir*minT, Main.cpp:118
        const float oy = addRn(f.posY, mulRn(ry, startT));
        const float oz = addRn(f.posZ, mulRn(rz, startT));

        float tHit;
        uint64_t vox;
        const int code = raymarch<FAST, false, IdxT, kTileThreads>(octree, ox, oy, oz, rx, ry, rz, 0.0f, stack, tHit, vox);
        // the warp is convergent again here: one material fetch and one pass through the shading code
        // for all its hits together (misses read some descriptor and discard the result)
        const uint32_t material = ldNode(octree + IdxT(vox));   // VoxelOctree.cpp:282
        uint32_t colour = packGrey(shadeMaterial(material, rx, ry, rz, f.lightX, f.lightY, f.lightZ));
        if (code == kMiss) colour = 0xFF000000u;                // Vec3() -> black, Main.cpp:117
        rgba[size_t(py)*size_t(plan.width) + px] = colour;
    }
}

// renderTile with stride > 1 (the reference's renderHalfSize preview: stride 3, Main.cpp:101-106,161):
// only pixels whose offsets inside the tile are multiples of the stride are traced; every other pixel
// copies its block's corner pixel. One block per tile: the corner colours go through shared memory. Every
// thread runs the traversal (its __syncwarp needs whole warps) -- the ones without a pixel of their own
// on a ray that starts behind the volume and leaves it in two trips.
template <bool FAST, typename IdxT>
__global__ void __launch_bounds__(kTileThreads)
finePassStridedKernel(const uint32_t *__restrict__ octree, FramePlanDev plan, FrameConsts f,
                      const TileRecord *__restrict__ tiles, const FrameCounters *__restrict__ counters,
                      uint32_t *__restrict__ rgba, int pixelStride) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t corner[64];
    SmemStack<IdxT, kTileThreads> stack;
    stack.init(smem);
    const unsigned tile = blockIdx.x;
    if (tile >= counters->tilesRendered) return;
    const uint4 rec = __ldg(reinterpret_cast<const uint4 *>(tiles) + tile);
    const int lx = int(threadIdx.x & 7), ly = int(threadIdx.x >> 3);
    const int px = int(rec.x & 0xFFFFu) + lx, py = int(rec.x >> 16) + ly;
    const bool inside = px < plan.width && py < int(rec.y);
    const bool traces = inside && lx % pixelStride == 0 && ly % pixelStride == 0;
    const float startT = __uint_as_float(rec.z);

    float rx = 1.0f, ry = 1.0f, rz = 1.0f, ox = 3.0f, oy = 3.0f, oz = 3.0f;
    if (traces) {
        rayDirection(f, __ldg(plan.dxFine + px), __ldg(plan.dyFine + py), rx, ry, rz);
        ox = addRn(f.posX, mulRn(rx, startT));
        oy = addRn(f.posY, mulRn(ry, startT));
        oz = addRn(f.posZ, mulRn(rz, startT));
    }
    float tHit;
    uint64_t vox;
    const int code = raymarch<FAST, false, IdxT, kTileThreads>(octree, ox, oy, oz, rx, ry, rz, 0.0f, stack, tHit, vox);
    const uint32_t material = ldNode(octree + IdxT(vox));
    uint32_t colour = packGrey(shadeMaterial(material, rx, ry, rz, f.lightX, f.lightY, f.lightZ));
    if (code == kMiss) colour = 0xFF000000u;
    if (traces) corner[threadIdx.x] = colour;
    __syncthreads();
    if (inside) rgba[size_t(py)*size_t(plan.width) + px] = corner[(ly - ly % pixelStride)*8 + (lx - lx % pixelStride)];
}

// 64-bit word indices are needed from 2^32 words (16 GiB) on; SVO_FORCE_WIDE_INDEX=1 selects that
// instantiation for any tree so that tests can cover it.
inline bool wideIndex(const TreeDev &tree) {
    static const bool force = [] {
        const char *e = getenv("SVO_FORCE_WIDE_INDEX");
        return e && e[0] == '1';
    }();
    return force || tree.nWords >= (1ull << 32);
}

inline uint32_t stackSlots(const TreeDev &tree) { return tree.depth > 1 ? tree.depth - 1 : 1; }

template <typename K>
cudaError_t ensureSmem(K kernel, size_t bytes) {
    if (bytes <= 48*1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
}

} // namespace

// Tile column tx belongs to rank (tx / run) % world: vertical stripes `run` tiles wide.
namespace {
std::atomic<int> &tileRunSetting() {
    static std::atomic<int> run([] {
        const char *e = getenv("SVO_TILE_RUN");
        int v = e ? atoi(e) : 0;
        return v > 0 ? v : 4;
    }());
    return run;
}
} // namespace

int tileRunLength(int tileWorld) { return tileWorld > 1 ? tileRunSetting().load(std::memory_order_relaxed) : 1; }

void setTileRunLength(int run) { tileRunSetting().store(run > 0 ? run : 4, std::memory_order_relaxed); }

int ownedTileColumns(int tileCols, int tileRank, int tileWorld) {
    int run = tileRunLength(tileWorld), n = 0;
    for (int start = tileRank*run; start < tileCols; start += tileWorld*run)
        n += (start + run <= tileCols) ? run : tileCols - start;
    return n;
}

namespace {

inline int ownedCols(const FramePlanDev &plan, int tileRank, int tileWorld) {
    return ownedTileColumns(plan.tileCols, tileRank, tileWorld);
}
inline int ownedTiles(const FramePlanDev &plan, int tileRank, int tileWorld) {
    return ownedCols(plan, tileRank, tileWorld)*plan.totalTileRows;
}

template <bool FAST, bool LOD, typename IdxT>
cudaError_t launchBatchT(const TreeDev &tree, uint64_t n, const float *o, const float *d, float rayScale,
                         uint8_t *hit, float *t, uint32_t *normal, uint64_t *voxel, const uint32_t *order,
                         cudaStream_t stream) {
    size_t smem = SmemStack<IdxT, kBatchThreads>::bytes(stackSlots(tree));
    auto kernel = raymarchBatchKernel<FAST, LOD, IdxT>;
    cudaError_t e = ensureSmem(kernel, smem);
    if (e != cudaSuccess) return e;
    uint64_t blocks = (n + kBatchThreads - 1)/kBatchThreads;
    if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    kernel<<<unsigned(blocks), kBatchThreads, smem, stream>>>(tree.words, n, o, d, rayScale, hit, t, normal, voxel, order);
    return cudaGetLastError();
}

template <typename IdxT>
cudaError_t launchCoarseT(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts, float *depth,
                          FrameCounters *counters, int tileRank, int tileWorld, cudaStream_t stream) {
    size_t smem = SmemStack<IdxT, kCoarseThreads>::bytes(stackSlots(tree));
    auto kernel = coarsePassKernel<IdxT>;
    cudaError_t e = ensureSmem(kernel, smem);
    if (e != cudaSuccess) return e;
    // a rank needs the corner columns on both sides of each of its runs: run + 1 per run
    int run = tileRunLength(tileWorld);
    int runsOwned = (plan.tileCols + tileWorld*run - 1)/(tileWorld*run);
    int colSlots = tileWorld == 1 ? plan.tilesX : runsOwned*(run + 1);
    int groupsPerStrip = ((colSlots + 7)/8)*((plan.tilesYFull + 3)/4);      // 8 x 4 corners per warp
    int totalSlots = groupsPerStrip*plan.nStrips*32;
    int blocks = (totalSlots + kCoarseThreads - 1)/kCoarseThreads;
    kernel<<<blocks, kCoarseThreads, smem, stream>>>(tree.words, plan, consts, depth, counters, tileRank, tileWorld,
                                                     run, colSlots, totalSlots);
    return cudaGetLastError();
}

template <bool FAST, typename IdxT>
cudaError_t launchFineT(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts,
                        const TileRecord *tiles, const FrameCounters *counters, uint32_t *rgba, int owned,
                        int pixelStride, cudaStream_t stream) {
    size_t smem = SmemStack<IdxT, kTileThreads>::bytes(stackSlots(tree));
    if (pixelStride > 1) {
        auto strided = finePassStridedKernel<FAST, IdxT>;
        cudaError_t es = ensureSmem(strided, smem);
        if (es != cudaSuccess) return es;
        strided<<<owned, kTileThreads, smem, stream>>>(tree.words, plan, consts, tiles, counters, rgba, pixelStride);
        return cudaGetLastError();
    }
    auto kernel = finePassKernel<FAST, IdxT>;
    cudaError_t e = ensureSmem(kernel, smem);
    if (e != cudaSuccess) return e;
    int blocks = (owned + int(kTilesPerBlock) - 1)/int(kTilesPerBlock);
    kernel<<<blocks, kTileThreads, smem, stream>>>(tree.words, plan, consts, tiles, counters, rgba);
    return cudaGetLastError();
}

} // namespace

cudaError_t launchRaymarchBatch(const TreeDev &tree, uint64_t n, const float *o, const float *d, float rayScale,
                                int flavour, uint8_t *hit, float *t, uint32_t *normal, uint64_t *voxel,
                                const uint32_t *order, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    bool wide = wideIndex(tree);
    bool lod = rayScale != 0.0f;
    bool fast = flavour != 0;
#define SVO_BATCH(F, L) \
    (wide ? launchBatchT<F, L, uint64_t>(tree, n, o, d, rayScale, hit, t, normal, voxel, order, stream) \
          : launchBatchT<F, L, uint32_t>(tree, n, o, d, rayScale, hit, t, normal, voxel, order, stream))
    if (fast) return lod ? SVO_BATCH(true, true) : SVO_BATCH(true, false);
    return lod ? SVO_BATCH(false, true) : SVO_BATCH(false, false);
#undef SVO_BATCH
}

// ---- coherence order for incoherent ray batches -------------------------------------------------
// Rays are binned by direction (4 x 4 x 4 cells of d / max|d|) with ONE stable 6-bit radix pass, so rays
// of a bin keep their submission order (neighbouring pixels stay neighbours). A warp then holds rays that
// share their octant mirroring and roughly their direction: on 16-spp ambient-occlusion rays the
// trace-driven model gives 1.65x fewer warp instructions (SIMT efficiency 0.21 -> 0.34).
__global__ void __launch_bounds__(256)
directionBinKernel(uint64_t n, const float *__restrict__ d, uint32_t *__restrict__ keys, uint32_t *__restrict__ index) {
    uint64_t i = uint64_t(blockIdx.x)*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = __ldg(d + 3*i), y = __ldg(d + 3*i + 1), z = __ldg(d + 3*i + 2);
    const float m = fmaxf(fmaxf(fabsf(x), fabsf(y)), fmaxf(fabsf(z), 1e-30f));
    auto cell = [m](float v) { return uint32_t(min(3, max(0, int((v/m + 1.0f)*2.0f)))); };
    keys[i] = cell(x) | (cell(y) << 2) | (cell(z) << 4);
    index[i] = uint32_t(i);
}

size_t coherenceOrderBytes(uint64_t n) {
    size_t temp = 0;
    cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, temp, k, v, int(n), 0, 6);
    return size_t(n)*16 + ((temp + 255) & ~size_t(255)) + 256;
}

// workspace: coherenceOrderBytes(n) bytes on the device. *orderOut points into it.
cudaError_t buildCoherenceOrder(uint64_t n, const float *d, void *workspace, const uint32_t **orderOut, cudaStream_t stream) {
    if (n >= (1ull << 31)) return cudaErrorInvalidValue;
    uint32_t *base = static_cast<uint32_t *>(workspace);
    uint32_t *keys0 = base, *keys1 = base + n, *idx0 = base + 2*n, *idx1 = base + 3*n;
    void *temp = base + 4*n;
    size_t tempBytes = 0;
    cub::DoubleBuffer<uint32_t> k(keys0, keys1), v(idx0, idx1);
    cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, k, v, int(n), 0, 6, stream);
    directionBinKernel<<<unsigned((n + 255)/256), 256, 0, stream>>>(n, d, keys0, idx0);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, tempBytes, k, v, int(n), 0, 6, stream);
    if (e != cudaSuccess) return e;
    *orderOut = v.Current();
    return cudaGetLastError();
}

cudaError_t launchShadeBatch(uint64_t n, const uint8_t *hit, const uint32_t *normal, const float *d, const float light[3],
                             uint32_t *rgba, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    uint64_t blocks = (n + kBatchThreads - 1)/kBatchThreads;
    if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    shadeBatchKernel<<<unsigned(blocks), kBatchThreads, 0, stream>>>(n, hit, normal, d, light[0], light[1], light[2], rgba);
    return cudaGetLastError();
}

cudaError_t launchCoarsePass(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts, int flavour,
                             float *depth, FrameCounters *counters, int tileRank, int tileWorld,
                             cudaStream_t stream) {
    // The beam pass is individually rounded in BOTH flavours: its depths feed
    // `minT - 0.03` (Main.cpp:197) and the tile-skip test (Main.cpp:191), so a
    // last-bit change here moves every ray origin of a tile or drops / adds a
    // whole tile. It is < 5 % of the rays; FAST only changes the fine pass.
    (void)flavour;
    bool wide = wideIndex(tree);
    return wide ? launchCoarseT<uint64_t>(tree, plan, consts, depth, counters, tileRank, tileWorld, stream)
                : launchCoarseT<uint32_t>(tree, plan, consts, depth, counters, tileRank, tileWorld, stream);
}

cudaError_t launchClassifyTiles(const FramePlanDev &plan, const FrameConsts &consts, const float *depth,
                                uint32_t *rgba, int tileRank, int tileWorld, int pixelStride, TileRecord *tiles,
                                FrameCounters *counters, unsigned long long *fineRaysTotal, cudaStream_t stream) {
    int owned = ownedTiles(plan, tileRank, tileWorld);
    if (owned <= 0) return cudaSuccess;
    classifyTilesKernel<<<(owned + kClassifyThreads - 1)/kClassifyThreads, kClassifyThreads, 0, stream>>>(
        plan, consts.beamBias, depth, rgba, tileRank, tileWorld, tileRunLength(tileWorld),
        ownedCols(plan, tileRank, tileWorld), owned, pixelStride > 1 ? pixelStride : 1, tiles, counters, fineRaysTotal);
    return cudaGetLastError();
}

cudaError_t launchFinePass(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts, int flavour,
                           const TileRecord *tiles, const FrameCounters *counters, uint32_t *rgba,
                           int tileRank, int tileWorld, int pixelStride, cudaStream_t stream) {
    int owned = ownedTiles(plan, tileRank, tileWorld);
    if (owned <= 0) return cudaSuccess;
    bool wide = wideIndex(tree);
    if (flavour != 0)
        return wide ? launchFineT<true, uint64_t>(tree, plan, consts, tiles, counters, rgba, owned, pixelStride, stream)
                    : launchFineT<true, uint32_t>(tree, plan, consts, tiles, counters, rgba, owned, pixelStride, stream);
    return wide ? launchFineT<false, uint64_t>(tree, plan, consts, tiles, counters, rgba, owned, pixelStride, stream)
                : launchFineT<false, uint32_t>(tree, plan, consts, tiles, counters, rgba, owned, pixelStride, stream);
}

// The pixels of the tile columns a rank owns, from one framebuffer to another of the same pitch: thread k of a
// row moves pixel k of the rank's columns laid side by side, so a warp moves one run of four tile columns =
// 32 pixels = 128 contiguous bytes on both sides -- one full-width PCIe write when `dst` is mapped host memory.
__global__ void __launch_bounds__(256)
copyOwnedColumnsKernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int width, int height,
                       int ownedPixels, int run, int tileRank, int tileWorld) {
    const uint32_t k = blockIdx.x*256u + threadIdx.x;
    const int y = int(blockIdx.y);
    if (k >= uint32_t(ownedPixels) || y >= height) return;
    const uint32_t slot = k >> 3;                        // which of the rank's tile columns
    const uint32_t tx = (slot/uint32_t(run))*uint32_t(tileWorld*run) + uint32_t(tileRank*run) + slot%uint32_t(run);
    const uint32_t x = tx*8u + (k & 7u);
    if (x >= uint32_t(width)) return;
    const size_t at = size_t(y)*size_t(width) + x;
    dst[at] = src[at];
}

cudaError_t launchCopyOwnedColumns(const FramePlanDev &plan, int width, int height, const uint32_t *src, uint32_t *dst,
                                   int tileRank, int tileWorld, cudaStream_t stream) {
    const int ownedPixels = ownedCols(plan, tileRank, tileWorld)*8;
    if (ownedPixels <= 0) return cudaSuccess;
    dim3 grid(unsigned((ownedPixels + 255)/256), unsigned(height));
    copyOwnedColumnsKernel<<<grid, 256, 0, stream>>>(src, dst, width, height, ownedPixels, tileRunLength(tileWorld),
                                                     tileRank, tileWorld);
    return cudaGetLastError();
}

} // namespace svo
