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
    SmemStack<IdxT, kBatchThreads, LOD> stack;
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

// K1 with lane refill: persistent warps pull rays from a cursor (one atomic per kRefillChunk rays and warp), every
// lane keeps its traversal state in registers, and as soon as kRefillIdle lanes of a warp have finished their rays
// those lanes write their results and take the next rays of the warp's chunk -- the other lanes' traversals are not
// disturbed. On incoherent batches (ambient-occlusion rays: trip counts within a warp differ by 3-5x) a warp
// otherwise runs at the pace of its longest ray with most lanes idle. Results are the same words at the same indices.
constexpr unsigned kRefillChunk = 256;   // rays per cursor atomic: contiguous in the (direction-binned) order, so a warp's rays stay coherent
constexpr int kRefillIdle = 8;           // refill once this many lanes are idle (SVO_REFILL_IDLE overrides: experiment switch)

template <bool FAST, bool LOD, typename IdxT>
__global__ void __launch_bounds__(kBatchThreads, 12)
raymarchBatchRefillKernel(const uint32_t *__restrict__ octree, uint64_t n, const float *__restrict__ o,
                          const float *__restrict__ d, float rayScale, uint8_t *__restrict__ hit,
                          float *__restrict__ t, uint32_t *__restrict__ normal, uint64_t *__restrict__ voxel,
                          const uint32_t *__restrict__ order, unsigned long long *__restrict__ cursor, int refillIdle) {
    extern __shared__ __align__(16) unsigned char smem[];
    SmemStack<IdxT, kBatchThreads, LOD> stack;
    stack.init(smem);
    constexpr unsigned kFull = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned below = (1u << lane) - 1u;

    // A lane is in flight while r.childShift < 8 (rayTrip leaves one of the kExit codes there when the ray ends), so
    // "alive" costs no register and no bookkeeping in the loop.
    RayState<IdxT> r;
    r.childShift = kExitMiss;
    uint64_t ray = ~uint64_t(0);        // index of the ray this lane holds (in flight, or finished and not yet written)
    bool finite = true;
    float tHit = kTreeMiss;
    uint64_t vox = 0;
    unsigned long long next = 0, end = 0;   // the warp's chunk [next, end) of the cursor's range
    bool exhausted = false;

    for (;;) {
        // ---- finished lanes write their results, then every idle lane takes the next ray of the warp's chunk
        const bool idleLane = r.childShift >= 8u;
        if (idleLane && ray != ~uint64_t(0)) {
            int code = exitCode(r.childShift, LOD);
            if (!finite) code = kMiss;
            if (code == kMiss) { tHit = kTreeMiss; vox = ~uint64_t(0); }
            const uint32_t material = code == kHitLeaf ? ldNode(octree + vox) : 0u;   // VoxelOctree.cpp:282
            if (hit) hit[ray] = uint8_t(code);
            if (t) t[ray] = tHit;
            if (normal) normal[ray] = material;
            if (voxel) voxel[ray] = vox;
            ray = ~uint64_t(0);
        }
        if (next >= end && !exhausted) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(cursor, (unsigned long long)kRefillChunk);
            base = __shfl_sync(kFull, base, 0);
            next = base;
            end = base + kRefillChunk < n ? base + kRefillChunk : n;
            if (base >= n) { exhausted = true; next = end = 0; }
        }
        const unsigned idle = __ballot_sync(kFull, idleLane);
        if (idleLane) {
            const unsigned long long i = next + __popc(idle & below);
            if (i < end) {
                ray = order ? uint64_t(__ldg(order + i)) : uint64_t(i);
                float ox = __ldg(o + 3*ray), oy = __ldg(o + 3*ray + 1), oz = __ldg(o + 3*ray + 2);
                float dx = __ldg(d + 3*ray), dy = __ldg(d + 3*ray + 1), dz = __ldg(d + 3*ray + 2);
                // non-finite rays never leave the loop (see raymarchBatchKernel): swapped for a two-trip ray, reported as misses
                const float magnitude = fabsf(ox) + fabsf(oy) + fabsf(oz) + fabsf(dx) + fabsf(dy) + fabsf(dz);
                finite = magnitude < __int_as_float(0x7f800000);
                if (!finite) { ox = oy = oz = 3.0f; dx = dy = dz = 1.0f; }
                tHit = kTreeMiss;
                rayBegin<FAST, IdxT>(octree, ox, oy, oz, dx, dy, dz, r);
            }
        }
        {
            const unsigned long long taken = next + __popc(idle);
            next = taken < end ? taken : end;
        }
        const unsigned live = __ballot_sync(kFull, r.childShift < 8u);
        if (live == 0) {
            if (exhausted) break;
            continue;                     // the chunk ran out exactly here: fetch the next one
        }
        // ---- trips, until enough lanes are idle again (or, once the cursor is exhausted, until all are done); the
        // vote is taken every second trip
        const int keep = (exhausted && next >= end) ? 0 : 32 - refillIdle;
        do {
            if (r.childShift < 8u) rayTrip<FAST, LOD, IdxT, kBatchThreads>(octree, r, rayScale, stack, tHit, vox);
            if (r.childShift < 8u) rayTrip<FAST, LOD, IdxT, kBatchThreads>(octree, r, rayScale, stack, tHit, vox);
        } while (__popc(__ballot_sync(kFull, r.childShift < 8u)) > keep);
    }
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
    SmemStack<IdxT, kCoarseThreads, true> stack;
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

// ---- K2c: the traversal prefix the rays of a tile share (FAST flavour only) ----------------------------------------
// The four corner pixels' rays of a tile are traced in lock step, four lanes per tile, for as long as all four are at
// the same loop-head state (same node, same child cell): pushes from the root down to the first empty cell near the
// tile's start point, and the first steps through empty space. Every other ray of the tile lies inside the cone of
// the four, and each decision of the traversal compares the t of two axis-aligned planes along a ray from a common
// origin -- (c1 - o.a)/d.a < (c2 - o.b)/d.b, linear in d once the octant is fixed -- so while the four corner rays
// agree, every ray between them takes the same decisions: the fine pass starts its rays at the last shared state
// (rayRestart) instead of at the root and skips those trips (8192^3 at 4K: 8.8 of 37 trips per ray, measured on the
// oracle's traces; all of them coherent pushes and steps, 14 % of the pass's warp instructions). What is not covered
// by the argument is rounding: a ray for which one of the skipped comparisons is within an ulp of a tie, or whose
// start point (at distance startT on its own, Quake-normalised direction) is on the other side of a cell face from
// all four corners' -- measured: 11 of 7.86 M rays take a different first cell, and those still find their voxel
// through the normal traversal from there. That is why this is the FAST flavour only (>= 99.99 % identical pixels
// required, 100 % measured); the VALIDATION flavour always starts at the root.
constexpr int kPrefixThreads = 64;      // 16 tiles per block
constexpr int kPrefixMaxTrips = 64;

__global__ void __launch_bounds__(kPrefixThreads)
tilePrefixKernel(const uint32_t *__restrict__ octree, FramePlanDev plan, FrameConsts f, const TileRecord *__restrict__ tiles,
                 const FrameCounters *__restrict__ counters, uint32_t *__restrict__ prefix, int recordWords) {
    extern __shared__ __align__(16) unsigned char smem[];
    typedef SmemStack<uint32_t, kPrefixThreads, false> Stack;
    Stack stack;
    stack.init(smem);
    constexpr unsigned kFull = 0xffffffffu;
    const unsigned nTiles = counters->tilesRendered;
    const unsigned tile = blockIdx.x*(kPrefixThreads/4) + (threadIdx.x >> 2);
    const unsigned corner = threadIdx.x & 3u;
    const unsigned groupShift = threadIdx.x & 28u;          // first lane of this tile's four
    bool active = tile < nTiles;
    if (__ballot_sync(kFull, active) == 0) return;

    RayState<uint32_t> r;
    r.parent = 0; r.scale = 0; r.childShift = kExitMiss; r.octantMask = 0;
    r.posX = r.posY = r.posZ = 0.0f;
    if (active) {
        const uint4 rec = __ldg(reinterpret_cast<const uint4 *>(tiles) + tile);
        const int x0 = int(rec.x & 0xFFFFu), y0 = int(rec.x >> 16);
        const int w = min(x0 + 8, plan.width) - x0, h = min(y0 + 8, int(rec.y)) - y0;
        const int px = x0 + ((corner & 1u) ? w - 1 : 0), py = y0 + ((corner & 2u) ? h - 1 : 0);
        const float startT = __uint_as_float(rec.z);
        float rx, ry, rz;
        rayDirection(f, __ldg(plan.dxFine + px), __ldg(plan.dyFine + py), rx, ry, rz);
        rayBegin<true, uint32_t>(octree, addRn(f.posX, mulRn(rx, startT)), addRn(f.posY, mulRn(ry, startT)),
                                 addRn(f.posZ, mulRn(rz, startT)), rx, ry, rz, r);
    }
    uint32_t snapParent = 0, snapPacked = 0;
    float snapX = 0.0f, snapY = 0.0f, snapZ = 0.0f;
    float tHit;
    uint64_t vox;
    for (int trip = 0; trip < kPrefixMaxTrips; ++trip) {
        // all four rays of the tile in flight and at the same loop-head state?
        const uint32_t key = uint32_t(r.scale) | (r.childShift << 8) | (r.octantMask << 16);
        const unsigned alive = __ballot_sync(kFull, active && r.childShift < 8u);
        const unsigned same = __match_any_sync(kFull, (uint64_t(r.parent) << 32) | key);
        active = active && ((alive >> groupShift) & 0xFu) == 0xFu && ((same >> groupShift) & 0xFu) == 0xFu;
        if (__ballot_sync(kFull, active) == 0) break;
        if (active) {
            snapParent = r.parent;
            snapPacked = key | (1u << 24);
            snapX = r.posX; snapY = r.posY; snapZ = r.posZ;
            rayTrip<true, false, uint32_t, kPrefixThreads>(octree, r, 0.0f, stack, tHit, vox);
        }
    }
    if (corner == 0 && tile < nTiles) {
        uint32_t *out = prefix + size_t(tile)*size_t(recordWords);
        const int scale = int(snapPacked & 0xFFu);
        if (scale >= kMaxScale - 1) snapPacked = 0;     // nothing shared below the root: the rays start there anyway
        *reinterpret_cast<uint4 *>(out) = make_uint4(snapParent, snapPacked, __float_as_uint(snapX), __float_as_uint(snapY));
        out[4] = __float_as_uint(snapZ);
        if (snapPacked)
            for (int sc = scale + 1; sc < kMaxScale; ++sc) {        // the parents above the shared state (this lane's stack holds them)
                uint32_t parent;
                float unused;
                Stack::load(stack.slot(sc), parent, unused);
                out[kPrefixHeaderWords + (kMaxScale - 1 - sc)] = parent;
            }
    }
}

// ---- K3 -------------------------------------------------------------------

template <bool FAST, typename IdxT>
__global__ void __launch_bounds__(kTileThreads, 32)
finePassKernel(const uint32_t *__restrict__ octree, FramePlanDev plan, FrameConsts f,
               const TileRecord *__restrict__ tiles, const FrameCounters *__restrict__ counters,
               uint32_t *__restrict__ rgba, const uint32_t *__restrict__ prefix, int prefixWords) {
    extern __shared__ __align__(16) unsigned char smem[];
    SmemStack<IdxT, kTileThreads, false> stack;
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
        int code;
        if (FAST && sizeof(IdxT) == 4 && prefix) {
            // start at the state the tile's corner rays shared last (tilePrefixKernel) instead of at the root
            RayState<uint32_t> r;
            raySetup<FAST, uint32_t>(ox, oy, oz, rx, ry, rz, r);
            const SmemStack<uint32_t, kTileThreads, false> &stack32 = reinterpret_cast<const SmemStack<uint32_t, kTileThreads, false> &>(stack);
            if (!rayRestart<FAST, kTileThreads>(octree, r, prefix + size_t(tile)*size_t(prefixWords), stack32)) rayRoot<FAST, uint32_t>(octree, r);
            for (;;)
                if (rayTrip<FAST, false, uint32_t, kTileThreads>(octree, r, 0.0f, stack32, tHit, vox)) break;
            __syncwarp();
            code = exitCode(r.childShift, false);
        } else {
            code = raymarch<FAST, false, IdxT, kTileThreads>(octree, ox, oy, oz, rx, ry, rz, 0.0f, stack, tHit, vox);
        }
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
    SmemStack<IdxT, kTileThreads, false> stack;
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

} // namespace
int prefixRecordWords(const TreeDev &tree);
namespace {

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

int defaultTileRun() { return tileRunSetting().load(std::memory_order_relaxed); }

void setDefaultTileRun(int run) { tileRunSetting().store(run > 0 ? run : 4, std::memory_order_relaxed); }

TileShare tileShare(int tileRank, int tileWorld, int run) {
    TileShare share;
    share.rank = tileRank;
    share.world = tileWorld;
    share.run = tileWorld > 1 ? (run > 0 ? run : defaultTileRun()) : 1;
    return share;
}

int ownedTileColumns(int tileCols, const TileShare &share) {
    int n = 0;
    for (int start = share.rank*share.run; start < tileCols; start += share.world*share.run)
        n += (start + share.run <= tileCols) ? share.run : tileCols - start;
    return n;
}

namespace {

inline int ownedCols(const FramePlanDev &plan, const TileShare &share) { return ownedTileColumns(plan.tileCols, share); }
inline int ownedTiles(const FramePlanDev &plan, const TileShare &share) { return ownedCols(plan, share)*plan.totalTileRows; }

template <bool FAST, bool LOD, typename IdxT>
cudaError_t launchBatchT(const TreeDev &tree, uint64_t n, const float *o, const float *d, float rayScale,
                         uint8_t *hit, float *t, uint32_t *normal, uint64_t *voxel, const uint32_t *order,
                         cudaStream_t stream) {
    size_t smem = SmemStack<IdxT, kBatchThreads, LOD>::bytes(stackSlots(tree));
    auto kernel = raymarchBatchKernel<FAST, LOD, IdxT>;
    cudaError_t e = ensureSmem(kernel, smem);
    if (e != cudaSuccess) return e;
    uint64_t blocks = (n + kBatchThreads - 1)/kBatchThreads;
    if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    kernel<<<unsigned(blocks), kBatchThreads, smem, stream>>>(tree.words, n, o, d, rayScale, hit, t, normal, voxel, order);
    return cudaGetLastError();
}

template <bool FAST, bool LOD, typename IdxT>
cudaError_t launchBatchRefillT(const TreeDev &tree, uint64_t n, const float *o, const float *d, float rayScale,
                               uint8_t *hit, float *t, uint32_t *normal, uint64_t *voxel, const uint32_t *order,
                               unsigned long long *cursor, cudaStream_t stream) {
    size_t smem = SmemStack<IdxT, kBatchThreads, LOD>::bytes(stackSlots(tree));
    auto kernel = raymarchBatchRefillKernel<FAST, LOD, IdxT>;
    cudaError_t e = ensureSmem(kernel, smem);
    if (e != cudaSuccess) return e;
    // persistent grid: as many blocks as fit the device at once (one wave), never more than the rays need. Asked for on
    // every call (microseconds): the answer depends on the current device and, through the stack, on the tree's depth, and
    // svo_multi_raymarch_batch launches from one thread per device.
    int device = 0, sms = 0, perSm = 0;
    if ((e = cudaGetDevice(&device)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, kBatchThreads, smem)) != cudaSuccess) return e;
    const int residentBlocks = (sms > 0 ? sms : 1)*(perSm > 0 ? perSm : 1);
    uint64_t blocks = (n + kBatchThreads - 1)/kBatchThreads;
    if (blocks > uint64_t(residentBlocks)) blocks = uint64_t(residentBlocks);
    if ((e = cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), stream)) != cudaSuccess) return e;
    static const int refillIdle = [] {
        const char *e = getenv("SVO_REFILL_IDLE");
        const int v = e ? atoi(e) : 0;
        return v >= 1 && v <= 31 ? v : kRefillIdle;
    }();
    kernel<<<unsigned(blocks), kBatchThreads, smem, stream>>>(tree.words, n, o, d, rayScale, hit, t, normal, voxel, order, cursor, refillIdle);
    return cudaGetLastError();
}

template <typename IdxT>
cudaError_t launchCoarseT(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts, float *depth,
                          FrameCounters *counters, const TileShare &share, cudaStream_t stream) {
    const int tileRank = share.rank, tileWorld = share.world, run = share.run;
    size_t smem = SmemStack<IdxT, kCoarseThreads, true>::bytes(stackSlots(tree));
    auto kernel = coarsePassKernel<IdxT>;
    cudaError_t e = ensureSmem(kernel, smem);
    if (e != cudaSuccess) return e;
    // a rank needs the corner columns on both sides of each of its runs: run + 1 per run
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
                        int pixelStride, uint32_t *prefix, cudaStream_t stream) {
    size_t smem = SmemStack<IdxT, kTileThreads, false>::bytes(stackSlots(tree));
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
    const int prefixWords = prefixRecordWords(tree);
    if (!FAST || sizeof(IdxT) != 4 || prefixWords == 0) prefix = nullptr;
    kernel<<<blocks, kTileThreads, smem, stream>>>(tree.words, plan, consts, tiles, counters, rgba, prefix, prefixWords);
    return cudaGetLastError();
}

} // namespace

cudaError_t launchRaymarchBatch(const TreeDev &tree, uint64_t n, const float *o, const float *d, float rayScale,
                                int flavour, uint8_t *hit, float *t, uint32_t *normal, uint64_t *voxel,
                                const uint32_t *order, unsigned long long *refillCursor, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    bool wide = wideIndex(tree);
    bool lod = rayScale != 0.0f;
    bool fast = flavour != 0;
    if (refillCursor) {
#define SVO_REFILL(F, L) \
    (wide ? launchBatchRefillT<F, L, uint64_t>(tree, n, o, d, rayScale, hit, t, normal, voxel, order, refillCursor, stream) \
          : launchBatchRefillT<F, L, uint32_t>(tree, n, o, d, rayScale, hit, t, normal, voxel, order, refillCursor, stream))
        if (fast) return lod ? SVO_REFILL(true, true) : SVO_REFILL(true, false);
        return lod ? SVO_REFILL(false, true) : SVO_REFILL(false, false);
#undef SVO_REFILL
    }
#define SVO_BATCH(F, L) \
    (wide ? launchBatchT<F, L, uint64_t>(tree, n, o, d, rayScale, hit, t, normal, voxel, order, stream) \
          : launchBatchT<F, L, uint32_t>(tree, n, o, d, rayScale, hit, t, normal, voxel, order, stream))
    if (fast) return lod ? SVO_BATCH(true, true) : SVO_BATCH(true, false);
    return lod ? SVO_BATCH(false, true) : SVO_BATCH(false, false);
#undef SVO_BATCH
}

// ---- coherence order for incoherent ray batches -------------------------------------------------
// Rays are binned by direction (4 x 4 x 4 cells of d / max|d|) with ONE stable counting sort over the 64 bins, so rays
// of a bin keep their submission order (neighbouring pixels stay neighbours). A warp then holds rays that share their
// octant mirroring and roughly their direction: on 16-spp ambient-occlusion rays the trace-driven model gives 1.65x
// fewer warp instructions (SIMT efficiency 0.21 -> 0.34). Three small kernels of our own (no library sort): bin keys +
// per-tile histograms, one scan over (bin, tile), stable scatter of the ray indices.
constexpr int kBinThreads = 256;                 // 8 warps
constexpr int kBinRaysPerWarp = 512;             // a warp bins a contiguous run of rays
constexpr int kBinTile = (kBinThreads/32)*kBinRaysPerWarp;   // rays per block
constexpr int kBins = 64;

__device__ __forceinline__ uint32_t directionBin(const float *__restrict__ d, uint64_t i) {
    const float x = __ldg(d + 3*i), y = __ldg(d + 3*i + 1), z = __ldg(d + 3*i + 2);
    const float m = fmaxf(fmaxf(fabsf(x), fabsf(y)), fmaxf(fabsf(z), 1e-30f));
    auto cell = [m](float v) { return uint32_t(min(3, max(0, int((v/m + 1.0f)*2.0f)))); };
    return cell(x) | (cell(y) << 2) | (cell(z) << 4);    // NaN directions land in cell 0 (int(NaN) == 0)
}

// keys[i] = bin of ray i; tileCounts[bin][tile] = rays of the tile in the bin
__global__ void __launch_bounds__(kBinThreads)
directionBinKernel(uint64_t n, const float *__restrict__ d, uint8_t *__restrict__ keys, uint32_t *__restrict__ tileCounts,
                   uint32_t nTiles) {
    __shared__ uint32_t hist[kBins];
    if (threadIdx.x < kBins) hist[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t first = uint64_t(blockIdx.x)*kBinTile;
    for (uint32_t k = threadIdx.x; k < uint32_t(kBinTile); k += kBinThreads) {
        const uint64_t i = first + k;
        if (i < n) {
            const uint32_t key = directionBin(d, i);
            keys[i] = uint8_t(key);
            atomicAdd(&hist[key], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < kBins) tileCounts[size_t(threadIdx.x)*nTiles + blockIdx.x] = hist[threadIdx.x];
}

// exclusive scan of tileCounts in (bin, tile) order, in place: one block, a running carry over chunks of its width
__global__ void __launch_bounds__(1024)
scanTileCountsKernel(uint32_t *__restrict__ counts, uint32_t total) {
    __shared__ uint32_t warpSums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < total; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < total ? counts[i] : 0u;
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= unsigned(o)) incl += up;
        }
        if (lane == 31) warpSums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warpSums[lane], wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= unsigned(o)) wi += up;
            }
            warpSums[lane] = wi - w;            // exclusive
        }
        __syncthreads();
        const uint32_t c = carry;
        if (i < total) counts[i] = c + warpSums[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + warpSums[31] + incl;
        __syncthreads();
    }
}

// order[position] = ray index, stable inside every bin: the tile's base per bin comes from the scan, warps of the tile
// take their rays in order (each warp a contiguous run), lanes with equal keys are ranked by lane (__match_any_sync)
__global__ void __launch_bounds__(kBinThreads)
scatterByBinKernel(uint64_t n, const uint8_t *__restrict__ keys, const uint32_t *__restrict__ tileBase, uint32_t nTiles,
                   uint32_t *__restrict__ order) {
    __shared__ uint32_t warpHist[kBinThreads/32][kBins];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t k = lane; k < uint32_t(kBins); k += 32) warpHist[warp][k] = 0;
    __syncwarp();
    const uint64_t first = uint64_t(blockIdx.x)*kBinTile + uint64_t(warp)*kBinRaysPerWarp;
    for (uint32_t k = 0; k < uint32_t(kBinRaysPerWarp); k += 32) {
        const uint64_t i = first + k + lane;
        if (i < n) atomicAdd(&warpHist[warp][keys[i]], 1u);
    }
    __syncthreads();
    // warp w's base in bin k: tile base + the counts of the warps before it
    for (uint32_t k = threadIdx.x; k < uint32_t(kBins); k += kBinThreads) {
        uint32_t run = __ldg(tileBase + size_t(k)*nTiles + blockIdx.x);
        for (int w = 0; w < kBinThreads/32; ++w) {
            const uint32_t c = warpHist[w][k];
            warpHist[w][k] = run;
            run += c;
        }
    }
    __syncthreads();
    for (uint32_t k = 0; k < uint32_t(kBinRaysPerWarp); k += 32) {
        const uint64_t i = first + k + lane;
        const bool valid = i < n;
        const uint32_t key = valid ? uint32_t(keys[i]) : 0xFFu;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (valid) {
            const uint32_t base = warpHist[warp][key];
            order[base + __popc(peers & ((1u << lane) - 1u))] = uint32_t(i);
        }
        __syncwarp();
        if (valid && (peers & ((1u << lane) - 1u)) == 0) warpHist[warp][key] += __popc(peers);   // the lowest lane of each group
        __syncwarp();
    }
}

size_t coherenceOrderBytes(uint64_t n) {
    const uint64_t nTiles = (n + kBinTile - 1)/kBinTile;
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    return up(size_t(n)*sizeof(uint32_t)) + up(size_t(n)) + up(size_t(nTiles)*kBins*sizeof(uint32_t)) + 256;
}

// workspace: coherenceOrderBytes(n) bytes on the device. *orderOut points into it.
cudaError_t buildCoherenceOrder(uint64_t n, const float *d, void *workspace, const uint32_t **orderOut, cudaStream_t stream) {
    if (n >= (1ull << 31)) return cudaErrorInvalidValue;
    const uint32_t nTiles = uint32_t((n + kBinTile - 1)/kBinTile);
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    unsigned char *base = static_cast<unsigned char *>(workspace);
    uint32_t *order = reinterpret_cast<uint32_t *>(base);
    uint8_t *keys = base + up(size_t(n)*sizeof(uint32_t));
    uint32_t *counts = reinterpret_cast<uint32_t *>(base + up(size_t(n)*sizeof(uint32_t)) + up(size_t(n)));
    directionBinKernel<<<nTiles, kBinThreads, 0, stream>>>(n, d, keys, counts, nTiles);
    scanTileCountsKernel<<<1, 1024, 0, stream>>>(counts, nTiles*uint32_t(kBins));
    scatterByBinKernel<<<nTiles, kBinThreads, 0, stream>>>(n, keys, counts, nTiles, order);
    *orderOut = order;
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
                             float *depth, FrameCounters *counters, const TileShare &share, cudaStream_t stream) {
    // The beam pass is individually rounded in BOTH flavours: its depths feed
    // `minT - 0.03` (Main.cpp:197) and the tile-skip test (Main.cpp:191), so a
    // last-bit change here moves every ray origin of a tile or drops / adds a
    // whole tile. It is < 5 % of the rays; FAST only changes the fine pass.
    (void)flavour;
    bool wide = wideIndex(tree);
    return wide ? launchCoarseT<uint64_t>(tree, plan, consts, depth, counters, share, stream)
                : launchCoarseT<uint32_t>(tree, plan, consts, depth, counters, share, stream);
}

cudaError_t launchClassifyTiles(const FramePlanDev &plan, const FrameConsts &consts, const float *depth,
                                uint32_t *rgba, const TileShare &share, int pixelStride, TileRecord *tiles,
                                FrameCounters *counters, unsigned long long *fineRaysTotal, cudaStream_t stream) {
    int owned = ownedTiles(plan, share);
    if (owned <= 0) return cudaSuccess;
    classifyTilesKernel<<<(owned + kClassifyThreads - 1)/kClassifyThreads, kClassifyThreads, 0, stream>>>(
        plan, consts.beamBias, depth, rgba, share.rank, share.world, share.run,
        ownedCols(plan, share), owned, pixelStride > 1 ? pixelStride : 1, tiles, counters, fineRaysTotal);
    return cudaGetLastError();
}

cudaError_t launchFinePass(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts, int flavour,
                           const TileRecord *tiles, const FrameCounters *counters, uint32_t *rgba,
                           const TileShare &share, int pixelStride, uint32_t *prefix, cudaStream_t stream) {
    int owned = ownedTiles(plan, share);
    if (owned <= 0) return cudaSuccess;
    bool wide = wideIndex(tree);
    if (flavour != 0)
        return wide ? launchFineT<true, uint64_t>(tree, plan, consts, tiles, counters, rgba, owned, pixelStride, prefix, stream)
                    : launchFineT<true, uint32_t>(tree, plan, consts, tiles, counters, rgba, owned, pixelStride, prefix, stream);
    return wide ? launchFineT<false, uint64_t>(tree, plan, consts, tiles, counters, rgba, owned, pixelStride, prefix, stream)
                : launchFineT<false, uint32_t>(tree, plan, consts, tiles, counters, rgba, owned, pixelStride, prefix, stream);
}

// K2c: the tile's shared traversal prefix, four corner rays per tile in lock step; a no-op (returns 0 launches) when
// launchFinePass would not use the records.
cudaError_t launchTilePrefix(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts, int flavour,
                             const TileRecord *tiles, const FrameCounters *counters, const TileShare &share,
                             int pixelStride, uint32_t *prefix, cudaStream_t stream) {
    const int owned = ownedTiles(plan, share);
    if (owned <= 0 || !prefix || !finePassUsesPrefix(tree, flavour, pixelStride)) return cudaSuccess;
    const size_t psmem = SmemStack<uint32_t, kPrefixThreads, false>::bytes(stackSlots(tree));
    cudaError_t e = ensureSmem(tilePrefixKernel, psmem);
    if (e != cudaSuccess) return e;
    const int tilesPerBlock = kPrefixThreads/4;
    tilePrefixKernel<<<(owned + tilesPerBlock - 1)/tilesPerBlock, kPrefixThreads, psmem, stream>>>(
        tree.words, plan, consts, tiles, counters, prefix, prefixRecordWords(tree));
    return cudaGetLastError();
}

// Words per tile of the shared-prefix records (header + one parent per stack slot, rounded to 16 bytes); 0 when the
// tree's kernels do not use them (64-bit word indices).
// Shallow trees do not use them either: what the restart saves is the descent from the root, and below ten levels the
// extra launch costs more than that (measured on one box, FAST frames with / without: 8192^3 (13 levels) at 4K 9.91 /
// 9.07 Grays/s, its fly-through 11.88 / 10.65, 2048^3 (11 levels) at 1080p 8.80 / 8.45, the 256^3 Dragon (8 levels) at
// 720p 12.15 / 12.57).
constexpr uint32_t kPrefixMinDepth = 10;

int prefixRecordWords(const TreeDev &tree) {
    if (wideIndex(tree) || tree.depth < kPrefixMinDepth) return 0;
    return (kPrefixHeaderWords + int(stackSlots(tree)) + 3) & ~3;
}

// 1 when launchFinePass with these arguments and a prefix buffer launches tilePrefixKernel as well (launch counts)
int finePassUsesPrefix(const TreeDev &tree, int flavour, int pixelStride) {
    return flavour != 0 && pixelStride <= 1 && prefixRecordWords(tree) > 0 ? 1 : 0;
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

// RGBA words -> (grey, alpha) byte pairs for the tile columns a rank owns (SVO_PIXELS_GREY8A8): the reference's pixel is
// 0xFF000000 | g << 16 | g << 8 | g or 0 (Main.cpp:128-132, :165), so byte 0 and byte 3 of the word carry all of it.
__global__ void __launch_bounds__(256)
packGrey8a8Kernel(const uint32_t *__restrict__ src, uint16_t *__restrict__ dst, int width, int height, int ownedPixels,
                  int run, int tileRank, int tileWorld) {
    const uint32_t k = blockIdx.x*256u + threadIdx.x;
    const int y = int(blockIdx.y);
    if (k >= uint32_t(ownedPixels) || y >= height) return;
    const uint32_t slot = k >> 3;
    const uint32_t tx = (slot/uint32_t(run))*uint32_t(tileWorld*run) + uint32_t(tileRank*run) + slot%uint32_t(run);
    const uint32_t x = tx*8u + (k & 7u);
    if (x >= uint32_t(width)) return;
    const size_t at = size_t(y)*size_t(width) + x;
    const uint32_t p = src[at];
    dst[at] = uint16_t((p & 0xFFu) | ((p >> 16) & 0xFF00u));
}

// four pixels per thread: a warp stores 256 contiguous bytes -- one full-width write burst when `dst` is mapped host memory
__global__ void __launch_bounds__(256)
packGrey8a8Vec4Kernel(const uint32_t *__restrict__ src, uint16_t *__restrict__ dst, int width, int height, int ownedQuads,
                      int run, int tileRank, int tileWorld) {
    const uint32_t k = blockIdx.x*256u + threadIdx.x;      // quad k of this rank's columns laid side by side (two per tile column)
    const int y = int(blockIdx.y);
    if (k >= uint32_t(ownedQuads) || y >= height) return;
    const uint32_t slot = k >> 1;
    const uint32_t tx = (slot/uint32_t(run))*uint32_t(tileWorld*run) + uint32_t(tileRank*run) + slot%uint32_t(run);
    const uint32_t x = tx*8u + (k & 1u)*4u;
    if (x >= uint32_t(width)) return;
    const size_t at = size_t(y)*size_t(width) + x;
    if (x + 4u <= uint32_t(width)) {
        const uint4 p = *reinterpret_cast<const uint4 *>(src + at);
        auto pack = [](uint32_t a, uint32_t b) {
            return ((a & 0xFFu) | ((a >> 16) & 0xFF00u)) | (((b & 0xFFu) | ((b >> 16) & 0xFF00u)) << 16);
        };
        *reinterpret_cast<uint2 *>(dst + at) = make_uint2(pack(p.x, p.y), pack(p.z, p.w));
    } else {
        for (uint32_t i = 0; x + i < uint32_t(width); ++i) {
            const uint32_t p = src[at + i];
            dst[at + i] = uint16_t((p & 0xFFu) | ((p >> 16) & 0xFF00u));
        }
    }
}

cudaError_t launchPackGrey8a8(const FramePlanDev &plan, int width, int height, const uint32_t *src, uint16_t *dst,
                              const TileShare &share, cudaStream_t stream) {
    const int cols = ownedCols(plan, share);
    if (cols <= 0) return cudaSuccess;
    const bool aligned = (width & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0;
    if (aligned) {
        const int quads = cols*2;
        dim3 grid(unsigned((quads + 255)/256), unsigned(height));
        packGrey8a8Vec4Kernel<<<grid, 256, 0, stream>>>(src, dst, width, height, quads, share.run, share.rank, share.world);
    } else {
        const int ownedPixels = cols*8;
        dim3 grid(unsigned((ownedPixels + 255)/256), unsigned(height));
        packGrey8a8Kernel<<<grid, 256, 0, stream>>>(src, dst, width, height, ownedPixels, share.run, share.rank, share.world);
    }
    return cudaGetLastError();
}

cudaError_t launchCopyOwnedColumns(const FramePlanDev &plan, int width, int height, const uint32_t *src, uint32_t *dst,
                                   const TileShare &share, cudaStream_t stream) {
    const int ownedPixels = ownedCols(plan, share)*8;
    if (ownedPixels <= 0) return cudaSuccess;
    dim3 grid(unsigned((ownedPixels + 255)/256), unsigned(height));
    copyOwnedColumnsKernel<<<grid, 256, 0, stream>>>(src, dst, width, height, ownedPixels, share.run, share.rank, share.world);
    return cudaGetLastError();
}

} // namespace svo
