// Kernel launch interface between the C ABI (svo_capi.cu) and the kernels
// (svo_kernels.cu). Host-only structs; no CUDA types beyond cudaStream_t.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "svo_traverse.cuh"

namespace svo {

// Geometry of the reference's strip / tile decomposition (src/Main.cpp:351-362)
// for one (width, height, strips) configuration, with the running-sum screen
// coordinates precomputed on the host exactly as renderBatch / renderTile
// accumulate them (Main.cpp:97-100, 167-170).
struct FramePlanDev {
    int32_t width, height;
    int32_t nStrips;        // strips that own at least one row
    int32_t stripRows;      // rows per strip ("stride", Main.cpp:351)
    int32_t tilesX;         // corner columns
    int32_t tilesYFull;     // corner rows of a full strip
    int32_t tilesYLast;     // corner rows of the last strip
    int32_t tileRowsFull;   // tilesYFull - 1
    int32_t tileRowsLast;   // tilesYLast - 1
    int32_t tileCols;       // tilesX - 1
    int32_t totalTileRows;
    int32_t totalTiles;
    int32_t totalCorners;
    const float *dxCoarse;  // [tilesX]
    const float *dyCoarse;  // [nStrips][tilesYFull]
    const float *dxFine;    // [width]
    const float *dyFine;    // [height]
};

struct FrameCounters {      // device-side; zeroed by the coarse pass, filled by the tile classifier
    unsigned int tilesRendered; // == length of the tile list
    unsigned int pad;
    unsigned long long fineRays;
};

// One rendered 8x8 tile (Main.cpp:186-197), produced by the classifier, consumed by the fine pass.
struct TileRecord {
    uint32_t xy;        // x0 | y0 << 16 (pixel origin)
    uint32_t yEnd;      // strip end row, exclusive (tiles are clipped to their strip, Main.cpp:194-195)
    float startT;       // max(min4 - 0.03, 0), Main.cpp:197
    uint32_t pad;
};

struct TreeDev {
    const uint32_t *words;
    uint64_t nWords;
    uint32_t depth;
};

// order == nullptr: thread k traces ray k. Otherwise thread k traces ray order[k] (see buildCoherenceOrder).
// refillCursor != nullptr (one device word, zeroed by the launch): the persistent lane-refill kernel -- warps pull
// rays from the cursor and a lane that finishes takes the next ray; nullptr: one ray per thread.
cudaError_t launchRaymarchBatch(const TreeDev &tree, uint64_t n, const float *o, const float *d, float rayScale,
                                int flavour, uint8_t *hit, float *t, uint32_t *normal, uint64_t *voxel,
                                const uint32_t *order, unsigned long long *refillCursor, cudaStream_t stream);

// Direction-binned submission order for incoherent batches (one stable 6-bit radix pass).
size_t coherenceOrderBytes(uint64_t n);
cudaError_t buildCoherenceOrder(uint64_t n, const float *d, void *workspace, const uint32_t **orderOut, cudaStream_t stream);

// shade + pack per ray (Main.cpp:81-90, 128-132); hit may be null (every ray shaded). Device pointers.
cudaError_t launchShadeBatch(uint64_t n, const uint8_t *hit, const uint32_t *normal, const float *d, const float light[3],
                             uint32_t *rgba, cudaStream_t stream);

// Multi-GPU tile interleave: tile column tx belongs to rank (tx / run) % world -- vertical stripes `run` tile columns
// wide, dealt round-robin. One value per call: nothing process-wide is read while a frame is being enqueued.
struct TileShare {
    int rank = 0, world = 1, run = 1;
};
// The stripe width of callers that do not name one (svo_frame_desc has no field for it): SVO_TILE_RUN or
// svo_frame_set_tile_run, else four. run <= 0: that default; world == 1: stripes do not exist (run 1).
int defaultTileRun();
void setDefaultTileRun(int run);     // <= 0 restores four
TileShare tileShare(int tileRank, int tileWorld, int run = 0);
int ownedTileColumns(int tileCols, const TileShare &share);

// Beam pass; also zeroes `counters` for the classifier that follows on the same stream.
// With share.world >= 3 only the corners next to this rank's tile columns are traced.
cudaError_t launchCoarsePass(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts, int flavour,
                             float *depth, FrameCounters *counters, const TileShare &share, cudaStream_t stream);

// Per owned tile: min of the four corner depths; skipped tiles are zero-filled (the strip memset,
// Main.cpp:165), rendered tiles are appended to `tiles` (capacity = owned tiles) and counted.
cudaError_t launchClassifyTiles(const FramePlanDev &plan, const FrameConsts &consts, const float *depth,
                                uint32_t *rgba, const TileShare &share, int pixelStride, TileRecord *tiles,
                                FrameCounters *counters, unsigned long long *fineRaysTotal, cudaStream_t stream);

// Fine pass over the tile list (grid covers the worst case; blocks past the list length exit).
// pixelStride > 1: renderTile's preview mode (Main.cpp:101-106), one ray per stride x stride block of a tile.
// prefix != nullptr (room for prefixRecordWords(tree) words per owned tile, filled by launchTilePrefix): in the FAST
// flavour the rays of a tile start at the traversal state its four corner rays share instead of at the root.
cudaError_t launchTilePrefix(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts, int flavour,
                             const TileRecord *tiles, const FrameCounters *counters, const TileShare &share,
                             int pixelStride, uint32_t *prefix, cudaStream_t stream);
cudaError_t launchFinePass(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts, int flavour,
                           const TileRecord *tiles, const FrameCounters *counters, uint32_t *rgba,
                           const TileShare &share, int pixelStride, uint32_t *prefix, cudaStream_t stream);
int prefixRecordWords(const TreeDev &tree);
int finePassUsesPrefix(const TreeDev &tree, int flavour, int pixelStride);

// The owned tile columns' pixels from `src` to `dst` (both width x height, pitch = width): the per-rank
// device -> host leg of a multi-GPU frame when `dst` is mapped page-locked host memory.
// The owned tile columns' pixels as (grey, alpha) byte pairs (SVO_PIXELS_GREY8A8), same pitch in pixels.
cudaError_t launchPackGrey8a8(const FramePlanDev &plan, int width, int height, const uint32_t *src, uint16_t *dst,
                              const TileShare &share, cudaStream_t stream);
cudaError_t launchCopyOwnedColumns(const FramePlanDev &plan, int width, int height, const uint32_t *src, uint32_t *dst,
                                   const TileShare &share, cudaStream_t stream);

} // namespace svo
