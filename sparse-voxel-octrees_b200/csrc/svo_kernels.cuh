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

struct FrameCounters {      // device-side, zeroed per frame
    unsigned long long fineRays;
    unsigned long long tilesRendered;
};

struct TreeDev {
    const uint32_t *words;
    uint64_t nWords;
    uint32_t depth;
};

cudaError_t launchRaymarchBatch(const TreeDev &tree, uint64_t n, const float *o, const float *d, float rayScale,
                                int flavour, uint8_t *hit, float *t, uint32_t *normal, uint64_t *voxel,
                                cudaStream_t stream);

cudaError_t launchCoarsePass(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts, int flavour,
                             float *depth, cudaStream_t stream);

cudaError_t launchFinePass(const TreeDev &tree, const FramePlanDev &plan, const FrameConsts &consts, int flavour,
                           const float *depth, uint32_t *rgba, int tileRank, int tileWorld, cudaStream_t stream);

// Counts the tiles / pixels the fine pass renders for this rank (from the coarse
// depth buffer alone); `counters` must be zeroed by the caller.
cudaError_t launchTileStats(const FramePlanDev &plan, const float *depth, int tileRank, int tileWorld,
                            FrameCounters *counters, cudaStream_t stream);

} // namespace svo
