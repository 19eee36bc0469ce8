// Triangle voxelisation on the GPU. See svo_voxelize.cu.
#pragma once

#include <cstdint>
#include <string>

#include "ply_io.hpp"
#include "svo_build.cuh"

namespace svo {

struct VoxelizeStats {
    uint64_t triangles = 0;
    uint64_t cellRecords = 0;     // (cell, triangle) overlaps found
    uint64_t voxels = 0;          // filled cells
    int dims[3] = {0, 0, 0};
    int cacheBlock = 0;           // edge of the reference's cache block for this memory budget
    int subBlock[3] = {0, 0, 0};  // per-thread sub-block of the reference's partition
    float overlapMs = 0.0f, sortMs = 0.0f, foldMs = 0.0f;
    uint32_t largeTriangles = 0;  // triangles handled by a whole block each (bounding box of more than 2^15 cells)
};

// The voxels the reference's PlyLoader + VoxelData(loader, sideLength, memoryBudget) hand to buildOctree,
// for a pool of `threadCount` threads (the partition the result depends on), fed into `builder`
// (begin() is called here with the volume the reference derives from the mesh).
bool voxelizeMesh(const Mesh &mesh, int sideLength, uint64_t memoryBudget, int threadCount, OctreeBuilder &builder,
                  VoxelizeStats &stats, std::string &err);

} // namespace svo
