// Octree construction on the GPU ("next" row f2): the node array the reference's
// VoxelOctree(VoxelData*) / buildOctree produces (reference src/VoxelOctree.cpp:125-205), word for word,
// built level by level in HBM instead of by a recursive host walk. See svo_build.cu.
#pragma once

#include <cstdint>
#include <string>

#include <cuda_runtime.h>

namespace svo {

struct BuildStats {
    uint64_t voxels = 0;        // non-empty voxels that made it into the tree
    uint64_t nodes = 0;         // descriptors (all levels)
    uint64_t farBlocks = 0;     // child blocks that carry far words
    float gatherMs = 0.0f;      // dense chunks -> (Morton key, value) list, or sparse list -> keys
    float sortMs = 0.0f;
    float levelsMs = 0.0f;      // bottom-up: nodes, masks, subtree sizes, far decisions
    float emitMs = 0.0f;        // top-down: addresses and words
};

struct BuildResult {
    uint32_t *dWords = nullptr; // cudaMalloc'ed, nWords + 1 words (one zero padding word)
    uint64_t nWords = 0;
    uint32_t depth = 0;
    float center[3] = {0.0f, 0.0f, 0.0f};
    BuildStats stats;
};

// Collects voxels on the current device, then builds. Not thread-safe; one build per object.
class OctreeBuilder {
public:
    OctreeBuilder() = default;
    ~OctreeBuilder();
    OctreeBuilder(const OctreeBuilder &) = delete;
    OctreeBuilder &operator=(const OctreeBuilder &) = delete;

    // Volume of w x h x d voxels (the .voxel header, reference src/VoxelData.cpp:41-43).
    bool begin(int w, int h, int d, std::string &err);
    // `count` voxels of the dense grid starting at linear index `first` (x fastest), in DEVICE memory.
    bool addDenseChunk(const uint32_t *dVoxels, uint64_t first, uint64_t count, std::string &err);
    // n voxels as (x, y, z) triples + values, in DEVICE memory; zero values and out-of-volume
    // coordinates are dropped like the dense path drops them. Coordinates must be unique.
    bool addSparse(const uint32_t *dXyz, const uint32_t *dValues, uint64_t n, std::string &err);
    // Sorts, builds the levels bottom-up, lays the words out top-down.
    bool finish(BuildResult &out, std::string &err);

private:
    bool reserve(uint64_t entries, std::string &err);
    int w_ = 0, h_ = 0, d_ = 0, side_ = 0, levels_ = 0;
    uint64_t *dKeys_ = nullptr;
    uint32_t *dVals_ = nullptr;
    unsigned long long *dCursor_ = nullptr;
    uint64_t capacity_ = 0, count_ = 0;
    float gatherMs_ = 0.0f;
};

// The builder's private stream-ordered memory pool on the current device (nullptr if pools are unavailable).
// The voxeliser allocates from it too, so the tens of GB it releases are what the build that follows reuses.
// OctreeBuilder::finish hands the pool to a background thread for trimming (giving 40 GB back to the driver
// takes up to 2 s); the next build waits for that thread first.
cudaMemPool_t buildScratchPool();

// The inverse of the builder: the filled voxels of a node array resident on the current device, in
// Morton order (ascending x + 2y + 4z per level). Outputs are cudaMalloc'ed: 3 coordinates per voxel and
// one material word per voxel.
bool extractVoxels(const uint32_t *dWords, uint32_t depth, uint32_t **dXyzOut, uint32_t **dValuesOut, uint64_t *nOut,
                   std::string &err);

} // namespace svo
