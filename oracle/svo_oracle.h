/*
 * TEST INFRASTRUCTURE ONLY -- the parity oracle, never the product.
 *
 * Plain-C restatement of the reference's hot path, instrumented with the
 * node-fetch counters the roofline accounting needs:
 *   - VoxelOctree::raymarch           reference src/VoxelOctree.cpp:207-346
 *   - shade / renderTile / renderBatch reference src/Main.cpp:81-202
 *   - the per-frame constants          reference src/Main.cpp:149-163,
 *                                       src/math/Mat4.cpp:59-86, MatrixStack.cpp:71-73
 *   - decompressMaterial / invSqrt     reference src/Util.hpp:47-58,86-100
 *
 * PINNED: tests/test_oracle_pins.py checks this restatement bit-for-bit
 * against the reference's own object code (oracle/_ref/libsvo_ref.so, built
 * from /root/reference/src by oracle/build_ref.sh) on the reference's sample
 * tree, and against the committed golden vectors under tests/golden/ that
 * were generated from that object code (tests/golden/make_golden.py).
 *
 * Must be compiled with -ffp-contract=off (no FMA contraction).
 */
#ifndef SVO_ORACLE_H_
#define SVO_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint64_t rays;
    uint64_t iterations;   /* trips round the while loop, VoxelOctree.cpp:252 */
    uint64_t desc_fetches; /* `current = _octree[parent]`, :253-254 */
    uint64_t far_fetches;  /* `_octree[parent + 1]`, :278-279 */
    uint64_t leaf_fetches; /* leaf material word, :282 */
    uint64_t pushes;       /* :287-306 */
    uint64_t pops;         /* :318-338 */
    uint64_t max_iterations;
    uint64_t hits;         /* raymarch returned true (leaf hit or LOD exit) */
    uint64_t lod_exits;
} svo_oracle_counters;

/* result codes */
enum { SVO_ORACLE_MISS = 0, SVO_ORACLE_LEAF = 1, SVO_ORACLE_LOD = 2 };

/* One ray. Returns SVO_ORACLE_MISS / _LEAF / _LOD (the reference returns
 * `true` for the latter two). `*normal` is written only on a leaf hit, `*t`
 * only on a hit, exactly like the reference. `*voxel` (may be NULL): leaf hit
 * -> word index of the leaf; LOD exit -> parent descriptor index |
 * (childShift << 60). `c` may be NULL. */
int svo_oracle_raymarch(const uint32_t *octree, const float *o, const float *d, float rayScale,
        uint32_t *normal, float *t, uint64_t *voxel, svo_oracle_counters *c);

/* n rays; o/d are n x 3 floats. hit: u8 result codes. Outputs keep their
 * previous contents where the reference leaves them untouched. Any output
 * pointer may be NULL. `threads` OS threads over contiguous ranges. */
void svo_oracle_raymarch_batch(const uint32_t *octree, uint64_t n, const float *o, const float *d,
        float rayScale, uint8_t *hit, float *t, uint32_t *normal, uint64_t *voxel,
        svo_oracle_counters *c, int threads);

/* Everything renderBatch derives from the matrix stacks before its loops
 * (Main.cpp:149-163). */
typedef struct {
    int32_t width, height, strips, tile_size;
    float pos[3];
    float a11, a12, a21, a22, a31, a32; /* tform columns 1 and 2 (rows 1-3) */
    float zx, zy, zz;                   /* planeDist * column 3 */
    float scale, tile_scale, coarse_scale, aspect;
    float light[3];
    float beam_bias;                    /* 0.03f, Main.cpp:197 */
} svo_oracle_frame;

void svo_oracle_frame_constants(const float *model16, const float *view16, const float *center,
        int width, int height, int strips, svo_oracle_frame *out);

/* Orbit camera, Main.cpp:212-213,244-245 (MODEL = rotXYZ(pitch,0,0)*rotXYZ(0,yaw,0),
 * VIEW = translate(0,0,-radius)). */
void svo_oracle_orbit_camera(float pitchDeg, float yawDeg, float radius, float *model16, float *view16);

/* Material word -> shaded grey value in [0, ...), Main.cpp:81-90. */
float svo_oracle_shade(uint32_t material, const float *dir, const float *light);
/* Grey value -> packed pixel, Main.cpp:128-132 (non-Apple branch). */
uint32_t svo_oracle_pack(float v);
void svo_oracle_decompress_material(uint32_t word, float *n, float *shade);
float svo_oracle_inv_sqrt(float x);

/* One frame = all strips of renderBatch. rgba: width*height words.
 * depth (optional): per strip tilesX*tilesY floats, strip after strip.
 * coarse/fine (optional): counters for the two ray classes. */
int svo_oracle_render_frame(const uint32_t *octree, const svo_oracle_frame *f, uint32_t *rgba,
        float *depth, svo_oracle_counters *coarse, svo_oracle_counters *fine, int threads);
/* Same with renderTile's pixel stride (Main.cpp:92-106): 1, or 3 = the reference's renderHalfSize preview. */
int svo_oracle_render_frame_strided(const uint32_t *octree, const svo_oracle_frame *f, int pixelStride, uint32_t *rgba,
        float *depth, svo_oracle_counters *coarse, svo_oracle_counters *fine, int threads);

/* Kernel design tool: per-ray loop-trip traces of the fine pass in GPU warp order (see svo_oracle.c). */
int64_t svo_oracle_trace_fine_warps(const uint32_t *octree, const svo_oracle_frame *f, int tileStride,
        uint8_t *ops, uint32_t maxOps, uint32_t *counts, int64_t maxWarps);

void svo_oracle_trace_rays(const uint32_t *octree, uint64_t n, const float *o, const float *d, float rayScale,
        uint8_t *ops, uint32_t maxOps, uint32_t *counts);

/* Tree statistics by a full walk from the root (App. A.1). */
typedef struct {
    uint64_t descriptors, leaves, far_words, far_blocks;
    uint32_t depth;              /* number of descriptor levels; voxel grid side = 1 << depth */
    uint64_t per_level[24];      /* descriptors per level */
    uint64_t max_index;          /* highest word index reached */
} svo_oracle_tree_stats;
int svo_oracle_tree_walk(const uint32_t *octree, uint64_t words, svo_oracle_tree_stats *out);

/* Row f2: VoxelOctree(VoxelData*) / buildOctree (reference src/VoxelOctree.cpp:125-205) over a dense
 * w*h*d grid (x fastest, 0 = empty). malloc'ed result, released with svo_oracle_free. */
uint32_t *svo_oracle_build_octree(const uint32_t *voxels, int w, int h, int d, uint64_t *nWordsOut, float center[3]);
void svo_oracle_free(void *p);

/* Row f3 (oracle/svo_oracle_ply.c): PlyLoader(path) + what VoxelData(loader, sideLength, mem) hands to
 * buildOctree (reference src/PlyLoader.cpp:64-474, src/VoxelData.cpp:50-56,178-181) with the whole volume in
 * one cache block; threadCount = size of the reference's thread pool (it shapes the sub-block partition the
 * result depends on). malloc'ed w*h*d volume (svo_oracle_free) or NULL. */
uint32_t *svo_oracle_voxelize_ply(const char *plyPath, int sideLength, int threadCount, int dims[3], uint64_t *nTrianglesOut);
/* PlyLoader's triangle list (33 floats each: pos[3][3], normal[3][3], color[3][3], lower[3], upper[3]). */
/* (triangle, sub-block) listings that fall outside the reference's sub-block grid (aliased flat indices): see
 * oracle/svo_oracle_ply.c. Block lists only; blockEdge 0 = one cache block. */
int64_t svo_oracle_block_list_aliases(const char *plyPath, int sideLength, int blockEdge, int threadCount,
                                      uint64_t *candidatesOut, int grid[3], int real[3]);
float *svo_oracle_ply_triangles(const char *plyPath, uint64_t *nOut, float lower[3], float upper[3]);
/* The filled voxels of a node array spanning side^3 voxels, written into vol (w*h*d, x fastest, pre-zeroed). */
void svo_oracle_tree_to_volume(const uint32_t *octree, int side, uint32_t *vol, int w, int h, int d);

#ifdef __cplusplus
}
#endif
#endif
