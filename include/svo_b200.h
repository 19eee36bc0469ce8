/*
 * svo_b200.h -- C ABI of the B200-native sparse-voxel-octree ray caster.
 *
 * This is the drop-in boundary for the ray-casting path of
 * tunabrain/sparse-voxel-octrees. The reference has no FFI layer; its de-facto
 * operator API is the public part of `class VoxelOctree`
 * (reference src/VoxelOctree.hpp:48-57) plus the per-frame call
 * `renderBatch(BatchData*)` (reference src/Main.cpp:139, called at :218).
 * Each entry point below names the reference interface it replaces. The C++
 * facade with the reference's own signatures is
 * sparse-voxel-octrees_b200/host/VoxelOctree.hpp; INTEGRATION.md shows the
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns an svo_status
 *    (0 = ok) unless noted; svo_last_error() gives the thread-local message.
 *  - handles are opaque; the caller owns every host buffer it passes in.
 *  - functions may be called from any host thread; one in-flight call per
 *    tree handle (the reference's raymarch is re-entrant across 16 threads,
 *    Main.cpp:364-367 -- here concurrency comes from batching instead).
 *  - there is NO CPU fallback: without a CUDA device every device entry
 *    point fails with SVO_ERR_NO_DEVICE.
 *  - node words are the reference's array unchanged.
 */
#ifndef SVO_B200_H_
#define SVO_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(_WIN32)
#  define SVO_API __declspec(dllexport)
#else
#  define SVO_API __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define SVO_ABI_VERSION 2

typedef enum svo_status {
    SVO_OK = 0,
    SVO_ERR_INVALID_ARGUMENT = 1,
    SVO_ERR_IO = 2,             /* fopen/fread/fwrite failed (the reference ignores these silently, VoxelOctree.cpp:60,95) */
    SVO_ERR_FORMAT = 3,         /* malformed .oct / LZ4 stream / node array */
    SVO_ERR_OUT_OF_MEMORY = 4,
    SVO_ERR_CUDA = 5,
    SVO_ERR_NO_DEVICE = 6,
    SVO_ERR_UNSUPPORTED = 7
} svo_status;

/* Arithmetic flavour (BASELINE.json north_star).
 * VALIDATION: every float op individually rounded (no FMA contraction, IEEE
 *   divide, the reference's (a<b)?b:a min/max forms): hit mask and hit voxel
 *   bit-exact against the reference built with -ffp-contract=off.
 * FAST: the fine-pass / batch traversal uses single FMAs wherever the product
 *   is exact (power-of-two factor) and FMNMX for min/max. The per-iteration
 *   corner planes pos*dT - bT stay two roundings: fusing them moved 0.017 % of
 *   the pixels to a neighbouring voxel in measurement, over the 0.01 % budget.
 *   Ray generation, shading and the beam (coarse) pass are individually
 *   rounded in both flavours. Bar: >= 99.99 % identical pixels. */
typedef enum svo_flavour { SVO_FLAVOUR_VALIDATION = 0, SVO_FLAVOUR_FAST = 1 } svo_flavour;
/* OR into `flavour` of svo_raymarch_batch[_device] for INCOHERENT batches (secondary rays: ambient
 * occlusion, reflections): threads take the rays in a direction-binned order (one stable 6-bit radix
 * pass over the directions) instead of submission order, so that a warp's rays share their octant and
 * roughly their direction. Results are identical and land at the rays' own indices; only the speed
 * changes (faster on incoherent batches, a few percent slower on already coherent ones). */
#define SVO_BATCH_COHERENCE_ORDER 0x100
/* Also for incoherent batches, and independent of the order above: LANE REFILL. Persistent warps pull rays from a
 * cursor with one atomic per 256 rays, every lane keeps its traversal state in registers, and once eight lanes of a
 * warp have finished their rays those lanes store their results and start the next rays while the other lanes carry on
 * -- instead of the whole warp waiting for its longest ray. Same words at the same indices. */
#define SVO_BATCH_LANE_REFILL 0x200

/* Ray result codes written to `hit[]`. Non-zero == the reference's `true`. */
enum { SVO_MISS = 0, SVO_HIT_LEAF = 1, SVO_HIT_LOD = 2 };
/* Values stored for rays whose outputs the reference leaves untouched. */
#define SVO_T_MISS 1e10f        /* same sentinel renderBatch uses, Main.cpp:140 */
#define SVO_VOXEL_NONE UINT64_MAX

typedef struct svo_tree svo_tree;

/* ---- library ------------------------------------------------------------- */

SVO_API int svo_abi_version(void);
/* Thread-local, never NULL; valid until the next failing call on this thread. */
SVO_API const char *svo_last_error(void);
SVO_API int svo_device_count(int *count);
/* free() for buffers this library returns (svo_oct_read). */
SVO_API void svo_free(void *p);
/* Page-locked host memory for frame / ray buffers (full-speed PCIe copies). */
SVO_API int svo_host_alloc(size_t bytes, void **out);
SVO_API int svo_host_free(void *p);

/* ---- .oct files, host only (no GPU needed) -------------------------------- *
 * Replaces VoxelOctree::VoxelOctree(const char*) (VoxelOctree.cpp:57-90) and
 * VoxelOctree::save (VoxelOctree.cpp:92-123). Same bytes on disk: float32
 * center[3], uint64 wordCount, then per 64 MiB slice uint64 compSize + one
 * LZ4 block of a single streaming context. Differences:
 * errors are reported, and slice sizes are 64-bit (the reference truncates to
 * int at VoxelOctree.cpp:79, App. E.1). */
SVO_API int svo_oct_read(const char *path, uint32_t **words, uint64_t *n_words, float center[3]);
/* compress != 0: greedy LZ4 matches (any reference build reads the result);
 * compress == 0: literal-only blocks (fastest, largest). */
SVO_API int svo_oct_write(const char *path, const uint32_t *words, uint64_t n_words, const float center[3], int compress);

/* ---- trees ----------------------------------------------------------------- */

typedef struct svo_tree_info {
    uint64_t n_words;
    float center[3];            /* VoxelOctree::center(), VoxelOctree.hpp:55 */
    uint32_t depth;             /* descriptor levels; voxel grid side = 1 << depth */
    int32_t device;
    uint64_t device_bytes;
} svo_tree_info;

/* Adopts a node array (e.g. the one the reference's builder produced,
 * VoxelOctree.cpp:125-137) and uploads it unchanged to `device`'s HBM.
 * The words are copied; the caller keeps ownership of `words`. */
/* Full check of a node array on the host (the reference has none: VoxelOctree.cpp:57-90 loads whatever the file holds
 * and raymarch follows child pointers unchecked, :253-293 -- and so do the kernels here, which size their stack from
 * the depth of the first-child chain). Walks every reachable descriptor: child blocks strictly behind their parent,
 * every descriptor, far word and leaf word inside the array, no branch deeper than the first-child chain, no node
 * reachable more often than the array has words. SVO_ERR_FORMAT + svo_last_error() name the first violation.
 * svo_tree_create_from_words and svo_tree_load_oct run it BY DEFAULT before a tree is handed out (overlapped with the
 * upload; the environment variable SVO_VALIDATE_TREES=0 opts out for trusted arrays); subtrees are walked on all host
 * cores (the 1.6 GB array of an 8192^3 tree: 1.9 s on 8). */
typedef struct svo_words_report {
    uint64_t descriptors;       /* reachable descriptors (Dragon: 29,156) */
    uint64_t leaves;            /* reachable leaf words (Dragon: 90,707) */
    uint64_t far_words;         /* descriptors that carry a far word (Dragon: 24) */
    uint32_t min_leaf_depth, max_leaf_depth;    /* level of the leaves' parents, root = 1: equal in a builder-made tree */
    uint32_t depth;             /* levels of the first-child chain: what the traversal sizes its stack from */
    uint32_t reserved;
} svo_words_report;
SVO_API int svo_words_validate(const uint32_t *words, uint64_t n_words, svo_words_report *report);

SVO_API int svo_tree_create_from_words(const uint32_t *words, uint64_t n_words, const float center[3],
                                       int device, svo_tree **out);
/* VoxelOctree(const char *path), VoxelOctree.cpp:57-90. */
SVO_API int svo_tree_load_oct(const char *path, int device, svo_tree **out);
/* VoxelOctree::save, VoxelOctree.cpp:92-123 (words read back from HBM). */
SVO_API int svo_tree_save_oct(const svo_tree *tree, const char *path, int compress);
SVO_API int svo_tree_get_info(const svo_tree *tree, svo_tree_info *out);
/* Copies the node array back to the host (n_words from svo_tree_get_info). */
SVO_API int svo_tree_download_words(const svo_tree *tree, uint32_t *words_out, uint64_t n_words);
SVO_API int svo_tree_destroy(svo_tree *tree);

/* ---- construction on the GPU -------------------- *
 * Replace VoxelOctree(VoxelData *voxels) / buildOctree (VoxelOctree.cpp:125-205): the node array is
 * built in HBM, word for word the array the reference builds from the same voxels, and stays there.
 * A voxel is filled iff its uint32 material word is non-zero (VoxelData.hpp:106-138). As in the
 * reference, the last z-plane of an odd-depth volume is not seen (VoxelData.cpp:160-163). */

/* Dense grid in host memory, w*h*d words, x fastest: the payload of a raw .voxel file. */
SVO_API int svo_tree_build_from_voxels(const uint32_t *voxels, int w, int h, int d, int device, svo_tree **out);
/* VoxelData(const char *path, size_t mem) + VoxelOctree(VoxelData*), Main.cpp:318-319: raw .voxel file
 * (int32 w, h, d, then the grid; VoxelData.cpp:36-48, 183-201), streamed; no memory budget argument. */
SVO_API int svo_tree_build_from_voxel_file(const char *path, int device, svo_tree **out);
/* The same volume given as n filled voxels: xyz = n (x, y, z) triples, values = n material words, both in
 * host memory. Zero words and coordinates outside w x h x d are ignored; a coordinate named twice is an
 * error. For volumes whose dense form does not fit anywhere (8192^3 = 2 TiB). */
SVO_API int svo_tree_build_from_sparse(const uint32_t *xyz, const uint32_t *values, uint64_t n, int w, int h, int d,
                                       int device, svo_tree **out);

/* PlyLoader(const char *path) + PlyLoader::tris() (PlyLoader.cpp:64-226, PlyLoader.hpp:110): the mesh as the
 * voxeliser sees it. Host only. *tris receives n x 33 floats (pos[3][3], normal[3][3], color[3][3],
 * lower[3], upper[3] per triangle; positions rescaled to the unit box), released with svo_free. */
SVO_API int svo_ply_read_triangles(const char *path, float **tris, uint64_t *n, float lower[3], float upper[3]);
/* Mesh -> tree: PlyLoader(path) + VoxelData(loader, resolution, mem) +
 * VoxelOctree(VoxelData*) -- the in-memory path of the reference's -builder mode (Main.cpp:320-325,
 * PlyLoader.cpp:64-474). The PLY file (ASCII or binary; x y z [nx ny nz] [red green blue]; faces as
 * vertex_indices lists, fans for polygons) is voxelised on the GPU at `resolution` (the reference's
 * --resolution, default 256) and the tree built in HBM. The reference's result also depends on its memory
 * budget (cache block size) and on the size of its thread pool (the cache block is cut into per-thread
 * sub-blocks, and triangle lists and cell centres are computed per sub-block): pass the values the reference
 * run would have used -- mem_budget 0 = its 1 GiB default (Main.cpp:269), threads 0 = this host's hardware
 * thread count (ThreadUtils::idealThreadCount) -- and the node array is the reference's, word for word. */
SVO_API int svo_tree_build_from_ply(const char *path, int resolution, uint64_t mem_budget, int threads, int device,
                                    svo_tree **out);

typedef struct svo_voxelize_stats {
    uint64_t triangles;
    uint64_t cell_records;      /* (cell, triangle) overlaps */
    uint64_t voxels;            /* filled cells */
    int32_t dims[3];            /* volume the reference derives from the mesh (PlyLoader::suggestedDimensions) */
    int32_t cache_block;        /* edge of the reference's cache block for the budget (VoxelData.cpp:203-262) */
    int32_t sub_block[3];       /* per-thread sub-block (PlyLoader.cpp:381-440) */
    int32_t large_triangles;    /* triangles whose bounding box holds more than 2^15 cells: voxelised by a block each */
    float overlap_ms, sort_ms, fold_ms;   /* device time of the voxeliser's three phases */
    float reserved2;
} svo_voxelize_stats;
/* Statistics of the calling thread's last successful svo_tree_build_from_ply call. */
SVO_API int svo_voxelize_last_stats(svo_voxelize_stats *out);

/* The inverse: the filled voxels of a tree, in the builder's (Morton) order. Call with xyz_out ==
 * values_out == NULL to get the count in *n_out, then with host buffers of `capacity` >= count voxels
 * (xyz_out: 3 words per voxel). Coordinates are in the 2^depth grid the tree spans. */
SVO_API int svo_tree_extract_voxels(const svo_tree *tree, uint32_t *xyz_out, uint32_t *values_out, uint64_t capacity,
                                    uint64_t *n_out);
/* build(extract(tree)) without leaving HBM, for a volume of w x h x d voxels (each <= 2^depth). A tree
 * made by the reference's builder or by svo_tree_build_* comes back word for word. */
SVO_API int svo_tree_rebuild(const svo_tree *tree, int w, int h, int d, svo_tree **out);

typedef struct svo_build_stats {
    uint64_t voxels;            /* filled voxels in the tree */
    uint64_t nodes;             /* descriptors */
    uint64_t far_blocks;        /* child blocks with far words (VoxelOctree.cpp:182-193) */
    uint64_t words;             /* node array length */
    float gather_ms;            /* device time: voxels -> (Morton key, material) list */
    float sort_ms;
    float levels_ms;            /* bottom-up sweep */
    float emit_ms;              /* top-down sweep */
} svo_build_stats;
/* Statistics of the calling thread's last successful svo_tree_build_* call. */
SVO_API int svo_build_last_stats(svo_build_stats *out);

/* ---- traversal: VoxelOctree::raymarch (VoxelOctree.cpp:207-346), batched ---- *
 * Ray i: origin o[3i..3i+2], direction d[3i..3i+2] (need not be unit; the
 * octree occupies [1,2]^3), common rayScale. Outputs (any may be NULL):
 *   hit[i]    SVO_MISS / SVO_HIT_LEAF / SVO_HIT_LOD
 *   t[i]      entry t of the leaf (:344) or exit t of the LOD cube (:266); SVO_T_MISS on a miss
 *   normal[i] leaf material word (:282); 0 on a miss or LOD exit
 *   voxel[i]  leaf word index (:282) | for LOD exits parent index + (childShift << 60); SVO_VOXEL_NONE on a miss
 * Host-buffer variant: copies in, runs, copies out, synchronises. */
SVO_API int svo_raymarch_batch(svo_tree *tree, uint64_t n, const float *o, const float *d, float ray_scale,
                               int flavour, uint8_t *hit, float *t, uint32_t *normal, uint64_t *voxel);
/* Device-buffer variant: pointers are device pointers on the tree's device;
 * asynchronous on `stream` (a cudaStream_t, NULL = default stream). */
SVO_API int svo_raymarch_batch_device(svo_tree *tree, uint64_t n, const float *d_o, const float *d_d,
                                      float ray_scale, int flavour, uint8_t *d_hit, float *d_t,
                                      uint32_t *d_normal, uint64_t *d_voxel, void *stream);
/* The reference's single-ray signature mapped to a batch of one; outputs are
 * left untouched exactly where the reference leaves them untouched. Returns
 * the status; *hit_out receives the reference's bool. Not a performance path. */
SVO_API int svo_raymarch(svo_tree *tree, const float o[3], const float d[3], float ray_scale,
                         uint32_t *normal, float *t, int *hit_out);

/* ---- shading: shade + pixel pack (Main.cpp:81-90, 128-132), batched ---------- *
 * For ray i with hit code hit[i] (NULL = every ray hit), material word normal[i] (Util.hpp:86-100) and
 * direction d[3i..] (the direction raymarch was called with), rgba[i] = the reference's pixel:
 * 0xFF000000 | g << 16 | g << 8 | g with g = uint32(min(shade, 1) * 255.0); misses give 0xFF000000, the
 * value renderTile stores for them (Main.cpp:115-118). `light` is the per-frame light vector
 * (Main.cpp:163; svo_frame_constants.light). Host buffers; runs on the tree's device. */
SVO_API int svo_shade_batch(svo_tree *tree, uint64_t n, const uint8_t *hit, const uint32_t *normal, const float *d,
                            const float light[3], uint32_t *rgba);

/* ---- camera: what renderBatch reads from the matrix stacks ------------------ */

typedef struct svo_camera {
    float model[16];            /* MODEL_STACK top, row-major a11..a44 (Mat4.hpp:29-38) */
    float view[16];             /* VIEW_STACK top */
} svo_camera;

/* Main.cpp:212-213,244-245: MODEL = rotXYZ(pitch,0,0)*rotXYZ(0,yaw,0), VIEW = translate(0,0,-radius). */
SVO_API void svo_orbit_camera(float pitch_deg, float yaw_deg, float radius, svo_camera *out);

/* ---- interactive viewer ------------------------
 * The camera control of the reference's `-viewer` mode without SDL: the mouse / key state of
 * src/Events.cpp:28-77 and renderLoop's event handling (src/Main.cpp:229-252) as a state machine the
 * host application feeds with its own window system's events (or with a script: svo_headless --events).
 * Event numbers are SDL 1.2's, the ones the reference receives. */
enum svo_event_type {
    SVO_EVENT_KEY_DOWN = 2, SVO_EVENT_KEY_UP = 3, SVO_EVENT_MOUSE_MOTION = 4, SVO_EVENT_BUTTON_DOWN = 5, SVO_EVENT_BUTTON_UP = 6
};
enum { SVO_BUTTON_LEFT = 1, SVO_BUTTON_RIGHT = 3, SVO_KEY_ESCAPE = 27 };
enum svo_viewer_action {
    SVO_VIEWER_WAIT = 0,        /* keep waiting: mouse motion with no button held (Main.cpp:229) */
    SVO_VIEWER_FRAME = 1,       /* render the next frame with state.camera / state.preview */
    SVO_VIEWER_QUIT = 2         /* Escape is down: doTerminate (Main.cpp:231-234) */
};
typedef struct svo_viewer_event {
    int32_t type;               /* svo_event_type */
    int32_t code;               /* button (SVO_BUTTON_*) or key */
    int32_t dx, dy;             /* relative motion of SVO_EVENT_MOUSE_MOTION (SDL's xrel / yrel) */
} svo_viewer_event;
typedef struct svo_viewer_state {
    float radius, pitch, yaw;   /* Main.cpp:207-209: 1, 0, 0 */
    int32_t mouse_down[2];      /* left, right (Events.cpp:34) */
    int32_t mouse_dx, mouse_dy; /* the last motion not yet read (Events.cpp:31-32, read-and-clear :110-124): motion
                                   that arrives while no button is held is applied by the next button press */
    int32_t escape_down;
    int32_t preview;            /* renderHalfSize (Main.cpp:246,251,253): render with svo_frame_desc.pixel_stride 3 */
    int32_t quit;
    svo_camera camera;          /* MODEL / VIEW for the next frame: identity / translate(0, 0, -1) at first
                                   (Main.cpp:212-213); a drag with the left button replaces MODEL (:244-245), with
                                   the right button VIEW (:250) */
} svo_viewer_state;

SVO_API void svo_viewer_init(svo_viewer_state *state);
/* One event. Returns a svo_viewer_action (negative svo_status on a null argument, sign flipped). */
SVO_API int svo_viewer_feed(svo_viewer_state *state, const svo_viewer_event *event);

/* The scalars renderBatch derives before its loops (Main.cpp:149-163). */
typedef struct svo_frame_constants {
    int32_t width, height, strips, tile_size;
    float pos[3];
    float a11, a12, a21, a22, a31, a32;
    float zx, zy, zz;
    float scale, tile_scale, coarse_scale, aspect;
    float light[3];
    float beam_bias;
} svo_frame_constants;

SVO_API int svo_frame_constants_from_camera(const svo_camera *cam, const float center[3], int width, int height,
                                            int strips, svo_frame_constants *out);

/* ---- frames: renderBatch over all strips (Main.cpp:139-202, 351-362) -------- */

typedef struct svo_frame_desc {
    int32_t width, height;
    int32_t strips;             /* the reference's NumThreads (Main.cpp:57): the image depends on it */
    int32_t flavour;            /* svo_flavour */
    /* Multi-GPU tile interleave: this call renders only the 8x8 tiles whose
     * column index tx satisfies (tx / run) % tile_world == tile_rank (vertical
     * stripes `run` tiles wide, dealt round-robin; run = 4 = 32 pixels unless
     * svo_frame_set_tile_run changed it) and touches no
     * other pixel; its beam pass traces only the tile corners on either side
     * of those stripes (5/4 of 1/tile_world of the corners).
     * Single GPU: tile_rank = 0, tile_world = 1. */
    int32_t tile_rank, tile_world;
    /* renderTile's stride (Main.cpp:92-106): 0 or 1 = every pixel is traced; 3 = the reference's
     * renderHalfSize preview (Main.cpp:161, while the mouse drags): inside every 8x8 tile only pixels at
     * offsets 0, 3, 6 are traced and each other pixel repeats the traced pixel up-left of it. */
    int32_t pixel_stride;
    /* svo_pixel_format of HOST frames delivered by svo_multi_render_sequence / svo_multi_render_frame (every other entry
     * point takes SVO_PIXELS_RGBA8 only). */
    int32_t pixel_format;
} svo_frame_desc;

/* The reference's image is grey with full or no coverage: renderTile stores 0xFF000000 | g << 16 | g << 8 | g
 * (Main.cpp:128-132) and skipped tiles keep the strip memset's 0x00000000 (Main.cpp:165). SVO_PIXELS_GREY8A8 ships exactly
 * that information in two bytes per pixel -- byte 0 = g, byte 1 = alpha (0 or 255) -- packed on the GPU before the
 * device -> host copy: half the bytes on links that are the limit of host-visible frame rates (one PCIe link for 720p
 * frames of a small tree, the box's host links at 4 and 8 GPUs). svo_pixels_expand_grey8a turns it back into the
 * reference's words, bit for bit. */
typedef enum svo_pixel_format { SVO_PIXELS_RGBA8 = 0, SVO_PIXELS_GREY8A8 = 1 } svo_pixel_format;
/* dst[i] = alpha << 24 | g << 16 | g << 8 | g for n pixels (host memory; dst may not overlap src). */
SVO_API void svo_pixels_expand_grey8a(const uint16_t *src, uint64_t n, uint32_t *dst);

typedef struct svo_frame_stats {
    uint64_t coarse_rays;       /* raymarch calls of the beam pass (Main.cpp:181) made by this rank (a subset of
                                 * svo_frame_layout.corners when tile_world >= 3) */
    uint64_t fine_rays;         /* raymarch calls of renderTile (Main.cpp:118) by this rank */
    uint64_t tiles_rendered;
    uint64_t tiles_total;       /* tiles owned by this rank */
    uint32_t kernel_launches;   /* kernels this call put on the stream */
    uint32_t reserved;
    float coarse_ms;            /* device time of the beam-pass kernel (CUDA events on the call's stream) */
    float fine_ms;              /* device time of the fine-pass kernel (+ the tile classifier when a caller-owned depth
                                 * buffer puts every kernel on the call's stream) */
} svo_frame_stats;

/* Geometry of the reference's strip / tile decomposition for one configuration
 * (Main.cpp:351-362), host only. The depth buffer has `corners` floats; tiles are numbered
 * strip by strip, row by row (tile t is in column t % tile_cols); with a tile interleave tile t
 * belongs to rank ((t % tile_cols) / run) % tile_world (svo_frame_tile_owner; run = 4 by default). */
typedef struct svo_frame_layout {
    int32_t n_strips;           /* strips that own at least one row */
    int32_t strip_rows;         /* rows per strip ("stride", Main.cpp:351) */
    int32_t tiles_x;            /* corner columns per strip (Main.cpp:359) */
    int32_t tiles_y_full;       /* corner rows of a full strip (Main.cpp:360) */
    int32_t tiles_y_last;       /* corner rows of the last strip */
    int32_t tile_cols;          /* tiles per tile row */
    int32_t tiles;              /* 8x8 tiles in the frame */
    int32_t corners;            /* beam-pass rays per frame == depth buffer length */
} svo_frame_layout;
SVO_API int svo_frame_get_layout(int width, int height, int strips, svo_frame_layout *out);
/* Pixel rectangle [x0,x1) x [y0,y1) of tile `tile` (clipped to its strip and the image). */
SVO_API int svo_frame_tile_rect(int width, int height, int strips, int tile, int32_t rect[4]);
/* Rank that renders tile `tile` under an interleave over `tile_world` ranks; < 0 on bad arguments. */
SVO_API int svo_frame_tile_owner(int width, int height, int strips, int tile, int tile_world);
/* Width, in 8-pixel tile columns, of the vertical stripes the ranks are dealt (default 4 = 32 pixels; the
 * environment variable SVO_TILE_RUN sets the initial value; run <= 0 restores the default). Process-wide, to be
 * called by every rank with the same value while no frame is in flight. The image does not depend on it. Wider
 * stripes make each rank's rows of pixels longer -- what svo_frame_copy_owned_tiles moves per PCIe write burst:
 * measured on B200, 128-byte runs (the default) reach 28 GB/s into mapped host memory, whole rows 50 GB/s.
 * Read once per call of the single-device entry points that take tile_rank / tile_world (one-process-per-GPU callers);
 * the multi-GPU handle (svo_multi_*) neither reads nor changes it: every sequence carries its own stripe width
 * (svo_sequence_stats.tile_run). */
SVO_API int svo_frame_set_tile_run(int run);

/* Host-buffer variant: rgba (width*height uint32, the reference's backBuffer
 * layout 0xFF000000|b<<16|g<<8|r, Main.cpp:128-134; pitch = width*4) and the
 * optional coarse depth buffer (per strip tilesX*tilesY floats, strip after
 * strip; miss = 1e10f) are HOST pointers; the call uploads the per-frame
 * constants, renders, copies the frame back and synchronises. Pass memory from
 * svo_host_alloc for full-speed copies. With tile_world > 1 pixels of tiles
 * owned by other ranks are returned as they were in the internal buffer
 * (zero on a fresh handle) and only this rank's corners of `depth` are valid. */
SVO_API int svo_render_frame(svo_tree *tree, const svo_camera *cam, const svo_frame_desc *desc,
                             uint32_t *rgba, float *depth, svo_frame_stats *stats);
/* Pipelined host-buffer variant: enqueues the frame and its device->host copies and returns at
 * once with a ticket; svo_frame_wait(ticket) blocks until `rgba` (and `depth`) hold the frame.
 * Up to FOUR frames may be in flight per (width, height, strips) configuration (a fifth is refused),
 * so the copy of frame i and the long-ray tail of its fine pass overlap the rendering of frames
 * i+1.. (and the beam passes of later frames overlap the fine passes of earlier ones). Host buffers must stay valid, and should be page-locked (svo_host_alloc),
 * until the wait returns. svo_render_frame == svo_render_frame_async + svo_frame_wait. */
SVO_API int svo_render_frame_async(svo_tree *tree, const svo_camera *cam, const svo_frame_desc *desc,
                                   uint32_t *rgba, float *depth, int want_stats, int *ticket);
SVO_API int svo_frame_wait(svo_tree *tree, const svo_frame_desc *desc, int ticket, svo_frame_stats *stats);
/* Device-buffer variant: d_rgba may be any pointer the tree's device can
 * write -- local HBM or a peer GPU's framebuffer mapped with svo_ipc_open, in
 * which case finished tiles travel over NVLink as the kernel stores them.
 * Asynchronous: the tile classifier and the fine pass run on `stream`; the beam
 * pass runs on one of the tree's two internal high-priority streams (it only
 * touches internal buffers, a ring of eight frames deep) and `stream` waits for
 * it, so consecutive calls overlap the beam passes of the next frames with the
 * fine pass of the current one. Frames issued on different streams into
 * different framebuffers may overlap entirely (bench.py alternates two). If
 * d_depth is given (device pointer, receives the coarse depth buffer; with
 * tile_world > 1 only the corners this rank needs are written) the beam pass
 * runs on `stream` too. `stats` (optional, host) is filled only when
 * `sync_stats` != 0, which synchronises the stream. */
SVO_API int svo_render_frame_device(svo_tree *tree, const svo_camera *cam, const svo_frame_desc *desc,
                                    uint32_t *d_rgba, float *d_depth, void *stream,
                                    svo_frame_stats *stats, int sync_stats);

/* The host-visible leg of a multi-GPU frame (bench.py `e2e` at N > 1; the reference has one address space and
 * no counterpart): every rank copies the pixels of the tiles IT owns (desc->tile_rank / tile_world) from its
 * framebuffer `d_src` to `d_dst`, a framebuffer of the same layout that may be local HBM, a peer mapping
 * (svo_ipc_open) or page-locked host memory shared by all ranks and mapped with svo_host_register -- then each
 * GPU ships 1 / world of the frame over its own PCIe link, 128 contiguous bytes per warp, instead of rank 0's
 * link carrying all of it. Asynchronous on `stream`. */
SVO_API int svo_frame_copy_owned_tiles(int device, const svo_frame_desc *desc, const uint32_t *d_src, uint32_t *d_dst,
                                       void *stream);
/* Page-locks and maps existing host memory (e.g. a shared-memory segment every rank has mapped) for `device`;
 * *device_ptr is the address kernels and copies use. */
SVO_API int svo_host_register(int device, void *p, size_t bytes, void **device_ptr);
SVO_API int svo_host_unregister(void *p);

/* ---- several GPUs of one node, ONE process ----------------------------------------------------------------
 * Replaces what the reference's viewer does with its strip threads: the strip set-up and thread spawn
 * (Main.cpp:351-367) and the two-phase barrier around every frame (Main.cpp:217-219, ThreadBarrier.cpp:41-59).
 * The node array is replicated into the HBM of every listed device; one worker thread per device enqueues that
 * device's share of every frame (the tile columns (tx / run) % n_devices == index: beam pass of the corners next
 * to them, tile classifier, fine pass); frames are ordered by CUDA events, across devices too -- there is no NCCL,
 * no second process and no host-side wait inside a frame sequence. Up to six frames are in flight.
 *   SVO_OUTPUT_DEVICE  every device's fine pass stores its finished pixels straight into a framebuffer in
 *                      devices[0]'s HBM over NVLink (peer access; the gather is fused into the kernel); the frame
 *                      barrier is devices[0]'s gather stream waiting for every device's frame event.
 *   SVO_OUTPUT_HOST    every device renders its tile columns into a local framebuffer and ships them itself into
 *                      the caller's page-locked host frame (svo_host_alloc): 1 / n_devices of the frame per PCIe
 *                      link, no device-to-device traffic at all.
 * A device may be listed more than once (replicas that share a GPU: how a one-GPU box exercises this path). */
typedef struct svo_multi svo_multi;
enum { SVO_OUTPUT_DEVICE = 0, SVO_OUTPUT_HOST = 1 };

SVO_API int svo_multi_create_from_words(const uint32_t *words, uint64_t n_words, const float center[3],
                                        const int *devices, int n_devices, svo_multi **out);
/* VoxelOctree(const char *path), VoxelOctree.cpp:57-90, once; then replicated. */
SVO_API int svo_multi_load_oct(const char *path, const int *devices, int n_devices, svo_multi **out);
SVO_API int svo_multi_destroy(svo_multi *m);
SVO_API int svo_multi_device_count(const svo_multi *m);
/* The replica on devices[index] (owned by the handle): for svo_raymarch_batch*, svo_tree_get_info, ... */
SVO_API svo_tree *svo_multi_tree(svo_multi *m, int index);

typedef struct svo_sequence_stats {
    uint64_t frames;
    uint64_t coarse_rays;       /* beam rays as the reference issues them: svo_frame_layout.corners per frame */
    uint64_t fine_rays;         /* renderTile's raymarch calls, all devices, all frames */
    uint64_t kernel_launches;   /* kernels enqueued, all devices */
    float device_ms;            /* SVO_OUTPUT_DEVICE: CUDA events on devices[0] around the whole sequence
                                 * (recorded when every worker is ready to enqueue ... last frame gathered); SVO_OUTPUT_HOST: 0 */
    float wall_ms;              /* host clock: every device's worker ready to enqueue ... last frame complete (in HBM / in
                                 * host memory); the workers' wake-up before that is not counted */
    int32_t lanes;              /* frames in flight */
    int32_t tile_run;           /* width of the devices' stripes in 8-pixel tile columns */
} svo_sequence_stats;

/* Called on the calling thread when frame `frame` is complete in host memory, before its buffer is reused. */
typedef void (*svo_frame_callback)(void *user, int frame, const uint32_t *rgba);

/* Renders cams[0 .. n_frames) back to back (the reference's renderLoop, Main.cpp:204-262, over a camera path).
 * desc->tile_rank / tile_world are ignored (the handle deals the tiles). SVO_OUTPUT_HOST: frame k lands in
 * host_frames[k % n_host_frames] (page-locked; width*height words each, or width*height uint16 when desc->pixel_format is
 * SVO_PIXELS_GREY8A8 -- the pointers are then really uint16_t *); min(6, n_host_frames) frames are in
 * flight; on_frame (optional) sees every frame. SVO_OUTPUT_DEVICE: host_frames is ignored; the last frames stay
 * in devices[0]'s HBM (svo_multi_device_frame). `stats` is optional. */
SVO_API int svo_multi_render_sequence(svo_multi *m, const svo_camera *cams, int n_frames, const svo_frame_desc *desc,
                                      int output, uint32_t *const *host_frames, int n_host_frames,
                                      svo_frame_callback on_frame, void *user, svo_sequence_stats *stats);
/* One frame into host memory (any host pointer; page-locked is faster): renderBatch over all strips, all devices. */
SVO_API int svo_multi_render_frame(svo_multi *m, const svo_camera *cam, const svo_frame_desc *desc, uint32_t *rgba,
                                   svo_frame_stats *stats);
/* devices[0]'s framebuffer holding frame (n_frames - 1 - back), back in [0, lanes), of the last
 * SVO_OUTPUT_DEVICE sequence; valid until the next sequence. */
SVO_API int svo_multi_device_frame(svo_multi *m, int back, uint32_t **d_rgba);
/* svo_raymarch_batch with the rays dealt to the devices in contiguous ranges (no exchange: results land in the
 * caller's arrays at the rays' own indices). */
SVO_API int svo_multi_raymarch_batch(svo_multi *m, uint64_t n, const float *o, const float *d, float ray_scale,
                                     int flavour, uint8_t *hit, float *t, uint32_t *normal, uint64_t *voxel);

/* ---- device memory + peer mapping (multi-GPU gather over NVLink) ----------- */

SVO_API int svo_device_alloc(int device, size_t bytes, void **out);
SVO_API int svo_device_free(int device, void *p);
SVO_API int svo_device_memset(int device, void *p, int value, size_t bytes);
SVO_API int svo_device_to_host(int device, void *host_dst, const void *device_src, size_t bytes);
SVO_API int svo_host_to_device(int device, void *device_dst, const void *host_src, size_t bytes);
/* Asynchronous on `stream` (a cudaStream_t); host memory should be page-locked. */
SVO_API int svo_device_to_host_async(int device, void *host_dst, const void *device_src, size_t bytes, void *stream);
SVO_API int svo_device_synchronize(int device);
#define SVO_IPC_HANDLE_BYTES 64
/* Export a device allocation (made with svo_device_alloc) so that another
 * process on the same node can map it; handle is SVO_IPC_HANDLE_BYTES bytes. */
SVO_API int svo_ipc_export(int device, void *p, uint8_t handle[SVO_IPC_HANDLE_BYTES]);
SVO_API int svo_ipc_open(int device, const uint8_t handle[SVO_IPC_HANDLE_BYTES], void **out);
SVO_API int svo_ipc_close(int device, void *p);

#ifdef __cplusplus
}
#endif
#endif /* SVO_B200_H_ */
