/*
 * TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from
 * the product (sparse-voxel-octrees_b200/). Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load the library this
 * file builds.
 *
 * What this is: a C-ABI harness around the *reference's own object code*.
 * oracle/build_ref.sh compiles the reference sources where they lie under
 * /root/reference/src (VoxelOctree.cpp, VoxelData.cpp, PlyLoader.cpp, Util.cpp,
 * Debug.cpp, thread/*.cpp, math/*.cpp, third-party/*.c) and this translation
 * unit, which textually includes src/Main.cpp behind the headless SDL shim
 * (oracle/ref_shim/SDL.h), into oracle/_ref/libsvo_ref.so. No reference source
 * is copied into the repository.
 *
 * The only edit applied to Main.cpp (by sed, on the fly, into a temp file that
 * is deleted after compiling) is dropping `const` from the three "adapt this
 * to your platform" constants NumThreads / GWidth / GHeight (Main.cpp:56-60)
 * and from AspectRatio (Main.cpp:62) so that one library serves every
 * (W, H, strips) configuration. The arithmetic is unchanged: the same IEEE
 * single-precision expressions are evaluated at run time instead of being
 * constant-folded.
 *
 * Exposed: tree load/save/build (reference loader, saver, builder), the
 * reference VoxelOctree::raymarch (VoxelOctree.cpp:207-346) one ray or a
 * batch, the reference frame loop renderBatch (Main.cpp:139-202) over the
 * reference's strip decomposition (Main.cpp:351-362), the orbit camera
 * (Main.cpp:212-213,244-250), and compressMaterial (Util.hpp:64-84).
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <ostream>
#include <stack>
#include <string>
#include <thread>
#include <vector>
#include <stdlib.h>
#include <stdio.h>

#include "SDL.h"

/* VoxelOctree keeps its node array private (class default access); the harness
 * needs to adopt / expose it. Same layout, different access keyword. */
#define class struct
#include "VoxelOctree.hpp"
#undef class

#define main reference_main
#include SVO_REF_MAIN_CPP
#undef main

/* ---- the scripted SDL behind oracle/ref_shim/SDL.h (row f4) ---------------
 * src/Events.cpp, src/ThreadBarrier.cpp and Main.cpp's `-viewer` main loop are
 * the reference's own object code; what they see of SDL is an event queue filled
 * from a script and a window whose every SDL_UpdateRect is recorded. */
namespace {
struct ViewerRun {
    std::deque<SDL_Event> queue;
    SDL_Surface *surface = nullptr;
    int W = 0, H = 0, maxFrames = 0, frames = 0;
    int eventsTaken = 0;
    uint32_t *rgba = nullptr;       /* maxFrames x W x H, optional */
    float *models = nullptr, *views = nullptr;
    int32_t *half = nullptr, *eventsAtFrame = nullptr;
};
ViewerRun *g_viewer = nullptr;

SDL_Event escapeKey(int type) {
    SDL_Event e;
    std::memset(&e, 0, sizeof e);
    e.key.type = (unsigned char)type;
    e.key.keysym.sym = SDLK_ESCAPE;
    return e;
}
} // namespace

extern "C" int svo_shim_wait_event(SDL_Event *event) {
    /* an exhausted script presses Escape: the only way out of renderLoop (Main.cpp:231-234) */
    if (!g_viewer || g_viewer->queue.empty()) { *event = escapeKey(SDL_KEYDOWN); return 1; }
    *event = g_viewer->queue.front();
    g_viewer->queue.pop_front();
    g_viewer->eventsTaken++;
    return 1;
}
extern "C" int svo_shim_poll_event(SDL_Event *event) {
    if (!g_viewer || g_viewer->queue.empty()) return 0;
    *event = g_viewer->queue.front();
    g_viewer->queue.pop_front();
    return 1;
}
extern "C" void svo_shim_surface_created(SDL_Surface *surface) { if (g_viewer) g_viewer->surface = surface; }
extern "C" void svo_shim_present(SDL_Surface *surface) {
    ViewerRun *v = g_viewer;
    if (!v || v->frames >= v->maxFrames) { if (v) v->frames++; return; }
    const int k = v->frames++;
    if (v->rgba) std::memcpy(v->rgba + size_t(k)*v->W*v->H, surface->pixels, size_t(v->W)*v->H*4);
    Mat4 model, view;
    MatrixStack::get(MODEL_STACK, model);
    MatrixStack::get(VIEW_STACK, view);
    if (v->models) std::memcpy(v->models + 16*k, model.a, 64);
    if (v->views) std::memcpy(v->views + 16*k, view.a, 64);
    if (v->half) v->half[k] = renderHalfSize ? 1 : 0;      /* what the frame was rendered with */
    if (v->eventsAtFrame) v->eventsAtFrame[k] = v->eventsTaken;
}

namespace {

struct SpinBarrier {
    std::atomic<int> count;
    std::atomic<int> generation;
    int n;
    explicit SpinBarrier(int n_) : count(0), generation(0), n(n_) {}
    void wait() {
        int gen = generation.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
            count.store(0, std::memory_order_relaxed);
            generation.fetch_add(1, std::memory_order_release);
        } else {
            int spins = 0;
            while (generation.load(std::memory_order_acquire) == gen)
                if (++spins > 2000) std::this_thread::yield();
        }
    }
};

void ensurePool() {
    if (!ThreadUtils::pool)
        ThreadUtils::startThreads(ThreadUtils::idealThreadCount()); /* Main.cpp:313 */
}

VoxelOctree *emptyTree() {
    /* The path ctor leaves the object untouched when fopen fails
     * (VoxelOctree.cpp:58-60); that is the only way to get an empty tree. */
    VoxelOctree *tree = new VoxelOctree("/nonexistent/svo_ref_harness");
    tree->_octreeSize = 0;
    return tree;
}

void setMat(StackName n, const float *m) {
    Mat4 mat;
    std::memcpy(mat.a, m, sizeof(float)*16);
    MatrixStack::set(n, mat);
}

} // namespace

extern "C" {

int svoref_abi_version() { return 1; }
int svoref_hardware_threads() { return int(ThreadUtils::idealThreadCount()); }

/* ---- trees ------------------------------------------------------------- */

void *svoref_tree_load(const char *path) {
    FILE *fp = fopen(path, "rb");
    if (!fp) return nullptr;
    fclose(fp);
    return new VoxelOctree(path); /* VoxelOctree.cpp:57-90 */
}

void *svoref_tree_from_words(const uint32_t *words, uint64_t n, const float *center) {
    VoxelOctree *tree = emptyTree();
    tree->_octree.reset(new uint32[n]);
    std::memcpy(tree->_octree.get(), words, n*sizeof(uint32));
    tree->_octreeSize = n;
    tree->_center = Vec3(center[0], center[1], center[2]);
    return tree;
}

/* Main.cpp:316-319 pattern (raw .voxel file -> VoxelData -> VoxelOctree). */
void *svoref_tree_build_voxel_file(const char *path, uint64_t mem) {
    ensurePool();
    FILE *fp = fopen(path, "rb");
    if (!fp) return nullptr;
    fclose(fp);
    std::unique_ptr<VoxelData> data(new VoxelData(path, size_t(mem)));
    return new VoxelOctree(data.get()); /* VoxelOctree.cpp:125-137 */
}

/* Main.cpp:323-325 pattern (PLY -> PlyLoader -> VoxelData -> VoxelOctree). */
void *svoref_tree_build_ply(const char *path, int resolution, uint64_t mem) {
    ensurePool();
    FILE *fp = fopen(path, "rb");
    if (!fp) return nullptr;
    fclose(fp);
    std::unique_ptr<PlyLoader> loader(new PlyLoader(path));
    std::unique_ptr<VoxelData> data(new VoxelData(loader.get(), resolution, size_t(mem)));
    return new VoxelOctree(data.get());
}

/* Main.cpp:316-317: PLY -> raw .voxel file. */
int svoref_ply_to_voxel_file(const char *ply, const char *out, int resolution, uint64_t mem) {
    ensurePool();
    FILE *fp = fopen(ply, "rb");
    if (!fp) return -1;
    fclose(fp);
    std::unique_ptr<PlyLoader> loader(new PlyLoader(ply));
    loader->convertToVolume(out, resolution, size_t(mem));
    return 0;
}

void svoref_tree_save(void *h, const char *path) { static_cast<VoxelOctree *>(h)->save(path); }
uint64_t svoref_tree_word_count(void *h) { return static_cast<VoxelOctree *>(h)->_octreeSize; }
const uint32_t *svoref_tree_words(void *h) { return static_cast<VoxelOctree *>(h)->_octree.get(); }
void svoref_tree_center(void *h, float *out) {
    Vec3 c = static_cast<VoxelOctree *>(h)->center();
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
void svoref_tree_destroy(void *h) { delete static_cast<VoxelOctree *>(h); }

/* ---- material codec (Util.hpp:64-100) ---------------------------------- */

uint32_t svoref_compress_material(const float *n, float shade) {
    return compressMaterial(Vec3(n[0], n[1], n[2]), shade);
}
void svoref_decompress_material(uint32_t word, float *n, float *shade) {
    Vec3 v;
    decompressMaterial(word, v, *shade);
    n[0] = v.x; n[1] = v.y; n[2] = v.z;
}
float svoref_inv_sqrt(float x) { return invSqrt(x); }

/* ---- traversal (VoxelOctree.cpp:207-346) ------------------------------- */

/* `normal` and `t` are preset to the caller's sentinel so that "untouched on a
 * miss / LOD exit" (App. E.3) is observable. */
int svoref_raymarch(void *h, const float *o, const float *d, float rayScale, uint32_t *normal, float *t) {
    uint32 n = *normal;
    float tt = *t;
    bool hit = static_cast<VoxelOctree *>(h)->raymarch(Vec3(o[0], o[1], o[2]), Vec3(d[0], d[1], d[2]), rayScale, n, tt);
    *normal = n;
    *t = tt;
    return hit ? 1 : 0;
}

/* Plain loop over raymarch on contiguous ray ranges, `threads` OS threads
 * (BASELINE.md section 3, config C4 procedure). Returns wall seconds. */
double svoref_raymarch_batch(void *h, uint64_t n, const float *o, const float *d, float rayScale,
        uint8_t *hit, float *t, uint32_t *normal, int threads) {
    VoxelOctree *tree = static_cast<VoxelOctree *>(h);
    if (threads < 1) threads = 1;
    auto body = [&](uint64_t a, uint64_t b) {
        for (uint64_t i = a; i < b; ++i) {
            uint32 nn = normal ? normal[i] : 0;
            float tt = t ? t[i] : 0.0f;
            bool r = tree->raymarch(Vec3(o[3*i], o[3*i + 1], o[3*i + 2]), Vec3(d[3*i], d[3*i + 1], d[3*i + 2]),
                    rayScale, nn, tt);
            if (hit) hit[i] = r;
            if (normal) normal[i] = nn;
            if (t) t[i] = tt;
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) {
        body(0, n);
    } else {
        std::vector<std::thread> pool;
        for (int k = 0; k < threads; ++k)
            pool.emplace_back(body, n*k/threads, n*(k + 1)/threads);
        for (auto &th : pool) th.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

/* ---- camera (Main.cpp:212-213, 244-250; Mat4.cpp:96-125) ---------------- */

void svoref_orbit_camera(float pitchDeg, float yawDeg, float radius, float *model16, float *view16) {
    Mat4 view = Mat4::translate(Vec3(0.0f, 0.0f, -radius));
    Mat4 model = Mat4::rotXYZ(Vec3(pitchDeg, 0.0f, 0.0f))*Mat4::rotXYZ(Vec3(0.0f, yawDeg, 0.0f));
    std::memcpy(model16, model.a, sizeof(float)*16);
    std::memcpy(view16, view.a, sizeof(float)*16);
}

/* The matrix renderBatch reads (Main.cpp:149-150). */
void svoref_inv_modelview(const float *model16, const float *view16, float *out16) {
    setMat(MODEL_STACK, model16);
    setMat(VIEW_STACK, view16);
    Mat4 tform;
    MatrixStack::get(INV_MODELVIEW_STACK, tform);
    std::memcpy(out16, tform.a, sizeof(float)*16);
}

/* ---- frame loop (Main.cpp:139-202 over the strips of Main.cpp:351-362) -- */

/* Renders `numFrames` frames (camera k = models[16k..], views[16k..]) of a
 * W x H image split into `strips` horizontal strips exactly as main() does,
 * on `threads` OS threads that pull strips from a shared counter. `rgba`
 * (W*H words) receives the last frame; `depth` (optional, per strip
 * tilesX*tilesY floats laid out strip after strip) the last coarse buffer;
 * `frameSeconds[k]` the wall time of the parallel section of frame k
 * (BASELINE.md section 3.3). Returns 0, or -1 on bad arguments. */
int svoref_render_frames_subset(void *h, int W, int H, int strips, int stripModulo, int numFrames, const float *models,
        const float *views, int threads, uint32_t *rgba, float *depth, double *frameSeconds);

/* Main.cpp:69,161: renderHalfSize (set while the mouse drags, Main.cpp:246-253) makes renderTile trace
 * every third pixel of a tile and replicate it (stride 3). Sticky until changed. */
static bool g_halfSize = false;
void svoref_set_half_size(int on) { g_halfSize = on != 0; }

int svoref_render_frames(void *h, int W, int H, int strips, int numFrames, const float *models,
        const float *views, int threads, uint32_t *rgba, float *depth, double *frameSeconds) {
    return svoref_render_frames_subset(h, W, H, strips, 1, numFrames, models, views, threads, rgba, depth, frameSeconds);
}

/* Same, but only strips s with s % stripModulo == 0 are rendered (a bounded
 * sample of the frame for timing; the other rows of `rgba` are left alone). */
int svoref_render_frames_subset(void *h, int W, int H, int strips, int stripModulo, int numFrames, const float *models,
        const float *views, int threads, uint32_t *rgba, float *depth, double *frameSeconds) {
    if (stripModulo < 1) return -1;
    if (!h || W < 1 || H < 1 || strips < 1 || numFrames < 1 || !rgba) return -1;
    VoxelOctree *tree = static_cast<VoxelOctree *>(h);
    if (threads < 1) threads = 1;
    if (threads > strips) threads = strips;

    NumThreads = strips;
    GWidth = W;
    GHeight = H;
    AspectRatio = GHeight/(float)GWidth; /* Main.cpp:62 */
    renderHalfSize = g_halfSize;
    doTerminate = false;

    SDL_Surface surface;
    surface.pixels = rgba;
    surface.pitch = W*4;
    surface.w = W;
    surface.h = H;
    backBuffer = &surface;

    /* Main.cpp:351-362 */
    std::vector<BatchData> td(strips);
    std::vector<std::unique_ptr<float[]>> depthBuffers(strips);
    int stride = (GHeight - 1)/NumThreads + 1;
    for (int i = 0; i < strips; i++) {
        td[i].id = i;
        td[i].tree = tree;
        td[i].x0 = 0;
        td[i].x1 = GWidth;
        td[i].y0 = i*stride;
        td[i].y1 = std::min((i + 1)*stride, GHeight);
        td[i].tilesX = (td[i].x1 - td[i].x0 - 1)/TileSize + 2;
        td[i].tilesY = (td[i].y1 - td[i].y0 - 1)/TileSize + 2;
        /* strips past the bottom edge (y0 >= H) have negative heights in the
         * reference too; give them an empty but valid buffer */
        int cells = td[i].y0 < td[i].y1 ? td[i].tilesX*td[i].tilesY : 0;
        depthBuffers[i].reset(new float[std::max(cells, 1)]);
        td[i].depthBuffer = depthBuffers[i].get();
    }

    std::atomic<int> nextStrip(0);
    SpinBarrier barrier(threads);
    auto worker = [&](int id) {
        for (int f = 0; f < numFrames; ++f) {
            std::chrono::steady_clock::time_point t0;
            if (id == 0) {
                setMat(MODEL_STACK, models + 16*f);
                setMat(VIEW_STACK, views + 16*f);
                nextStrip.store(0);
            }
            barrier.wait();
            if (id == 0) t0 = std::chrono::steady_clock::now();
            for (;;) {
                int s = nextStrip.fetch_add(1);
                if (s >= strips) break;
                if (td[s].y0 < td[s].y1 && s % stripModulo == 0)
                    renderBatch(&td[s]);
            }
            barrier.wait();
            if (id == 0 && frameSeconds)
                frameSeconds[f] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
    };
    std::vector<std::thread> pool;
    for (int k = 1; k < threads; ++k) pool.emplace_back(worker, k);
    worker(0);
    for (auto &th : pool) th.join();

    if (depth) {
        size_t off = 0;
        for (int i = 0; i < strips; i++) {
            if (td[i].y0 >= td[i].y1) continue; /* strip owns no rows: no cells */
            int cells = td[i].tilesX*td[i].tilesY;
            std::memcpy(depth + off, td[i].depthBuffer, sizeof(float)*cells);
            off += cells;
        }
    }
    backBuffer = nullptr;
    return 0;
}

/* ---- row f4: the reference's interactive viewer, run headless ------------
 * Runs the reference's own `main -viewer <path>` (Main.cpp:332-376): its loader, its NumThreads render
 * threads and ThreadBarrier, renderLoop's event handling (Main.cpp:204-258) and Events.cpp's mouse state,
 * against the scripted SDL above. `events` is nEvents x 4 int32: {SDL type (2 key down, 3 key up, 4 mouse
 * motion, 5 button down, 6 button up), button (1 left, 3 right) or key, xrel, yrel}; when the script runs
 * out Escape is pressed. Every frame the viewer presents is recorded (up to maxFrames): pixels, the MODEL
 * and VIEW matrices and renderHalfSize it was rendered with, and how many script events had been consumed
 * when it was shown. Returns the number of frames presented, or -1. */
int svoref_viewer_run(const char *octPath, int W, int H, int strips, int nEvents, const int32_t *events, int maxFrames,
                      uint32_t *rgba, float *models, float *views, int32_t *half, int32_t *eventsAtFrame) {
    if (!octPath || W < 1 || H < 1 || strips < 1 || nEvents < 0 || maxFrames < 1 || g_viewer) return -1;
    FILE *fp = fopen(octPath, "rb");
    if (!fp) return -1;
    fclose(fp);
    ViewerRun run;
    run.W = W; run.H = H; run.maxFrames = maxFrames;
    run.rgba = rgba; run.models = models; run.views = views; run.half = half; run.eventsAtFrame = eventsAtFrame;
    for (int i = 0; i < nEvents; ++i) {
        SDL_Event e;
        std::memset(&e, 0, sizeof e);
        const int32_t *ev = events + 4*i;
        switch (ev[0]) {
        case SDL_MOUSEMOTION: e.motion.type = SDL_MOUSEMOTION; e.motion.xrel = ev[2]; e.motion.yrel = ev[3]; break;
        case SDL_MOUSEBUTTONDOWN: case SDL_MOUSEBUTTONUP: e.button.type = (unsigned char)ev[0]; e.button.button = (unsigned char)ev[1]; break;
        case SDL_KEYDOWN: case SDL_KEYUP:
            if (ev[1] < 0 || ev[1] >= SDLK_LAST) return -1;
            e.key.type = (unsigned char)ev[0]; e.key.keysym.sym = ev[1]; break;
        default: return -1;       /* SDL_QUIT would exit() the process (Events.cpp:75-76) */
        }
        run.queue.push_back(e);
    }
    NumThreads = strips;
    GWidth = W;
    GHeight = H;
    AspectRatio = GHeight/(float)GWidth;     /* Main.cpp:62 */
    renderHalfSize = false;                  /* a fresh process: static storage */
    g_viewer = &run;
    std::string path(octPath);
    char arg0[] = "sparse-voxel-octrees", arg1[] = "-viewer";
    char *argv[] = {arg0, arg1, &path[0], nullptr};
    reference_main(3, argv);
    /* Events.cpp keeps its state in statics: release Escape and the buttons, read the speeds away */
    run.queue.clear();
    run.queue.push_back(escapeKey(SDL_KEYUP));
    for (int b : {SDL_BUTTON_LEFT, SDL_BUTTON_RIGHT}) {
        SDL_Event e;
        std::memset(&e, 0, sizeof e);
        e.button.type = SDL_MOUSEBUTTONUP;
        e.button.button = (unsigned char)b;
        run.queue.push_back(e);
    }
    checkEvents();
    getMouseXSpeed();
    getMouseYSpeed();
    g_viewer = nullptr;
    if (run.surface) { free(run.surface->pixels); free(run.surface); }
    backBuffer = nullptr;
    delete barrier;
    barrier = nullptr;
    return run.frames;
}

} // extern "C"
