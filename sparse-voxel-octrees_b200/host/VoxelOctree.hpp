// Drop-in C++ facade with the reference's own class surface
// (reference src/VoxelOctree.hpp:48-57), forwarding to the C ABI in
// include/svo_b200.h. Header-only; link with -lsvo_b200.
//
//   VoxelOctree(const char *path)                  reference src/VoxelOctree.cpp:57
//   VoxelOctree(VoxelData *voxels)                 reference src/VoxelOctree.cpp:125  (with the VoxelData / PlyLoader
//                                                  stand-ins below: the reference's two VoxelData constructors name the
//                                                  source, the tree is voxelised / built in HBM; also fromVoxelFile /
//                                                  fromPly / fromVoxels / fromSparse, and adopt() for a finished array)
//   void save(const char *path)                    reference src/VoxelOctree.cpp:92
//   bool raymarch(o, d, rayScale, normal&, t&)     reference src/VoxelOctree.cpp:207
//   Vec3 center() const                            reference src/VoxelOctree.hpp:55
//
// When this header is included after the reference's math/Vec3.hpp and
// IntTypes.hpp it uses their Vec3 / uint32; otherwise it supplies minimal
// stand-ins. The builder constructor VoxelOctree(VoxelData*) builds the same
// node array on the GPU from the source the VoxelData stand-in names -- a raw
// .voxel file (VoxelData(path, mem), VoxelData.cpp:36-48) or a mesh
// (VoxelData(PlyLoader*, sideLength, mem), VoxelData.cpp:50-56); the factories
// take the same sources directly, or a dense grid / a list of filled voxels
// (what a voxeliser produces), and adopt(words, count, center) uploads an
// array built elsewhere unchanged.
//
// Differences a caller can observe: failures throw std::runtime_error carrying
// svo_last_error() instead of being ignored (the reference silently continues
// after a failed fopen, VoxelOctree.cpp:60,95); raymarch() is a batch of one
// (a PCIe round trip per call) -- use raymarchBatch()/renderFrame() for speed.
#ifndef SVO_B200_VOXELOCTREE_HPP_
#define SVO_B200_VOXELOCTREE_HPP_

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "svo_b200.h"

#ifndef MATH_VEC3_HPP_
struct Vec3 {
    float x, y, z;
    Vec3() : x(0.0f), y(0.0f), z(0.0f) {}
    Vec3(float a) : x(a), y(a), z(a) {}
    Vec3(float _x, float _y, float _z) : x(_x), y(_y), z(_z) {}
};
#endif
#ifndef INTTYPES_HPP_
typedef std::uint32_t uint32;
typedef std::uint64_t uint64;
#endif

// Stand-ins with the reference's constructor signatures for the two ways its `-builder` mode names a voxel source
// (Main.cpp:315-325): they only record the source; VoxelOctree(VoxelData*) below does the work on the GPU. Not
// defined when the reference's own PlyLoader.hpp / VoxelData.hpp were included first (their objects are CPU-side
// voxel caches this library has no use for: build with the factories then).
#ifndef PLYLOADER_HPP_
class PlyLoader {                                   // reference src/PlyLoader.hpp:96, PlyLoader.cpp:64-79
    std::string _path;
public:
    explicit PlyLoader(const char *path) : _path(path) {}
    const std::string &path() const { return _path; }
};
#define SVO_B200_PLYLOADER_STANDIN 1
#endif
#ifndef VOXELDATA_HPP_
class VoxelData {
    std::string _path;
    bool _isPly;
    int _sideLength;
    std::size_t _mem;
public:
    VoxelData(const char *path, std::size_t mem)                      // raw .voxel file, reference src/VoxelData.cpp:37-48
        : _path(path), _isPly(false), _sideLength(0), _mem(mem) {}
#ifdef SVO_B200_PLYLOADER_STANDIN
    VoxelData(PlyLoader *loader, std::size_t sideLength, std::size_t mem)   // mesh, reference src/VoxelData.cpp:50-56
        : _path(loader->path()), _isPly(true), _sideLength(int(sideLength)), _mem(mem) {}
#endif
    const std::string &path() const { return _path; }
    bool isPly() const { return _isPly; }
    int sideLength() const { return _sideLength; }
    std::size_t memoryBudget() const { return _mem; }
};
#define SVO_B200_VOXELDATA_STANDIN 1
#endif

class VoxelOctree {
    svo_tree *_tree;
    svo_tree_info _info;

    static void check(int status, const char *what) {
        if (status != SVO_OK) throw std::runtime_error(std::string(what) + ": " + svo_last_error());
    }
    VoxelOctree(const VoxelOctree &);
    VoxelOctree &operator=(const VoxelOctree &);
    explicit VoxelOctree(svo_tree *tree) : _tree(tree) { check(svo_tree_get_info(_tree, &_info), "svo_tree_get_info"); }

public:
    explicit VoxelOctree(const char *path, int device = 0) : _tree(0) {
        check(svo_tree_load_oct(path, device, &_tree), "VoxelOctree(path)");
        check(svo_tree_get_info(_tree, &_info), "svo_tree_get_info");
    }
#ifdef SVO_B200_VOXELDATA_STANDIN
    // VoxelOctree(VoxelData *voxels), reference src/VoxelOctree.cpp:125-137: the tree of the source `voxels` names,
    // word for word the reference builder's (the memory budget shapes a mesh's result, so it is passed on).
    explicit VoxelOctree(VoxelData *voxels, int device = 0) : _tree(0) {
        if (voxels->isPly())
            check(svo_tree_build_from_ply(voxels->path().c_str(), voxels->sideLength(), voxels->memoryBudget(), 0, device, &_tree),
                  "VoxelOctree(VoxelData*)");
        else
            check(svo_tree_build_from_voxel_file(voxels->path().c_str(), device, &_tree), "VoxelOctree(VoxelData*)");
        check(svo_tree_get_info(_tree, &_info), "svo_tree_get_info");
    }
#endif
    // Node array from the reference's builder (or anywhere else), uploaded unchanged.
    static VoxelOctree *adopt(const uint32 *words, uint64 count, const Vec3 &center, int device = 0) {
        float c[3] = {center.x, center.y, center.z};
        svo_tree *tree = 0;
        check(svo_tree_create_from_words(words, count, c, device, &tree), "VoxelOctree::adopt");
        return new VoxelOctree(tree);
    }
    // VoxelData(path, mem) + VoxelOctree(VoxelData*), Main.cpp:318-319: raw .voxel file -> tree in HBM.
    static VoxelOctree *fromVoxelFile(const char *path, int device = 0) {
        svo_tree *tree = 0;
        check(svo_tree_build_from_voxel_file(path, device, &tree), "VoxelOctree::fromVoxelFile");
        return new VoxelOctree(tree);
    }
    // PlyLoader(path) + VoxelData(loader, resolution, mem) + VoxelOctree(VoxelData*), Main.cpp:320-325. memBudget 0 =
    // the reference's 1 GiB, threads 0 = this host's hardware threads (the reference's pool size shapes the result).
    static VoxelOctree *fromPly(const char *path, int resolution = 256, uint64 memBudget = 0, int threads = 0, int device = 0) {
        svo_tree *tree = 0;
        check(svo_tree_build_from_ply(path, resolution, memBudget, threads, device, &tree), "VoxelOctree::fromPly");
        return new VoxelOctree(tree);
    }
    // Dense w*h*d grid of material words (x fastest, 0 = empty).
    static VoxelOctree *fromVoxels(const uint32 *voxels, int w, int h, int d, int device = 0) {
        svo_tree *tree = 0;
        check(svo_tree_build_from_voxels(voxels, w, h, d, device, &tree), "VoxelOctree::fromVoxels");
        return new VoxelOctree(tree);
    }
    // n filled voxels: xyz = n (x, y, z) triples, values = n material words.
    static VoxelOctree *fromSparse(const uint32 *xyz, const uint32 *values, uint64 n, int w, int h, int d, int device = 0) {
        svo_tree *tree = 0;
        check(svo_tree_build_from_sparse(xyz, values, n, w, h, d, device, &tree), "VoxelOctree::fromSparse");
        return new VoxelOctree(tree);
    }
    ~VoxelOctree() { svo_tree_destroy(_tree); }

    void save(const char *path) { check(svo_tree_save_oct(_tree, path, 1), "VoxelOctree::save"); }

    bool raymarch(const Vec3 &o, const Vec3 &d, float rayScale, uint32 &normal, float &t) {
        float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
        int hit = 0;
        check(svo_raymarch(_tree, oo, dd, rayScale, &normal, &t, &hit), "VoxelOctree::raymarch");
        return hit != 0;
    }

    Vec3 center() const { return Vec3(_info.center[0], _info.center[1], _info.center[2]); }

    // ---- beyond the reference surface: what a GPU needs to be fast ----------
    uint64 wordCount() const { return _info.n_words; }
    uint32 depth() const { return _info.depth; }
    svo_tree *handle() { return _tree; }

    // n rays, host arrays (n x 3 floats each). See svo_raymarch_batch.
    void raymarchBatch(uint64 n, const float *o, const float *d, float rayScale, uint8_t *hit, float *t,
                       uint32 *normal, uint64 *voxel = 0, int flavour = SVO_FLAVOUR_VALIDATION) {
        check(svo_raymarch_batch(_tree, n, o, d, rayScale, flavour, hit, t, normal, voxel), "VoxelOctree::raymarchBatch");
    }

    // For call sites written like the reference's per-pixel loop (`tree->raymarch(...)` per ray, Main.cpp:118,181): queue
    // the rays where raymarch() was called, flush() once (one batch instead of a PCIe round trip per ray), read the
    // results where they were used. Outputs keep the reference's semantics (normal / t untouched where it leaves them).
    class RayQueue {
        std::vector<float> _o, _d;
        std::vector<uint8_t> _hit;
        std::vector<float> _t;
        std::vector<uint32> _normal;
    public:
        std::size_t push(const Vec3 &o, const Vec3 &d) {
            _o.push_back(o.x); _o.push_back(o.y); _o.push_back(o.z);
            _d.push_back(d.x); _d.push_back(d.y); _d.push_back(d.z);
            return _o.size()/3 - 1;
        }
        std::size_t size() const { return _o.size()/3; }
        void flush(VoxelOctree &tree, float rayScale, int flavour = SVO_FLAVOUR_VALIDATION) {
            const std::size_t n = size();
            _hit.assign(n, 0); _t.assign(n, 0.0f); _normal.assign(n, 0);
            if (n) tree.raymarchBatch(n, _o.data(), _d.data(), rayScale, _hit.data(), _t.data(), _normal.data(), 0, flavour);
        }
        // the reference's bool raymarch(..., uint32 &normal, float &t) for queued ray i
        bool result(std::size_t i, uint32 &normal, float &t) const {
            if (_hit[i] != SVO_MISS) t = _t[i];
            if (_hit[i] == SVO_HIT_LEAF) normal = _normal[i];
            return _hit[i] != SVO_MISS;
        }
        void clear() { _o.clear(); _d.clear(); }
    };

    // One frame of the reference's renderBatch loop over `strips` strips (Main.cpp:139-202, 351-362).
    // pixelStride 3 = renderTile's preview mode while dragging (Main.cpp:101-106, 161).
    svo_frame_stats renderFrame(const svo_camera &cam, int width, int height, int strips, uint32 *rgba,
                                int flavour = SVO_FLAVOUR_FAST, int pixelStride = 1) {
        svo_frame_desc desc = {width, height, strips, flavour, 0, 1, pixelStride, 0};
        svo_frame_stats stats;
        check(svo_render_frame(_tree, &cam, &desc, rgba, 0, &stats), "VoxelOctree::renderFrame");
        return stats;
    }
};

#endif
