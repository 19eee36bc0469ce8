// PLY meshes -> the triangle list the reference's PlyLoader holds after its constructor
// (reference src/PlyLoader.cpp:64-226). Host only. See ply_io.cpp.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

namespace svo {

// struct Triangle of the reference (src/PlyLoader.hpp:37-55), flattened: three vertices with position
// (already rescaled to the unit box), normal and colour, plus the bounding box of the positions.
struct MeshTriangle {
    float pos[3][3];
    float normal[3][3];
    float color[3][3];
    float lower[3], upper[3];
};

// struct Vertex of the reference (src/PlyLoader.hpp:30-35): position (rescaled), normal, colour
struct MeshVertex { float pos[3], normal[3], color[3]; };

// malloc'ed array of trivially copyable records that is NOT zero-filled: a 10 M-triangle mesh is hundreds of
// MB, and std::vector's value-initialisation would touch every page once more on one thread
template <typename T>
class PodArray {
public:
    PodArray() = default;
    PodArray(const PodArray &) = delete;
    PodArray &operator=(const PodArray &) = delete;
    ~PodArray() { free(p_); }
    bool resize(size_t n) {                     // contents are kept up to min(old, new) elements
        void *q = realloc(p_, (n ? n : 1)*sizeof(T));
        if (!q) return false;
        p_ = static_cast<T *>(q);
        n_ = n;
        return true;
    }
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    T *data() { return p_; }
    const T *data() const { return p_; }
    T &operator[](size_t i) { return p_[i]; }
    const T &operator[](size_t i) const { return p_[i]; }

private:
    T *p_ = nullptr;
    size_t n_ = 0;
};

// The mesh as the file holds it -- vertices and index triples (polygons already cut into fans) -- which is what
// the voxeliser uploads: 36 B per vertex + 12 B per triangle instead of 132 B per assembled triangle (the
// triangles are assembled on the device, svo_voxelize.cu::assembleTrianglesKernel).
struct Mesh {
    PodArray<MeshVertex> verts;
    PodArray<uint32_t> indices;                          // 3 per triangle, in file order
    bool hasNormals = false;                             // else every triangle gets its face normal (:213-218)
    float lower[3] = {0, 0, 0}, upper[3] = {0, 0, 0};    // rescaled bounds (PlyLoader::_lower / _upper)
    size_t triangleCount() const { return indices.size()/3; }
};

// status: 0 ok, else the svo_status value (2 io, 3 format)
bool readPlyMesh(const char *path, Mesh &out, std::string &err, int &status);

// PlyLoader's `_tris` (Triangle::Triangle, :40-54, with the face normal of :213-218): the host-side assembly,
// on all cores. Used by svo_ply_read_triangles (tests, tools); the build path assembles on the device.
void assembleTriangles(const Mesh &mesh, MeshTriangle *out);

} // namespace svo
