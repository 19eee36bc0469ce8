// PLY meshes -> the triangle list the reference's PlyLoader holds after its constructor
// (reference src/PlyLoader.cpp:64-226). Host only. See ply_io.cpp.
#pragma once

#include <cstdint>
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

struct Mesh {
    std::vector<MeshTriangle> tris;
    float lower[3] = {0, 0, 0}, upper[3] = {0, 0, 0};   // rescaled bounds (PlyLoader::_lower / _upper)
};

// status: 0 ok, else the svo_status value (2 io, 3 format)
bool readPlyMesh(const char *path, Mesh &out, std::string &err, int &status);

} // namespace svo
