// PLY reader for the voxeliser.
//
// Replaces PlyLoader::PlyLoader / openPly / readVertices / rescaleVertices / readTriangles (reference
// src/PlyLoader.cpp:64-226), which sit on the third-party `plyfile` library. The container is parsed
// here from its specification: header (format ascii | binary_little_endian | binary_big_endian,
// elements, scalar and list properties of the eight PLY types), then element data in header order.
// What is kept from the reference, because it decides the voxels:
//   * vertex properties x y z nx ny nz red green blue, any stored type, converted to float (plyfile's
//     PLY_FLOAT conversion); missing ones default to 0 0 0 / 0 0 0 / 255 255 255 (:131-134);
//   * bounds from the raw positions, then pos = (pos - lower) * (1 / largest extent), bounds scaled by
//     the same factor WITHOUT subtracting lower (:167-181);
//   * faces with more than three indices become triangle fans (v0, v_{k-1}, v_k) (:207-221);
//   * without vertex normals every triangle gets its face normal (:213-218).
// Unlike the reference (ASSERT / silent failure) every problem is reported.
#include "ply_io.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

namespace svo {

namespace {

enum PlyType { kChar, kUchar, kShort, kUshort, kInt, kUint, kFloat, kDouble, kBadType };
const int kSize[] = {1, 1, 2, 2, 4, 4, 4, 8, 0};

PlyType parseType(const std::string &s) {
    static const char *names[8][2] = {{"char", "int8"}, {"uchar", "uint8"}, {"short", "int16"}, {"ushort", "uint16"},
                                      {"int", "int32"}, {"uint", "uint32"}, {"float", "float32"}, {"double", "float64"}};
    for (int t = 0; t < 8; ++t) if (s == names[t][0] || s == names[t][1]) return PlyType(t);
    return kBadType;
}

struct Property { std::string name; PlyType type = kBadType, countType = kBadType; bool isList = false; };
struct Element { std::string name; long long count = 0; std::vector<Property> props; };

struct FileCloser { void operator()(FILE *f) const { if (f) fclose(f); } };

class Reader {
public:
    Reader(FILE *fp, int format) : fp_(fp), format_(format) {}
    bool ok = true;
    double scalar(PlyType type) {
        if (format_ == 0) {
            double v = 0.0;
            if (fscanf(fp_, "%lf", &v) != 1) ok = false;
            return v;
        }
        unsigned char b[8];
        const int n = kSize[type];
        if (fread(b, 1, size_t(n), fp_) != size_t(n)) { ok = false; return 0.0; }
        if (format_ == 2) for (int i = 0; i < n/2; ++i) std::swap(b[i], b[n - 1 - i]);
        switch (type) {
        case kChar: return double(static_cast<signed char>(b[0]));
        case kUchar: return double(b[0]);
        case kShort: { int16_t v; memcpy(&v, b, 2); return v; }
        case kUshort: { uint16_t v; memcpy(&v, b, 2); return v; }
        case kInt: { int32_t v; memcpy(&v, b, 4); return v; }
        case kUint: { uint32_t v; memcpy(&v, b, 4); return v; }
        case kFloat: { float v; memcpy(&v, b, 4); return v; }
        default: { double v; memcpy(&v, b, 8); return v; }
        }
    }
    void skip(const Property &p) {
        if (p.isList) {
            const long long n = (long long)scalar(p.countType);
            for (long long k = 0; k < n && ok; ++k) scalar(p.type);
        } else {
            scalar(p.type);
        }
    }

private:
    FILE *fp_;
    int format_;
};

inline float minStd(float a, float b) { return (b < a) ? b : a; }   // std::min(a, b)
inline float maxStd(float a, float b) { return (a < b) ? b : a; }   // std::max(a, b)

struct Vertex { float pos[3], normal[3], color[3]; };

} // namespace

bool readPlyMesh(const char *path, Mesh &out, std::string &err, int &status) {
    status = 0;
    std::unique_ptr<FILE, FileCloser> fp(fopen(path, "rb"));
    if (!fp) { err = std::string("cannot open ") + path; status = 2; return false; }
    char line[1024];
    if (!fgets(line, sizeof line, fp.get()) || strncmp(line, "ply", 3) != 0) { err = std::string(path) + " is not a PLY file"; status = 3; return false; }
    int format = -1;
    std::vector<Element> elements;
    bool headerEnded = false;
    while (fgets(line, sizeof line, fp.get())) {
        char a[64], b[64], c[64];
        if (!strncmp(line, "end_header", 10)) { headerEnded = true; break; }
        if (sscanf(line, "format %63s", a) == 1) {
            format = !strcmp(a, "ascii") ? 0 : !strcmp(a, "binary_little_endian") ? 1 : !strcmp(a, "binary_big_endian") ? 2 : -1;
        } else if (sscanf(line, "element %63s %63s", a, b) == 2) {
            Element e;
            e.name = a;
            e.count = atoll(b);
            elements.push_back(e);
        } else if (!elements.empty() && sscanf(line, "property list %63s %63s %63s", a, b, c) == 3) {
            Property p;
            p.name = c; p.isList = true; p.countType = parseType(a); p.type = parseType(b);
            if (p.type == kBadType || p.countType == kBadType) { err = std::string("unknown PLY type in: ") + line; status = 3; return false; }
            elements.back().props.push_back(p);
        } else if (!elements.empty() && sscanf(line, "property %63s %63s", a, b) == 2) {
            Property p;
            p.name = b; p.type = parseType(a);
            if (p.type == kBadType) { err = std::string("unknown PLY type in: ") + line; status = 3; return false; }
            elements.back().props.push_back(p);
        }
    }
    if (!headerEnded || format < 0) { err = std::string(path) + ": incomplete PLY header"; status = 3; return false; }
    bool hasVerts = false, hasFaces = false;
    for (const Element &e : elements) { hasVerts |= e.name == "vertex"; hasFaces |= e.name == "face"; }
    if (!hasVerts || !hasFaces) { err = "PLY file has to have triangles and vertices"; status = 3; return false; }   // :100

    static const char *vpNames[9] = {"x", "y", "z", "nx", "ny", "nz", "red", "green", "blue"};
    const float vertDefault[9] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 255.0f, 255.0f, 255.0f};
    std::vector<Vertex> verts;
    bool vertsRead = false, hasNormals = false;
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    Reader in(fp.get(), format);
    out.tris.clear();

    for (const Element &e : elements) {
        if (e.name == "vertex") {
            std::vector<int> slot(e.props.size(), -1);
            bool avail[9] = {false};
            for (size_t p = 0; p < e.props.size(); ++p)
                for (int t = 0; t < 9; ++t)
                    if (!e.props[p].isList && e.props[p].name == vpNames[t]) { slot[p] = t; avail[t] = true; break; }
            hasNormals = avail[3] && avail[4] && avail[5];
            if (e.count < 0 || e.count > 0x7FFFFFFF) { err = "bad vertex count"; status = 3; return false; }
            verts.resize(size_t(e.count));
            for (long long i = 0; i < e.count; ++i) {
                float data[9];
                memcpy(data, vertDefault, sizeof data);
                for (size_t p = 0; p < e.props.size(); ++p) {
                    if (slot[p] >= 0) data[slot[p]] = float(in.scalar(e.props[p].type));
                    else in.skip(e.props[p]);
                }
                if (!in.ok) { err = std::string(path) + ": short read in the vertex data"; status = 3; return false; }
                memcpy(verts[size_t(i)].pos, data, 12);
                memcpy(verts[size_t(i)].normal, data + 3, 12);
                memcpy(verts[size_t(i)].color, data + 6, 12);
                for (int t = 0; t < 3; ++t) { lo[t] = minStd(lo[t], data[t]); hi[t] = maxStd(hi[t], data[t]); }
            }
            // rescaleVertices, :167-181
            const float diff[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
            int largest = 2;
            if (diff[0] > diff[1] && diff[0] > diff[2]) largest = 0;
            else if (diff[1] > diff[2]) largest = 1;
            const float factor = 1.0f/diff[largest];
            for (Vertex &v : verts)
                for (int t = 0; t < 3; ++t) v.pos[t] = (v.pos[t] - lo[t])*factor;
            for (int t = 0; t < 3; ++t) { hi[t] *= factor; lo[t] *= factor; }
            vertsRead = true;
        } else if (e.name == "face") {
            if (!vertsRead) { err = "PLY faces before vertices are not supported"; status = 3; return false; }
            for (long long i = 0; i < e.count; ++i) {
                for (const Property &pr : e.props) {
                    if (!(pr.isList && pr.name == "vertex_indices")) { in.skip(pr); continue; }
                    const long long cnt = (long long)in.scalar(pr.countType);
                    long long v0 = 0, v1 = 0;
                    for (long long k = 0; k < cnt && in.ok; ++k) {
                        const long long idx = (long long)in.scalar(pr.type);
                        if (idx < 0 || size_t(idx) >= verts.size()) { err = "PLY face refers to a vertex that does not exist"; status = 3; return false; }
                        if (k == 0) { v0 = idx; continue; }
                        if (k == 1) { v1 = idx; continue; }
                        MeshTriangle t;
                        const Vertex *vs[3] = {&verts[size_t(v0)], &verts[size_t(v1)], &verts[size_t(idx)]};
                        for (int w = 0; w < 3; ++w) {
                            memcpy(t.pos[w], vs[w]->pos, 12);
                            memcpy(t.normal[w], vs[w]->normal, 12);
                            memcpy(t.color[w], vs[w]->color, 12);
                        }
                        for (int q = 0; q < 3; ++q) {                                   // Triangle::Triangle, :40-54
                            t.lower[q] = minStd(t.pos[0][q], minStd(t.pos[1][q], t.pos[2][q]));
                            t.upper[q] = maxStd(t.pos[0][q], maxStd(t.pos[1][q], t.pos[2][q]));
                        }
                        if (!hasNormals) {                                              // :213-218
                            float e1[3], e2[3];
                            for (int q = 0; q < 3; ++q) { e1[q] = t.pos[1][q] - t.pos[0][q]; e2[q] = t.pos[2][q] - t.pos[0][q]; }
                            const float n[3] = {e1[1]*e2[2] - e1[2]*e2[1], e1[2]*e2[0] - e1[0]*e2[2], e1[0]*e2[1] - e1[1]*e2[0]};
                            const float inv = 1.0f/std::sqrt(n[0]*n[0] + n[1]*n[1] + n[2]*n[2]);
                            for (int w = 0; w < 3; ++w) { t.normal[w][0] = n[0]*inv; t.normal[w][1] = n[1]*inv; t.normal[w][2] = n[2]*inv; }
                        }
                        out.tris.push_back(t);
                        v1 = idx;
                    }
                }
                if (!in.ok) { err = std::string(path) + ": short read in the face data"; status = 3; return false; }
            }
        } else {
            for (long long i = 0; i < e.count && in.ok; ++i)
                for (const Property &pr : e.props) in.skip(pr);
            if (!in.ok) { err = std::string(path) + ": short read in element " + e.name; status = 3; return false; }
        }
    }
    if (out.tris.empty()) { err = std::string(path) + ": no triangles"; status = 3; return false; }
    memcpy(out.lower, lo, 12);
    memcpy(out.upper, hi, 12);
    return true;
}

} // namespace svo
