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
#include <algorithm>
#include <memory>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

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

inline bool isSpace(uint8_t c) { return c == ' ' || c == '\n' || c == '\r' || c == '\t'; }

// strtod for the tokens mesh files are made of, exactly: [-+]digits[.digits][e[-+]digits] with at most 15 significant
// digits and a decimal exponent within +-22. Then the digits are an integer below 2^53 and the power of ten is a double
// too, so ONE correctly rounded multiply or divide gives the correctly rounded result -- what strtod returns
// (Clinger's fast path). Anything else (more digits, inf / nan, hex, junk) is left to strtod.
inline bool fastDecimal(const uint8_t *p, double &out, const uint8_t **end = nullptr) {
    static const double pow10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                     1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    bool negative = false;
    if (*p == '-' || *p == '+') { negative = *p == '-'; ++p; }
    uint64_t mantissa = 0;
    int digits = 0, significant = 0, exponent = 0;
    for (; *p >= '0' && *p <= '9'; ++p, ++digits) {
        if (significant > 0 || *p != '0') { if (++significant > 15) return false; }
        mantissa = mantissa*10 + uint64_t(*p - '0');
    }
    if (*p == '.') {
        ++p;
        for (; *p >= '0' && *p <= '9'; ++p, ++digits, --exponent) {
            if (significant > 0 || *p != '0') { if (++significant > 15) return false; }
            mantissa = mantissa*10 + uint64_t(*p - '0');
        }
    }
    if (digits == 0) return false;
    if (*p == 'e' || *p == 'E') {
        ++p;
        bool expNegative = false;
        if (*p == '-' || *p == '+') { expNegative = *p == '-'; ++p; }
        int e = 0, expDigits = 0;
        for (; *p >= '0' && *p <= '9' && expDigits < 4; ++p, ++expDigits) e = e*10 + (*p - '0');
        if (expDigits == 0 || (*p >= '0' && *p <= '9')) return false;
        exponent += expNegative ? -e : e;
    }
    if (*p != 0 && !isSpace(*p)) return false;
    if (exponent < -22 || exponent > 22) return false;
    double v = double(mantissa);                          // exact: below 10^15 < 2^53
    v = exponent < 0 ? v/pow10[-exponent] : v*pow10[exponent];
    out = negative ? -v : v;
    if (end) *end = p;
    return true;
}

// Cursor over the element data, which is held in memory as a whole: a 10 M-triangle file is 55 M scalars,
// and one fread per scalar costs more than everything the GPU does with the mesh afterwards.
class Reader {
public:
    Reader(const uint8_t *begin, const uint8_t *end, int format) : p_(begin), end_(end), format_(format) {}
    bool ok = true;
    const uint8_t *position() const { return p_; }
    double scalar(PlyType type) {
        if (format_ == 0) {
            while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\r' || *p_ == '\t')) ++p_;
            if (p_ >= end_) { ok = false; return 0.0; }
            double fast;
            const uint8_t *after;
            if (fastDecimal(p_, fast, &after)) { p_ = after; return fast; }
            char *stop = nullptr;
            const double v = strtod(reinterpret_cast<const char *>(p_), &stop);   // the buffer is NUL-terminated
            if (stop == reinterpret_cast<const char *>(p_)) { ok = false; return 0.0; }
            p_ = reinterpret_cast<const uint8_t *>(stop);
            return v;
        }
        const int n = kSize[type];
        if (end_ - p_ < n) { ok = false; return 0.0; }
        const double v = decode(p_, type, format_);
        p_ += n;
        return v;
    }
    void skip(const Property &p) {
        if (p.isList) {
            const long long n = (long long)scalar(p.countType);
            for (long long k = 0; k < n && ok; ++k) scalar(p.type);
        } else {
            scalar(p.type);
        }
    }
    static double decode(const uint8_t *src, PlyType type, int format) {
        unsigned char b[8];
        const int n = kSize[type];
        memcpy(b, src, size_t(n));
        if (format == 2) for (int i = 0; i < n/2; ++i) std::swap(b[i], b[n - 1 - i]);
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

private:
    const uint8_t *p_, *end_;
    int format_;
};

inline float minStd(float a, float b) { return (b < a) ? b : a; }   // std::min(a, b)
inline float maxStd(float a, float b) { return (a < b) ? b : a; }   // std::max(a, b)

int workerCount(size_t items, size_t worthIt = 65536) {
    int n = int(std::thread::hardware_concurrency());
    if (n > 16) n = 16;
    if (n < 1 || items < worthIt) n = 1;
    if (size_t(n) > items && items > 0) n = int(items);
    return n;
}

template <class Fn>
void parallelRanges(size_t items, Fn fn, size_t worthIt = 65536) {      // fn(threadIndex, begin, end)
    const int n = workerCount(items, worthIt);
    std::vector<std::thread> pool;
    for (int t = 1; t < n; ++t) pool.emplace_back(fn, t, items*size_t(t)/size_t(n), items*size_t(t + 1)/size_t(n));
    fn(0, size_t(0), items/size_t(n));
    for (auto &th : pool) th.join();
}

// Triangle::Triangle (:40-54) + the face normal of meshes without vertex normals (:213-218)
inline void makeTriangle(const MeshVertex &a, const MeshVertex &b, const MeshVertex &c, bool hasNormals, MeshTriangle &t) {
    const MeshVertex *vs[3] = {&a, &b, &c};
    for (int w = 0; w < 3; ++w) {
        memcpy(t.pos[w], vs[w]->pos, 12);
        memcpy(t.normal[w], vs[w]->normal, 12);
        memcpy(t.color[w], vs[w]->color, 12);
    }
    for (int q = 0; q < 3; ++q) {
        t.lower[q] = minStd(t.pos[0][q], minStd(t.pos[1][q], t.pos[2][q]));
        t.upper[q] = maxStd(t.pos[0][q], maxStd(t.pos[1][q], t.pos[2][q]));
    }
    if (!hasNormals) {
        float e1[3], e2[3];
        for (int q = 0; q < 3; ++q) { e1[q] = t.pos[1][q] - t.pos[0][q]; e2[q] = t.pos[2][q] - t.pos[0][q]; }
        const float n[3] = {e1[1]*e2[2] - e1[2]*e2[1], e1[2]*e2[0] - e1[0]*e2[2], e1[0]*e2[1] - e1[1]*e2[0]};
        const float inv = 1.0f/std::sqrt(n[0]*n[0] + n[1]*n[1] + n[2]*n[2]);
        for (int w = 0; w < 3; ++w) { t.normal[w][0] = n[0]*inv; t.normal[w][1] = n[1]*inv; t.normal[w][2] = n[2]*inv; }
    }
}

// The element data, whole: binary files are mapped (no copy; the kernel's page cache is the buffer), ASCII
// files are read into a NUL-terminated buffer for strtod.
class FileData {
public:
    ~FileData() {
        if (map_ && map_ != MAP_FAILED) munmap(map_, mapBytes_);
    }
    bool open(const char *path, off_t dataStart, off_t fileEnd, bool binary) {
        const size_t n = size_t(fileEnd - dataStart);
        if (binary && n > 0) {
            const int fd = ::open(path, O_RDONLY);
            if (fd < 0) return false;
            map_ = mmap(nullptr, size_t(fileEnd), PROT_READ, MAP_PRIVATE, fd, 0);
            ::close(fd);
            if (map_ == MAP_FAILED) { map_ = nullptr; return false; }
            mapBytes_ = size_t(fileEnd);
            madvise(map_, mapBytes_, MADV_WILLNEED);
            begin_ = static_cast<const uint8_t *>(map_) + dataStart;
            end_ = begin_ + n;
            return true;
        }
        std::unique_ptr<FILE, FileCloser> fp(fopen(path, "rb"));
        if (!fp || fseeko(fp.get(), dataStart, SEEK_SET) != 0) return false;
        text_.assign(n + 1, 0);
        if (fread(text_.data(), 1, n, fp.get()) != n) return false;
        begin_ = text_.data();
        end_ = begin_ + n;
        return true;
    }
    const uint8_t *begin() const { return begin_; }
    const uint8_t *end() const { return end_; }

private:
    void *map_ = nullptr;
    size_t mapBytes_ = 0;
    std::vector<uint8_t> text_;
    const uint8_t *begin_ = nullptr, *end_ = nullptr;
};


// ASCII element data as a stream of whitespace-separated tokens, cut into chunks at token boundaries with the
// global index of every chunk's first token known (one parallel counting sweep). An element whose records have a
// fixed number of tokens -- scalar properties only, or faces that all turn out to be triangles -- can then be
// parsed on all cores: a token's global index says which record and which property it is.

class AsciiTokens {
public:
    AsciiTokens(const uint8_t *begin, const uint8_t *end) {
        const size_t bytes = size_t(end - begin);
        const int chunks = workerCount(bytes/8)*4;
        starts_.push_back(begin);
        for (int k = 1; k < chunks; ++k) {
            const uint8_t *p = begin + bytes*size_t(k)/size_t(chunks);
            while (p < end && !isSpace(*p)) ++p;          // a token belongs to the chunk it starts in
            if (p > starts_.back()) starts_.push_back(p);
        }
        starts_.push_back(end);
        const size_t n = starts_.size() - 1;
        first_.assign(n + 1, 0);
        parallelRanges(n, [&](int, size_t a, size_t b) {
            for (size_t c = a; c < b; ++c) {
                uint64_t count = 0;
                bool inToken = false;                     // every chunk starts at the data start or on whitespace
                for (const uint8_t *p = starts_[c]; p < starts_[c + 1]; ++p) {
                    const bool sp = isSpace(*p);
                    count += (!sp && !inToken);
                    inToken = !sp;
                }
                first_[c + 1] = count;
            }
        }, 1);
        for (size_t c = 0; c < n; ++c) first_[c + 1] += first_[c];
    }
    uint64_t total() const { return first_.back(); }
    // fn(globalIndex, pointer to the token) for every token with index in [t0, t1), on all cores
    template <class Fn>
    void forRange(uint64_t t0, uint64_t t1, Fn fn) const {
        const size_t n = starts_.size() - 1;
        parallelRanges(n, [&](int th, size_t a, size_t b) {
            for (size_t c = a; c < b; ++c) {
                if (first_[c + 1] <= t0 || first_[c] >= t1) continue;
                uint64_t g = first_[c];
                bool inToken = false;
                for (const uint8_t *p = starts_[c]; p < starts_[c + 1] && g < t1; ++p) {
                    const bool sp = isSpace(*p);
                    if (!sp && !inToken) {
                        if (g >= t0) fn(th, g, p);
                        ++g;
                    }
                    inToken = !sp;
                }
            }
        }, 1);
    }
    // where token `t` starts (the end of the data for t == total())
    const uint8_t *position(uint64_t t) const {
        if (t >= total()) return starts_.back();
        size_t c = 0;
        while (first_[c + 1] <= t) ++c;
        uint64_t g = first_[c];
        bool inToken = false;
        for (const uint8_t *p = starts_[c]; p < starts_[c + 1]; ++p) {
            const bool sp = isSpace(*p);
            if (!sp && !inToken) {
                if (g == t) return p;
                ++g;
            }
            inToken = !sp;
        }
        return starts_.back();
    }

private:
    std::vector<const uint8_t *> starts_;
    std::vector<uint64_t> first_;
};

} // namespace

void assembleTriangles(const Mesh &mesh, MeshTriangle *out) {
    const uint32_t *idx = mesh.indices.data();
    const MeshVertex *v = mesh.verts.data();
    parallelRanges(mesh.triangleCount(), [&](int, size_t begin, size_t end) {
        for (size_t i = begin; i < end; ++i)
            makeTriangle(v[idx[3*i]], v[idx[3*i + 1]], v[idx[3*i + 2]], mesh.hasNormals, out[i]);
    });
}

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

    const off_t dataStart = ftello(fp.get());
    fseeko(fp.get(), 0, SEEK_END);
    const off_t fileEnd = ftello(fp.get());
    fp.reset();
    FileData data;
    if (dataStart < 0 || fileEnd < dataStart || !data.open(path, dataStart, fileEnd, format != 0)) {
        err = std::string("cannot read ") + path; status = 2; return false;
    }
    const uint8_t *const dataEnd = data.end();

    static const char *vpNames[9] = {"x", "y", "z", "nx", "ny", "nz", "red", "green", "blue"};
    const float vertDefault[9] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 255.0f, 255.0f, 255.0f};
    PodArray<MeshVertex> &verts = out.verts;
    bool vertsRead = false;
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    Reader in(data.begin(), dataEnd, format);
    out.indices.resize(0);
    out.hasNormals = false;
    size_t nIndices = 0;
    const char *const noMemory = "out of host memory for the mesh";
    // ASCII: the token index of the current element's first token, for as long as every element so far had a fixed
    // number of tokens per record (see AsciiTokens)
    std::unique_ptr<AsciiTokens> tokens;
    uint64_t tokenCursor = 0;
    bool tokenCursorValid = format == 0;
    if (format == 0) tokens.reset(new AsciiTokens(data.begin(), dataEnd));
    auto parseToken = [](const uint8_t *p, double &v) {       // one whole token, like Reader::scalar would take it
        if (fastDecimal(p, v)) return true;
        char *stop = nullptr;
        v = strtod(reinterpret_cast<const char *>(p), &stop);
        return stop != reinterpret_cast<const char *>(p) && (*stop == 0 || isSpace(uint8_t(*stop)));
    };

    for (const Element &e : elements) {
        if (e.name == "vertex") {
            std::vector<int> slot(e.props.size(), -1);
            bool avail[9] = {false};
            bool fixedSize = format != 0, hasList = false;
            size_t stride = 0;
            std::vector<size_t> offsets(e.props.size(), 0);
            for (size_t p = 0; p < e.props.size(); ++p) {
                for (int t = 0; t < 9; ++t)
                    if (!e.props[p].isList && e.props[p].name == vpNames[t]) { slot[p] = t; avail[t] = true; break; }
                if (e.props[p].isList) { fixedSize = false; hasList = true; }
                offsets[p] = stride;
                stride += size_t(kSize[e.props[p].type]);
            }
            out.hasNormals = avail[3] && avail[4] && avail[5];
            if (e.count < 0 || e.count > 0x7FFFFFFF) { err = "bad vertex count"; status = 3; return false; }
            if (!verts.resize(size_t(e.count))) { err = noMemory; status = 4; return false; }
            auto store = [&](size_t i, const float *v9, float *tlo, float *thi) {
                memcpy(verts[i].pos, v9, 12);
                memcpy(verts[i].normal, v9 + 3, 12);
                memcpy(verts[i].color, v9 + 6, 12);
                for (int t = 0; t < 3; ++t) { tlo[t] = minStd(tlo[t], v9[t]); thi[t] = maxStd(thi[t], v9[t]); }   // :160-163
            };
            if (fixedSize) {
                // binary records of one size: decoded on all cores (min / max do not depend on the order)
                const uint8_t *base = in.position();
                if (size_t(dataEnd - base) < stride*size_t(e.count)) { err = std::string(path) + ": short read in the vertex data"; status = 3; return false; }
                std::vector<float> los(16*3, 1e30f), his(16*3, -1e30f);
                parallelRanges(size_t(e.count), [&](int th, size_t begin, size_t end) {
                    float tlo[3] = {1e30f, 1e30f, 1e30f}, thi[3] = {-1e30f, -1e30f, -1e30f};
                    for (size_t i = begin; i < end; ++i) {
                        float v9[9];
                        memcpy(v9, vertDefault, sizeof v9);
                        const uint8_t *rec = base + i*stride;
                        for (size_t p = 0; p < e.props.size(); ++p)
                            if (slot[p] >= 0) v9[slot[p]] = float(Reader::decode(rec + offsets[p], e.props[p].type, format));
                        store(i, v9, tlo, thi);
                    }
                    memcpy(&los[size_t(th)*3], tlo, 12);
                    memcpy(&his[size_t(th)*3], thi, 12);
                });
                for (int th = 0; th < 16; ++th)
                    for (int t = 0; t < 3; ++t) { lo[t] = minStd(lo[t], los[size_t(th)*3 + t]); hi[t] = maxStd(hi[t], his[size_t(th)*3 + t]); }
                in = Reader(base + stride*size_t(e.count), dataEnd, format);
            } else if (tokenCursorValid && !hasList && tokens->total() >= tokenCursor + uint64_t(e.count)*e.props.size()) {
                // ASCII, scalar properties only: token g of the element is property g % n of vertex g / n
                const uint64_t np = e.props.size(), t0 = tokenCursor, t1 = t0 + uint64_t(e.count)*np;
                parallelRanges(verts.size(), [&](int, size_t begin, size_t end) {
                    for (size_t i = begin; i < end; ++i) memcpy(&verts[i], vertDefault, sizeof vertDefault);
                });
                static_assert(sizeof(MeshVertex) == 9*sizeof(float), "MeshVertex is nine packed floats");
                std::vector<char> badNumber(16, 0);
                tokens->forRange(t0, t1, [&](int th, uint64_t g, const uint8_t *p) {
                    const uint64_t rel = g - t0;
                    const int sl = slot[size_t(rel % np)];
                    if (sl < 0) return;
                    double v;
                    if (!parseToken(p, v)) { badNumber[size_t(th)] = 1; return; }
                    reinterpret_cast<float *>(&verts[size_t(rel/np)])[sl] = float(v);
                });
                for (char b : badNumber) if (b) { err = std::string(path) + ": short read in the vertex data"; status = 3; return false; }
                std::vector<float> los(16*3, 1e30f), his(16*3, -1e30f);
                parallelRanges(verts.size(), [&](int th, size_t begin, size_t end) {
                    float tlo[3] = {1e30f, 1e30f, 1e30f}, thi[3] = {-1e30f, -1e30f, -1e30f};
                    for (size_t i = begin; i < end; ++i)
                        for (int t = 0; t < 3; ++t) { tlo[t] = minStd(tlo[t], verts[i].pos[t]); thi[t] = maxStd(thi[t], verts[i].pos[t]); }
                    memcpy(&los[size_t(th)*3], tlo, 12);
                    memcpy(&his[size_t(th)*3], thi, 12);
                });
                for (int th = 0; th < 16; ++th)
                    for (int t = 0; t < 3; ++t) { lo[t] = minStd(lo[t], los[size_t(th)*3 + t]); hi[t] = maxStd(hi[t], his[size_t(th)*3 + t]); }
                tokenCursor = t1;
                in = Reader(tokens->position(tokenCursor), dataEnd, format);
            } else {
                tokenCursorValid = false;
                for (long long i = 0; i < e.count; ++i) {
                    float v9[9];
                    memcpy(v9, vertDefault, sizeof v9);
                    for (size_t p = 0; p < e.props.size(); ++p) {
                        if (slot[p] >= 0) v9[slot[p]] = float(in.scalar(e.props[p].type));
                        else in.skip(e.props[p]);
                    }
                    if (!in.ok) { err = std::string(path) + ": short read in the vertex data"; status = 3; return false; }
                    store(size_t(i), v9, lo, hi);
                }
            }
            // rescaleVertices, :167-181
            const float diff[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
            int largest = 2;
            if (diff[0] > diff[1] && diff[0] > diff[2]) largest = 0;
            else if (diff[1] > diff[2]) largest = 1;
            const float factor = 1.0f/diff[largest];
            parallelRanges(verts.size(), [&](int, size_t begin, size_t end) {
                for (size_t i = begin; i < end; ++i)
                    for (int t = 0; t < 3; ++t) verts[i].pos[t] = (verts[i].pos[t] - lo[t])*factor;
            });
            for (int t = 0; t < 3; ++t) { hi[t] *= factor; lo[t] *= factor; }
            vertsRead = true;
        } else if (e.name == "face") {
            if (!vertsRead) { err = "PLY faces before vertices are not supported"; status = 3; return false; }
            if (e.count < 0) { err = "bad face count"; status = 3; return false; }
            const Property *listProp = nullptr;
            for (const Property &pr : e.props) if (pr.isList && pr.name == "vertex_indices") listProp = &pr;
            std::vector<char> bad(16, 0);
            const size_t nVerts = verts.size();

            // Fast path -- what mesh exporters write: binary, the index list is the only property and every face
            // is a triangle. If the face data is exactly count * (count field + 3 indices) bytes long and every
            // count field AT THOSE STRIDES reads 3, then (by induction from the first record) that is the parse;
            // checked and decoded on all cores in one sweep. Anything else takes the general path below.
            bool fastDone = false;
            if (format != 0 && listProp && e.props.size() == 1 && &e == &elements.back() && e.count > 0) {
                const size_t cs = size_t(kSize[listProp->countType]), is = size_t(kSize[listProp->type]), rec = cs + 3*is;
                const uint8_t *base = in.position();
                if (size_t(dataEnd - base) == rec*size_t(e.count)) {
                    if (!out.indices.resize(nIndices + 3*size_t(e.count))) { err = noMemory; status = 4; return false; }
                    uint32_t *dst = out.indices.data() + nIndices;
                    std::vector<char> notTriangles(16, 0);
                    parallelRanges(size_t(e.count), [&](int th, size_t begin, size_t end) {
                        for (size_t i = begin; i < end; ++i) {
                            const uint8_t *r = base + i*rec;
                            if (Reader::decode(r, listProp->countType, format) != 3.0) { notTriangles[size_t(th)] = 1; return; }
                            for (int k = 0; k < 3; ++k) {
                                const long long idx = (long long)Reader::decode(r + cs + size_t(k)*is, listProp->type, format);
                                if (idx < 0 || size_t(idx) >= nVerts) { bad[size_t(th)] = 1; return; }
                                dst[3*i + size_t(k)] = uint32_t(idx);
                            }
                        }
                    });
                    bool allTriangles = true;
                    for (char c : notTriangles) allTriangles &= !c;
                    if (allTriangles) {
                        nIndices += 3*size_t(e.count);
                        in = Reader(dataEnd, dataEnd, format);
                        fastDone = true;
                    } else {
                        out.indices.resize(nIndices);
                        std::fill(bad.begin(), bad.end(), 0);
                    }
                }
            }
            // The same idea for ASCII files: if every face is a triangle, token g of the element is field g % 4 of face
            // g / 4 -- true by induction as soon as every token at a multiple of 4 reads 3.
            if (!fastDone && tokenCursorValid && listProp && e.props.size() == 1 && e.count > 0 &&
                tokens->total() >= tokenCursor + 4*uint64_t(e.count)) {
                const uint64_t t0 = tokenCursor, t1 = t0 + 4*uint64_t(e.count);
                if (!out.indices.resize(nIndices + 3*size_t(e.count))) { err = noMemory; status = 4; return false; }
                uint32_t *dst = out.indices.data() + nIndices;
                std::vector<char> notTriangles(16, 0);
                tokens->forRange(t0, t1, [&](int th, uint64_t g, const uint8_t *p) {
                    const uint64_t rel = g - t0;
                    double v;
                    if (!parseToken(p, v)) { notTriangles[size_t(th)] = 1; return; }     // let the general path report it
                    if ((rel & 3) == 0) {
                        if (v != 3.0) notTriangles[size_t(th)] = 1;
                        return;
                    }
                    const long long idx = (long long)v;
                    if (idx < 0 || size_t(idx) >= nVerts) { bad[size_t(th)] = 1; return; }
                    dst[3*size_t(rel >> 2) + size_t(rel & 3) - 1] = uint32_t(idx);
                });
                bool allTriangles = true;
                for (char c : notTriangles) allTriangles &= !c;
                if (allTriangles) {
                    nIndices += 3*size_t(e.count);
                    tokenCursor = t1;
                    in = Reader(tokens->position(tokenCursor), dataEnd, format);
                    fastDone = true;
                } else {
                    out.indices.resize(nIndices);
                    std::fill(bad.begin(), bad.end(), 0);
                }
            }
            if (!fastDone) {
                tokenCursorValid = false;
                // pass 1 (sequential, cheap): where every face's index list starts, how long it is, and where
                // its triangles go (a polygon of k vertices is a fan of k - 2 triangles, :207-221)
                struct FaceRef { const uint8_t *indices; uint32_t count; uint64_t firstTriangle; };
                std::vector<FaceRef> faces;
                std::vector<long long> asciiIndices;        // ASCII: indices parsed in pass 1
                faces.reserve(size_t(e.count));
                uint64_t nTriangles = 0;
                for (long long i = 0; i < e.count; ++i) {
                    FaceRef f = {nullptr, 0, nTriangles};
                    for (const Property &pr : e.props) {
                        if (&pr != listProp) { in.skip(pr); continue; }
                        const long long cnt = (long long)in.scalar(pr.countType);
                        if (!in.ok || cnt < 0 || cnt > 0x7FFFFFFF) { in.ok = false; break; }
                        f.count = uint32_t(cnt);
                        if (format == 0) {
                            f.indices = reinterpret_cast<const uint8_t *>(uintptr_t(asciiIndices.size()));
                            for (long long k = 0; k < cnt && in.ok; ++k) asciiIndices.push_back((long long)in.scalar(pr.type));
                        } else {
                            f.indices = in.position();
                            const size_t bytes = size_t(cnt)*size_t(kSize[pr.type]);
                            if (size_t(dataEnd - in.position()) < bytes) { in.ok = false; break; }
                            in = Reader(in.position() + bytes, dataEnd, format);
                        }
                    }
                    if (!in.ok) { err = std::string(path) + ": short read in the face data"; status = 3; return false; }
                    if (f.count >= 3) nTriangles += f.count - 2;
                    faces.push_back(f);
                }
                // pass 2 (all cores): the fans' index triples, in file order
                if (!out.indices.resize(nIndices + 3*size_t(nTriangles))) { err = noMemory; status = 4; return false; }
                uint32_t *dst = out.indices.data() + nIndices;
                parallelRanges(faces.size(), [&](int th, size_t begin, size_t end) {
                    for (size_t i = begin; i < end; ++i) {
                        const FaceRef &f = faces[i];
                        long long v0 = 0, v1 = 0;
                        for (uint32_t k = 0; k < f.count; ++k) {
                            long long idx;
                            if (format == 0) idx = asciiIndices[size_t(uintptr_t(f.indices)) + k];
                            else idx = (long long)Reader::decode(f.indices + size_t(k)*size_t(kSize[listProp->type]), listProp->type, format);
                            if (idx < 0 || size_t(idx) >= nVerts) { bad[size_t(th)] = 1; break; }
                            if (k == 0) { v0 = idx; continue; }
                            if (k == 1) { v1 = idx; continue; }
                            uint32_t *tri = dst + 3*(size_t(f.firstTriangle) + (k - 2));
                            tri[0] = uint32_t(v0); tri[1] = uint32_t(v1); tri[2] = uint32_t(idx);
                            v1 = idx;
                        }
                    }
                });
                nIndices += 3*size_t(nTriangles);
            }
            for (char b : bad) if (b) { err = "PLY face refers to a vertex that does not exist"; status = 3; return false; }
        } else {
            bool scalarOnly = true;
            for (const Property &pr : e.props) scalarOnly &= !pr.isList;
            if (tokenCursorValid && scalarOnly && e.count >= 0 && tokens->total() >= tokenCursor + uint64_t(e.count)*e.props.size()) {
                tokenCursor += uint64_t(e.count)*e.props.size();
                in = Reader(tokens->position(tokenCursor), dataEnd, format);
                continue;
            }
            tokenCursorValid = false;
            for (long long i = 0; i < e.count && in.ok; ++i)
                for (const Property &pr : e.props) in.skip(pr);
            if (!in.ok) { err = std::string(path) + ": short read in element " + e.name; status = 3; return false; }
        }
    }
    if (out.indices.empty()) { err = std::string(path) + ": no triangles"; status = 3; return false; }
    memcpy(out.lower, lo, 12);
    memcpy(out.upper, hi, 12);
    return true;
}

} // namespace svo
