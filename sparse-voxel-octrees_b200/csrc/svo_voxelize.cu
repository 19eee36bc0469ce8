// Triangle voxelisation on the GPU -- SURVEY.md section 8, "next" row f3.
//
// What it computes is fixed by the reference: the material word of every cell a triangle touches, as
// PlyLoader::processBlock / triangleToVolume / writeTriangleCell leave it in VoxelData's cache block
// (reference src/PlyLoader.cpp:296-374, 442-472) for buildOctree to consume, including everything that
// result depends on besides the mesh:
//   * the cache block edge VoxelData::init derives from the memory budget (src/VoxelData.cpp:203-262) and
//     the per-thread sub-blocks findBestBlockPartition cuts it into (src/PlyLoader.cpp:381-440): a
//     triangle is offered to a sub-block by a block-level triBoxOverlap (:235-277), and inside a
//     sub-block the cell centres are running sums that start at the triangle's bounding box CLIPPED to
//     that sub-block (:337-374);
//   * _sideLength = sideLength - 2 (:415) with pointToGrid / hx subtracting 2 again (:228-233, :352);
//   * per cell, a sequential fold over its triangles in triangle order: compress -> decompress -> blend
//     -> compress, counts saturating at 255 (:296-335).
// How it computes it is not: one thread per triangle walks the sub-blocks and cells exactly like the
// reference's loops (same float operations in the same order; the library is built with -fmad=false,
// IEEE divide and square root), first counting, then writing (cell, normal, shade) records at offsets
// from a prefix sum -- so records are in triangle order; one STABLE radix sort by cell keeps that order
// inside every cell; one thread per cell folds its run. The filled cells go straight into the octree
// builder (svo_build.cu) as a sparse list. Nothing of the volume ever exists densely.
//
// Not mirrored: a triangle whose sub-block range reaches past the sub-block grid, so that the flat index
// x + gridW*(y + gridH*z) names a block of the next row (the reference would blend it into that block as well).
// It cannot happen for a mesh this loader or the reference's produced: positions are rescaled into [0, 1]^3 and
// pointToGrid maps 1 to sideLength - 3; tests/test_oracle_pins.py::test_block_lists_never_alias_another_sub_block
// counts such listings in the restatement of the reference's loop (zero for every mesh, block size and pool size).
#include "svo_voxelize.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "svo_traverse.cuh"   // invSqrtQuake

namespace svo {

namespace {

constexpr int kThreads = 128;

struct Partition {
    int sideM2;                 // the member _sideLength = sideLength - 2 (:415)
    int volumeW, volumeH, volumeD;
    int block;                  // cache block edge (cubic: VoxelData.cpp:53-54)
    int subW, subH, subD, partW, partH, partD;
    int gridW, gridH, gridD;    // the reference's (over-estimated) sub-block grid, :431-433
    int realW, realH, realD;    // sub-blocks that exist: partW * number of cache blocks per axis
};

struct CellRecord { float nx, ny, nz, shade; };

__device__ __forceinline__ float minStd(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float maxStd(float a, float b) { return (a < b) ? b : a; }

__device__ __forceinline__ void cross3(const float *a, const float *b, float *out) {
    out[0] = a[1]*b[2] - a[2]*b[1];
    out[1] = a[2]*b[0] - a[0]*b[2];
    out[2] = a[0]*b[1] - a[1]*b[0];
}
__device__ __forceinline__ float length3(const float *a) { return sqrtf(a[0]*a[0] + a[1]*a[1] + a[2]*a[2]); }

// ---- triBoxOverlap (reference src/third-party/tribox3.c): separating-axis test, same operations ----

__device__ __forceinline__ bool axisSeparates(float pa, float pb, float rad) {
    float mn, mx;
    if (pa < pb) { mn = pa; mx = pb; } else { mn = pb; mx = pa; }
    return mn > rad || mx < -rad;
}

__device__ bool triBoxOverlap(const float *c, const float *h, const float (*tv)[3]) {
    float v0[3], v1[3], v2[3], e0[3], e1[3], e2[3];
    for (int q = 0; q < 3; ++q) { v0[q] = tv[0][q] - c[q]; v1[q] = tv[1][q] - c[q]; v2[q] = tv[2][q] - c[q]; }
    for (int q = 0; q < 3; ++q) { e0[q] = v1[q] - v0[q]; e1[q] = v2[q] - v1[q]; e2[q] = v0[q] - v2[q]; }

    float fex = fabsf(e0[0]), fey = fabsf(e0[1]), fez = fabsf(e0[2]);
    if (axisSeparates(e0[2]*v0[1] - e0[1]*v0[2], e0[2]*v2[1] - e0[1]*v2[2], fez*h[1] + fey*h[2])) return false;
    if (axisSeparates(-e0[2]*v0[0] + e0[0]*v0[2], -e0[2]*v2[0] + e0[0]*v2[2], fez*h[0] + fex*h[2])) return false;
    if (axisSeparates(e0[1]*v2[0] - e0[0]*v2[1], e0[1]*v1[0] - e0[0]*v1[1], fey*h[0] + fex*h[1])) return false;

    fex = fabsf(e1[0]); fey = fabsf(e1[1]); fez = fabsf(e1[2]);
    if (axisSeparates(e1[2]*v0[1] - e1[1]*v0[2], e1[2]*v2[1] - e1[1]*v2[2], fez*h[1] + fey*h[2])) return false;
    if (axisSeparates(-e1[2]*v0[0] + e1[0]*v0[2], -e1[2]*v2[0] + e1[0]*v2[2], fez*h[0] + fex*h[2])) return false;
    if (axisSeparates(e1[1]*v0[0] - e1[0]*v0[1], e1[1]*v1[0] - e1[0]*v1[1], fey*h[0] + fex*h[1])) return false;

    fex = fabsf(e2[0]); fey = fabsf(e2[1]); fez = fabsf(e2[2]);
    if (axisSeparates(e2[2]*v0[1] - e2[1]*v0[2], e2[2]*v1[1] - e2[1]*v1[2], fez*h[1] + fey*h[2])) return false;
    if (axisSeparates(-e2[2]*v0[0] + e2[0]*v0[2], -e2[2]*v1[0] + e2[0]*v1[2], fez*h[0] + fex*h[2])) return false;
    if (axisSeparates(e2[1]*v2[0] - e2[0]*v2[1], e2[1]*v1[0] - e2[0]*v1[1], fey*h[0] + fex*h[1])) return false;

    for (int q = 0; q < 3; ++q) {
        float mn = v0[q], mx = v0[q];
        if (v1[q] < mn) mn = v1[q];
        if (v1[q] > mx) mx = v1[q];
        if (v2[q] < mn) mn = v2[q];
        if (v2[q] > mx) mx = v2[q];
        if (mn > h[q] || mx < -h[q]) return false;
    }
    float normal[3], vmin[3], vmax[3];
    cross3(e0, e1, normal);
    for (int q = 0; q < 3; ++q) {                       // planeBoxOverlap
        const float v = v0[q];
        if (normal[q] > 0.0f) { vmin[q] = -h[q] - v; vmax[q] = h[q] - v; }
        else { vmin[q] = h[q] - v; vmax[q] = -h[q] - v; }
    }
    if (normal[0]*vmin[0] + normal[1]*vmin[1] + normal[2]*vmin[2] > 0.0f) return false;
    return normal[0]*vmax[0] + normal[1]*vmax[1] + normal[2]*vmax[2] >= 0.0f;
}

// The last of triBoxOverlap's thirteen tests (planeBoxOverlap) on its own, same operations: true = the triangle's plane
// misses the box, so triBoxOverlap is false. The separating-axis test is a conjunction of independent tests, so running
// this one first changes no result; for a large triangle it rejects nearly every cell of the bounding box at a quarter
// of the cost.
__device__ __forceinline__ bool planeMissesBox(const float *c, const float *h, const float (*tv)[3]) {
    float v0[3], v1[3], v2[3], e0[3], e1[3];
    for (int q = 0; q < 3; ++q) { v0[q] = tv[0][q] - c[q]; v1[q] = tv[1][q] - c[q]; v2[q] = tv[2][q] - c[q]; }
    for (int q = 0; q < 3; ++q) { e0[q] = v1[q] - v0[q]; e1[q] = v2[q] - v1[q]; }
    float normal[3], vmin[3], vmax[3];
    cross3(e0, e1, normal);
    for (int q = 0; q < 3; ++q) {
        const float v = v0[q];
        if (normal[q] > 0.0f) { vmin[q] = -h[q] - v; vmax[q] = h[q] - v; }
        else { vmin[q] = h[q] - v; vmax[q] = -h[q] - v; }
    }
    if (normal[0]*vmin[0] + normal[1]*vmin[1] + normal[2]*vmin[2] > 0.0f) return true;
    return !(normal[0]*vmax[0] + normal[1]*vmax[1] + normal[2]*vmax[2] >= 0.0f);
}

// ---- material codec (reference src/Util.hpp:64-100) ----

__device__ uint32_t compressMaterial(const float *n, float shade) {
    uint32_t face = 0;
    float dominant = fabsf(n[0]);
    if (fabsf(n[1]) > dominant) { dominant = fabsf(n[1]); face = 1; }
    if (fabsf(n[2]) > dominant) { dominant = fabsf(n[2]); face = 2; }
    const float nf = face == 0 ? n[0] : face == 1 ? n[1] : n[2];
    const float na = face == 0 ? n[1] : face == 1 ? n[2] : n[0];
    const float nb = face == 0 ? n[2] : face == 1 ? n[0] : n[1];
    const uint32_t sign = nf < 0.0f;
    const float n1 = na/dominant, n2 = nb/dominant;
    const int ui = int((n1*0.5f + 0.5f)*2047.0f), vi = int((n2*0.5f + 0.5f)*2047.0f), ci = int(shade*127.0f);
    const uint32_t u = uint32_t(min(ui, 0x7FF)), v = uint32_t(min(vi, 0x7FF)), c = uint32_t(min(ci, 0x7F));
    return (sign << 31) | (face << 29) | (u << 18) | (v << 7) | c;
}

__device__ void decompressMaterial(uint32_t word, float *n, float &shade) {
    const uint32_t face = (word >> 29) & 3u;
    const float s = (word & 0x80000000u) ? -1.0f : 1.0f;
    const float u = float((word >> 18) & 0x7FFu)*4.8852e-4f*2.0f - 1.0f;
    const float v = float((word >> 7) & 0x7FFu)*4.8852e-4f*2.0f - 1.0f;
    n[0] = (face == 0) ? s : ((face == 1) ? v : u);
    n[1] = (face == 0) ? u : ((face == 1) ? s : v);
    n[2] = (face == 0) ? v : ((face == 1) ? u : s);
    const float inv = invSqrtQuake(n[0]*n[0] + n[1]*n[1] + n[2]*n[2]);       // fastNormalization
    n[0] *= inv; n[1] *= inv; n[2] *= inv;
    shade = float(word & 0x7Fu)*1.0f/127.0f;
}

// ---- the reference's loops, one triangle per thread ----

__device__ __forceinline__ void pointToGrid(const Partition &P, const float *p, int &x, int &y, int &z) {   // :228-233
    x = int(p[0]*float(P.sideM2 - 2) + 1.0f);
    y = int(p[1]*float(P.sideM2 - 2) + 1.0f);
    z = int(p[2]*float(P.sideM2 - 2) + 1.0f);
}

// writeTriangleCell up to the point where the cell's previous content matters (:296-316)
__device__ void cellContribution(const MeshTriangle &t, float cx, float cy, float cz, CellRecord &rec) {
    float f1[3], f2[3], f3[3], a[3], b[3], c[3];
    const float p[3] = {cx, cy, cz};
    for (int q = 0; q < 3; ++q) { f1[q] = t.pos[0][q] - p[q]; f2[q] = t.pos[1][q] - p[q]; f3[q] = t.pos[2][q] - p[q]; }
    for (int q = 0; q < 3; ++q) { a[q] = t.pos[0][q] - t.pos[1][q]; b[q] = t.pos[0][q] - t.pos[2][q]; }
    cross3(a, b, c);
    const float area = length3(c);
    cross3(f2, f3, c);
    float l1 = length3(c)/area;
    cross3(f3, f1, c);
    float l2 = length3(c)/area;
    if (!(l1 >= 0.0f && l2 >= 0.0f && l1 + l2 <= 1.0f)) {
        l1 = minStd(maxStd(l1, 0.0f), 1.0f);
        l2 = minStd(maxStd(l2, 0.0f), 1.0f);
        const float tau = l1 + l2;
        if (tau > 1.0f) { l1 /= tau; l2 /= tau; }
    }
    const float l3 = 1.0f - l1 - l2;
    float n[3], col[3];
    for (int q = 0; q < 3; ++q) {
        n[q] = t.normal[0][q]*l1 + t.normal[1][q]*l2 + t.normal[2][q]*l3;
        col[q] = t.color[0][q]*l1 + t.color[1][q]*l2 + t.color[2][q]*l3;
    }
    const float inv = 1.0f/sqrtf(n[0]*n[0] + n[1]*n[1] + n[2]*n[2]);
    rec.nx = n[0]*inv; rec.ny = n[1]*inv; rec.nz = n[2]*inv;
    rec.shade = (col[0]*0.2126f + col[1]*0.7152f + col[2]*0.0722f)*(1.0f/256.0f);
}

// triangleToVolume for the sub-block (gx, gy, gz) (:337-374). WRITE == false only counts.
template <bool WRITE>
__device__ uint32_t cellsOfSubBlock(const Partition &P, const MeshTriangle &t, int gx, int gy, int gz, uint64_t *keys,
                                    CellRecord *records, uint64_t at) {
    if (gx < 0 || gy < 0 || gz < 0 || gx >= P.realW || gy >= P.realH || gz >= P.realD) return 0;
    const int bufferX = (gx/P.partW)*P.block, bufferY = (gy/P.partH)*P.block, bufferZ = (gz/P.partD)*P.block;
    const int bufferW = min(P.block, P.volumeW - bufferX), bufferH = min(P.block, P.volumeH - bufferY),
              bufferD = min(P.block, P.volumeD - bufferZ);
    if (bufferW <= 0 || bufferH <= 0 || bufferD <= 0) return 0;
    const int offX = (gx % P.partW)*P.subW, offY = (gy % P.partH)*P.subH, offZ = (gz % P.partD)*P.subD;
    int lx, ly, lz, ux, uy, uz;
    pointToGrid(P, t.lower, lx, ly, lz);
    pointToGrid(P, t.upper, ux, uy, uz);
    lx = max(lx, bufferX + offX);
    ly = max(ly, bufferY + offY);
    lz = max(lz, bufferZ + offZ);
    ux = min(ux, bufferX + min(offX + P.subW, bufferW) - 1);
    uy = min(uy, bufferY + min(offY + P.subH, bufferH) - 1);
    uz = min(uz, bufferZ + min(offZ + P.subD, bufferD) - 1);
    if (lx > ux || ly > uy || lz > uz) return 0;
    const float hx = 1.0f/float(P.sideM2 - 2);
    const float half[3] = {0.5f*hx, 0.5f*hx, 0.5f*hx};
    float center[3];
    uint32_t n = 0;
    center[2] = (float(lz) - 0.5f)*hx;
    for (int z = lz; z <= uz; z++, center[2] += hx) {
        center[1] = (float(ly) - 0.5f)*hx;
        for (int y = ly; y <= uy; y++, center[1] += hx) {
            center[0] = (float(lx) - 0.5f)*hx;
            for (int x = lx; x <= ux; x++, center[0] += hx) {
                if (!triBoxOverlap(center, half, t.pos)) continue;
                if (WRITE) {
                    keys[at + n] = uint64_t(x) + uint64_t(P.volumeW)*(uint64_t(y) + uint64_t(P.volumeH)*uint64_t(z));
                    cellContribution(t, center[0], center[1], center[2], records[at + n]);
                }
                ++n;
            }
        }
    }
    return n;
}

// iterateOverlappingBlocks (:235-277) with triangleToVolume as its body
template <bool WRITE>
__device__ uint32_t cellsOfTriangle(const Partition &P, const MeshTriangle &t, uint64_t *keys, CellRecord *records, uint64_t at) {
    int lx, ly, lz, ux, uy, uz;
    pointToGrid(P, t.lower, lx, ly, lz);
    pointToGrid(P, t.upper, ux, uy, uz);
    const int lgx = lx/P.subW, lgy = ly/P.subH, lgz = lz/P.subD;
    const int ugx = (ux + 1)/P.subW, ugy = (uy + 1)/P.subH, ugz = (uz + 1)/P.subD;
    const int maxSide = max(ugx - lgx, max(ugy - lgy, ugz - lgz));
    uint32_t n = 0;
    if (maxSide > 0) {
        const float hx = float(P.subW)/float(P.sideM2 - 2), hy = float(P.subH)/float(P.sideM2 - 2), hz = float(P.subD)/float(P.sideM2 - 2);
        const float half[3] = {0.5f*hx, 0.5f*hy, 0.5f*hz};
        float center[3];
        center[2] = (float(lgz) + 0.5f)*hz;
        for (int z = lgz; z <= ugz; ++z, center[2] += hz) {
            center[1] = (float(lgy) + 0.5f)*hy;
            for (int y = lgy; y <= ugy; ++y, center[1] += hy) {
                center[0] = (float(lgx) + 0.5f)*hx;
                for (int x = lgx; x <= ugx; ++x, center[0] += hx)
                    if (triBoxOverlap(center, half, t.pos)) n += cellsOfSubBlock<WRITE>(P, t, x, y, z, keys, records, at + n);
            }
        }
    } else {
        n += cellsOfSubBlock<WRITE>(P, t, lgx, lgy, lgz, keys, records, at);
    }
    return n;
}

// ---- large triangles: a block per triangle --------------------------------------------------------------------------
// One thread walking every cell of a triangle's bounding box (cellsOfTriangle) is fine for meshes of small triangles --
// the benchmark's are a few cells each -- but a low-poly mesh at a high resolution has triangles whose boxes hold 10^6 to
// 10^9 cells: one thread would run for minutes with the rest of the GPU idle. Triangles whose box holds more than
// kLargeCells cells are listed by countCellsKernel and handled by a whole block each: the block's threads test the
// sub-blocks of the triangle's range in parallel, then, sub-block by sub-block, its cells. The reference's arithmetic is
// kept to the bit: cell (and sub-block) centres are RUNNING SUMS along each axis (center += hx, :337-374 / :235-277), so
// one thread per axis accumulates them into a shared table the way the reference's loops do and every test reads its three
// coordinates from the tables. Records of one triangle may land in any order inside the triangle's range (a cell appears
// at most once per triangle; the fold only depends on the order of TRIANGLES within a cell).
constexpr uint64_t kLargeCells = 1u << 15;
constexpr int kLargeThreads = 256;
constexpr int kAxisTable = 1024;          // longest running-sum table per axis (sub-block range / cells of a sub-block)
constexpr int kSubList = 2048;            // overlapping sub-blocks collected per round

__device__ __forceinline__ uint64_t boxCells(const Partition &P, const MeshTriangle &t) {
    int lx, ly, lz, ux, uy, uz;
    pointToGrid(P, t.lower, lx, ly, lz);
    pointToGrid(P, t.upper, ux, uy, uz);
    if (ux < lx || uy < ly || uz < lz) return 0;
    return uint64_t(ux - lx + 1)*uint64_t(uy - ly + 1)*uint64_t(uz - lz + 1);
}

// running sums first + k*step, k = 0 .. n-1, accumulated like `for (...; v += step)` -- by ONE thread
__device__ __forceinline__ void runningSums(float *table, float first, float step, int n) {
    float v = first;
    for (int k = 0; k < n; ++k, v += step) table[k] = v;
}

template <bool WRITE>
__global__ void __launch_bounds__(kLargeThreads)
largeTrianglesKernel(Partition P, const MeshTriangle *__restrict__ tris, const uint32_t *__restrict__ largeList,
                     const uint32_t *__restrict__ largeCount, int slabs, uint64_t *itemCounts, const uint64_t *__restrict__ itemOffsets,
                     const uint64_t *__restrict__ offsets, uint64_t *keys, CellRecord *records) {
    __shared__ float subX[kAxisTable], subY[kAxisTable], subZ[kAxisTable];     // sub-block centres of the current window
    __shared__ float tabX[kAxisTable], tabY[kAxisTable], tabZ[kAxisTable];     // cell centres of the current sub-block (window)
    __shared__ int subList[kSubList];
    __shared__ int subCount;
    __shared__ unsigned long long written, blockTotal;
    // a work item is (large triangle, slab): the triangle's cell range along z is cut into `slabs` pieces so that a mesh of a
    // few huge triangles still fills the GPU; every block of a triangle repeats the (cheap) sub-block tests and takes the cells
    // of its slab only -- with the running sums still started where the reference starts them
    const uint32_t nItems = *largeCount*uint32_t(slabs);
    for (uint32_t item = blockIdx.x; item < nItems; item += gridDim.x) {
        const uint32_t ti = largeList[item/uint32_t(slabs)];
        const int slab = int(item % uint32_t(slabs));
        const MeshTriangle &t = tris[ti];
        int lx, ly, lz, ux, uy, uz;
        pointToGrid(P, t.lower, lx, ly, lz);
        pointToGrid(P, t.upper, ux, uy, uz);
        const long long zCells = (long long)uz - lz + 1;
        const int slabLo = lz + int(zCells*slab/slabs), slabHi = lz + int(zCells*(slab + 1)/slabs) - 1;
        const int lgx = lx/P.subW, lgy = ly/P.subH, lgz = lz/P.subD;
        const int ugx = (ux + 1)/P.subW, ugy = (uy + 1)/P.subH, ugz = (uz + 1)/P.subD;
        const bool ranged = max(ugx - lgx, max(ugy - lgy, ugz - lgz)) > 0;       // else: the one sub-block, untested (:272-276)
        const int nx = ranged ? ugx - lgx + 1 : 1, ny = ranged ? ugy - lgy + 1 : 1, nz = ranged ? ugz - lgz + 1 : 1;
        __syncthreads();
        if (threadIdx.x == 0) { written = 0; blockTotal = 0; }
        unsigned long long mine = 0;
        const float shx = float(P.subW)/float(P.sideM2 - 2), shy = float(P.subH)/float(P.sideM2 - 2), shz = float(P.subD)/float(P.sideM2 - 2);
        const float subHalf[3] = {0.5f*shx, 0.5f*shy, 0.5f*shz};
        const float hx = 1.0f/float(P.sideM2 - 2);
        const float half[3] = {0.5f*hx, 0.5f*hx, 0.5f*hx};
        // Ranges longer than a table are walked window by window. The reference's running sum of an axis restarts at the
        // range's first entry for every row, so a window's first value is that start plus (window offset) additions of the
        // step -- accumulated one by one by the thread that fills the table.
        for (int wz = 0; wz < nz; wz += kAxisTable) for (int wy = 0; wy < ny; wy += kAxisTable) for (int wx = 0; wx < nx; wx += kAxisTable) {
            const int cxn = min(kAxisTable, nx - wx), cyn = min(kAxisTable, ny - wy), czn = min(kAxisTable, nz - wz);
            __syncthreads();
            if (ranged) {
                if (threadIdx.x == 0) { float v = (float(lgx) + 0.5f)*shx; for (int k = 0; k < wx; ++k) v += shx; runningSums(subX, v, shx, cxn); }
                if (threadIdx.x == 32) { float v = (float(lgy) + 0.5f)*shy; for (int k = 0; k < wy; ++k) v += shy; runningSums(subY, v, shy, cyn); }
                if (threadIdx.x == 64) { float v = (float(lgz) + 0.5f)*shz; for (int k = 0; k < wz; ++k) v += shz; runningSums(subZ, v, shz, czn); }
            }
            __syncthreads();
            const long long windowSubs = (long long)cxn*cyn*czn;
            for (long long base = 0; base < windowSubs; base += kSubList) {
                // phase A: which sub-blocks of this slice does the triangle overlap (:235-277)? all threads, one test each
                __syncthreads();
                if (threadIdx.x == 0) subCount = 0;
                __syncthreads();
                const long long sliceEnd = min(windowSubs, base + (long long)kSubList);
                for (long long q = base + threadIdx.x; q < sliceEnd; q += kLargeThreads) {
                    bool overlaps = true;
                    if (ranged) {
                        const float c[3] = {subX[int(q % cxn)], subY[int((q/cxn) % cyn)], subZ[int(q/((long long)cxn*cyn))]};
                        overlaps = !planeMissesBox(c, subHalf, t.pos) && triBoxOverlap(c, subHalf, t.pos);
                    }
                    if (overlaps) subList[atomicAdd(&subCount, 1)] = int(q - base);
                }
                __syncthreads();
                const int listed = subCount;
                // phase B: the cells of every listed sub-block (:337-374), all threads on one sub-block at a time
                for (int e = 0; e < listed; ++e) {
                    const long long q = base + subList[e];
                    const int gx = lgx + wx + int(q % cxn), gy = lgy + wy + int((q/cxn) % cyn), gz = lgz + wz + int(q/((long long)cxn*cyn));
                    if (gx < 0 || gy < 0 || gz < 0 || gx >= P.realW || gy >= P.realH || gz >= P.realD) continue;
                    const int bufferX = (gx/P.partW)*P.block, bufferY = (gy/P.partH)*P.block, bufferZ = (gz/P.partD)*P.block;
                    const int bufferW = min(P.block, P.volumeW - bufferX), bufferH = min(P.block, P.volumeH - bufferY),
                              bufferD = min(P.block, P.volumeD - bufferZ);
                    if (bufferW <= 0 || bufferH <= 0 || bufferD <= 0) continue;
                    const int offX = (gx % P.partW)*P.subW, offY = (gy % P.partH)*P.subH, offZ = (gz % P.partD)*P.subD;
                    const int clx = max(lx, bufferX + offX), cly = max(ly, bufferY + offY), clz = max(lz, bufferZ + offZ);
                    const int cux = min(ux, bufferX + min(offX + P.subW, bufferW) - 1), cuy = min(uy, bufferY + min(offY + P.subH, bufferH) - 1),
                              cuz = min(uz, bufferZ + min(offZ + P.subD, bufferD) - 1);
                    if (clx > cux || cly > cuy || clz > cuz) continue;      // (block-uniform: every thread takes the same branch)
                    const int zlo = max(clz, slabLo), zhi = min(cuz, slabHi);
                    if (zlo > zhi) continue;
                    for (int vz = zlo; vz <= zhi; vz += kAxisTable) for (int vy = cly; vy <= cuy; vy += kAxisTable) for (int vx = clx; vx <= cux; vx += kAxisTable) {
                        const int mx = min(kAxisTable, cux - vx + 1), my = min(kAxisTable, cuy - vy + 1), mz = min(kAxisTable, zhi - vz + 1);
                        __syncthreads();        // the cell tables of the previous sub-block are no longer read
                        if (threadIdx.x == 96) { float v = (float(clx) - 0.5f)*hx; for (int k = clx; k < vx; ++k) v += hx; runningSums(tabX, v, hx, mx); }
                        if (threadIdx.x == 128) { float v = (float(cly) - 0.5f)*hx; for (int k = cly; k < vy; ++k) v += hx; runningSums(tabY, v, hx, my); }
                        if (threadIdx.x == 160) { float v = (float(clz) - 0.5f)*hx; for (int k = clz; k < vz; ++k) v += hx; runningSums(tabZ, v, hx, mz); }
                        __syncthreads();
                        const long long cells = (long long)mx*my*mz;
                        for (long long c = threadIdx.x; c < cells; c += kLargeThreads) {
                            const int ix = int(c % mx), iy = int((c/mx) % my), iz = int(c/((long long)mx*my));
                            const float center[3] = {tabX[ix], tabY[iy], tabZ[iz]};
                            if (planeMissesBox(center, half, t.pos) || !triBoxOverlap(center, half, t.pos)) continue;
                            if (WRITE) {
                                const unsigned long long at = offsets[ti] + itemOffsets[item] + atomicAdd(&written, 1ull);
                                keys[at] = uint64_t(vx + ix) + uint64_t(P.volumeW)*(uint64_t(vy + iy) + uint64_t(P.volumeH)*uint64_t(vz + iz));
                                cellContribution(t, center[0], center[1], center[2], records[at]);
                            }
                            ++mine;
                        }
                    }
                }
            }
        }
        if (!WRITE) {
            atomicAdd(&blockTotal, mine);
            __syncthreads();
            if (threadIdx.x == 0) itemCounts[item] = blockTotal;
        }
    }
}

// per large triangle: where each of its slabs' records start inside the triangle's range, and the triangle's total
__global__ void __launch_bounds__(kThreads)
largeOffsetsKernel(const uint32_t *__restrict__ largeList, const uint32_t *__restrict__ largeCount, int slabs,
                   const uint64_t *__restrict__ itemCounts, uint64_t *itemOffsets, uint64_t *counts) {
    const uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= *largeCount) return;
    uint64_t run = 0;
    for (int sIdx = 0; sIdx < slabs; ++sIdx) {
        itemOffsets[size_t(j)*slabs + sIdx] = run;
        run += itemCounts[size_t(j)*slabs + sIdx];
    }
    counts[largeList[j]] = run;
}

__global__ void __launch_bounds__(kThreads)
countCellsKernel(Partition P, const MeshTriangle *__restrict__ tris, uint32_t nTris, uint64_t *counts, uint32_t *largeList,
                 uint32_t *largeCount) {
    const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= nTris) return;
    if (boxCells(P, tris[i]) > kLargeCells) {       // a block's job (largeTrianglesKernel)
        largeList[atomicAdd(largeCount, 1u)] = i;
        counts[i] = 0;
        return;
    }
    counts[i] = cellsOfTriangle<false>(P, tris[i], nullptr, nullptr, 0);
}

__global__ void __launch_bounds__(kThreads)
writeCellsKernel(Partition P, const MeshTriangle *__restrict__ tris, uint32_t nTris, const uint64_t *__restrict__ offsets,
                 uint64_t *keys, CellRecord *records) {
    const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= nTris) return;
    if (boxCells(P, tris[i]) > kLargeCells) return; // written by largeTrianglesKernel<true>
    cellsOfTriangle<true>(P, tris[i], keys, records, offsets[i]);
}

__global__ void __launch_bounds__(256)
iotaKernel(uint32_t n, uint32_t *out) {
    const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

// One thread per run of equal cell keys (sorted, triangle order kept): writeTriangleCell's fold (:318-334).
__global__ void __launch_bounds__(256)
foldCellsKernel(Partition P, uint32_t n, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ order,
                const CellRecord *__restrict__ records, uint32_t *xyz, uint32_t *values, unsigned long long *cursor) {
    const uint32_t k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint64_t key = keys[k];
    if (k > 0 && keys[k - 1] == key) return;
    uint32_t data = 0, count = 0;
    for (uint32_t j = k; j < n && keys[j] == key; ++j) {
        const CellRecord r = records[order[j]];
        const float nrm[3] = {r.nx, r.ny, r.nz};
        if (data == 0) {
            count = 1;
            data = compressMaterial(nrm, r.shade);
        } else {
            const float currentRatio = float(count)/(float(count) + 1.0f);
            const float newRatio = 1.0f - currentRatio;
            float cn[3], cs;
            decompressMaterial(data, cn, cs);
            float nn[3] = {cn[0]*currentRatio + nrm[0]*newRatio, cn[1]*currentRatio + nrm[1]*newRatio, cn[2]*currentRatio + nrm[2]*newRatio};
            const float ns = cs*currentRatio + r.shade*newRatio;
            if (nn[0]*nn[0] + nn[1]*nn[1] + nn[2]*nn[2] < 1e-3f) { nn[0] = cn[0]; nn[1] = cn[1]; nn[2] = cn[2]; }
            data = compressMaterial(nn, ns);
            count = min(count + 1u, 255u);
        }
    }
    if (data == 0) return;
    const unsigned long long at = atomicAdd(cursor, 1ull);
    const uint64_t plane = uint64_t(P.volumeW)*uint64_t(P.volumeH);
    const uint32_t z = uint32_t(key/plane);
    const uint32_t rem = uint32_t(key - uint64_t(z)*plane);
    xyz[3*at] = rem % uint32_t(P.volumeW);
    xyz[3*at + 1] = rem/uint32_t(P.volumeW);
    xyz[3*at + 2] = z;
    values[at] = data;
}

// Triangle::Triangle (reference src/PlyLoader.hpp:40-54) and, for meshes without vertex normals, the face
// normal of PlyLoader::readTriangles (:213-218), one thread per triangle from the uploaded vertices and index
// triples: the same float operations in the same order as ply_io.cpp::makeTriangle (explicitly rounded, IEEE
// square root and divide), so the triangles are the ones the host-side assembly produces, bit for bit.
__global__ void __launch_bounds__(kThreads)
assembleTrianglesKernel(const MeshVertex *__restrict__ verts, const uint32_t *__restrict__ indices, uint32_t nTris,
                        int hasNormals, MeshTriangle *__restrict__ tris) {
    const uint32_t i = blockIdx.x*kThreads + threadIdx.x;
    if (i >= nTris) return;
    MeshTriangle t;
    for (int w = 0; w < 3; ++w) {
        const MeshVertex v = verts[indices[3*size_t(i) + w]];
        for (int q = 0; q < 3; ++q) { t.pos[w][q] = v.pos[q]; t.normal[w][q] = v.normal[q]; t.color[w][q] = v.color[q]; }
    }
    for (int q = 0; q < 3; ++q) {
        t.lower[q] = minStd(t.pos[0][q], minStd(t.pos[1][q], t.pos[2][q]));
        t.upper[q] = maxStd(t.pos[0][q], maxStd(t.pos[1][q], t.pos[2][q]));
    }
    if (!hasNormals) {
        float e1[3], e2[3], n[3];
        for (int q = 0; q < 3; ++q) { e1[q] = __fsub_rn(t.pos[1][q], t.pos[0][q]); e2[q] = __fsub_rn(t.pos[2][q], t.pos[0][q]); }
        n[0] = __fsub_rn(__fmul_rn(e1[1], e2[2]), __fmul_rn(e1[2], e2[1]));
        n[1] = __fsub_rn(__fmul_rn(e1[2], e2[0]), __fmul_rn(e1[0], e2[2]));
        n[2] = __fsub_rn(__fmul_rn(e1[0], e2[1]), __fmul_rn(e1[1], e2[0]));
        const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(n[0], n[0]), __fmul_rn(n[1], n[1])), __fmul_rn(n[2], n[2]));
        const float inv = __fdiv_rn(1.0f, __fsqrt_rn(len2));
        for (int w = 0; w < 3; ++w)
            for (int q = 0; q < 3; ++q) t.normal[w][q] = __fmul_rn(n[q], inv);
    }
    tris[i] = t;
}

// Scratch arrays, stream-ordered on the default stream, from the builder's pool (svo_build.cuh): what the
// voxeliser releases (27 GB for the 8192^3 mesh) is what the build that follows allocates from, instead of going
// back to the driver and being mapped again. Measured on the 8192^3 mesh, three builds in a row, wall time of
// svo_tree_build_from_ply: 0.98 / 0.61 / 2.22 s with the shared pool against 3.75 / 1.80 / 1.07 s with
// cudaMalloc / cudaFree here (the builder's sort and level stages then spend 0.3 s each growing its pool); the
// slow third build was the synchronous trim of the pool, which now runs on a background thread.
template <typename T>
struct Dev {
    T *p = nullptr;
    Dev() = default;
    Dev(const Dev &) = delete;
    Dev &operator=(const Dev &) = delete;
    ~Dev() { release(); }
    void release() { if (p) cudaFreeAsync(p, 0); p = nullptr; }
    cudaError_t alloc(uint64_t n) {
        release();
        const size_t bytes = size_t(n ? n : 1)*sizeof(T);
        cudaMemPool_t pool = buildScratchPool();
        if (!pool) return cudaMalloc(&p, bytes);
        return cudaMallocFromPoolAsync(reinterpret_cast<void **>(&p), bytes, pool, 0);
    }
};

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~Timer() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    void start() { cudaEventRecord(a, 0); }
    float stop() { cudaEventRecord(b, 0); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
};

uint64_t countCellsInHierarchicalGrid(int numLevels) {           // VoxelData.cpp:59-68
    uint64_t result = 0, size = 1;
    for (int i = 0; i < numLevels; ++i) { result += size; size *= 8; }
    return result;
}

int *pickMax(int &w, int &h, int &d) { if (w > h && w > d) return &w; else if (h > d) return &h; else return &d; }
int *pickMin(int &w, int &h, int &d) { if (w < h && w < d) return &w; else if (h < d) return &h; else return &d; }
int *pickMedian(int &w, int &h, int &d) {
    const int mx = std::max(w, std::max(h, d)), mn = std::min(w, std::min(h, d));
    if (w != mn && w != mx) return &w; else if (h != mn && h != mx) return &h; else return &d;
}

#define SVO_VOX_CUDA(call)                                                            \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            err = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
            return false;                                                             \
        }                                                                             \
    } while (0)

} // namespace

bool voxelizeMesh(const Mesh &mesh, int sideLength, uint64_t memoryBudget, int threadCount, OctreeBuilder &builder,
                  VoxelizeStats &stats, std::string &err) {
    if (sideLength < 8 || sideLength > (1 << 21)) { err = "resolution must be in [8, 2^21]"; return false; }
    if (threadCount < 1) { err = "thread count must be positive"; return false; }
    if (mesh.triangleCount() == 0) { err = "mesh has no triangles"; return false; }
    if (mesh.triangleCount() >= (1ull << 31)) { err = "more than 2^31 - 1 triangles are not supported"; return false; }

    // PlyLoader::suggestedDimensions (:498-503)
    const float sx = (mesh.upper[0] - mesh.lower[0])*float(sideLength - 2), sy = (mesh.upper[1] - mesh.lower[1])*float(sideLength - 2),
                sz = (mesh.upper[2] - mesh.lower[2])*float(sideLength - 2);
    const int w = int(sx) + 2, h = int(sy) + 2, d = int(sz) + 2;
    if (!builder.begin(w, h, d, err)) return false;

    // VoxelData::init (VoxelData.cpp:203-262): the largest cache block the budget allows
    int side = 1, highestBit = 0;
    while (side < w || side < h || side < d) { side <<= 1; ++highestBit; }
    const uint64_t cellCost = 4 + 1;
    int largestLowerLevel = -1;
    for (int i = highestBit; i >= 0; --i) {
        const int lowBit = highestBit - i;
        const uint64_t cost = countCellsInHierarchicalGrid(i + 1) + countCellsInHierarchicalGrid(lowBit) + (uint64_t(1) << uint64_t(lowBit*3))*cellCost;
        if (cost < memoryBudget) largestLowerLevel = highestBit - i;
    }
    if (largestLowerLevel < 0) { err = "memory budget too small for the reference's smallest cache block"; return false; }

    // PlyLoader::setupBlockProcessing (:407-440)
    Partition P;
    P.sideM2 = sideLength - 2;
    P.volumeW = w; P.volumeH = h; P.volumeD = d;
    P.block = 1 << largestLowerLevel;
    P.subW = P.subH = P.subD = P.block;
    for (int used = 1; used < threadCount; used *= 2) {                          // findBestBlockPartition, :381-405
        if ((*pickMax(P.subW, P.subH, P.subD) % 2) == 0) *pickMax(P.subW, P.subH, P.subD) /= 2;
        else if ((*pickMedian(P.subW, P.subH, P.subD) % 2) == 0) *pickMedian(P.subW, P.subH, P.subD) /= 2;
        else if ((*pickMin(P.subW, P.subH, P.subD) % 2) == 0) *pickMin(P.subW, P.subH, P.subD) /= 2;
        else break;
    }
    P.partW = P.block/P.subW; P.partH = P.block/P.subH; P.partD = P.block/P.subD;
    P.gridW = P.partW*(w + P.block - 1)/P.block;
    P.gridH = P.partH*(h + P.block - 1)/P.block;
    P.gridD = P.partD*(d + P.block - 1)/P.block;
    P.realW = P.partW*((w + P.block - 1)/P.block);
    P.realH = P.partH*((h + P.block - 1)/P.block);
    P.realD = P.partD*((d + P.block - 1)/P.block);
    stats.triangles = mesh.triangleCount();
    stats.dims[0] = w; stats.dims[1] = h; stats.dims[2] = d;
    stats.cacheBlock = P.block;
    stats.subBlock[0] = P.subW; stats.subBlock[1] = P.subH; stats.subBlock[2] = P.subD;

    const bool debugTiming = getenv("SVO_BUILD_DEBUG") != nullptr;
    const auto wall0 = std::chrono::steady_clock::now();
    auto wallMs = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count(); };
    const uint32_t nTris = uint32_t(mesh.triangleCount());
    const unsigned blocks = (nTris + kThreads - 1)/kThreads;
    Dev<MeshTriangle> dTris;
    Dev<uint64_t> dCounts;
    SVO_VOX_CUDA(dTris.alloc(nTris));
    {   // vertices + index triples up (48 B per triangle instead of 132), triangles assembled in HBM
        Dev<MeshVertex> dVerts;
        Dev<uint32_t> dIndices;
        SVO_VOX_CUDA(dVerts.alloc(mesh.verts.size()));
        SVO_VOX_CUDA(dIndices.alloc(mesh.indices.size()));
        SVO_VOX_CUDA(cudaMemcpyAsync(dVerts.p, mesh.verts.data(), mesh.verts.size()*sizeof(MeshVertex), cudaMemcpyHostToDevice, 0));
        SVO_VOX_CUDA(cudaMemcpyAsync(dIndices.p, mesh.indices.data(), mesh.indices.size()*sizeof(uint32_t), cudaMemcpyHostToDevice, 0));
        assembleTrianglesKernel<<<blocks, kThreads>>>(dVerts.p, dIndices.p, nTris, mesh.hasNormals ? 1 : 0, dTris.p);
        SVO_VOX_CUDA(cudaGetLastError());
        SVO_VOX_CUDA(cudaStreamSynchronize(0));      // the host arrays may go once this returns
    }
    const double uploadedAt = wallMs();
    SVO_VOX_CUDA(dCounts.alloc(uint64_t(nTris) + 1));
    SVO_VOX_CUDA(cudaMemset(dCounts.p + nTris, 0, sizeof(uint64_t)));

    Timer timer;
    timer.start();
    Dev<uint32_t> largeList, largeCount;
    SVO_VOX_CUDA(largeList.alloc(nTris));
    SVO_VOX_CUDA(largeCount.alloc(1));
    SVO_VOX_CUDA(cudaMemset(largeCount.p, 0, sizeof(uint32_t)));
    countCellsKernel<<<blocks, kThreads>>>(P, dTris.p, nTris, dCounts.p, largeList.p, largeCount.p);
    SVO_VOX_CUDA(cudaGetLastError());
    uint32_t nLarge = 0;
    SVO_VOX_CUDA(cudaMemcpy(&nLarge, largeCount.p, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    // few large triangles: cut each into slabs along z so that there are about four blocks per SM
    int slabs = 1;
    if (nLarge) slabs = int(std::min<uint64_t>(64, std::max<uint64_t>(1, (148u*4u + nLarge - 1)/nLarge)));
    const uint64_t nItems = uint64_t(nLarge)*uint64_t(slabs);
    const unsigned largeBlocks = unsigned(std::min<uint64_t>(nItems, 148u*8u));
    Dev<uint64_t> itemCounts, itemOffsets;
    if (nLarge) {
        SVO_VOX_CUDA(itemCounts.alloc(nItems));
        SVO_VOX_CUDA(itemOffsets.alloc(nItems));
        largeTrianglesKernel<false><<<largeBlocks, kLargeThreads>>>(P, dTris.p, largeList.p, largeCount.p, slabs, itemCounts.p, nullptr,
                                                                     nullptr, nullptr, nullptr);
        SVO_VOX_CUDA(cudaGetLastError());
        largeOffsetsKernel<<<(nLarge + kThreads - 1)/kThreads, kThreads>>>(largeList.p, largeCount.p, slabs, itemCounts.p, itemOffsets.p, dCounts.p);
        SVO_VOX_CUDA(cudaGetLastError());
    }
    stats.largeTriangles = nLarge;
    Dev<uint8_t> temp;
    size_t tempBytes = 0;
    SVO_VOX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tempBytes, dCounts.p, dCounts.p, int(nTris) + 1));
    SVO_VOX_CUDA(temp.alloc(tempBytes));
    SVO_VOX_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tempBytes, dCounts.p, dCounts.p, int(nTris) + 1));
    uint64_t nRecords = 0;
    SVO_VOX_CUDA(cudaMemcpy(&nRecords, dCounts.p + nTris, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (nRecords == 0) { err = "no triangle touches any cell"; return false; }
    if (nRecords >= (1ull << 31)) { err = "more than 2^31 - 1 (cell, triangle) overlaps are not supported"; return false; }
    stats.cellRecords = nRecords;
    const uint32_t n = uint32_t(nRecords);
    Dev<uint64_t> keys, keysAlt;
    Dev<CellRecord> records;
    Dev<uint32_t> order, orderAlt;
    SVO_VOX_CUDA(keys.alloc(n));
    SVO_VOX_CUDA(records.alloc(n));
    writeCellsKernel<<<blocks, kThreads>>>(P, dTris.p, nTris, dCounts.p, keys.p, records.p);
    SVO_VOX_CUDA(cudaGetLastError());
    if (nLarge) {
        largeTrianglesKernel<true><<<largeBlocks, kLargeThreads>>>(P, dTris.p, largeList.p, largeCount.p, slabs, nullptr, itemOffsets.p,
                                                                    dCounts.p, keys.p, records.p);
        SVO_VOX_CUDA(cudaGetLastError());
    }
    stats.overlapMs = timer.stop();
    dTris.release();
    dCounts.release();

    // stable sort of record indices by cell: triangle order survives inside every cell
    timer.start();
    SVO_VOX_CUDA(keysAlt.alloc(n));
    SVO_VOX_CUDA(order.alloc(n));
    SVO_VOX_CUDA(orderAlt.alloc(n));
    iotaKernel<<<(n + 255)/256, 256>>>(n, order.p);
    int keyBits = 1;
    while ((uint64_t(1) << keyBits) < uint64_t(w)*uint64_t(h)*uint64_t(d)) ++keyBits;
    cub::DoubleBuffer<uint64_t> kb(keys.p, keysAlt.p);
    cub::DoubleBuffer<uint32_t> vb(order.p, orderAlt.p);
    tempBytes = 0;
    SVO_VOX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, kb, vb, int(n), 0, keyBits));
    SVO_VOX_CUDA(temp.alloc(tempBytes));
    SVO_VOX_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tempBytes, kb, vb, int(n), 0, keyBits));
    stats.sortMs = timer.stop();

    timer.start();
    Dev<uint32_t> xyz, values;
    Dev<unsigned long long> cursor;
    SVO_VOX_CUDA(xyz.alloc(uint64_t(n)*3));
    SVO_VOX_CUDA(values.alloc(n));
    SVO_VOX_CUDA(cursor.alloc(1));
    SVO_VOX_CUDA(cudaMemset(cursor.p, 0, sizeof(unsigned long long)));
    foldCellsKernel<<<(n + 255)/256, 256>>>(P, n, kb.Current(), vb.Current(), records.p, xyz.p, values.p, cursor.p);
    SVO_VOX_CUDA(cudaGetLastError());
    unsigned long long nVoxels = 0;
    SVO_VOX_CUDA(cudaMemcpy(&nVoxels, cursor.p, sizeof nVoxels, cudaMemcpyDeviceToHost));
    stats.foldMs = timer.stop();
    stats.voxels = nVoxels;
    if (nVoxels == 0) { err = "the mesh produced no filled cell"; return false; }
    const double foldedAt = wallMs();
    const bool ok = builder.addSparse(xyz.p, values.p, nVoxels, err);
    if (debugTiming)
        fprintf(stderr, "[svo] voxelizeMesh: %.1f ms upload + assembly, %.1f ms to the fold's end (device %.1f ms), %.1f ms handing %llu voxels to the builder\n",
                uploadedAt, foldedAt - uploadedAt, stats.overlapMs + stats.sortMs + stats.foldMs, wallMs() - foldedAt, nVoxels);
    return ok;
}

} // namespace svo
