/*
 * TEST INFRASTRUCTURE ONLY -- the parity oracle, never the product.
 *
 * Row f3 of SURVEY.md section 8 (prepared for the next round: no product code calls or mirrors this yet):
 * plain-C restatement of the reference's triangle voxeliser,
 *   PlyLoader::PlyLoader / readVertices / rescaleVertices / readTriangles   reference src/PlyLoader.cpp:64-226
 *   pointToGrid, iterateOverlappingBlocks, buildBlockLists                  :228-293
 *   writeTriangleCell, triangleToVolume                                     :296-374
 *   findBestBlockPartition, setupBlockProcessing, processBlock              :381-472
 *   suggestedDimensions, convertToVolume                                    :498-534
 *   Triangle::Triangle, Triangle::barycentric                               :40-62
 *   compressMaterial / decompressMaterial                                   reference src/Util.hpp:64-100
 *   triBoxOverlap (Akenine-Moller's separating-axis test, 2001)             reference src/third-party/tribox3.c
 * The PLY container itself (third-party plyfile in the reference) is read by a small parser of our own:
 * ASCII / binary little / big endian, scalar vertex properties of any type converted to float like
 * ply_get_property(..., PLY_FLOAT) does, one list property `vertex_indices` on `face`.
 *
 * What the result depends on besides the mesh: the resolution, the memory budget
 * (slab depth of convertToVolume) and the THREAD COUNT of the reference's pool, because the cache block is
 * partitioned into per-thread sub-blocks (findBestBlockPartition) and both the block-level triangle lists
 * and the cell centres' running sums start at sub-block boundaries.
 *
 * PINNED: tests/test_oracle_pins.py::test_voxeliser_port_equals_reference compares the volume written here
 * byte for byte with the one the reference's own object code writes (oracle/_ref, svoref_ply_to_voxel_file)
 * with the same thread count. Must be compiled with -ffp-contract=off.
 */
#include "svo_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float pos[3], normal[3], color[3]; } Vtx;
typedef struct { Vtx v[3]; float lower[3], upper[3]; } Tri;

static float minStdF(float a, float b) { return (b < a) ? b : a; }
static float maxStdF(float a, float b) { return (a < b) ? b : a; }
static int minI(int a, int b) { return b < a ? b : a; }
static int maxI(int a, int b) { return a < b ? b : a; }

static void cross3(const float *a, const float *b, float *out) {            /* Vec3::cross, Vec3.hpp */
    out[0] = a[1]*b[2] - a[2]*b[1];
    out[1] = a[2]*b[0] - a[0]*b[2];
    out[2] = a[0]*b[1] - a[1]*b[0];
}
static float length3(const float *a) { return sqrtf(a[0]*a[0] + a[1]*a[1] + a[2]*a[2]); }
static void normalize3(const float *a, float *out) {                        /* Vec3::normalize */
    float inv = 1.0f/sqrtf(a[0]*a[0] + a[1]*a[1] + a[2]*a[2]);
    out[0] = a[0]*inv; out[1] = a[1]*inv; out[2] = a[2]*inv;
}

/* ---- PLY container ------------------------------------------------------------------------------- */

enum { T_CHAR, T_UCHAR, T_SHORT, T_USHORT, T_INT, T_UINT, T_FLOAT, T_DOUBLE, T_BAD };
static const int kTypeSize[] = {1, 1, 2, 2, 4, 4, 4, 8, 0};

static int typeOf(const char *s) {
    static const char *names[][2] = {{"char", "int8"}, {"uchar", "uint8"}, {"short", "int16"}, {"ushort", "uint16"},
                                      {"int", "int32"}, {"uint", "uint32"}, {"float", "float32"}, {"double", "float64"}};
    for (int t = 0; t < 8; ++t) if (!strcmp(s, names[t][0]) || !strcmp(s, names[t][1])) return t;
    return T_BAD;
}

static double readScalar(FILE *fp, int type, int format /* 0 ascii, 1 LE, 2 BE */, int *ok) {
    if (format == 0) {
        double v = 0.0;
        if (fscanf(fp, "%lf", &v) != 1) *ok = 0;
        return v;
    }
    unsigned char b[8];
    int n = kTypeSize[type];
    if (fread(b, 1, (size_t)n, fp) != (size_t)n) { *ok = 0; return 0.0; }
    if (format == 2) for (int i = 0; i < n/2; ++i) { unsigned char t = b[i]; b[i] = b[n - 1 - i]; b[n - 1 - i] = t; }
    switch (type) {
    case T_CHAR: return (signed char)b[0];
    case T_UCHAR: return b[0];
    case T_SHORT: { short v; memcpy(&v, b, 2); return v; }
    case T_USHORT: { unsigned short v; memcpy(&v, b, 2); return v; }
    case T_INT: { int v; memcpy(&v, b, 4); return v; }
    case T_UINT: { unsigned v; memcpy(&v, b, 4); return v; }
    case T_FLOAT: { float v; memcpy(&v, b, 4); return v; }
    default: { double v; memcpy(&v, b, 8); return v; }
    }
}

typedef struct { char name[64]; int type, isList, countType; } Prop;
typedef struct { char name[64]; long long count; Prop props[32]; int nProps; } Elem;

/* Returns triangles (malloc'ed) in the reference's order; lower/upper are the RESCALED bounds. */
static Tri *loadPly(const char *path, size_t *nTrisOut, float lower[3], float upper[3]) {
    FILE *fp = fopen(path, "rb");
    if (!fp) return NULL;
    char line[1024], a[64], b[64], c[64], d[64];
    int format = -1;
    Elem elems[8];
    int nElems = 0;
    if (!fgets(line, sizeof line, fp) || strncmp(line, "ply", 3)) { fclose(fp); return NULL; }
    while (fgets(line, sizeof line, fp)) {
        if (!strncmp(line, "end_header", 10)) break;
        if (sscanf(line, "format %63s", a) == 1) {
            format = !strcmp(a, "ascii") ? 0 : !strcmp(a, "binary_little_endian") ? 1 : !strcmp(a, "binary_big_endian") ? 2 : -1;
        } else if (sscanf(line, "element %63s %63s", a, b) == 2 && nElems < 8) {
            memset(&elems[nElems], 0, sizeof(Elem));
            strcpy(elems[nElems].name, a);
            elems[nElems].count = atoll(b);
            ++nElems;
        } else if (nElems > 0 && sscanf(line, "property list %63s %63s %63s", a, b, c) == 3) {
            Elem *e = &elems[nElems - 1];
            if (e->nProps < 32) { Prop *p = &e->props[e->nProps++]; strcpy(p->name, c); p->isList = 1; p->countType = typeOf(a); p->type = typeOf(b); }
        } else if (nElems > 0 && sscanf(line, "property %63s %63s", a, d) == 2) {
            Elem *e = &elems[nElems - 1];
            if (e->nProps < 32) { Prop *p = &e->props[e->nProps++]; strcpy(p->name, d); p->isList = 0; p->type = typeOf(a); }
        }
    }
    if (format < 0) { fclose(fp); return NULL; }

    static const char *vpNames[9] = {"x", "y", "z", "nx", "ny", "nz", "red", "green", "blue"};
    const float vertDefault[9] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 255.0f, 255.0f, 255.0f};
    Vtx *verts = NULL;
    long long nVerts = 0;
    Tri *tris = NULL;
    size_t nTris = 0, capTris = 0;
    int hasNormals = 0, ok = 1;
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};

    for (int ei = 0; ei < nElems && ok; ++ei) {
        Elem *e = &elems[ei];
        if (!strcmp(e->name, "vertex")) {
            int avail[9] = {0}, slot[32];
            for (int p = 0; p < e->nProps; ++p) {
                slot[p] = -1;
                for (int t = 0; t < 9; ++t) if (!e->props[p].isList && !strcmp(e->props[p].name, vpNames[t])) { slot[p] = t; avail[t] = 1; break; }
            }
            hasNormals = avail[3] && avail[4] && avail[5];
            nVerts = e->count;
            verts = (Vtx *)malloc(sizeof(Vtx)*(size_t)(nVerts > 0 ? nVerts : 1));
            for (long long i = 0; i < nVerts && ok; ++i) {
                float data[9];
                memcpy(data, vertDefault, sizeof data);
                for (int p = 0; p < e->nProps && ok; ++p) {
                    if (e->props[p].isList) {
                        int cnt = (int)readScalar(fp, e->props[p].countType, format, &ok);
                        for (int k = 0; k < cnt && ok; ++k) readScalar(fp, e->props[p].type, format, &ok);
                    } else {
                        double v = readScalar(fp, e->props[p].type, format, &ok);
                        if (slot[p] >= 0) data[slot[p]] = (float)v;          /* plyfile: stored type -> PLY_FLOAT */
                    }
                }
                memcpy(verts[i].pos, data, 12); memcpy(verts[i].normal, data + 3, 12); memcpy(verts[i].color, data + 6, 12);
                for (int t = 0; t < 3; ++t) { lo[t] = minStdF(lo[t], data[t]); hi[t] = maxStdF(hi[t], data[t]); }   /* :160-163 */
            }
            /* rescaleVertices, :167-181 */
            float diff[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
            int largest = 2;
            if (diff[0] > diff[1] && diff[0] > diff[2]) largest = 0;
            else if (diff[1] > diff[2]) largest = 1;
            float factor = 1.0f/diff[largest];
            for (long long i = 0; i < nVerts; ++i)
                for (int t = 0; t < 3; ++t) verts[i].pos[t] = (verts[i].pos[t] - lo[t])*factor;
            for (int t = 0; t < 3; ++t) { hi[t] *= factor; lo[t] *= factor; }
        } else if (!strcmp(e->name, "face")) {
            if (!verts) { ok = 0; break; }
            for (long long i = 0; i < e->count && ok; ++i) {
                for (int p = 0; p < e->nProps && ok; ++p) {
                    Prop *pr = &e->props[p];
                    if (pr->isList && !strcmp(pr->name, "vertex_indices")) {
                        int cnt = (int)readScalar(fp, pr->countType, format, &ok);
                        int v0 = 0, v1 = 0;
                        for (int k = 0; k < cnt && ok; ++k) {                    /* triangle fan, :207-221 */
                            int idx = (int)readScalar(fp, pr->type, format, &ok);
                            if (idx < 0 || idx >= nVerts) { ok = 0; break; }
                            if (k == 0) v0 = idx;
                            else if (k == 1) v1 = idx;
                            else {
                                if (nTris == capTris) { capTris = capTris ? capTris*2 : 1024; tris = (Tri *)realloc(tris, capTris*sizeof(Tri)); }
                                Tri *t = &tris[nTris++];
                                t->v[0] = verts[v0]; t->v[1] = verts[v1]; t->v[2] = verts[idx];
                                for (int q = 0; q < 3; ++q) {                    /* Triangle::Triangle, :40-54 */
                                    t->lower[q] = minStdF(t->v[0].pos[q], minStdF(t->v[1].pos[q], t->v[2].pos[q]));
                                    t->upper[q] = maxStdF(t->v[0].pos[q], maxStdF(t->v[1].pos[q], t->v[2].pos[q]));
                                }
                                if (!hasNormals) {                               /* :213-218 */
                                    float e1[3], e2[3], n[3];
                                    for (int q = 0; q < 3; ++q) { e1[q] = verts[v1].pos[q] - verts[v0].pos[q]; e2[q] = verts[idx].pos[q] - verts[v0].pos[q]; }
                                    cross3(e1, e2, n);
                                    normalize3(n, n);
                                    for (int w = 0; w < 3; ++w) memcpy(t->v[w].normal, n, 12);
                                }
                                v1 = idx;
                            }
                        }
                    } else if (pr->isList) {
                        int cnt = (int)readScalar(fp, pr->countType, format, &ok);
                        for (int k = 0; k < cnt && ok; ++k) readScalar(fp, pr->type, format, &ok);
                    } else {
                        readScalar(fp, pr->type, format, &ok);
                    }
                }
            }
        } else {
            for (long long i = 0; i < e->count && ok; ++i)
                for (int p = 0; p < e->nProps && ok; ++p) {
                    if (e->props[p].isList) {
                        int cnt = (int)readScalar(fp, e->props[p].countType, format, &ok);
                        for (int k = 0; k < cnt && ok; ++k) readScalar(fp, e->props[p].type, format, &ok);
                    } else readScalar(fp, e->props[p].type, format, &ok);
                }
        }
    }
    fclose(fp);
    free(verts);
    if (!ok) { free(tris); return NULL; }
    memcpy(lower, lo, 12); memcpy(upper, hi, 12);
    *nTrisOut = nTris;
    return tris ? tris : (Tri *)calloc(1, sizeof(Tri));
}

/* ---- triBoxOverlap: separating-axis test of a triangle against an axis-aligned box ---------------- */

static int planeBoxOverlap(const float *normal, const float *vert, const float *maxbox) {
    float vmin[3], vmax[3];
    for (int q = 0; q < 3; ++q) {
        float v = vert[q];
        if (normal[q] > 0.0f) { vmin[q] = -maxbox[q] - v; vmax[q] = maxbox[q] - v; }
        else { vmin[q] = maxbox[q] - v; vmax[q] = -maxbox[q] - v; }
    }
    if (normal[0]*vmin[0] + normal[1]*vmin[1] + normal[2]*vmin[2] > 0.0f) return 0;
    if (normal[0]*vmax[0] + normal[1]*vmax[1] + normal[2]*vmax[2] >= 0.0f) return 1;
    return 0;
}

/* one of the nine edge-cross-axis tests: projections pa, pb of the two vertices that differ, radius rad */
static int axisSeparates(float pa, float pb, float rad) {
    float mn, mx;
    if (pa < pb) { mn = pa; mx = pb; } else { mn = pb; mx = pa; }
    return mn > rad || mx < -rad;
}

static int triBoxOverlap(const float *c, const float *h, float tv[3][3]) {
    float v0[3], v1[3], v2[3], e0[3], e1[3], e2[3], fex, fey, fez, mn, mx;
    for (int q = 0; q < 3; ++q) { v0[q] = tv[0][q] - c[q]; v1[q] = tv[1][q] - c[q]; v2[q] = tv[2][q] - c[q]; }
    for (int q = 0; q < 3; ++q) { e0[q] = v1[q] - v0[q]; e1[q] = v2[q] - v1[q]; e2[q] = v0[q] - v2[q]; }

    fex = fabsf(e0[0]); fey = fabsf(e0[1]); fez = fabsf(e0[2]);
    if (axisSeparates(e0[2]*v0[1] - e0[1]*v0[2], e0[2]*v2[1] - e0[1]*v2[2], fez*h[1] + fey*h[2])) return 0;       /* X01 */
    if (axisSeparates(-e0[2]*v0[0] + e0[0]*v0[2], -e0[2]*v2[0] + e0[0]*v2[2], fez*h[0] + fex*h[2])) return 0;     /* Y02 */
    if (axisSeparates(e0[1]*v2[0] - e0[0]*v2[1], e0[1]*v1[0] - e0[0]*v1[1], fey*h[0] + fex*h[1])) return 0;       /* Z12 */

    fex = fabsf(e1[0]); fey = fabsf(e1[1]); fez = fabsf(e1[2]);
    if (axisSeparates(e1[2]*v0[1] - e1[1]*v0[2], e1[2]*v2[1] - e1[1]*v2[2], fez*h[1] + fey*h[2])) return 0;       /* X01 */
    if (axisSeparates(-e1[2]*v0[0] + e1[0]*v0[2], -e1[2]*v2[0] + e1[0]*v2[2], fez*h[0] + fex*h[2])) return 0;     /* Y02 */
    if (axisSeparates(e1[1]*v0[0] - e1[0]*v0[1], e1[1]*v1[0] - e1[0]*v1[1], fey*h[0] + fex*h[1])) return 0;       /* Z0 */

    fex = fabsf(e2[0]); fey = fabsf(e2[1]); fez = fabsf(e2[2]);
    if (axisSeparates(e2[2]*v0[1] - e2[1]*v0[2], e2[2]*v1[1] - e2[1]*v1[2], fez*h[1] + fey*h[2])) return 0;       /* X2 */
    if (axisSeparates(-e2[2]*v0[0] + e2[0]*v0[2], -e2[2]*v1[0] + e2[0]*v1[2], fez*h[0] + fex*h[2])) return 0;     /* Y1 */
    if (axisSeparates(e2[1]*v2[0] - e2[0]*v2[1], e2[1]*v1[0] - e2[0]*v1[1], fey*h[0] + fex*h[1])) return 0;       /* Z12 */

    for (int q = 0; q < 3; ++q) {                       /* the box's own axes */
        mn = mx = v0[q];
        if (v1[q] < mn) mn = v1[q];
        if (v1[q] > mx) mx = v1[q];
        if (v2[q] < mn) mn = v2[q];
        if (v2[q] > mx) mx = v2[q];
        if (mn > h[q] || mx < -h[q]) return 0;
    }
    float normal[3];
    cross3(e0, e1, normal);                             /* the triangle's plane */
    return planeBoxOverlap(normal, v0, h);
}

/* ---- material codec (Util.hpp:47-100) --------------------------------------------------------------- */

static uint32_t compressMaterial(const float *n, float shade) {
    uint32_t face = 0;
    float dominant = fabsf(n[0]);
    if (fabsf(n[1]) > dominant) { dominant = fabsf(n[1]); face = 1; }
    if (fabsf(n[2]) > dominant) { dominant = fabsf(n[2]); face = 2; }
    uint32_t sign = n[face] < 0.0f;
    static const int mod3[5] = {0, 1, 2, 0, 1};
    float n1 = n[mod3[face + 1]]/dominant;
    float n2 = n[mod3[face + 2]]/dominant;
    int32_t ui = (int32_t)((n1*0.5f + 0.5f)*2047), vi = (int32_t)((n2*0.5f + 0.5f)*2047), ci = (int32_t)(shade*127.0f);
    uint32_t u = (uint32_t)(ui < 0x7FF ? ui : 0x7FF), v = (uint32_t)(vi < 0x7FF ? vi : 0x7FF), c = (uint32_t)(ci < 0x7F ? ci : 0x7F);
    return (sign << 31) | (face << 29) | (u << 18) | v << 7 | c;
}

/* ---- block processing ------------------------------------------------------------------------------- */

typedef struct {
    const Tri *tris; size_t nTris;
    float lower[3], upper[3];
    int sideLength;                 /* the member _sideLength = sideLength - 2, :415 */
    int volumeW, volumeH, volumeD, blockW, blockH, blockD, subW, subH, subD, partW, partH, partD, numPartitions;
    int gridW, gridH, gridD;
    uint32_t *blockOffsets, *blockLists;
    uint8_t *counts;
    int bufferX, bufferY, bufferZ, bufferW, bufferH, bufferD;
} Loader;

static void pointToGrid(const Loader *L, const float *p, int *x, int *y, int *z) {   /* :228-233 */
    *x = (int)(p[0]*(L->sideLength - 2) + 1.0f);
    *y = (int)(p[1]*(L->sideLength - 2) + 1.0f);
    *z = (int)(p[2]*(L->sideLength - 2) + 1.0f);
}

typedef void (*BlockFn)(Loader *, size_t idx, uint32_t tri, int pass);

/* Sub-blocks a triangle was LISTED for whose (x, y, z) lies outside the sub-block grid: the flat index
 * x + gridW*(y + gridH*z) then names a block of the next row / slice (the reference blends the triangle into that
 * block as well) or, for z, lies past _blockOffsets. The GPU voxeliser does not mirror that aliasing
 * (csrc/svo_voxelize.cu header); svo_oracle_block_list_aliases counts it so that tests can show it never happens:
 * the loader rescales every mesh into [0, 1]^3, pointToGrid maps 1 to sideLength - 3 (:228-233 on top of :415), so
 * (u + 1)/sub stays inside the grid of a volume of sideLength cells (:431-433); and a sub-block past the volume would
 * start beyond 1, where triBoxOverlap rejects every triangle. */
static uint64_t g_listedOutsideGrid = 0, g_candidatesOutsideGrid = 0;

static void iterateOverlappingBlocks(Loader *L, uint32_t ti, BlockFn body, int pass) {   /* :235-277 */
    const Tri *t = &L->tris[ti];
    int lx, ly, lz, ux, uy, uz;
    pointToGrid(L, t->lower, &lx, &ly, &lz);
    pointToGrid(L, t->upper, &ux, &uy, &uz);
    int lgx = lx/L->subW, lgy = ly/L->subH, lgz = lz/L->subD;
    int ugx = (ux + 1)/L->subW, ugy = (uy + 1)/L->subH, ugz = (uz + 1)/L->subD;
    int maxSide = maxI(ugx - lgx, maxI(ugy - lgy, ugz - lgz));
    if (maxSide > 0) {
        float hx = L->subW/(float)(L->sideLength - 2), hy = L->subH/(float)(L->sideLength - 2), hz = L->subD/(float)(L->sideLength - 2);
        float tv[3][3];
        for (int k = 0; k < 3; ++k) memcpy(tv[k], t->v[k].pos, 12);
        float half[3] = {0.5f*hx, 0.5f*hy, 0.5f*hz}, center[3];
        center[2] = (lgz + 0.5f)*hz;
        for (int z = lgz; z <= ugz; ++z, center[2] += hz) {
            center[1] = (lgy + 0.5f)*hy;
            for (int y = lgy; y <= ugy; ++y, center[1] += hy) {
                center[0] = (lgx + 0.5f)*hx;
                for (int x = lgx; x <= ugx; ++x, center[0] += hx) {
                    int outside = x >= L->gridW || y >= L->gridH || z >= L->gridD;
                    if (outside && pass == 0) ++g_candidatesOutsideGrid;
                    if (triBoxOverlap(center, half, tv)) {
                        if (outside) { if (pass == 0) ++g_listedOutsideGrid; if (z >= L->gridD) continue; }
                        body(L, (size_t)(x + L->gridW*(y + L->gridH*z)), ti, pass);
                    }
                }
            }
        }
    } else {
        body(L, (size_t)(lgx + L->gridW*(lgy + L->gridH*lgz)), ti, pass);
    }
}

static void blockBody(Loader *L, size_t idx, uint32_t tri, int pass) {
    if (pass == 0) L->blockOffsets[1 + idx]++;
    else L->blockLists[L->blockOffsets[idx]++] = tri;
}

static void buildBlockLists(Loader *L) {                                                 /* :279-293 */
    size_t n = (size_t)L->gridW*L->gridH*L->gridD + 1;
    L->blockOffsets = (uint32_t *)calloc(n, sizeof(uint32_t));
    for (size_t i = 0; i < L->nTris; ++i) iterateOverlappingBlocks(L, (uint32_t)i, blockBody, 0);
    for (size_t i = 1; i < n; ++i) L->blockOffsets[i] += L->blockOffsets[i - 1];
    L->blockLists = (uint32_t *)malloc(sizeof(uint32_t)*(L->blockOffsets[n - 1] ? L->blockOffsets[n - 1] : 1));
    for (size_t i = 0; i < L->nTris; ++i) iterateOverlappingBlocks(L, (uint32_t)i, blockBody, 1);
    for (size_t i = n - 1; i >= 1; --i) L->blockOffsets[i] = L->blockOffsets[i - 1];
    L->blockOffsets[0] = 0;
}

static int barycentric(const Tri *t, const float *p, float *l1, float *l2) {            /* :56-66 */
    float f1[3], f2[3], f3[3], a[3], b[3], c[3];
    for (int q = 0; q < 3; ++q) { f1[q] = t->v[0].pos[q] - p[q]; f2[q] = t->v[1].pos[q] - p[q]; f3[q] = t->v[2].pos[q] - p[q]; }
    for (int q = 0; q < 3; ++q) { a[q] = t->v[0].pos[q] - t->v[1].pos[q]; b[q] = t->v[0].pos[q] - t->v[2].pos[q]; }
    cross3(a, b, c);
    float area = length3(c);
    cross3(f2, f3, c);
    *l1 = length3(c)/area;
    cross3(f3, f1, c);
    *l2 = length3(c)/area;
    return *l1 >= 0.0f && *l2 >= 0.0f && *l1 + *l2 <= 1.0f;
}

static void writeTriangleCell(Loader *L, uint32_t *data, int x, int y, int z, float cx, float cy, float cz, const Tri *t) {  /* :296-335 */
    size_t idx = (size_t)(x - L->bufferX) + (size_t)L->bufferW*((size_t)(y - L->bufferY) + (size_t)L->bufferH*(size_t)(z - L->bufferZ));
    float p[3] = {cx, cy, cz}, l1, l2, l3;
    if (!barycentric(t, p, &l1, &l2)) {
        l1 = minStdF(maxStdF(l1, 0.0f), 1.0f);
        l2 = minStdF(maxStdF(l2, 0.0f), 1.0f);
        float tau = l1 + l2;
        if (tau > 1.0f) { l1 /= tau; l2 /= tau; }
    }
    l3 = 1.0f - l1 - l2;
    float n[3], col[3];
    for (int q = 0; q < 3; ++q) {
        n[q] = t->v[0].normal[q]*l1 + t->v[1].normal[q]*l2 + t->v[2].normal[q]*l3;
        col[q] = t->v[0].color[q]*l1 + t->v[1].color[q]*l2 + t->v[2].color[q]*l3;
    }
    normalize3(n, n);
    float shade = (col[0]*0.2126f + col[1]*0.7152f + col[2]*0.0722f)*(1.0f/256.0f);
    if (data[idx] == 0) {
        L->counts[idx] = 1;
        data[idx] = compressMaterial(n, shade);
    } else {
        float currentRatio = L->counts[idx]/(L->counts[idx] + 1.0f);
        float newRatio = 1.0f - currentRatio;
        float cn[3], cs;
        svo_oracle_decompress_material(data[idx], cn, &cs);
        float nn[3] = {cn[0]*currentRatio + n[0]*newRatio, cn[1]*currentRatio + n[1]*newRatio, cn[2]*currentRatio + n[2]*newRatio};
        float ns = cs*currentRatio + shade*newRatio;
        if (nn[0]*nn[0] + nn[1]*nn[1] + nn[2]*nn[2] < 1e-3f) memcpy(nn, cn, 12);
        data[idx] = compressMaterial(nn, ns);
        L->counts[idx] = (uint8_t)minI((int)L->counts[idx] + 1, 255);
    }
}

static void triangleToVolume(Loader *L, uint32_t *data, const Tri *t, int offX, int offY, int offZ) {   /* :337-374 */
    int lx, ly, lz, ux, uy, uz;
    pointToGrid(L, t->lower, &lx, &ly, &lz);
    pointToGrid(L, t->upper, &ux, &uy, &uz);
    lx = maxI(lx, L->bufferX + offX);
    ly = maxI(ly, L->bufferY + offY);
    lz = maxI(lz, L->bufferZ + offZ);
    ux = minI(ux, L->bufferX + minI(offX + L->subW, L->bufferW) - 1);
    uy = minI(uy, L->bufferY + minI(offY + L->subH, L->bufferH) - 1);
    uz = minI(uz, L->bufferZ + minI(offZ + L->subD, L->bufferD) - 1);
    if (lx > ux || ly > uy || lz > uz) return;
    float hx = 1.0f/(L->sideLength - 2);
    float tv[3][3];
    for (int k = 0; k < 3; ++k) memcpy(tv[k], t->v[k].pos, 12);
    float half[3] = {0.5f*hx, 0.5f*hx, 0.5f*hx}, center[3];
    center[2] = (lz - 0.5f)*hx;
    for (int z = lz; z <= uz; z++, center[2] += hx) {
        center[1] = (ly - 0.5f)*hx;
        for (int y = ly; y <= uy; y++, center[1] += hx) {
            center[0] = (lx - 0.5f)*hx;
            for (int x = lx; x <= ux; x++, center[0] += hx)
                if (triBoxOverlap(center, half, tv)) writeTriangleCell(L, data, x, y, z, center[0], center[1], center[2], t);
        }
    }
}

static int *pickMax(int *w, int *h, int *d) { if (*w > *h && *w > *d) return w; else if (*h > *d) return h; else return d; }
static int *pickMin(int *w, int *h, int *d) { if (*w < *h && *w < *d) return w; else if (*h < *d) return h; else return d; }
static int *pickMedian(int *w, int *h, int *d) {
    int mx = maxI(*w, maxI(*h, *d)), mn = minI(*w, minI(*h, *d));
    if (*w != mn && *w != mx) return w; else if (*h != mn && *h != mx) return h; else return d;
}
static void findBestBlockPartition(int *w, int *h, int *d, int numThreads) {             /* :381-405 */
    int used = 1;
    while (used < numThreads) {
        if ((*pickMax(w, h, d) % 2) == 0) *pickMax(w, h, d) /= 2;
        else if ((*pickMedian(w, h, d) % 2) == 0) *pickMedian(w, h, d) /= 2;
        else if ((*pickMin(w, h, d) % 2) == 0) *pickMin(w, h, d) /= 2;
        else break;
        used *= 2;
    }
}

static void setupBlockProcessing(Loader *L, int sideLength, int blockW, int blockH, int blockD, int volumeW, int volumeH,
                                 int volumeD, int threadCount) {                        /* :407-440 */
    L->counts = (uint8_t *)malloc((size_t)blockW*(size_t)blockH*(size_t)blockD);
    L->sideLength = sideLength - 2;
    L->blockW = L->subW = blockW; L->blockH = L->subH = blockH; L->blockD = L->subD = blockD;
    findBestBlockPartition(&L->subW, &L->subH, &L->subD, threadCount);
    L->partW = L->blockW/L->subW; L->partH = L->blockH/L->subH; L->partD = L->blockD/L->subD;
    L->numPartitions = L->partW*L->partH*L->partD;
    L->volumeW = volumeW; L->volumeH = volumeH; L->volumeD = volumeD;
    L->gridW = L->partW*(L->volumeW + L->blockW - 1)/L->blockW;
    L->gridH = L->partH*(L->volumeH + L->blockH - 1)/L->blockH;
    L->gridD = L->partD*(L->volumeD + L->blockD - 1)/L->blockD;
    buildBlockLists(L);
}

static void processBlock(Loader *L, uint32_t *data, int x, int y, int z, int w, int h, int d) {   /* :442-472 */
    L->bufferX = x; L->bufferY = y; L->bufferZ = z; L->bufferW = w; L->bufferH = h; L->bufferD = d;
    for (int i = 0; i < L->numPartitions; ++i) {        /* the reference runs these on its pool; they write disjoint cells */
        int px = i % L->partW, py = (i/L->partW) % L->partH, pz = i/(L->partW*L->partH);
        int blockIdx = (x/L->subW + px) + L->gridW*((y/L->subH + py) + L->gridH*(z/L->subD + pz));
        int start = (int)L->blockOffsets[blockIdx], end = (int)L->blockOffsets[blockIdx + 1];
        for (int k = start; k < end; ++k)
            triangleToVolume(L, data, &L->tris[L->blockLists[k]], px*L->subW, py*L->subH, pz*L->subD);
    }
}

/* PlyLoader(path) + VoxelData(loader, sideLength, mem) (reference src/VoxelData.cpp:50-56) with a memory
 * budget large enough for the whole volume to be ONE cache block (edge = the power-of-two side, VoxelData.cpp:
 * 203-262), i.e. what VoxelData::cacheData -> PlyLoader::processBlock (VoxelData.cpp:178-181) leaves in the
 * buffer for buildOctree: the w*h*d volume, x fastest (malloc'ed, svo_oracle_free), or NULL if the file
 * cannot be read. threadCount = the reference pool's size (ThreadUtils::idealThreadCount()).
 * (The reference's other entry, convertToVolume (:505-534), divides by _numNonZeroBlocks == 0 in processBlock's
 * progress message (:460-462) and dies with SIGFPE, so it cannot serve as a comparator.) */
uint32_t *svo_oracle_voxelize_ply(const char *plyPath, int sideLength, int threadCount, int dims[3], uint64_t *nTrianglesOut) {
    Loader L;
    memset(&L, 0, sizeof L);
    size_t nTris = 0;
    Tri *tris = loadPly(plyPath, &nTris, L.lower, L.upper);
    if (!tris) return NULL;
    L.tris = tris; L.nTris = nTris;
    if (nTrianglesOut) *nTrianglesOut = nTris;
    float sizes[3];
    for (int q = 0; q < 3; ++q) sizes[q] = (L.upper[q] - L.lower[q])*(float)(sideLength - 2);     /* suggestedDimensions, :498-503 */
    int w = (int)sizes[0] + 2, h = (int)sizes[1] + 2, d = (int)sizes[2] + 2;
    int side = 1;
    while (side < w || side < h || side < d) side <<= 1;                                          /* VoxelData::init, :203-207 */
    uint32_t *data = (uint32_t *)calloc((size_t)w*(size_t)h*(size_t)d + 1, sizeof(uint32_t));
    setupBlockProcessing(&L, sideLength, side, side, side, w, h, d, threadCount);                 /* VoxelData.cpp:53-54 */
    processBlock(&L, data, 0, 0, 0, w, h, d);                                                     /* cacheData(0, 0, 0, w, h, d) */
    free(L.blockOffsets); free(L.blockLists); free(L.counts); free(tris);
    dims[0] = w; dims[1] = h; dims[2] = d;
    return data;
}

/* Block lists only (setupBlockProcessing, :407-440) for a cubic cache block of edge `blockEdge` (0: the power-of-two
 * side, one block) and `threadCount` pool threads: -> number of (triangle, sub-block) listings outside the sub-block
 * grid (see g_listedOutsideGrid); *candidatesOut: sub-blocks outside the grid that a triangle's index range reached
 * and triBoxOverlap then rejected; grid[3]: gridW, gridH, gridD; real[3]: sub-blocks that exist per axis.
 * < 0 when the file cannot be read. */
int64_t svo_oracle_block_list_aliases(const char *plyPath, int sideLength, int blockEdge, int threadCount,
                                      uint64_t *candidatesOut, int grid[3], int real[3]) {
    Loader L;
    memset(&L, 0, sizeof L);
    size_t nTris = 0;
    Tri *tris = loadPly(plyPath, &nTris, L.lower, L.upper);
    if (!tris) return -1;
    L.tris = tris; L.nTris = nTris;
    float sizes[3];
    for (int q = 0; q < 3; ++q) sizes[q] = (L.upper[q] - L.lower[q])*(float)(sideLength - 2);
    int w = (int)sizes[0] + 2, h = (int)sizes[1] + 2, d = (int)sizes[2] + 2;
    int side = 1;
    while (side < w || side < h || side < d) side <<= 1;
    if (blockEdge <= 0 || blockEdge > side) blockEdge = side;
    g_listedOutsideGrid = g_candidatesOutsideGrid = 0;
    /* setupBlockProcessing without the counts array (a block of 8192^3 bytes is not needed to list triangles) */
    L.sideLength = sideLength - 2;
    L.blockW = L.subW = L.blockH = L.subH = L.blockD = L.subD = blockEdge;
    findBestBlockPartition(&L.subW, &L.subH, &L.subD, threadCount);
    L.partW = L.blockW/L.subW; L.partH = L.blockH/L.subH; L.partD = L.blockD/L.subD;
    L.numPartitions = L.partW*L.partH*L.partD;
    L.volumeW = w; L.volumeH = h; L.volumeD = d;
    L.gridW = L.partW*(L.volumeW + L.blockW - 1)/L.blockW;
    L.gridH = L.partH*(L.volumeH + L.blockH - 1)/L.blockH;
    L.gridD = L.partD*(L.volumeD + L.blockD - 1)/L.blockD;
    buildBlockLists(&L);
    free(L.blockOffsets); free(L.blockLists); free(tris);
    if (candidatesOut) *candidatesOut = g_candidatesOutsideGrid;
    if (grid) { grid[0] = L.gridW; grid[1] = L.gridH; grid[2] = L.gridD; }
    if (real) { real[0] = (w + L.subW - 1)/L.subW; real[1] = (h + L.subH - 1)/L.subH; real[2] = (d + L.subD - 1)/L.subD; }
    return (int64_t)g_listedOutsideGrid;
}

/* The triangle list PlyLoader holds after its constructor, 33 floats per triangle: pos[3][3], normal[3][3],
 * color[3][3], lower[3], upper[3] (malloc'ed, svo_oracle_free); lower/upper: the rescaled mesh bounds. */
float *svo_oracle_ply_triangles(const char *plyPath, uint64_t *nOut, float lower[3], float upper[3]) {
    size_t n = 0;
    Tri *tris = loadPly(plyPath, &n, lower, upper);
    if (!tris) return NULL;
    float *out = (float *)malloc(sizeof(float)*33*(n ? n : 1));
    for (size_t i = 0; i < n; ++i) {
        float *o = out + 33*i;
        for (int v = 0; v < 3; ++v) { memcpy(o + 3*v, tris[i].v[v].pos, 12); memcpy(o + 9 + 3*v, tris[i].v[v].normal, 12); memcpy(o + 18 + 3*v, tris[i].v[v].color, 12); }
        memcpy(o + 27, tris[i].lower, 12); memcpy(o + 30, tris[i].upper, 12);
    }
    free(tris);
    *nOut = n;
    return out;
}

/* ---- node array -> voxels (to compare a reference-built tree with a volume) ------------------------- */

static void walkVoxels(const uint32_t *oct, uint64_t p, int x, int y, int z, int size, uint32_t *vol, int w, int h, int d) {
    uint32_t desc = oct[p];
    uint64_t off = desc >> 18;
    if (desc & 0x20000u) off = (off << 32) | oct[p + 1];
    uint64_t stride = (desc & 0x10000u) ? 2 : 1;
    uint32_t mask = (desc >> 8) & 0xFFu;
    int half = size >> 1, i = 0;
    for (int o = 0; o < 8; ++o) {
        if (!((mask >> o) & 1u)) continue;
        int cx = x + (o & 1)*half, cy = y + ((o >> 1) & 1)*half, cz = z + ((o >> 2) & 1)*half;
        if (half == 1) {
            if (cx < w && cy < h && cz < d) vol[(size_t)cx + (size_t)w*((size_t)cy + (size_t)h*(size_t)cz)] = oct[p + off + (uint64_t)i];
        } else {
            walkVoxels(oct, p + off + (uint64_t)i*stride, cx, cy, cz, half, vol, w, h, d);
        }
        ++i;
    }
}

/* Fills vol (w*h*d words, zeroed by the caller) with the material words of a tree that spans side^3 voxels. */
void svo_oracle_tree_to_volume(const uint32_t *octree, int side, uint32_t *vol, int w, int h, int d) {
    walkVoxels(octree, 0, 0, 0, 0, side, vol, w, h, d);
}
