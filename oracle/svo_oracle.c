/*
 * TEST INFRASTRUCTURE ONLY -- the parity oracle, never the product. See
 * svo_oracle.h for what is restated and how it is pinned to the reference's
 * own object code (oracle/_ref) and to tests/golden/.
 *
 * Written from SURVEY.md Appendix A-C against the reference source; every
 * floating-point expression keeps the reference's operand order and rounding
 * points (two roundings for a*b - c, IEEE divide, the (a<b)?b:a forms of
 * std::max/std::min). Compile with -ffp-contract=off.
 */
#include "svo_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#define SVO_MAX_SCALE 23 /* VoxelOctree.hpp:38 */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
/* std::max / std::min as libstdc++ defines them (operand order matters for -0.0f / NaN) */
static inline float maxStd(float a, float b) { return (a < b) ? b : a; }
static inline float minStd(float a, float b) { return (b < a) ? b : a; }
static inline uint32_t popc8(uint32_t v) { return (uint32_t)__builtin_popcount(v & 0xFFu); } /* == BitCount[], VoxelOctree.cpp:36-53 */

/* Util.hpp:47-58 */
float svo_oracle_inv_sqrt(float x) {
    float halfX = x*0.5f;
    float y = u2f(0x5f3759dfu - (f2u(x) >> 1));
    return y*(1.5f - halfX*y*y);
}

/* VoxelOctree.cpp:207-346. `ops` (optional) records what each trip round the loop did:
 * 'P' push, 'A' advance, 'Q' advance + pop, 'L' leaf hit, 'D' LOD exit, 'X' advance + pop out of the root. */
static int raymarchImpl(const uint32_t *octree, const float *o, const float *d, float rayScale,
        uint32_t *normal, float *t, uint64_t *voxel, svo_oracle_counters *c,
        uint8_t *ops, uint32_t maxOps, uint32_t *nOps) {
    uint32_t opCount = 0;
#define SVO_OP(ch) do { if (ops && opCount < maxOps) ops[opCount] = (uint8_t)(ch); ++opCount; } while (0)
    uint64_t stackParent[SVO_MAX_SCALE + 1];
    float stackMaxT[SVO_MAX_SCALE + 1];
    float dT[3], bT[3], pos[3];
    uint32_t octantMask = 7;

    for (int a = 0; a < 3; ++a) {
        float da = d[a];
        if (fabsf(da) < 1e-4f) da = 1e-4f;          /* :217-219 (sign is dropped) */
        dT[a] = 1.0f/-fabsf(da);                    /* :221-223 */
        bT[a] = dT[a]*o[a];                         /* :225-227 */
        if (da > 0.0f) {                            /* :229-232 */
            octantMask ^= 1u << a;
            bT[a] = 3.0f*dT[a] - bT[a];
        }
    }

    float minT = maxStd(2.0f*dT[0] - bT[0], maxStd(2.0f*dT[1] - bT[1], 2.0f*dT[2] - bT[2]));
    float maxT = minStd(dT[0] - bT[0], minStd(dT[1] - bT[1], dT[2] - bT[2]));
    minT = maxStd(minT, 0.0f);

    uint32_t current = 0;
    uint64_t parent = 0;
    uint32_t idx = 0;
    int scale = SVO_MAX_SCALE - 1;
    float scaleExp2 = 0.5f;
    for (int a = 0; a < 3; ++a) {                   /* :248-250 */
        pos[a] = 1.0f;
        if (1.5f*dT[a] - bT[a] > minT) { idx ^= 1u << a; pos[a] = 1.5f; }
    }

    uint64_t iterations = 0, descF = 0, farF = 0, leafF = 0, pushes = 0, pops = 0;
    int result = SVO_ORACLE_MISS;
    float tOut = 0.0f;

    while (scale < SVO_MAX_SCALE) {
        ++iterations;
        if (current == 0) { current = octree[parent]; ++descF; }

        float cornerT[3];
        for (int a = 0; a < 3; ++a) cornerT[a] = pos[a]*dT[a] - bT[a];
        float maxTC = minStd(cornerT[0], minStd(cornerT[1], cornerT[2]));

        uint32_t childShift = idx ^ octantMask;
        uint32_t childMasks = current << childShift;

        if ((childMasks & 0x8000u) && minT <= maxT) {
            if (maxTC*rayScale >= scaleExp2) {      /* LOD exit :265-268 */
                result = SVO_ORACLE_LOD;
                tOut = maxTC;
                if (voxel) *voxel = parent | ((uint64_t)childShift << 60);
                SVO_OP('D');
                break;
            }
            float maxTV = minStd(maxT, maxTC);
            float half = scaleExp2*0.5f;
            float centerT[3];
            for (int a = 0; a < 3; ++a) centerT[a] = half*dT[a] + cornerT[a];

            if (minT <= maxTV) {
                uint64_t childOffset = current >> 18;
                if (current & 0x20000u) {           /* far word :278-279 */
                    childOffset = (childOffset << 32) | (uint64_t)octree[parent + 1];
                    ++farF;
                }
                if (!(childMasks & 0x80u)) {        /* leaf :281-285 */
                    uint64_t leaf = childOffset + parent + popc8(((childMasks >> (8 + childShift)) << childShift) & 127u);
                    *normal = octree[leaf];
                    ++leafF;
                    if (voxel) *voxel = leaf;
                    result = SVO_ORACLE_LEAF;
                    tOut = minT;
                    SVO_OP('L');
                    break;
                }
                stackParent[scale] = parent;        /* push :287-306 */
                stackMaxT[scale] = maxT;
                ++pushes;
                uint32_t siblings = popc8(childMasks & 127u);
                parent += childOffset + siblings;
                if (current & 0x10000u) parent += siblings;
                idx = 0;
                --scale;
                scaleExp2 = half;
                for (int a = 0; a < 3; ++a)
                    if (centerT[a] > minT) { idx ^= 1u << a; pos[a] += scaleExp2; }
                maxT = maxTV;
                current = 0;
                SVO_OP('P');
                continue;
            }
        }

        uint32_t stepMask = 0;                      /* advance :310-316 */
        for (int a = 0; a < 3; ++a)
            if (cornerT[a] <= maxTC) { stepMask ^= 1u << a; pos[a] -= scaleExp2; }
        minT = maxTC;
        idx ^= stepMask;

        if ((idx & stepMask) != 0) {                /* pop :318-338 */
            ++pops;
            int32_t differingBits = 0;
            for (int a = 0; a < 3; ++a)
                if (stepMask & (1u << a))
                    differingBits |= (int32_t)(f2u(pos[a]) ^ f2u(pos[a] + scaleExp2));
            scale = (int)(f2u((float)differingBits) >> 23) - 127;
            scaleExp2 = u2f((uint32_t)(scale - SVO_MAX_SCALE + 127) << 23);
            /* scale == 23 (ray left the root) reads slot 23, which the
             * reference never wrote either; give it a defined value */
            parent = (scale < SVO_MAX_SCALE) ? stackParent[scale] : 0;
            maxT = (scale < SVO_MAX_SCALE) ? stackMaxT[scale] : 0.0f;
            idx = 0;
            for (int a = 0; a < 3; ++a) {
                uint32_t sh = f2u(pos[a]) >> scale;
                pos[a] = u2f(sh << scale);
                idx |= (sh & 1u) << a;
            }
            current = 0;
            SVO_OP(scale < SVO_MAX_SCALE ? 'Q' : 'X');
        } else {
            SVO_OP('A');
        }
    }
#undef SVO_OP
    if (nOps) *nOps = opCount;

    if (c) {
        c->rays += 1;
        c->iterations += iterations;
        c->desc_fetches += descF;
        c->far_fetches += farF;
        c->leaf_fetches += leafF;
        c->pushes += pushes;
        c->pops += pops;
        if (iterations > c->max_iterations) c->max_iterations = iterations;
        if (result != SVO_ORACLE_MISS) c->hits += 1;
        if (result == SVO_ORACLE_LOD) c->lod_exits += 1;
    }
    if (result != SVO_ORACLE_MISS) *t = tOut;       /* :266 / :344 */
    return result;
}

int svo_oracle_raymarch(const uint32_t *octree, const float *o, const float *d, float rayScale,
        uint32_t *normal, float *t, uint64_t *voxel, svo_oracle_counters *c) {
    return raymarchImpl(octree, o, d, rayScale, normal, t, voxel, c, NULL, 0, NULL);
}

static void mergeCounters(svo_oracle_counters *dst, const svo_oracle_counters *src) {
    dst->rays += src->rays;
    dst->iterations += src->iterations;
    dst->desc_fetches += src->desc_fetches;
    dst->far_fetches += src->far_fetches;
    dst->leaf_fetches += src->leaf_fetches;
    dst->pushes += src->pushes;
    dst->pops += src->pops;
    if (src->max_iterations > dst->max_iterations) dst->max_iterations = src->max_iterations;
    dst->hits += src->hits;
    dst->lod_exits += src->lod_exits;
}

/* ---- batch --------------------------------------------------------------- */

typedef struct {
    const uint32_t *octree;
    uint64_t begin, end;
    const float *o, *d;
    float rayScale;
    uint8_t *hit; float *t; uint32_t *normal; uint64_t *voxel;
    svo_oracle_counters counters;
} BatchJob;

static void *batchWorker(void *arg) {
    BatchJob *j = (BatchJob *)arg;
    for (uint64_t i = j->begin; i < j->end; ++i) {
        uint32_t n = j->normal ? j->normal[i] : 0;
        float t = j->t ? j->t[i] : 0.0f;
        uint64_t v = j->voxel ? j->voxel[i] : 0;
        int r = svo_oracle_raymarch(j->octree, j->o + 3*i, j->d + 3*i, j->rayScale, &n, &t, &v, &j->counters);
        if (j->hit) j->hit[i] = (uint8_t)r;
        if (j->t) j->t[i] = t;
        if (j->normal) j->normal[i] = n;
        if (j->voxel) j->voxel[i] = v;
    }
    return NULL;
}

void svo_oracle_raymarch_batch(const uint32_t *octree, uint64_t n, const float *o, const float *d,
        float rayScale, uint8_t *hit, float *t, uint32_t *normal, uint64_t *voxel,
        svo_oracle_counters *c, int threads) {
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n) threads = n ? (int)n : 1;
    BatchJob *jobs = (BatchJob *)calloc((size_t)threads, sizeof(BatchJob));
    pthread_t *tids = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int k = 0; k < threads; ++k) {
        BatchJob *j = &jobs[k];
        j->octree = octree;
        j->begin = n*(uint64_t)k/(uint64_t)threads;
        j->end = n*(uint64_t)(k + 1)/(uint64_t)threads;
        j->o = o; j->d = d; j->rayScale = rayScale;
        j->hit = hit; j->t = t; j->normal = normal; j->voxel = voxel;
        if (k > 0) pthread_create(&tids[k], NULL, batchWorker, j);
    }
    batchWorker(&jobs[0]);
    for (int k = 1; k < threads; ++k) pthread_join(tids[k], NULL);
    if (c) for (int k = 0; k < threads; ++k) mergeCounters(c, &jobs[k].counters);
    free(jobs);
    free(tids);
}

/* ---- camera / per-frame constants ---------------------------------------- */

/* Mat4.cpp:67-78, row-major a[i*4 + t] */
static void mat4Mul(const float *a, const float *b, float *out) {
    float r[16];
    for (int i = 0; i < 4; ++i)
        for (int t = 0; t < 4; ++t)
            r[i*4 + t] = a[i*4 + 0]*b[0*4 + t] + a[i*4 + 1]*b[1*4 + t] + a[i*4 + 2]*b[2*4 + t] + a[i*4 + 3]*b[3*4 + t];
    memcpy(out, r, sizeof r);
}

static void mat4Translate(float x, float y, float z, float *m) {
    static const float id[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    memcpy(m, id, sizeof id);
    m[3] = x; m[7] = y; m[11] = z;
}

/* Mat4.cpp:59-65 */
static void mat4PseudoInvert(const float *m, float *out) {
    float trans[16], rot[16];
    mat4Translate(-m[3], -m[7], -m[11], trans);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            rot[i*4 + j] = m[j*4 + i];
    rot[12] = rot[13] = rot[14] = 0.0f;
    mat4Mul(rot, trans, out);
}

/* Mat4.cpp:114-125 */
static void mat4RotXYZ(float rx, float ry, float rz, float *m) {
    const float pi = (float)3.14159265358979323846;
    float r[3] = {rx*pi/180.0f, ry*pi/180.0f, rz*pi/180.0f};
    float c[3] = {cosf(r[0]), cosf(r[1]), cosf(r[2])};
    float s[3] = {sinf(r[0]), sinf(r[1]), sinf(r[2])};
    float v[16] = {
        c[1]*c[2], -c[0]*s[2] + s[0]*s[1]*c[2],  s[0]*s[2] + c[0]*s[1]*c[2], 0.0f,
        c[1]*s[2],  c[0]*c[2] + s[0]*s[1]*s[2], -s[0]*c[2] + c[0]*s[1]*s[2], 0.0f,
            -s[1],                   s[0]*c[1],                   c[0]*c[1], 0.0f,
             0.0f,                        0.0f,                        0.0f, 1.0f};
    memcpy(m, v, sizeof v);
}

void svo_oracle_orbit_camera(float pitchDeg, float yawDeg, float radius, float *model16, float *view16) {
    float a[16], b[16];
    mat4Translate(0.0f, 0.0f, -radius, view16);
    mat4RotXYZ(pitchDeg, 0.0f, 0.0f, a);
    mat4RotXYZ(0.0f, yawDeg, 0.0f, b);
    mat4Mul(a, b, model16);
}

/* Main.cpp:149-163 */
void svo_oracle_frame_constants(const float *model16, const float *view16, const float *center,
        int width, int height, int strips, svo_oracle_frame *out) {
    float inv[16], m[16];
    mat4PseudoInvert(model16, inv);                 /* MatrixStack.cpp:71-73 */
    mat4Mul(inv, view16, m);

    out->width = width;
    out->height = height;
    out->strips = strips;
    out->tile_size = 8;                             /* Main.cpp:63 */
    for (int i = 0; i < 3; ++i) {
        /* tform*Vec3() + center + Vec3(1.0), Mat4.cpp:80-86 */
        float tp = m[i*4 + 0]*0.0f + m[i*4 + 1]*0.0f + m[i*4 + 2]*0.0f + m[i*4 + 3];
        out->pos[i] = tp + center[i] + 1.0f;
    }
    m[3] = m[7] = m[11] = 0.0f;

    out->a11 = m[0]; out->a12 = m[1];
    out->a21 = m[4]; out->a22 = m[5];
    out->a31 = m[8]; out->a32 = m[9];
    out->scale = 2.0f/width;
    out->tile_scale = out->tile_size*out->scale;
    float planeDist = 1.0f/tanf((float)3.14159265358979323846/6.0f);
    out->zx = planeDist*m[2];
    out->zy = planeDist*m[6];
    out->zz = planeDist*m[10];
    out->coarse_scale = 2.0f*out->tile_size/(planeDist*height);
    out->aspect = height/(float)width;              /* Main.cpp:62 */

    float l[3];
    for (int i = 0; i < 3; ++i)
        l[i] = m[i*4 + 0]*-1.0f + m[i*4 + 1]*1.0f + m[i*4 + 2]*-1.0f + m[i*4 + 3];
    float inv_len = 1.0f/sqrtf(l[0]*l[0] + l[1]*l[1] + l[2]*l[2]); /* Vec3.hpp:57-61 */
    for (int i = 0; i < 3; ++i) out->light[i] = l[i]*inv_len;
    out->beam_bias = 0.03f;                         /* Main.cpp:197 */
}

/* ---- shading -------------------------------------------------------------- */

/* Util.hpp:86-100 */
void svo_oracle_decompress_material(uint32_t word, float *n, float *shade) {
    static const int mod3[] = {0, 1, 2, 0, 1};
    uint32_t sign = (word & 0x80000000u) >> 31;
    uint32_t face = (word & 0x60000000u) >> 29;
    uint32_t u = (word & 0x1FFC0000u) >> 18;
    uint32_t v = (word & 0x0003FF80u) >> 7;
    uint32_t c = word & 0x7Fu;
    float tmp[4] = {0.0f, 0.0f, 0.0f, 0.0f}; /* face == 3 never occurs in valid data; stay in bounds */
    tmp[face & 3] = sign ? -1.0f : 1.0f;
    if (face < 3) {
        tmp[mod3[face + 1]] = u*4.8852e-4f*2.0f - 1.0f;
        tmp[mod3[face + 2]] = v*4.8852e-4f*2.0f - 1.0f;
    }
    float s = svo_oracle_inv_sqrt(tmp[0]*tmp[0] + tmp[1]*tmp[1] + tmp[2]*tmp[2]);
    n[0] = tmp[0]*s; n[1] = tmp[1]*s; n[2] = tmp[2]*s;
    *shade = c*1.0f/127.0f;
}

/* Main.cpp:81-90 with Vec3::dot / Vec3::reflect (Vec3.hpp:49-51,63-71) */
float svo_oracle_shade(uint32_t material, const float *ray, const float *light) {
    float n[3], c;
    svo_oracle_decompress_material(material, n, &c);
    float proj = (n[0]*ray[0] + n[1]*ray[1] + n[2]*ray[2])*2.0f;
    float r[3] = {ray[0] - n[0]*proj, ray[1] - n[1]*proj, ray[2] - n[2]*proj};
    float d = maxStd(light[0]*r[0] + light[1]*r[1] + light[2]*r[2], 0.0f);
    float specular = d*d;
    return c*0.9f*fabsf(light[0]*n[0] + light[1]*n[1] + light[2]*n[2]) + specular*0.2f;
}

/* Main.cpp:128-132: float -> double, double multiply, truncate */
uint32_t svo_oracle_pack(float v) {
    uint32_t g = (uint32_t)(minStd(v, 1.0f)*255.0);
    return g | (g << 8) | (g << 16) | 0xFF000000u;
}

/* ---- frame ----------------------------------------------------------------- */

typedef struct {
    const uint32_t *octree;
    const svo_oracle_frame *f;
    uint32_t *rgba;
    float *depth;            /* may be NULL */
    const uint64_t *depthOffset;
    atomic_int *next;
    svo_oracle_counters coarse, fine;
    int pixelStride;
} FrameJob;

static void rayDir(const svo_oracle_frame *f, float dx, float dy, float *dir) {
    dir[0] = dx*f->a11 + dy*f->a12 + f->zx;
    dir[1] = dx*f->a21 + dy*f->a22 + f->zy;
    dir[2] = dx*f->a31 + dy*f->a32 + f->zz;
    float s = svo_oracle_inv_sqrt(dir[0]*dir[0] + dir[1]*dir[1] + dir[2]*dir[2]);
    dir[0] *= s; dir[1] *= s; dir[2] *= s;
}

/* Main.cpp:92-137; stride 1, or 3 while renderHalfSize is set (Main.cpp:161) */
static void renderTile(FrameJob *j, int x0, int y0, int x1, int y1, float minT) {
    const svo_oracle_frame *f = j->f;
    const int stride = j->pixelStride > 1 ? j->pixelStride : 1;
    float dy = f->aspect - y0*f->scale;
    for (int y = y0; y < y1; ++y, dy -= f->scale) {
        float dx = -1.0f + x0*f->scale;
        for (int x = x0; x < x1; ++x, dx += f->scale) {
            int cornerX = x - ((x - x0) % stride);          /* :101-106 */
            int cornerY = y - ((y - y0) % stride);
            if (cornerX != x || cornerY != y) {
                j->rgba[x + (size_t)y*(size_t)f->width] = j->rgba[cornerX + (size_t)cornerY*(size_t)f->width];
                continue;
            }
            float dir[3], org[3];
            rayDir(f, dx, dy, dir);
            for (int a = 0; a < 3; ++a) org[a] = f->pos[a] + dir[a]*minT;
            uint32_t material = 0;
            float t = 0.0f;
            float v = 0.0f;
            if (svo_oracle_raymarch(j->octree, org, dir, 0.0f, &material, &t, NULL, &j->fine))
                v = svo_oracle_shade(material, dir, f->light);
            j->rgba[x + (size_t)y*(size_t)f->width] = svo_oracle_pack(v);
        }
    }
}

/* Main.cpp:139-202 for the strip [y0, y1) */
static void renderStrip(FrameJob *j, int strip) {
    const svo_oracle_frame *f = j->f;
    const float TreeMiss = 1e10f;
    int stride = (f->height - 1)/f->strips + 1;     /* Main.cpp:351 */
    int x0 = 0, x1 = f->width;
    int y0 = strip*stride;
    int y1 = (strip + 1)*stride < f->height ? (strip + 1)*stride : f->height;
    if (y0 >= y1) return;
    int tile = f->tile_size;
    int tilesX = (x1 - x0 - 1)/tile + 2;
    int tilesY = (y1 - y0 - 1)/tile + 2;
    float *depth = (float *)malloc(sizeof(float)*(size_t)tilesX*(size_t)tilesY);

    memset(j->rgba + (size_t)y0*(size_t)f->width, 0, sizeof(uint32_t)*(size_t)(y1 - y0)*(size_t)f->width);

    float dy = f->aspect - y0*f->scale;
    for (int y = 0, idx = 0; y < tilesY; ++y, dy -= f->tile_scale) {
        float dx = -1.0f + x0*f->scale;
        for (int x = 0; x < tilesX; ++x, dx += f->tile_scale, ++idx) {
            float dir[3];
            rayDir(f, dx, dy, dir);
            uint32_t material = 0;
            float t = 0.0f;
            if (svo_oracle_raymarch(j->octree, f->pos, dir, f->coarse_scale, &material, &t, NULL, &j->coarse))
                depth[idx] = t;
            else
                depth[idx] = TreeMiss;

            if (x > 0 && y > 0) {
                float minT = minStd(minStd(depth[idx], depth[idx - 1]), minStd(depth[idx - tilesX], depth[idx - tilesX - 1]));
                if (minT != TreeMiss) {
                    int tx0 = (x - 1)*tile + x0;
                    int ty0 = (y - 1)*tile + y0;
                    int tx1 = tx0 + tile < x1 ? tx0 + tile : x1;
                    int ty1 = ty0 + tile < y1 ? ty0 + tile : y1;
                    renderTile(j, tx0, ty0, tx1, ty1, maxStd(minT - f->beam_bias, 0.0f));
                }
            }
        }
    }
    if (j->depth) memcpy(j->depth + j->depthOffset[strip], depth, sizeof(float)*(size_t)tilesX*(size_t)tilesY);
    free(depth);
}

static void *frameWorker(void *arg) {
    FrameJob *j = (FrameJob *)arg;
    for (;;) {
        int s = atomic_fetch_add(j->next, 1);
        if (s >= j->f->strips) break;
        renderStrip(j, s);
    }
    return NULL;
}

int svo_oracle_render_frame(const uint32_t *octree, const svo_oracle_frame *f, uint32_t *rgba,
        float *depth, svo_oracle_counters *coarse, svo_oracle_counters *fine, int threads) {
    return svo_oracle_render_frame_strided(octree, f, 1, rgba, depth, coarse, fine, threads);
}

int svo_oracle_render_frame_strided(const uint32_t *octree, const svo_oracle_frame *f, int pixelStride, uint32_t *rgba,
        float *depth, svo_oracle_counters *coarse, svo_oracle_counters *fine, int threads) {
    if (!octree || !f || !rgba || f->width < 1 || f->height < 1 || f->strips < 1) return -1;
    if (threads < 1) threads = 1;
    if (threads > f->strips) threads = f->strips;

    uint64_t *offsets = (uint64_t *)calloc((size_t)f->strips + 1, sizeof(uint64_t));
    int stride = (f->height - 1)/f->strips + 1;
    for (int s = 0; s < f->strips; ++s) {
        int y0 = s*stride;
        int y1 = (s + 1)*stride < f->height ? (s + 1)*stride : f->height;
        int tilesX = (f->width - 1)/f->tile_size + 2;
        int tilesY = (y1 - y0 - 1)/f->tile_size + 2;
        uint64_t cells = y0 < y1 ? (uint64_t)tilesX*(uint64_t)tilesY : 0; /* strips past the bottom edge own nothing */
        offsets[s + 1] = offsets[s] + cells;
    }

    atomic_int next;
    atomic_init(&next, 0);
    FrameJob *jobs = (FrameJob *)calloc((size_t)threads, sizeof(FrameJob));
    pthread_t *tids = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int k = 0; k < threads; ++k) {
        jobs[k].octree = octree;
        jobs[k].f = f;
        jobs[k].rgba = rgba;
        jobs[k].depth = depth;
        jobs[k].depthOffset = offsets;
        jobs[k].next = &next;
        jobs[k].pixelStride = pixelStride;
        if (k > 0) pthread_create(&tids[k], NULL, frameWorker, &jobs[k]);
    }
    frameWorker(&jobs[0]);
    for (int k = 1; k < threads; ++k) pthread_join(tids[k], NULL);
    for (int k = 0; k < threads; ++k) {
        if (coarse) mergeCounters(coarse, &jobs[k].coarse);
        if (fine) mergeCounters(fine, &jobs[k].fine);
    }
    free(jobs);
    free(tids);
    free(offsets);
    return 0;
}

/* ---- per-ray operation traces of the fine pass (kernel design tool) ------------ */

/* For every `tileStride`-th rendered tile of the frame, traces the 64 fine rays in GPU warp order
 * (two warps per tile, lane l -> pixel (l & 7, (l >> 3) + 4*half)). ops: [maxWarps*32][maxOps],
 * counts: [maxWarps*32] (0 for clipped pixels). Returns the number of warps written. */
int64_t svo_oracle_trace_fine_warps(const uint32_t *octree, const svo_oracle_frame *f, int tileStride,
        uint8_t *ops, uint32_t maxOps, uint32_t *counts, int64_t maxWarps) {
    const float TreeMiss = 1e10f;
    int64_t nWarps = 0, rendered = 0;
    int stride = (f->height - 1)/f->strips + 1;
    int tile = f->tile_size;
    for (int strip = 0; strip < f->strips; ++strip) {
        int y0 = strip*stride;
        int y1 = (strip + 1)*stride < f->height ? (strip + 1)*stride : f->height;
        if (y0 >= y1) continue;
        int tilesX = (f->width - 1)/tile + 2, tilesY = (y1 - y0 - 1)/tile + 2;
        float *depth = (float *)malloc(sizeof(float)*(size_t)tilesX*(size_t)tilesY);
        float dy = f->aspect - y0*f->scale;
        for (int y = 0, idx = 0; y < tilesY; ++y, dy -= f->tile_scale) {
            float dx = -1.0f + 0*f->scale;
            for (int x = 0; x < tilesX; ++x, dx += f->tile_scale, ++idx) {
                float dir[3], t = 0.0f;
                uint32_t material = 0;
                rayDir(f, dx, dy, dir);
                depth[idx] = svo_oracle_raymarch(octree, f->pos, dir, f->coarse_scale, &material, &t, NULL, NULL) ? t : TreeMiss;
            }
        }
        for (int y = 1; y < tilesY; ++y) {
            for (int x = 1; x < tilesX; ++x) {
                int idx = y*tilesX + x;
                float minT = minStd(minStd(depth[idx], depth[idx - 1]), minStd(depth[idx - tilesX], depth[idx - tilesX - 1]));
                if (minT == TreeMiss) continue;
                if ((rendered++ % tileStride) != 0) continue;
                float startT = maxStd(minT - f->beam_bias, 0.0f);
                int tx0 = (x - 1)*tile, ty0 = (y - 1)*tile + y0;
                for (int half = 0; half < 2; ++half) {
                    if (nWarps >= maxWarps) { free(depth); return nWarps; }
                    for (int lane = 0; lane < 32; ++lane) {
                        int px = tx0 + (lane & 7), py = ty0 + (lane >> 3) + 4*half;
                        size_t slot = (size_t)nWarps*32 + (size_t)lane;
                        counts[slot] = 0;
                        if (px >= f->width || py >= y1) continue;
                        float fdy = f->aspect - ty0*f->scale;
                        for (int k = ty0; k < py; ++k) fdy -= f->scale;
                        float fdx = -1.0f + tx0*f->scale;
                        for (int k = tx0; k < px; ++k) fdx += f->scale;
                        float dir[3], org[3], t = 0.0f;
                        uint32_t material = 0, n = 0;
                        rayDir(f, fdx, fdy, dir);
                        for (int a = 0; a < 3; ++a) org[a] = f->pos[a] + dir[a]*startT;
                        raymarchImpl(octree, org, dir, 0.0f, &material, &t, NULL, NULL, ops + slot*maxOps, maxOps, &n);
                        counts[slot] = n < maxOps ? n : maxOps;
                    }
                    ++nWarps;
                }
            }
        }
        free(depth);
    }
    return nWarps;
}

/* Kernel design tool: loop-trip traces of arbitrary rays (batch mode). ops: [n][maxOps], counts: [n]. */
void svo_oracle_trace_rays(const uint32_t *octree, uint64_t n, const float *o, const float *d, float rayScale,
        uint8_t *ops, uint32_t maxOps, uint32_t *counts) {
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t material = 0, cnt = 0;
        float t = 0.0f;
        raymarchImpl(octree, o + 3*i, d + 3*i, rayScale, &material, &t, NULL, NULL, ops + i*maxOps, maxOps, &cnt);
        counts[i] = cnt < maxOps ? cnt : maxOps;
    }
}

/* ---- tree walk (App. A.1) --------------------------------------------------- */

static int walkNode(const uint32_t *oct, uint64_t words, uint64_t p, uint32_t level, svo_oracle_tree_stats *st) {
    if (p >= words || level >= 24) return -1;
    uint32_t D = oct[p];
    uint32_t valid = (D >> 8) & 0xFFu;
    uint32_t nonLeaf = D & 0xFFu;
    uint64_t off = D >> 18;
    st->descriptors += 1;
    st->per_level[level] += 1;
    if (level + 1 > st->depth) st->depth = level + 1;
    if (p > st->max_index) st->max_index = p;
    if (D & 0x20000u) {
        if (p + 1 >= words) return -1;
        off = (off << 32) | oct[p + 1];
        st->far_words += 1;
        if (p + 1 > st->max_index) st->max_index = p + 1;
    }
    uint32_t count = popc8(valid);
    if (nonLeaf == 0) {
        if (count == 0) return 0;
        uint64_t last = p + off + count - 1;
        if (last >= words) return -1;
        st->leaves += count;
        if (last > st->max_index) st->max_index = last;
        return 0;
    }
    uint64_t strideWords = (D & 0x10000u) ? 2 : 1;
    if (D & 0x10000u) st->far_blocks += 1;
    for (uint32_t s = 0; s < count; ++s)
        if (walkNode(oct, words, p + off + s*strideWords, level + 1, st) != 0) return -1;
    return 0;
}

int svo_oracle_tree_walk(const uint32_t *octree, uint64_t words, svo_oracle_tree_stats *out) {
    memset(out, 0, sizeof *out);
    if (words == 0) return -1;
    return walkNode(octree, words, 0, 0, out);
}

/* ---- octree construction (row f2) --------------------------------------------
 * Restatement of VoxelOctree::VoxelOctree(VoxelData*) / buildOctree (reference
 * src/VoxelOctree.cpp:125-205) over a dense in-memory grid, with the voxel source
 * reduced to what buildOctree observes of VoxelData (src/VoxelData.hpp:106-138,
 * src/VoxelData.cpp:158-175): a cube "contains voxels" iff the occupancy pyramid
 * says so, and the pyramid's finest level is filled from z-planes
 * [0, (depth/2)*2) only (buildLowLut's thread partition, VoxelData.cpp:160-163),
 * so the last plane of an odd-depth volume is invisible. ChunkedAllocator's
 * deferred insert()/finalize() (src/ChunkedAllocator.hpp:73-116) is restated as
 * an insertion list merged at the end.                                           */

typedef struct {
    uint32_t *data; uint64_t size, cap;
    uint64_t *insIdx; uint32_t *insVal; uint64_t insCount, insCap;
    const uint32_t *vox; int w, h, d;
    uint8_t **lut; int levels;          /* lut[k]: occupancy of cubes of edge 2^k, (side >> k)^3 cells */
    int side;
} builder;

static void bPush(builder *b, uint32_t v) {
    if (b->size == b->cap) { b->cap = b->cap ? b->cap*2 : 4096; b->data = (uint32_t *)realloc(b->data, b->cap*sizeof(uint32_t)); }
    b->data[b->size++] = v;
}
static void bInsert(builder *b, uint64_t idx, uint32_t v) {
    if (b->insCount == b->insCap) {
        b->insCap = b->insCap ? b->insCap*2 : 1024;
        b->insIdx = (uint64_t *)realloc(b->insIdx, b->insCap*sizeof(uint64_t));
        b->insVal = (uint32_t *)realloc(b->insVal, b->insCap*sizeof(uint32_t));
    }
    b->insIdx[b->insCount] = idx; b->insVal[b->insCount++] = v;
}
static uint32_t bVoxel(const builder *b, int x, int y, int z) {            /* VoxelData.hpp:106-111 */
    if (x >= b->w || y >= b->h || z >= b->d) return 0;
    return b->vox[(size_t)x + (size_t)b->w*((size_t)y + (size_t)b->h*(size_t)z)];
}
static int bContains(const builder *b, int x, int y, int z, int size) {    /* VoxelData.hpp:122-138 */
    if (x >= b->w || y >= b->h || z >= b->d) return 0;
    if (size == 1) return bVoxel(b, x, y, z) != 0;
    int k = 0; while ((1 << k) < size) ++k;
    size_t n = (size_t)(b->side >> k);
    return b->lut[k][(size_t)(x >> k) + n*((size_t)(y >> k) + n*(size_t)(z >> k))] != 0;
}

static uint64_t bBuild(builder *b, int x, int y, int z, int size, uint64_t descriptorIndex) {   /* VoxelOctree.cpp:139-205 */
    int half = size >> 1;
    int posX[8] = {x + half, x, x + half, x, x + half, x, x + half, x};
    int posY[8] = {y + half, y + half, y, y, y + half, y + half, y, y};
    int posZ[8] = {z + half, z + half, z + half, z + half, z, z, z, z};
    uint64_t childOffset = b->size - descriptorIndex;
    int childCount = 0, childIndices[8];
    uint32_t childMask = 0;
    for (int i = 0; i < 8; ++i)
        if (bContains(b, posX[i], posY[i], posZ[i], half)) { childMask |= 128u >> i; childIndices[childCount++] = i; }
    int hasLarge = 0;
    uint32_t leafMask;
    if (half == 1) {
        leafMask = 0;
        for (int i = 0; i < childCount; ++i) {
            int idx = childIndices[childCount - i - 1];
            bPush(b, bVoxel(b, posX[idx], posY[idx], posZ[idx]));
        }
    } else {
        leafMask = childMask;
        for (int i = 0; i < childCount; ++i) bPush(b, 0);
        uint64_t grand[8], delta = 0, ins = b->insCount;
        for (int i = 0; i < childCount; ++i) {
            int idx = childIndices[childCount - i - 1];
            grand[i] = delta + bBuild(b, posX[idx], posY[idx], posZ[idx], half, descriptorIndex + childOffset + (uint64_t)i);
            delta += b->insCount - ins;
            ins = b->insCount;
            if (grand[i] > 0x3FFF) hasLarge = 1;
        }
        for (int i = 0; i < childCount; ++i) {
            uint64_t childIndex = descriptorIndex + childOffset + (uint64_t)i, offset = grand[i];
            if (hasLarge) {
                offset += (uint64_t)(childCount - i);
                bInsert(b, childIndex + 1, (uint32_t)offset);
                b->data[childIndex] |= 0x20000u;
                offset >>= 32;
            }
            b->data[childIndex] |= (uint32_t)(offset << 18);
        }
    }
    b->data[descriptorIndex] = (childMask << 8) | leafMask;
    if (hasLarge) b->data[descriptorIndex] |= 0x10000u;
    return childOffset;
}

static int cmpU64Pair(const void *a, const void *b) {
    uint64_t x = ((const uint64_t *)a)[0], y = ((const uint64_t *)b)[0];
    return x < y ? -1 : x > y;
}

/* voxels: w*h*d words, x fastest (the raw .voxel layout, VoxelData.cpp:183-201). Returns a malloc'ed
 * node array (free with svo_oracle_free) and its length; NULL on allocation failure. */
uint32_t *svo_oracle_build_octree(const uint32_t *voxels, int w, int h, int d, uint64_t *nWordsOut, float center[3]) {
    builder b;
    memset(&b, 0, sizeof b);
    b.vox = voxels; b.w = w; b.h = h; b.d = d;
    int side = 1;
    while (side < w || side < h || side < d) side <<= 1;       /* roundToPow2 + sideLength, VoxelData.cpp:204-207,293-295 */
    b.side = side;
    int levels = 0; while ((1 << levels) < side) ++levels;
    b.levels = levels;
    b.lut = (uint8_t **)calloc((size_t)levels + 1, sizeof(uint8_t *));
    for (int k = 1; k <= levels; ++k) {
        size_t n = (size_t)(side >> k);
        b.lut[k] = (uint8_t *)calloc(n*n*n, 1);
    }
    if (levels >= 1) {
        size_t n = (size_t)(side >> 1);
        int zEnd = (d/2)*2;                                     /* VoxelData.cpp:160-163 */
        for (int z = 0; z < zEnd; ++z) for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x)
            if (bVoxel(&b, x, y, z)) b.lut[1][(size_t)(x/2) + n*((size_t)(y/2) + n*(size_t)(z/2))] = 1;
    }
    for (int k = 2; k <= levels; ++k) {                         /* upsampleLutLevel, VoxelData.cpp:87-120 */
        size_t n = (size_t)(side >> k), m = n*2;
        for (size_t z = 0; z < n; ++z) for (size_t y = 0; y < n; ++y) for (size_t x = 0; x < n; ++x) {
            int v = 0;
            for (int c = 0; c < 8; ++c)
                v |= b.lut[k - 1][(2*x + (c & 1)) + m*((2*y + ((c >> 1) & 1)) + m*(2*z + (c >> 2)))];
            b.lut[k][x + n*(y + n*z)] = (uint8_t)(v != 0);
        }
    }
    bPush(&b, 0);                                               /* VoxelOctree.cpp:128-132 */
    bBuild(&b, 0, 0, 0, side, 0);
    b.data[0] |= 1u << 18;

    /* ChunkedAllocator::finalize */
    uint64_t total = b.size + b.insCount;
    uint32_t *out = (uint32_t *)malloc((total + 1)*sizeof(uint32_t));
    uint64_t *pairs = (uint64_t *)malloc((b.insCount + 1)*2*sizeof(uint64_t));
    for (uint64_t i = 0; i < b.insCount; ++i) { pairs[2*i] = b.insIdx[i]; pairs[2*i + 1] = b.insVal[i]; }
    qsort(pairs, b.insCount, 2*sizeof(uint64_t), cmpU64Pair);
    uint64_t o = 0, k = 0;
    for (uint64_t i = 0; i < b.size; ++i) {
        while (k < b.insCount && pairs[2*k] == i) out[o++] = (uint32_t)pairs[2*k++ + 1];
        out[o++] = b.data[i];
    }
    free(pairs);
    for (int kk = 1; kk <= levels; ++kk) free(b.lut[kk]);
    free(b.lut); free(b.data); free(b.insIdx); free(b.insVal);
    *nWordsOut = total;
    if (center) {                                               /* VoxelData::getCenter, VoxelData.cpp:297-303 */
        center[0] = (float)w*0.5f/(float)side; center[1] = (float)h*0.5f/(float)side; center[2] = (float)d*0.5f/(float)side;
    }
    return out;
}

void svo_oracle_free(void *p) { free(p); }
