/*
 * Synthetic scene generators for the benchmark configurations (BASELINE.json
 * configs[1] and configs[2], concretised in SURVEY.md section 8d). BENCH TOOLING,
 * not part of the product: the outputs are INPUTS to the reference's own builder
 * (VoxelData / PlyLoader, run through oracle/_ref), whose node array is then
 * uploaded unchanged.
 *
 *  C2: f(p) = |p - c| - 0.35 - 0.05*fBm(p) on a res^3 grid over the unit cube,
 *      voxel filled iff |f(centre)| < half a voxel diagonal, material =
 *      compressMaterial(grad f / |grad f|, 0.8)  (reference src/Util.hpp:64-84),
 *      written as the reference's raw `.voxel` format (3 x int32 dims, uint32
 *      voxels x-fastest; reference src/PlyLoader.cpp:521-529, src/VoxelData.cpp:41-43)
 *      as a SPARSE file through a shared mapping (only touched pages exist).
 *  C3: geodesic icosphere of frequency n (20*n^2 triangles) displaced radially by
 *      the same fBm, as a binary little-endian PLY with x y z and vertex_indices
 *      (what reference src/PlyLoader.cpp:128,190 reads).
 *
 * fBm: 5 octaves of trilinear value noise with smoothstep fade on an integer-hash
 * lattice, base frequency 4, gain 0.5, normalised to [-1, 1].
 *
 *   gcc -O2 -shared -fPIC -o tools/libscene_gen.so tools/scene_gen.c -lm
 */
#define _GNU_SOURCE
#define _FILE_OFFSET_BITS 64
#include <fcntl.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

static inline uint32_t hash3(int32_t x, int32_t y, int32_t z, uint32_t seed) {
    uint32_t h = seed;
    h ^= (uint32_t)x*0x9E3779B1u; h = (h << 13) | (h >> 19); h *= 0x85EBCA6Bu;
    h ^= (uint32_t)y*0xC2B2AE35u; h = (h << 13) | (h >> 19); h *= 0x85EBCA6Bu;
    h ^= (uint32_t)z*0x27D4EB2Fu; h = (h << 13) | (h >> 19); h *= 0x85EBCA6Bu;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

static inline double lattice(int32_t x, int32_t y, int32_t z, uint32_t seed) {
    return (double)(hash3(x, y, z, seed) >> 8)*(2.0/16777216.0) - 1.0;
}

static inline double fade(double t) { return t*t*(3.0 - 2.0*t); }

static double valueNoise(double x, double y, double z, uint32_t seed) {
    double fx = floor(x), fy = floor(y), fz = floor(z);
    int32_t ix = (int32_t)fx, iy = (int32_t)fy, iz = (int32_t)fz;
    double u = fade(x - fx), v = fade(y - fy), w = fade(z - fz);
    double c000 = lattice(ix, iy, iz, seed),     c100 = lattice(ix + 1, iy, iz, seed);
    double c010 = lattice(ix, iy + 1, iz, seed), c110 = lattice(ix + 1, iy + 1, iz, seed);
    double c001 = lattice(ix, iy, iz + 1, seed),     c101 = lattice(ix + 1, iy, iz + 1, seed);
    double c011 = lattice(ix, iy + 1, iz + 1, seed), c111 = lattice(ix + 1, iy + 1, iz + 1, seed);
    double x00 = c000 + (c100 - c000)*u, x10 = c010 + (c110 - c010)*u;
    double x01 = c001 + (c101 - c001)*u, x11 = c011 + (c111 - c011)*u;
    double y0 = x00 + (x10 - x00)*v, y1 = x01 + (x11 - x01)*v;
    return y0 + (y1 - y0)*w;
}

#define FBM_OCTAVES 5
double svo_scene_fbm(double x, double y, double z, uint32_t seed) {
    double sum = 0.0, amp = 1.0, freq = 4.0, norm = 0.0;
    for (int i = 0; i < FBM_OCTAVES; ++i) {
        sum += amp*valueNoise(x*freq + 17.0*i, y*freq + 31.0*i, z*freq + 47.0*i, seed + (uint32_t)i);
        norm += amp;
        amp *= 0.5;
        freq *= 2.0;
    }
    return sum/norm;
}

#define SDF_RADIUS 0.35
#define SDF_AMPLITUDE 0.05
/* |grad f| <= 1 + 0.05*sum_i a_i*f_i*3*sqrt(3) with a_i*f_i = 4/1.9375 per octave */
#define SDF_LIPSCHITZ (1.0 + SDF_AMPLITUDE*FBM_OCTAVES*(4.0/1.9375)*5.1962)

double svo_scene_sdf(double x, double y, double z, uint32_t seed) {
    double dx = x - 0.5, dy = y - 0.5, dz = z - 0.5;
    return sqrt(dx*dx + dy*dy + dz*dz) - SDF_RADIUS - SDF_AMPLITUDE*svo_scene_fbm(x, y, z, seed);
}

/* reference src/Util.hpp:64-84, restated (builder-side codec; checked against the
 * reference's own compressMaterial in tests/test_scene_tools.py) */
uint32_t svo_scene_compress_material(const float n[3], float shade) {
    const int32_t uScale = (1 << 11) - 1, vScale = (1 << 11) - 1;
    uint32_t face = 0;
    float dominant = fabsf(n[0]);
    if (fabsf(n[1]) > dominant) { dominant = fabsf(n[1]); face = 1; }
    if (fabsf(n[2]) > dominant) { dominant = fabsf(n[2]); face = 2; }
    uint32_t sign = n[face] < 0.0f;
    static const int mod3[] = {0, 1, 2, 0, 1};
    float n1 = n[mod3[face + 1]]/dominant;
    float n2 = n[mod3[face + 2]]/dominant;
    int32_t ui = (int32_t)((n1*0.5f + 0.5f)*uScale), vi = (int32_t)((n2*0.5f + 0.5f)*vScale);
    int32_t ci = (int32_t)(shade*127.0f);
    uint32_t u = (uint32_t)(ui < 0x7FF ? ui : 0x7FF), v = (uint32_t)(vi < 0x7FF ? vi : 0x7FF);
    uint32_t c = (uint32_t)(ci < 0x7F ? ci : 0x7F);
    return (sign << 31) | (face << 29) | (u << 18) | (v << 7) | c;
}

typedef struct {
    uint32_t *vox;      /* mapping of the voxel payload */
    int res;
    uint32_t seed;
    double thr;         /* half a voxel diagonal */
    uint64_t filled;
} SdfJob;

static void sdfCell(SdfJob *j, int x, int y, int z) {
    double inv = 1.0/j->res;
    double px = (x + 0.5)*inv, py = (y + 0.5)*inv, pz = (z + 0.5)*inv;
    double f = svo_scene_sdf(px, py, pz, j->seed);
    if (fabs(f) >= j->thr) return;
    double h = 0.5*inv;
    float n[3];
    double gx = svo_scene_sdf(px + h, py, pz, j->seed) - svo_scene_sdf(px - h, py, pz, j->seed);
    double gy = svo_scene_sdf(px, py + h, pz, j->seed) - svo_scene_sdf(px, py - h, pz, j->seed);
    double gz = svo_scene_sdf(px, py, pz + h, j->seed) - svo_scene_sdf(px, py, pz - h, j->seed);
    double len = sqrt(gx*gx + gy*gy + gz*gz);
    if (len < 1e-30) { gx = 1.0; gy = gz = 0.0; len = 1.0; }
    n[0] = (float)(gx/len); n[1] = (float)(gy/len); n[2] = (float)(gz/len);
    uint32_t word = svo_scene_compress_material(n, 0.8f);
    if (word == 0) word = 1; /* a zero word would read as "empty"  */
    j->vox[(size_t)x + (size_t)j->res*((size_t)y + (size_t)j->res*(size_t)z)] = word;
    j->filled++;
}

static void sdfDescend(SdfJob *j, int x, int y, int z, int size) {
    double inv = 1.0/j->res;
    double half = 0.5*size;
    double f = svo_scene_sdf((x + half)*inv, (y + half)*inv, (z + half)*inv, j->seed);
    /* no voxel centre inside the cube can come within thr of the surface */
    if (fabs(f) > SDF_LIPSCHITZ*half*1.7320508*inv + j->thr) return;
    if (size == 1) { sdfCell(j, x, y, z); return; }
    if (size <= 4) {
        for (int dz = 0; dz < size; ++dz)
            for (int dy = 0; dy < size; ++dy)
                for (int dx = 0; dx < size; ++dx) sdfCell(j, x + dx, y + dy, z + dz);
        return;
    }
    int h = size/2;
    for (int k = 0; k < 8; ++k) sdfDescend(j, x + (k & 1)*h, y + ((k >> 1) & 1)*h, z + ((k >> 2) & 1)*h, h);
}

/* Writes the sparse raw .voxel file. Returns the number of filled voxels, or -1. */
int64_t svo_scene_sdf_voxel_file(const char *path, int res, uint32_t seed) {
    if (res < 8 || (res & (res - 1))) return -1;
    int fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return -1;
    size_t payload = (size_t)res*(size_t)res*(size_t)res*4;
    size_t total = 12 + payload;
    if (ftruncate(fd, (off_t)total) != 0) { close(fd); return -1; }
    int32_t dims[3] = {res, res, res};
    if (pwrite(fd, dims, 12, 0) != 12) { close(fd); return -1; }
    /* the payload starts at byte 12: map from 0 and offset the pointer */
    unsigned char *map = (unsigned char *)mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (map == MAP_FAILED) { close(fd); return -1; }
    SdfJob job;
    job.vox = (uint32_t *)(map + 12);
    job.res = res;
    job.seed = seed;
    job.thr = 0.5*1.7320508075688772/res;
    job.filled = 0;
    sdfDescend(&job, 0, 0, 0, res);
    munmap(map, total);
    close(fd);
    return (int64_t)job.filled;
}

/* ---- icosphere PLY --------------------------------------------------------- */

static void normalize3(double *v) {
    double l = sqrt(v[0]*v[0] + v[1]*v[1] + v[2]*v[2]);
    v[0] /= l; v[1] /= l; v[2] /= l;
}

/* Returns the triangle count, or -1. Vertices are emitted per icosahedron face
 * (edge vertices duplicated, bit-identical positions are not required by PLY). */
int64_t svo_scene_icosphere_ply(const char *path, int n, uint32_t seed) {
    if (n < 1) return -1;
    const double phi = (1.0 + sqrt(5.0))/2.0;
    double V[12][3] = {{-1, phi, 0}, {1, phi, 0}, {-1, -phi, 0}, {1, -phi, 0}, {0, -1, phi}, {0, 1, phi},
                       {0, -1, -phi}, {0, 1, -phi}, {phi, 0, -1}, {phi, 0, 1}, {-phi, 0, -1}, {-phi, 0, 1}};
    static const int F[20][3] = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4},
                                 {11, 10, 2}, {10, 7, 6}, {7, 1, 8}, {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8},
                                 {3, 8, 9}, {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
    for (int i = 0; i < 12; ++i) normalize3(V[i]);

    int64_t vertsPerFace = (int64_t)(n + 1)*(n + 2)/2;
    int64_t nVerts = 20*vertsPerFace, nTris = 20*(int64_t)n*n;
    if (nVerts > 0x7FFFFFFF) return -1;
    FILE *fp = fopen(path, "wb");
    if (!fp) return -1;
    fprintf(fp, "ply\nformat binary_little_endian 1.0\ncomment displaced geodesic icosphere, frequency %d, seed %u\n"
                "element vertex %lld\nproperty float x\nproperty float y\nproperty float z\n"
                "element face %lld\nproperty list uchar int vertex_indices\nend_header\n",
            n, seed, (long long)nVerts, (long long)nTris);

    float *row = (float *)malloc(sizeof(float)*3*(size_t)(n + 1));
    for (int f = 0; f < 20; ++f) {
        const double *A = V[F[f][0]], *B = V[F[f][1]], *C = V[F[f][2]];
        for (int i = 0; i <= n; ++i) {          /* row i: j = 0 .. n - i */
            for (int j = 0; j <= n - i; ++j) {
                double a = (double)i/n, b = (double)j/n;
                double p[3];
                for (int k = 0; k < 3; ++k) p[k] = A[k] + (B[k] - A[k])*a + (C[k] - A[k])*b;
                normalize3(p);
                double bx = 0.5 + SDF_RADIUS*p[0], by = 0.5 + SDF_RADIUS*p[1], bz = 0.5 + SDF_RADIUS*p[2];
                double r = SDF_RADIUS + SDF_AMPLITUDE*svo_scene_fbm(bx, by, bz, seed);
                row[3*j] = (float)(0.5 + r*p[0]);
                row[3*j + 1] = (float)(0.5 + r*p[1]);
                row[3*j + 2] = (float)(0.5 + r*p[2]);
            }
            fwrite(row, sizeof(float)*3, (size_t)(n - i + 1), fp);
        }
    }
    free(row);

    /* index of (i, j) inside a face: rows 0..i-1 hold (n+1) + n + ... entries */
    unsigned char *rec = (unsigned char *)malloc(13*2*(size_t)n);
    for (int f = 0; f < 20; ++f) {
        int64_t base = f*vertsPerFace;
        for (int i = 0; i < n; ++i) {
            int64_t r0 = base + (int64_t)i*(n + 1) - (int64_t)i*(i - 1)/2;   /* start of row i */
            int64_t r1 = r0 + (n - i + 1);                                    /* start of row i + 1 */
            size_t cnt = 0;
            for (int j = 0; j < n - i; ++j) {
                int32_t t0[3] = {(int32_t)(r0 + j), (int32_t)(r1 + j), (int32_t)(r0 + j + 1)};
                rec[cnt] = 3; memcpy(rec + cnt + 1, t0, 12); cnt += 13;
                if (j < n - i - 1) {
                    int32_t t1[3] = {(int32_t)(r0 + j + 1), (int32_t)(r1 + j), (int32_t)(r1 + j + 1)};
                    rec[cnt] = 3; memcpy(rec + cnt + 1, t1, 12); cnt += 13;
                }
            }
            fwrite(rec, 1, cnt, fp);
        }
    }
    free(rec);
    int bad = ferror(fp);
    fclose(fp);
    return bad ? -1 : nTris;
}
