// Device-side ESVO traversal, ray generation and shading for sm_100a.
//
// What it computes is fixed by the reference (tunabrain/sparse-voxel-octrees):
//   raymarch            reference src/VoxelOctree.cpp:207-346
//   shade               reference src/Main.cpp:81-90
//   decompressMaterial  reference src/Util.hpp:86-100, invSqrt :47-58
//   pixel pack          reference src/Main.cpp:128-132
// How it computes it is not: one thread per ray, traversal state in registers,
// the scale-indexed (parent, maxT) stack in shared memory (conflict-free
// [slot][thread] layout), read-only node fetches with the possible far word
// fetched speculatively next to its descriptor, 32-bit word indices whenever
// the tree has < 2^32 words.
//
// Arithmetic flavours (svo_flavour in include/svo_b200.h): every float
// operation below goes through an explicitly rounded intrinsic, so neither
// flavour depends on the compiler's contraction setting. FAST fuses the
// traversal's a*b+-c forms into single FMAs where the product is exact (a power
// of two times a float) and turns min/max into FMNMX; see Arith below for why
// the general a*b-c stays unfused.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace svo {

constexpr int kMaxScale = 23;      // reference src/VoxelOctree.hpp:38
constexpr float kTreeMiss = 1e10f; // reference src/Main.cpp:140

template <bool FAST>
struct Arith {
    // a*b - c with an arbitrary product: two roundings in BOTH flavours. Fusing
    // this one (the per-iteration corner planes pos*dT - bT) was measured to move
    // 0.016-0.018 % of the pixels to a neighbouring voxel (grazing rays decided by
    // the last bit) -- over the 0.01 % the FAST flavour is allowed.
    static __device__ __forceinline__ float mulsub(float a, float b, float c) {
        return __fsub_rn(__fmul_rn(a, b), c);
    }
    // p*b - c and p*b + c where p is a power of two: the product is exact, so the
    // single FMA rounds exactly like the reference's multiply-then-add.
    static __device__ __forceinline__ float pow2mulsub(float p, float b, float c) {
        if (FAST) return __fmaf_rn(p, b, -c);
        return __fsub_rn(__fmul_rn(p, b), c);
    }
    static __device__ __forceinline__ float pow2muladd(float p, float b, float c) {
        if (FAST) return __fmaf_rn(p, b, c);
        return __fadd_rn(__fmul_rn(p, b), c);
    }
    // std::min(a, b) == (b < a) ? b : a ; std::max(a, b) == (a < b) ? b : a
    static __device__ __forceinline__ float min2(float a, float b) {
        if (FAST) return fminf(a, b);
        return (b < a) ? b : a;
    }
    static __device__ __forceinline__ float max2(float a, float b) {
        if (FAST) return fmaxf(a, b);
        return (a < b) ? b : a;
    }
};

// Exactly rounded helpers shared by both flavours (ray generation, shading).
__device__ __forceinline__ float mulRn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float addRn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float subRn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float minStd(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float maxStd(float a, float b) { return (a < b) ? b : a; }

// reference src/Util.hpp:47-58
__device__ __forceinline__ float invSqrtQuake(float x) {
    float halfX = mulRn(x, 0.5f);
    float y = __uint_as_float(0x5f3759dfu - (__float_as_uint(x) >> 1));
    return mulRn(y, subRn(1.5f, mulRn(mulRn(halfX, y), y)));
}

// x*x + y*y + z*z, left to right (reference src/math/Vec3.hpp:49-51)
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return addRn(addRn(mulRn(ax, bx), mulRn(ay, by)), mulRn(az, bz));
}

struct FrameConsts {
    float posX, posY, posZ;
    float a11, a12, a21, a22, a31, a32;
    float zx, zy, zz;
    float lightX, lightY, lightZ;
    float coarseScale;
    float beamBias;
};

// dir = normalize(dx*col1 + dy*col2 + z), reference src/Main.cpp:108-113 / :171-176
__device__ __forceinline__ void rayDirection(const FrameConsts &f, float dx, float dy, float &x, float &y, float &z) {
    x = addRn(addRn(mulRn(dx, f.a11), mulRn(dy, f.a12)), f.zx);
    y = addRn(addRn(mulRn(dx, f.a21), mulRn(dy, f.a22)), f.zy);
    z = addRn(addRn(mulRn(dx, f.a31), mulRn(dy, f.a32)), f.zz);
    float s = invSqrtQuake(dot3(x, y, z, x, y, z));
    x = mulRn(x, s);
    y = mulRn(y, s);
    z = mulRn(z, s);
}

// reference src/Main.cpp:81-90 + src/Util.hpp:86-100; returns the grey value
__device__ __forceinline__ float shadeMaterial(uint32_t word, float rx, float ry, float rz, float lx, float ly, float lz) {
    uint32_t face = (word >> 29) & 3u;
    float u = subRn(mulRn(mulRn(float((word >> 18) & 0x7FFu), 4.8852e-4f), 2.0f), 1.0f);
    float v = subRn(mulRn(mulRn(float((word >> 7) & 0x7FFu), 4.8852e-4f), 2.0f), 1.0f);
    float s = (word & 0x80000000u) ? -1.0f : 1.0f;
    // n[face] = s, n[(face+1)%3] = u, n[(face+2)%3] = v
    float nx = (face == 0) ? s : ((face == 1) ? v : u);
    float ny = (face == 0) ? u : ((face == 1) ? s : v);
    float nz = (face == 0) ? v : ((face == 1) ? u : s);
    float inv = invSqrtQuake(dot3(nx, ny, nz, nx, ny, nz));
    nx = mulRn(nx, inv);
    ny = mulRn(ny, inv);
    nz = mulRn(nz, inv);
    float c = __fdiv_rn(mulRn(float(word & 0x7Fu), 1.0f), 127.0f);

    float proj = mulRn(dot3(nx, ny, nz, rx, ry, rz), 2.0f);     // Vec3::reflect, Vec3.hpp:63-71
    float qx = subRn(rx, mulRn(nx, proj));
    float qy = subRn(ry, mulRn(ny, proj));
    float qz = subRn(rz, mulRn(nz, proj));
    float d = maxStd(dot3(lx, ly, lz, qx, qy, qz), 0.0f);
    float specular = mulRn(d, d);
    return addRn(mulRn(mulRn(c, 0.9f), fabsf(dot3(lx, ly, lz, nx, ny, nz))), mulRn(specular, 0.2f));
}

// uint32(std::min(v, 1.0f)*255.0) replicated to r,g,b with alpha 0xFF (Main.cpp:128-132).
// The reference multiplies in double; the product of a 24-bit and an 8-bit
// significand is exact there, and rounding the same product toward zero in
// float keeps every bit above 2^0, so truncation gives the same integer.
__device__ __forceinline__ uint32_t packGrey(float v) {
    uint32_t g = __float2uint_rz(__fmul_rz(minStd(v, 1.0f), 255.0f));
    return g | (g << 8) | (g << 16) | 0xFF000000u;
}

// The loop is bound by the SM's ALU pipe (integer / logic / compare / select: one warp instruction
// per two cycles; ncu: 73 % of its peak, against 24 % on the FMA pipe, which issues every cycle).
// So the per-axis "compare, conditionally move the position, collect a 3-bit mask" idiom spends
// one ALU instruction per axis -- FSET, a compare that writes 1.0f / 0.0f -- and does the rest on
// the FMA pipe: the position moves by flag*delta (exact: the product is 0 or delta) and the mask is
// accumulated as flagX + 2*flagY + 4*flagZ + 2^23, whose low mantissa bits are the mask, instead of
// SEL / IADD / FSEL chains.
__device__ __forceinline__ float flagGreater(float a, float b) {
    float r;
    asm("set.gt.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float flagNotGreater(float a, float b) {
    float r;
    asm("set.le.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
// p + flag*delta with flag in {0, 1}: exact product, one rounding-free add in [1, 2)
template <bool FAST>
__device__ __forceinline__ float moveIf(float flag, float delta, float p) {
    if (FAST) return __fmaf_rn(flag, delta, p);
    return __fadd_rn(__fmul_rn(flag, delta), p);
}
// bits 0-2 of the result = fx | fy << 1 | fz << 2 (flags are 0.0f / 1.0f; all sums are exact integers < 2^24)
__device__ __forceinline__ uint32_t flagMask(float fx, float fy, float fz) {
    float m = __fmaf_rn(fz, 4.0f, __fmaf_rn(fy, 2.0f, __fadd_rn(fx, 8388608.0f)));
    return __float_as_uint(m);
}

// The scale-indexed stack (reference: StackEntry rayStack[24], VoxelOctree.cpp:208-212) lives in shared memory, one
// entry per (slot, thread): slot s of thread t is at byte s*STRIDE + t*ENTRY, so a warp touching one slot makes one
// conflict-free (vector) access. Addresses are 32-bit shared-window addresses computed once per thread; push / pop are
// a single st.shared / ld.shared.
//
// WITH_MAXT == false (rays without the LOD test: the fine pass, plain batches): the entry is the parent index alone.
// The reference also saves maxT, the exit t of the parent's cell, and tests `minT <= maxT && minT <= min(maxT, maxTC)`
// (:263, :270-273) -- but maxTC <= maxT always holds: the child cell lies inside the parent's, its corner is >= the
// parent's corner on every axis, and p -> fl(fl(p*dT) - bT) is monotone non-increasing in p (dT < 0; IEEE rounding is
// monotone), so min(maxT, maxTC) == maxTC bit for bit and the two tests are `minT <= maxTC`. (Checked on 277 M loop
// trips of the oracle -- frame rays, ambient-occlusion rays and random rays that miss the cube: 0 exceptions.) With
// the LOD test in between (:265-268) `minT <= maxT` alone gates the LOD exit, so those rays keep maxT.
template <typename IdxT, int THREADS, bool WITH_MAXT>
struct SmemStack {
    static constexpr uint32_t kEntry = (sizeof(IdxT) == 8 ? 8u : 4u)*(WITH_MAXT ? 2u : 1u);
    static constexpr uint32_t kStride = THREADS*kEntry;
    uint32_t top;   // address of the slot for scale 22 (the root's children)
    __device__ __forceinline__ void init(const void *smem) {
        top = uint32_t(__cvta_generic_to_shared(smem)) + threadIdx.x*kEntry;
        asm volatile("" : "+r"(top));   // keep it in a register; ptxas would re-derive it at every pop
    }
    // The slot of `scale` is top + (22 - scale)*kStride = slot(scale) + kBias: one IMAD for the register
    // part, the constant part rides in the instruction's immediate offset.
    static constexpr uint32_t kBias = uint32_t(kMaxScale - 1)*kStride;
    __device__ __forceinline__ uint32_t slot(int scale) const { return top - uint32_t(scale)*kStride; }
    static __device__ __forceinline__ void store(uint32_t addr, IdxT parent, float maxT) {
        if (sizeof(IdxT) == 4 && WITH_MAXT)
            asm volatile("st.shared.v2.b32 [%0+%3], {%1, %2};" ::"r"(addr), "r"(uint32_t(parent)), "r"(__float_as_uint(maxT)), "n"(kBias) : "memory");
        else if (sizeof(IdxT) == 4)
            asm volatile("st.shared.b32 [%0+%2], %1;" ::"r"(addr), "r"(uint32_t(parent)), "n"(kBias) : "memory");
        else if (WITH_MAXT)
            asm volatile("st.shared.v4.b32 [%0+%5], {%1, %2, %3, %4};" ::"r"(addr), "r"(uint32_t(parent)),
                         "r"(uint32_t(uint64_t(parent) >> 32)), "r"(__float_as_uint(maxT)), "r"(0u), "n"(kBias) : "memory");
        else
            asm volatile("st.shared.v2.b32 [%0+%3], {%1, %2};" ::"r"(addr), "r"(uint32_t(parent)), "r"(uint32_t(uint64_t(parent) >> 32)), "n"(kBias) : "memory");
    }
    static __device__ __forceinline__ void load(uint32_t addr, IdxT &parent, float &maxT) {
        uint32_t lo = 0, hi = 0, m = 0, pad;
        if (sizeof(IdxT) == 4 && WITH_MAXT)
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2+%3];" : "=r"(lo), "=r"(m) : "r"(addr), "n"(kBias) : "memory");
        else if (sizeof(IdxT) == 4)
            asm volatile("ld.shared.b32 %0, [%1+%2];" : "=r"(lo) : "r"(addr), "n"(kBias) : "memory");
        else if (WITH_MAXT)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(lo), "=r"(hi), "=r"(m), "=r"(pad) : "r"(addr), "n"(kBias) : "memory");
        else
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2+%3];" : "=r"(lo), "=r"(hi) : "r"(addr), "n"(kBias) : "memory");
        parent = sizeof(IdxT) == 8 ? IdxT((uint64_t(hi) << 32) | lo) : IdxT(lo);
        if (WITH_MAXT) maxT = __uint_as_float(m);
    }
    static size_t bytes(uint32_t slots) { return size_t(slots)*kStride; }
};

__device__ __forceinline__ uint32_t ldNode(const uint32_t *__restrict__ p) { return __ldg(p); }

enum : int { kMiss = 0, kHitLeaf = 1, kHitLod = 2 };

// ---- the traversal, in three pieces ------------------------------------------------------------------------------
// reference src/VoxelOctree.cpp:207-346. RayState is the loop's state (registers), rayBegin the set-up (:214-250) with
// the first node fetch, rayTrip ONE trip round the loop (:252-339). raymarch() below runs them to completion for one
// ray per thread; the persistent kernels (svo_kernels.cu) keep a RayState per lane and hand a lane its next ray when it
// finishes one ("lane refill"), so both share every instruction of the traversal itself.
template <typename IdxT>
struct RayState {
    float dTx, dTy, dTz;        // 1 / -|d|, :221-223
    float bTx, bTy, bTz;        // :225-232
    uint32_t octantMask;
    float minT, maxT;
    uint32_t current, farWord;  // descriptor of `parent` and the word behind it
    IdxT parent;
    float posX, posY, posZ;
    int scale;
    float scaleExp2;
    uint32_t childShift;        // idx ^ octantMask (:261) while the ray is alive; kExitLeaf / kExitLod / kExitMiss after
};
constexpr uint32_t kExitLeaf = 8u, kExitLod = 16u, kExitMiss = 32u;

// Descriptor and its possible far word (bit 17) are fetched together (the node array carries one padding word so
// parent + 1 is always readable), right where `parent` changes: at the start, at the end of a push and at the end of
// a pop -- the reference's `current == 0` refetch flag (:253-254, :305, :337) never has to be tested.
template <typename IdxT>
__device__ __forceinline__ void fetchNode(const uint32_t *__restrict__ octree, RayState<IdxT> &r) {
    const uint32_t *node = octree + r.parent;
    r.current = ldNode(node);
    r.farWord = ldNode(node + 1);
}

// dT, bT and the octant mirroring of a ray (:214-232): everything about the ray that does not depend on where in the
// tree it is.
template <bool FAST, typename IdxT>
__device__ __forceinline__ void raySetup(float ox, float oy, float oz, float dx, float dy, float dz, RayState<IdxT> &r) {
    typedef Arith<FAST> A;
    if (fabsf(dx) < 1e-4f) dx = 1e-4f;      // :217-219, sign dropped on purpose
    if (fabsf(dy) < 1e-4f) dy = 1e-4f;
    if (fabsf(dz) < 1e-4f) dz = 1e-4f;

    r.dTx = __fdiv_rn(1.0f, -fabsf(dx));    // :221-223
    r.dTy = __fdiv_rn(1.0f, -fabsf(dy));
    r.dTz = __fdiv_rn(1.0f, -fabsf(dz));

    r.bTx = mulRn(r.dTx, ox);               // :225-227
    r.bTy = mulRn(r.dTy, oy);
    r.bTz = mulRn(r.dTz, oz);

    uint32_t octantMask = 7;                // :229-232
    if (dx > 0.0f) { octantMask ^= 1; r.bTx = A::mulsub(3.0f, r.dTx, r.bTx); }
    if (dy > 0.0f) { octantMask ^= 2; r.bTy = A::mulsub(3.0f, r.dTy, r.bTy); }
    if (dz > 0.0f) { octantMask ^= 4; r.bTz = A::mulsub(3.0f, r.dTz, r.bTz); }
    // keep the mask in a register: ptxas otherwise re-derives it from the
    // direction signs on every trip round the loop (12 extra instructions)
    asm volatile("" : "+r"(octantMask));
    r.octantMask = octantMask;
}

// The state at the root (:234-250) with the root's descriptor fetched.
template <bool FAST, typename IdxT>
__device__ __forceinline__ void rayRoot(const uint32_t *__restrict__ octree, RayState<IdxT> &r) {
    typedef Arith<FAST> A;
    float minT = maxStd(A::pow2mulsub(2.0f, r.dTx, r.bTx), maxStd(A::pow2mulsub(2.0f, r.dTy, r.bTy), A::pow2mulsub(2.0f, r.dTz, r.bTz)));
    r.maxT = minStd(subRn(r.dTx, r.bTx), minStd(subRn(r.dTy, r.bTy), subRn(r.dTz, r.bTz)));
    minT = maxStd(minT, 0.0f);
    r.minT = minT;

    r.current = 0;
    r.farWord = 0;
    r.parent = 0;
    uint32_t idx = 0;
    r.posX = 1.0f; r.posY = 1.0f; r.posZ = 1.0f;
    r.scale = kMaxScale - 1;
    r.scaleExp2 = 0.5f;

    if (A::mulsub(1.5f, r.dTx, r.bTx) > minT) { idx ^= 1; r.posX = 1.5f; }   // :248-250
    if (A::mulsub(1.5f, r.dTy, r.bTy) > minT) { idx ^= 2; r.posY = 1.5f; }
    if (A::mulsub(1.5f, r.dTz, r.bTz) > minT) { idx ^= 4; r.posZ = 1.5f; }
    // the loop tracks the reference's `idx` as childShift = idx ^ octantMask (:261), which is what it uses
    r.childShift = idx ^ r.octantMask;
    fetchNode(octree, r);
}

template <bool FAST, typename IdxT>
__device__ __forceinline__ void rayBegin(const uint32_t *__restrict__ octree, float ox, float oy, float oz,
                                         float dx, float dy, float dz, RayState<IdxT> &r) {
    raySetup<FAST, IdxT>(ox, oy, oz, dx, dy, dz, r);
    rayRoot<FAST, IdxT>(octree, r);
}

// ---- shared traversal prefix of a tile (FAST flavour of the fine pass only; see tilePrefixKernel in svo_kernels.cu) ----
// A record is kPrefixHeaderWords words -- parent, scale | childShift << 8 | octantMask << 16 | valid << 24, posX, posY,
// posZ -- followed by the parents saved on the stack for the scales above `scale`: word[5 + 22 - sc] for sc in (scale, 22].
constexpr int kPrefixHeaderWords = 5;

// Puts a ray into the loop-head state a tile's four corner rays shared last: the ray computes its own entry t into that
// cell (the reference's minT there is the exit plane of the cell it came from, which is this cell's entry plane on that
// axis -- the same float expression -- or 0 when the ray starts inside) and takes over the parents above it. Returns
// false, with `r` untouched beyond its set-up, when the record does not apply to this ray (other octant, or the ray's
// own interval in the cell is empty): the caller then starts at the root.
template <bool FAST, int THREADS>
__device__ __forceinline__ bool rayRestart(const uint32_t *__restrict__ octree, RayState<uint32_t> &r,
                                           const uint32_t *__restrict__ rec, const SmemStack<uint32_t, THREADS, false> &stack) {
    typedef Arith<FAST> A;
    typedef SmemStack<uint32_t, THREADS, false> Stack;
    const uint4 head = __ldg(reinterpret_cast<const uint4 *>(rec));
    const uint32_t packed = head.y;
    if ((packed >> 24) == 0u || ((packed >> 16) & 7u) != r.octantMask) return false;
    const int scale = int(packed & 0xFFu);
    const float se = __uint_as_float(uint32_t(scale - kMaxScale + 127) << 23);
    const float px = __uint_as_float(head.z), py = __uint_as_float(head.w), pz = __uint_as_float(__ldg(rec + 4));
    const float exitT = fminf(A::mulsub(px, r.dTx, r.bTx), fminf(A::mulsub(py, r.dTy, r.bTy), A::mulsub(pz, r.dTz, r.bTz)));
    const float entryT = fmaxf(fmaxf(A::mulsub(addRn(px, se), r.dTx, r.bTx), A::mulsub(addRn(py, se), r.dTy, r.bTy)),
                               fmaxf(A::mulsub(addRn(pz, se), r.dTz, r.bTz), 0.0f));
    if (!(entryT <= exitT)) return false;
    r.parent = head.x;
    r.scale = scale;
    r.scaleExp2 = se;
    r.childShift = (packed >> 8) & 0xFFu;
    r.posX = px; r.posY = py; r.posZ = pz;
    r.minT = entryT;
    r.maxT = 0.0f;      // not carried by rays without the LOD test
    for (int sc = scale + 1; sc < kMaxScale; ++sc)       // the same for every ray of the tile: a uniform loop of broadcast loads
        Stack::store(stack.slot(sc), __ldg(rec + kPrefixHeaderWords + (kMaxScale - 1 - sc)), 0.0f);
    fetchNode(octree, r);
    return true;
}

// One trip round the reference's loop (:252-339). Returns true when the ray is finished: r.childShift then says how
// (kExitLeaf: tOut / voxelOut = leaf word index; kExitLod: tOut / voxelOut = parent | childShift << 60; kExitMiss:
// voxelOut = some valid word index, so that callers can fetch octree[voxelOut] without a select).
// How the loop was left is recorded in childShift (always < 8 while the ray is alive): a separate result flag would
// have to be set on the pop path, which every pop executes, for the sake of the one pop per ray that leaves the root.
// LOD == false elides the rayScale test (rayScale == 0 can never pass it: maxTC*0 is +-0 or NaN, scaleExp2 > 0).
//
// min/max are FMNMX in both flavours here: every operand is a finite p*dT - bT with p*dT != 0, which can be +0 but
// never -0 or NaN, so FMNMX and the reference's (b < a) ? b : a select the same bits.
template <bool FAST, bool LOD, typename IdxT, int THREADS>
__device__ __forceinline__ bool rayTrip(const uint32_t *__restrict__ octree, RayState<IdxT> &r, float rayScale,
                                        const SmemStack<IdxT, THREADS, LOD> &stack, float &tOut, uint64_t &voxelOut) {
    typedef Arith<FAST> A;
    typedef SmemStack<IdxT, THREADS, LOD> Stack;

    const float cornerTX = A::mulsub(r.posX, r.dTx, r.bTx);   // :256-259
    const float cornerTY = A::mulsub(r.posY, r.dTy, r.bTy);
    const float cornerTZ = A::mulsub(r.posZ, r.dTz, r.bTz);
    const float maxTC = fminf(cornerTX, fminf(cornerTY, cornerTZ));

    const uint32_t childMasks = r.current << r.childShift;

    // :263-273. Without the LOD test in between, `minT <= maxT && minT <= min(maxT, maxTC)` is `minT <= maxTC`
    // (maxTC <= maxT always, see SmemStack), so the rays of the fine pass make one comparison and carry no maxT.
    const float maxTV = LOD ? fminf(r.maxT, maxTC) : maxTC;
    bool descend = (childMasks & 0x8000u) != 0;
    if (LOD) {
        descend = descend && r.minT <= r.maxT;
        if (descend && mulRn(maxTC, rayScale) >= r.scaleExp2) {   // :265-268
            tOut = maxTC;
            voxelOut = uint64_t(r.parent) | (uint64_t(r.childShift) << 60);
            r.childShift = kExitLod;
            return true;
        }
    }
    if (descend && r.minT <= maxTV) {
        IdxT childOffset = IdxT(r.current >> 18);
        if (r.current & 0x20000u) {                      // :278-279
            if (sizeof(IdxT) == 8)
                childOffset = IdxT((uint64_t(childOffset) << 32) | uint64_t(r.farWord));
            else
                childOffset = IdxT(r.farWord);           // high 14 bits are zero below 2^32 words
        }

        if (!(childMasks & 0x80u)) {                     // leaf, :281-285
            IdxT leaf = childOffset + r.parent + IdxT(__popc(((childMasks >> (8 + r.childShift)) << r.childShift) & 127u));
            voxelOut = uint64_t(leaf);
            tOut = r.minT;
            r.childShift = kExitLeaf;
            return true;
        }

        Stack::store(stack.slot(r.scale), r.parent, r.maxT);  // :287-288

        // siblings before this child, doubled when the block is far-interleaved (bit 16), :290-293
        uint32_t siblings = uint32_t(__popc(childMasks & 127u));
        // one predicate test + one predicated shift (the C forms compile to three or four instructions)
        asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %1, 0x10000;\n\tsetp.ne.u32 p, t, 0;\n\t@p shl.b32 %0, %0, 1;\n\t}"
            : "+r"(siblings) : "r"(r.current));
        r.parent += childOffset + IdxT(siblings);

        const float half = mulRn(r.scaleExp2, 0.5f);
        const float centerTX = A::pow2muladd(half, r.dTx, cornerTX);   // half is a power of two
        const float centerTY = A::pow2muladd(half, r.dTy, cornerTY);
        const float centerTZ = A::pow2muladd(half, r.dTz, cornerTZ);
        r.scale--;
        r.scaleExp2 = half;

        // idx = axes with centerT > minT, each of which moves to the upper half (:297-301)
        const float upX = flagGreater(centerTX, r.minT), upY = flagGreater(centerTY, r.minT), upZ = flagGreater(centerTZ, r.minT);
        r.posX = moveIf<FAST>(upX, half, r.posX);
        r.posY = moveIf<FAST>(upY, half, r.posY);
        r.posZ = moveIf<FAST>(upZ, half, r.posZ);
        r.childShift = (flagMask(upX, upY, upZ) & 7u) ^ r.octantMask;

        if (LOD) r.maxT = maxTV;
        fetchNode(octree, r);
        return false;
    }

    // :310-316: every axis whose plane is reached at maxTC steps down by one cell
    const float stX = flagNotGreater(cornerTX, maxTC), stY = flagNotGreater(cornerTY, maxTC), stZ = flagNotGreater(cornerTZ, maxTC);
    r.posX = moveIf<FAST>(stX, -r.scaleExp2, r.posX);
    r.posY = moveIf<FAST>(stY, -r.scaleExp2, r.posY);
    r.posZ = moveIf<FAST>(stZ, -r.scaleExp2, r.posZ);
    const uint32_t stepMask = flagMask(stX, stY, stZ) & 7u;
    r.minT = maxTC;
    // idx ^= stepMask; pop if (idx & stepMask) != 0 (:316-318)  <=>  a stepped axis had its idx bit clear
    const bool leavesParent = (~(r.childShift ^ r.octantMask) & stepMask) != 0;
    r.childShift ^= stepMask;

    if (leavesParent) {                                     // :318-338
        // pos ^ (pos + scaleExp2) over the stepped axes (:320-322); an axis that did not step adds
        // 0*scaleExp2 and contributes nothing. (Keeping the pre-step positions instead costs three
        // register moves on EVERY trip: ptxas copies them at the loop head.)
        const uint32_t differingBits =
            (__float_as_uint(r.posX) ^ __float_as_uint(moveIf<FAST>(stX, r.scaleExp2, r.posX))) |
            (__float_as_uint(r.posY) ^ __float_as_uint(moveIf<FAST>(stY, r.scaleExp2, r.posY))) |
            (__float_as_uint(r.posZ) ^ __float_as_uint(moveIf<FAST>(stZ, r.scaleExp2, r.posZ)));
        // reference: exponent of (float)differingBits. differingBits < 2^24
        // always (positions stay in [0.5, 2)), so that is the index of the
        // highest set bit; bit 23 set <=> the ray left the root (:341-342)
        if (differingBits > 0x7FFFFFu) {
            voxelOut = uint64_t(r.parent);   // any valid word index: the caller may fetch it unconditionally
            r.childShift = kExitMiss;
            return true;
        }
        asm("bfind.u32 %0, %1;" : "=r"(r.scale) : "r"(differingBits));   // FLO: index of the highest set bit
        r.scaleExp2 = __uint_as_float(uint32_t(r.scale - kMaxScale + 127) << 23);

        Stack::load(stack.slot(r.scale), r.parent, r.maxT);
        fetchNode(octree, r);

        // Truncate the positions to the `scale` grid (:329-334) on the FMA pipe: adding 2^scale with
        // round-toward-zero leaves exactly the mantissa bits >= `scale` (positions are in [1, 2), the
        // sum is in [2^scale, 2^(scale+1))), its lowest mantissa bit is the new idx bit, and
        // subtracting 2^scale again is exact.
        const float big = __uint_as_float(uint32_t(r.scale + 127) << 23);
        const float tX = __fadd_rz(r.posX, big), tY = __fadd_rz(r.posY, big), tZ = __fadd_rz(r.posZ, big);
        r.posX = subRn(tX, big);
        r.posY = subRn(tY, big);
        r.posZ = subRn(tZ, big);
        // idx = bit 0 of tX | bit 0 of tY << 1 | bit 0 of tZ << 2, by two bit-selects
        const uint32_t xy = (__float_as_uint(tX) & 1u) | ((__float_as_uint(tY) << 1) & ~1u);
        const uint32_t xyz = (xy & 3u) | ((__float_as_uint(tZ) << 2) & ~3u);
        r.childShift = (xyz ^ r.octantMask) & 7u;
    }
    return false;
}

__device__ __forceinline__ int exitCode(uint32_t childShift, bool lod) {
    return childShift == kExitLeaf ? kHitLeaf : (lod && childShift == kExitLod) ? kHitLod : kMiss;
}

// One ray per thread, run to completion. Returns kMiss / kHitLeaf / kHitLod.
//   tOut      written on a hit only
//   voxelOut  leaf word index (the caller fetches the material word octree[voxelOut], :282, AFTER the
//             warp has reconverged), or parent | childShift << 60 for LOD exits; on a miss some valid word index
// The loop has ONE exit: every way out records its result and breaks, and a warp barrier follows the
// loop. Without it the compiler threads whatever the caller does with a hit (material fetch + ~100
// instructions of shading) into the leaf branch inside the loop, where it runs once per distinct exit
// trip of the warp (8-10 times per warp, with 3 lanes active) instead of once with all lanes.
template <bool FAST, bool LOD, typename IdxT, int THREADS>
__device__ __forceinline__ int raymarch(const uint32_t *__restrict__ octree, float ox, float oy, float oz,
                                        float dx, float dy, float dz, float rayScale,
                                        const SmemStack<IdxT, THREADS, LOD> &stack,
                                        float &tOut, uint64_t &voxelOut) {
    RayState<IdxT> r;
    rayBegin<FAST, IdxT>(octree, ox, oy, oz, dx, dy, dz, r);
    for (;;)
        if (rayTrip<FAST, LOD, IdxT, THREADS>(octree, r, rayScale, stack, tOut, voxelOut)) break;
    __syncwarp();   // exited lanes do not take part; see the note on the single exit above
    return exitCode(r.childShift, LOD);
}

} // namespace svo
