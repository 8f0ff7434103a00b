// rrt_kernels.cu -- hand-written sm_100a kernels + the C ABI of include/rrt_b200.h.
//
// One hot path of lebek/reversible-raytracer, rebuilt B200-first (paths below are
// relative to the reference checkout; nothing here is translated from it -- the
// reference is a dense Theano graph, this is a per-ray register-resident design):
//   Camera.make_rays        scene.py:61-75 (+ orbit_experiments/scene.py:55-80)
//   Transform.__call__      transform.py:40-47
//   Sphere / Square         shape.py:25-69, 78-83, 109-138
//   Phong / DepthMap        shader.py:14-20, 28-53
//   Scene.build             scene.py:18-52
//   T.grad(loss, params)    optimize.py:25,73
//
// Design (DESIGN.md has the long form):
//   * one thread owns PIX pixels x SPT anti-alias samples = 8 rays, kept in
//     registers as 4 packed pairs; the ray-object sweep is issued as packed
//     FFMA2/FMUL2 (fma.rn.f32x2 / mul.rn.f32x2, sm_100+) with the object constants
//     broadcast from shared memory (LDS.128, scalar-broadcast operand form);
//   * the object table is transformed once per CTA into 64-byte sweep records in
//     shared memory, streamed in chunks when N is large;
//   * hits are rare per (ray, object): the hot loop is branch-free over groups of 4
//     objects (packed discriminants folded into a running max, FMNMX3), one warp-uniform
//     compare-and-branch per group into an out-of-line routine that re-tests the group
//     and runs the scalar canonical-order hit computation (sqrt, div, strict '<');
//   * shading and the reverse pass run once per winning ray after the sweep; hit
//     records are recomputed, never stored; per-object gradients are reduced
//     thread -> warp (shuffle) -> CTA (shared-memory slots) -> one atomic per CTA.
//   * every mask-determining operation is an explicit IEEE round-to-nearest
//     intrinsic or PTX instruction in the canonical order shared with
//     oracle/oracle_c.c, so hit masks are bit-exact against the oracle.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rrt_b200.h"

namespace {

typedef unsigned long long u64;

#ifndef RRT_RAYS
#define RRT_RAYS 8
#endif
constexpr int kRays = RRT_RAYS;   // rays per thread (kRays/2 packed pairs)
// Tuning knobs (measured on B200, C5: DESIGN.md 4.3).  8 rays/thread, 128-thread CTAs at
// 5 CTAs/SM (96 registers), 512-object chunks (32 KB) and groups of 4 were the best of the
// variants tried; 16 rays/thread, 256-thread CTAs, groups of 2/3/5/6/8 were all slower.
#ifndef RRT_OBJ_CHUNK
#define RRT_OBJ_CHUNK 512
#endif
#ifndef RRT_GROUP
#define RRT_GROUP 4
#endif
#ifndef RRT_SWEEP_UNROLL
#define RRT_SWEEP_UNROLL 1
#endif
#ifndef RRT_MIN_BLOCKS
#define RRT_MIN_BLOCKS 5
#endif
#ifndef RRT_MAX_WARPS
#define RRT_MAX_WARPS 4
#endif
constexpr int kObjChunk = RRT_OBJ_CHUNK;   // objects staged in shared memory at a time (64 B each)
constexpr int kSlots = 32;        // per-CTA gradient slots (object -> 19 floats)
constexpr int kSlotStride = 20;
constexpr int kMaxWarps = RRT_MAX_WARPS;

enum { MODE_FWD = 0, MODE_BWD = 1, MODE_FUSED = 2 };

struct KParams {
    rrt_scene sc;
    int rows;
    float* image;
    int32_t* hit_out;
    float* tmin_out;
    const float* dl_dimage;
    const int32_t* hit_in;
    const float* target;
    float cw[3];
    double* loss;
    float* grad;
    // host-precomputed prologue constants
    double lin_step;   // 1/(n-1): np.linspace step magnitude (0 when n == 1)
    float inv_s, inv_n; // exact reciprocals when S and n are powers of two
    int pow2;          // 1: (u+s)/S/n may be evaluated as exact multiplications
    int vec_ok;        // 1: image/target rows are 16-byte aligned (n % 4 == 0, aligned base pointers)
    rrt_step step;     // rrt_small_step_mse only (whole optimise step in one launch)
};

// ---------------------------------------------------------------- packed f32x2
__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 bc(float v) { return pk(v, v); }  // ptxas folds this into the .F32 broadcast operand
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// ---------------------------------------------------------------- jitter RNG
// Counter-based 32-bit hash (lowbias32 finaliser) -> 24-bit uniform in [0,1).
// Same function (by specification) as orc_rng in oracle/oracle_c.c.
__device__ __forceinline__ float rrt_rng(u64 seed, uint32_t scene, uint32_t pix, uint32_t s, uint32_t axis) {
    uint32_t x = (pix * 0x9E3779B1u) ^ (scene * 0x85EBCA77u) ^ ((s * 2u + axis) * 0xC2B2AE3Du) ^
                 (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x27D4EB2Fu);
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return (float)(x >> 8) * 5.9604644775390625e-08f;
}

// ---------------------------------------------------------------- primary rays
// np.linspace(start, stop, n)[i] = fl(fl(i*step) + start), last element = stop; the
// step (stop-start)/(n-1) = +-1/(n-1) is computed once on the host in IEEE double.
__device__ __forceinline__ double lin(int i, int n, double start, double stop, double step) {
    if (n == 1) return start;
    if (i == n - 1) return stop;
    return __dadd_rn(__dmul_rn((double)i, step), start);
}

// Camera.make_rays scene.py:66-72: float64 grid, normalise, cast to float32.
__device__ __forceinline__ void base_ray(int n, double step, int i, int j, float& rx, float& ry, float& rz) {
    double x = lin(i, n, 0.5, -0.5, -step);
    double y = lin(j, n, -0.5, 0.5, step);
    double s = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), 1.0);
    double nrm = __dsqrt_rn(s);
    rx = __double2float_rn(__ddiv_rn(x, nrm));
    ry = __double2float_rn(__ddiv_rn(y, nrm));
    rz = __double2float_rn(__ddiv_rn(1.0, nrm));
}

// scene.py:31-32 then :73-74, all float32 round-to-nearest.
__device__ __forceinline__ float jitter_offset(float u, int s, int S, int n) {
    return __fdiv_rn(__fdiv_rn(__fadd_rn(u, (float)s), (float)S), (float)n);
}
// same value when S and n are powers of two (division by 2^k == multiplication by 2^-k, exact)
__device__ __forceinline__ float jitter_offset_pow2(float u, int s, float inv_s, float inv_n) {
    return __fmul_rn(__fmul_rn(__fadd_rn(u, (float)s), inv_s), inv_n);
}

__device__ __forceinline__ float dot3_canon(float a0, float a1, float a2, float v0, float v1, float v2) {
    return __fmaf_rn(a2, v2, __fmaf_rn(a1, v1, __fmul_rn(a0, v0)));
}

// ---------------------------------------------------------------- object records
// 64-byte sweep record (4 x float4) in shared memory:
//   q0 = (a00, a11, a22, o'x)   q1 = (o'y, o'z, -cc, flags)
//   q2 = (a01, a02, a10, a12)   q3 = (a20, a21, |A|_F, 0)
// flags bit0 = square, bit1 = general (some off-diagonal of A is non-zero).
struct Obj {
    float a[9];
    float o[3];
    float ncc;
    int flags;
    float afro;   // Frobenius norm of A (culling bound only)
};

struct Globals {       // per-scene constants, held in shared memory
    float C[9], ct[3]; // camera.o2w rows 0..2
    float look[3];
    float L[3], I[3];
    float Lh[3], Ln;
    float U[3];        // -Lhat in canonical float32 order (shadow mask only)
};

// canonical -Lhat (bit-identical to orc_prep in oracle/oracle_c.c): RN sqrt and div
__device__ __forceinline__ void canon_to_light(const float* L, float* U) {
    const float ln = __fsqrt_rn(__fmaf_rn(L[2], L[2], __fmaf_rn(L[1], L[1], __fmul_rn(L[0], L[0]))));
    U[0] = -__fdiv_rn(L[0], ln); U[1] = -__fdiv_rn(L[1], ln); U[2] = -__fdiv_rn(L[2], ln);
}

__device__ __forceinline__ void make_obj_rows(const float (&m)[12], int type, const float* ct, Obj& ob);

__device__ __forceinline__ void make_obj(const float* __restrict__ w, int type, const float* ct, Obj& ob,
                                         bool want_afro = false) {
    float m[12];
    const float4* w4 = reinterpret_cast<const float4*>(w);
    float4 r0 = __ldg(w4), r1 = __ldg(w4 + 1), r2 = __ldg(w4 + 2);
    m[0] = r0.x; m[1] = r0.y; m[2] = r0.z; m[3] = r0.w;
    m[4] = r1.x; m[5] = r1.y; m[6] = r1.z; m[7] = r1.w;
    m[8] = r2.x; m[9] = r2.y; m[10] = r2.z; m[11] = r2.w;
    (void)want_afro;
    make_obj_rows(m, type, ct, ob);
}

// from the 12 floats of w2o rows 0..2 (already in registers)
__device__ __forceinline__ void make_obj_rows(const float (&m)[12], int type, const float* ct, Obj& ob) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
        ob.a[r * 3 + 0] = m[r * 4 + 0];
        ob.a[r * 3 + 1] = m[r * 4 + 1];
        ob.a[r * 3 + 2] = m[r * 4 + 2];
        // o' = A.c + b  (transform.py:44)
        ob.o[r] = __fmaf_rn(m[r * 4 + 2], ct[2], __fmaf_rn(m[r * 4 + 1], ct[1], __fmaf_rn(m[r * 4 + 0], ct[0], m[r * 4 + 3])));
    }
    float cc = __fsub_rn(__fmaf_rn(ob.o[2], ob.o[2], __fmaf_rn(ob.o[1], ob.o[1], __fmul_rn(ob.o[0], ob.o[0]))), 1.0f);
    ob.ncc = -cc;
    bool general = (ob.a[1] != 0.f) || (ob.a[2] != 0.f) || (ob.a[3] != 0.f) || (ob.a[5] != 0.f) || (ob.a[6] != 0.f) || (ob.a[7] != 0.f);
    ob.flags = (type == RRT_OBJ_SQUARE ? 1 : 0) | (general ? 2 : 0);
    float f2 = 0.f;                                    // (dead-code eliminated where afro is unused)
#pragma unroll
    for (int q = 0; q < 9; q++) f2 += ob.a[q] * ob.a[q];
    ob.afro = sqrtf(f2);                               // culling bound only
}

__device__ __forceinline__ void store_rec(float4* rec, const Obj& ob) {
    rec[0] = make_float4(ob.a[0], ob.a[4], ob.a[8], ob.o[0]);
    rec[1] = make_float4(ob.o[1], ob.o[2], ob.ncc, __int_as_float(ob.flags));
    rec[2] = make_float4(ob.a[1], ob.a[2], ob.a[3], ob.a[5]);
    rec[3] = make_float4(ob.a[6], ob.a[7], ob.afro, 0.f);
}

__device__ __forceinline__ void load_rec(const float4* rec, Obj& ob) {
    float4 q0 = rec[0], q1 = rec[1], q2 = rec[2], q3 = rec[3];
    ob.a[0] = q0.x; ob.a[4] = q0.y; ob.a[8] = q0.z; ob.o[0] = q0.w;
    ob.o[1] = q1.x; ob.o[2] = q1.y; ob.ncc = q1.z; ob.flags = __float_as_int(q1.w);
    ob.a[1] = q2.x; ob.a[2] = q2.y; ob.a[3] = q2.z; ob.a[5] = q2.w;
    ob.a[6] = q3.x; ob.a[7] = q3.y; ob.afro = q3.z;
}

// ---------------------------------------------------------------- one ray-object test
// Canonical order: the scalar twin of the packed sweep; bit-identical
// to orc_test in oracle/oracle_c.c.
struct HitRec {
    float d[3];
    float vn, pd, det, t;
};

template <bool DIAG_SHORTCUT = false>
__device__ __forceinline__ float obj_test(const Obj& ob, float dwx, float dwy, float dwz, HitRec& h) {
    if (DIAG_SHORTCUT && !(ob.flags & 2)) {
        // diagonal A (translate*scale objects): the fma chain with exact-zero off-diagonals
        // returns a_ii*d_i up to the sign of a zero, which nothing downstream can see
        h.d[0] = __fmul_rn(ob.a[0], dwx);
        h.d[1] = __fmul_rn(ob.a[4], dwy);
        h.d[2] = __fmul_rn(ob.a[8], dwz);
    } else {
        h.d[0] = dot3_canon(ob.a[0], ob.a[1], ob.a[2], dwx, dwy, dwz);
        h.d[1] = dot3_canon(ob.a[3], ob.a[4], ob.a[5], dwx, dwy, dwz);
        h.d[2] = dot3_canon(ob.a[6], ob.a[7], ob.a[8], dwx, dwy, dwz);
    }
    const float inf = __int_as_float(0x7f800000);
    if (!(ob.flags & 1)) {  // Sphere.distance shape.py:109-126
        h.vn = dot3_canon(h.d[0], h.d[1], h.d[2], h.d[0], h.d[1], h.d[2]);
        h.pd = dot3_canon(h.d[0], h.d[1], h.d[2], ob.o[0], ob.o[1], ob.o[2]);
        h.det = __fmaf_rn(h.pd, h.pd, __fmul_rn(h.vn, ob.ncc));
        if (!(h.det > 0.0f)) return h.t = inf;
        float sq = __fsqrt_rn(h.det);
        return h.t = __fdiv_rn(__fsub_rn(-h.pd, sq), h.vn);
    } else {                // Square._hit shape.py:25-40
        float t = __fdiv_rn(-ob.o[2], h.d[2]);
        float px = __fmaf_rn(t, h.d[0], ob.o[0]);
        float py = __fmaf_rn(t, h.d[1], ob.o[1]);
        bool m = (h.d[2] != 0.0f) && (t > 0.0f) && (px > -0.5f) && (px < 0.5f) && (py > -0.5f) && (py < 0.5f);
        h.vn = h.pd = h.det = 0.f;
        return h.t = m ? t : inf;
    }
}

// ---------------------------------------------------------------- packed sweep
template <bool GENERAL>
__device__ __forceinline__ u64 pair_det(const float4& q0, const float4& q1, const float4& q2, const float4& q3,
                                        u64 dx, u64 dy, u64 dz) {
    u64 ex, ey, ez;
    if (GENERAL) {
        ex = fma2(bc(q2.y), dz, fma2(bc(q2.x), dy, mul2(bc(q0.x), dx)));
        ey = fma2(bc(q2.w), dz, fma2(bc(q0.y), dy, mul2(bc(q2.z), dx)));
        ez = fma2(bc(q0.z), dz, fma2(bc(q3.y), dy, mul2(bc(q3.x), dx)));
    } else {
        ex = mul2(bc(q0.x), dx);
        ey = mul2(bc(q0.y), dy);
        ez = mul2(bc(q0.z), dz);
    }
    u64 vn = fma2(ez, ez, fma2(ey, ey, mul2(ex, ex)));
    u64 pd = fma2(ez, bc(q1.y), fma2(ey, bc(q1.x), mul2(ex, bc(q0.w))));
    return fma2(pd, pd, mul2(vn, bc(q1.z)));
}

struct RayPack {
    u64 dx[kRays / 2], dy[kRays / 2], dz[kRays / 2];
};

// Rare path of the sweep, out of line on purpose (keeps the hot loop's register and
// code footprint small).  Re-tests `cnt` staged objects (chunk-local k0..k0+cnt-1,
// global index kbase+k) against the thread's 8 rays, re-read from local memory
// (SoA [x0..x7|y0..y7|z0..z7], 16-byte aligned): packed discriminants first, then
// the scalar canonical-order routine only for the (ray, object) pairs with det > 0.
// List order + strict '<' == scene.py:46-47 (the earlier shape wins ties).
__device__ __noinline__ void rare_group(const float4* __restrict__ tab, int k0, int cnt, int kbase,
                                        const float* dw, float* tmin, int* idx) {
    const u64* dp = reinterpret_cast<const u64*>(dw);
    u64 dx[kRays / 2], dy[kRays / 2], dz[kRays / 2];
#pragma unroll
    for (int p = 0; p < kRays / 2; p++) { dx[p] = dp[p]; dy[p] = dp[kRays / 2 + p]; dz[p] = dp[kRays + p]; }
#pragma unroll 1
    for (int j = 0; j < cnt; j++) {
        const float4* rec = tab + 4 * (k0 + j);
        const float4 q0 = rec[0], q1 = rec[1], q2 = rec[2], q3 = rec[3];
        const bool square = __float_as_int(q1.w) & 1;
        float det[kRays];
#pragma unroll
        for (int p = 0; p < kRays / 2; p++) upk(pair_det<true>(q0, q1, q2, q3, dx[p], dy[p], dz[p]), det[2 * p], det[2 * p + 1]);
        unsigned hits = 0;
#pragma unroll
        for (int r = 0; r < kRays; r++) hits |= ((square || det[r] > 0.0f) ? 1u : 0u) << r;
        if (!hits) continue;
        Obj ob;
        load_rec(rec, ob);
#pragma unroll 1
        while (hits) {
            const int r = __ffs(hits) - 1;
            hits &= hits - 1;
            HitRec h;
            const float t = obj_test(ob, dw[r], dw[kRays + r], dw[2 * kRays + r], h);
            if (t < tmin[r]) { tmin[r] = t; idx[r] = kbase + k0 + j; }
        }
    }
}

constexpr int kGroup = RRT_GROUP;  // objects per branch in the hot loop

// max over the 8 dets of one object, folded into the running group max (FMNMX3 chain;
// fmaxf drops NaN, and NaN is a miss: shape.py:124-125)
// LDS.128 from a 32-bit shared-window address: keeps the hot loop free of the
// generic->shared address arithmetic (S2UR/ULEA per iteration) a float4* would cost.
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}

// ---------------------------------------------------------------- TMA staging of precomputed records
// One elected thread arms an mbarrier with the chunk's byte count and issues ONE bulk copy
// global -> shared (cp.async.bulk, SASS UBLKCP); every thread then waits on the barrier's
// phase.  Replaces ~90 instructions per object and thread of in-CTA record building.
__device__ __forceinline__ void mbar_init(uint32_t mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_bulk_load(uint32_t dst, const void* src, unsigned bytes, uint32_t mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, unsigned phase) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(phase) : "memory");
    } while (!ok);
}

// Out of line on purpose, like rare_group: keeps the render kernel's register allocation
// around the hot loop exactly as it is without the record table.  The barrier's phase lives in
// shared memory (flipped by thread 0 after the CTA-wide barrier that follows every staging).
__device__ __noinline__ void stage_records_tma(float4* smem_tab, const float* src, int cnt, unsigned long long* bar,
                                               const unsigned* phase_s, int* chunk_class, int tid) {
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(bar);
    const unsigned phase = *phase_s;
    if (tid == 0) tma_bulk_load((uint32_t)__cvta_generic_to_shared(smem_tab), src, (unsigned)cnt * 64u, mbar);
    mbar_wait(mbar, phase);
    // rrt_build_records left the chunk's class bits in the spare slot of its first record
    if (tid == 0 && chunk_class) *chunk_class |= __float_as_int(smem_tab[3].w);
}

template <bool GENERAL>
__device__ __forceinline__ float object_max_det(uint32_t rec, const RayPack& rp, float gmax) {
    float4 q0 = lds128(rec), q1 = lds128(rec + 16);
    float4 q2 = q0, q3 = q0;
    if (GENERAL) { q2 = lds128(rec + 32); q3 = lds128(rec + 48); }
#pragma unroll
    for (int p = 0; p < kRays / 2; p++) {
        float lo, hi;
        upk(pair_det<GENERAL>(q0, q1, q2, q3, rp.dx[p], rp.dy[p], rp.dz[p]), lo, hi);
        gmax = fmaxf(gmax, fmaxf(lo, hi));
    }
    return gmax;
}

// Sweep `count` staged SPHERES over the thread's 8 rays.  The hot loop is branch-free
// over groups of kGroup objects: packed FFMA2 discriminants, a running max, ONE
// compare-and-branch per group; a group with any det > 0 (rare: ~1e-3 per object and
// warp) is re-evaluated by the scalar canonical routine.  GENERAL=false is the
// diagonal fast path (translate*scale objects): exact-zero off-diagonals make it
// bit-identical to the general form.
template <bool GENERAL>
__device__ __forceinline__ void sweep_spheres(const float4* __restrict__ tab, int count, int kbase, const RayPack& rp,
                                              const float* dw, float* tmin, int* idx) {
    // one induction variable (the shared-window address) and a warp-uniform branch keep the
    // loop control at compare+branch; the object index is only reconstructed on the rare path
    constexpr int kUnroll = RRT_SWEEP_UNROLL;   // groups per loop trip (loop control amortised over kUnroll*kGroup objects)
    uint32_t rec0 = (uint32_t)__cvta_generic_to_shared(tab);
    uint32_t rec_end = rec0 + 64u * (uint32_t)(count - count % (kGroup * kUnroll));
    // launder both through an opaque move: otherwise ptxas rematerialises the shared-window
    // arithmetic (S2UR/ULEA) inside the loop instead of keeping two registers live
    asm volatile("mov.u32 %0, %0;" : "+r"(rec0));
    asm volatile("mov.u32 %0, %0;" : "+r"(rec_end));
    uint32_t rec = rec0;
#pragma unroll 1
    for (; rec != rec_end; rec += 64 * kGroup * kUnroll) {
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            float gmax = 0.0f;
#pragma unroll
            for (int j = 0; j < kGroup; j++) gmax = object_max_det<GENERAL>(rec + 64 * (u * kGroup + j), rp, gmax);
            if (__builtin_expect(__any_sync(0xffffffffu, gmax > 0.0f), 0))
                rare_group(tab, (int)((rec - rec0) >> 6) + u * kGroup, kGroup, kbase, dw, tmin, idx);
        }
    }
    const int k = count - count % (kGroup * kUnroll);
    if (k < count) rare_group(tab, k, count - k, kbase, dw, tmin, idx);
}

// Chunks that contain squares: spheres get the packed pre-test one by one, squares
// always take the scalar routine.
__device__ __forceinline__ void sweep_mixed(const float4* __restrict__ tab, int count, int kbase, const RayPack& rp,
                                            const float* dw, float* tmin, int* idx) {
#pragma unroll 1
    for (int k = 0; k < count; k++) {
        const int flags = __float_as_int(tab[4 * k + 1].w);
        float gmax = 1.0f;
        if (!(flags & 1)) gmax = object_max_det<true>((uint32_t)__cvta_generic_to_shared(tab + 4 * k), rp, 0.0f);
        if (gmax > 0.0f) rare_group(tab, k, 1, kbase, dw, tmin, idx);
    }
}

// ---------------------------------------------------------------- hard shadows (RRT_FLAG_SHADOWS)
// Sphere.shadow, shape.py:85-97, at the (commented-out) call site scene.py:41-45, in the
// caster's object space and canonical float32 order -- bit-identical to orc_shadowed in
// oracle/oracle_c.c.  `t` is the winner's ray parameter.
__device__ __forceinline__ bool shadow_test(const float4* __restrict__ rec, float wx, float wy, float wz, float t,
                                            const float* U) {
    Obj ob;
    load_rec(rec, ob);
    if (ob.flags & 1) return false;            // Square has no shadow method: casts none
    const float d0 = dot3_canon(ob.a[0], ob.a[1], ob.a[2], wx, wy, wz);
    const float d1 = dot3_canon(ob.a[3], ob.a[4], ob.a[5], wx, wy, wz);
    const float d2 = dot3_canon(ob.a[6], ob.a[7], ob.a[8], wx, wy, wz);
    const float y0 = __fmaf_rn(t, d0, ob.o[0]), y1 = __fmaf_rn(t, d1, ob.o[1]), y2 = __fmaf_rn(t, d2, ob.o[2]);
    const float x = dot3_canon(y0, y1, y2, U[0], U[1], U[2]);
    const float yy = dot3_canon(y0, y1, y2, y0, y1, y2);
    const float dec = __fadd_rn(__fmaf_rn(x, x, -yy), 1.0f);
    return dec > 0.0f && __fsub_rn(-x, __fsqrt_rn(dec)) >= 0.0f;
}

// General kernel: tests the thread's winning rays against one staged chunk of objects.
// Scalar and divergent on purpose -- shadows are an opt-in extension outside the
// roofline-accountable sweep; out of line so that the hot loop's registers are untouched.
__device__ __noinline__ unsigned shadow_chunk(const float4* __restrict__ tab, int cnt, int kbase, const float* dw,
                                              const float* tmin, const int* idx, const float* U, unsigned shadowed) {
#pragma unroll 1
    for (int r = 0; r < kRays; r++) {
        const int win = idx[r];
        if (win < 0 || (shadowed >> r & 1u)) continue;
        const float t = tmin[r], wx = dw[r], wy = dw[kRays + r], wz = dw[2 * kRays + r];
#pragma unroll 1
        for (int k = 0; k < cnt; k++) {
            if (kbase + k == win) continue;
            if (shadow_test(tab + 4 * k, wx, wy, wz, t, U)) { shadowed |= 1u << r; break; }
        }
    }
    return shadowed;
}

// ---------------------------------------------------------------- shading (float32)
// x ** y like C pow() (Theano's T.pow, shader.py:45): integer-valued exponents up to
// 1024 (shininess = 50 in every reference script) take square-and-multiply -- a
// negative base is fine there, as in pow(); everything else goes to powf.
__device__ __noinline__ float powf_general(float x, float y) { return powf(x, y); }

__device__ __forceinline__ float pow_shininess(float x, float y) {
    const int e = (int)y;
    if ((float)e == y && e >= 0 && e <= 1024) {
        float r = 1.0f, b = x;
        int k = e;
#pragma unroll 1
        while (k) {
            if (k & 1) r *= b;
            b *= b;
            k >>= 1;
        }
        return r;
    }
    return powf_general(x, y);
}


struct ShadeRec {
    float t, d[3], o[3], pn, nrm[3], ndl, rm[3], rv, pw, ph;
    bool inside[3];
};

__device__ __forceinline__ void shade(int shader, float max_depth, const Obj& ob, const float* mat, const Globals& g,
                                      const HitRec& h, ShadeRec& r, float rgb[3]) {
    r.t = h.t;
#pragma unroll
    for (int c = 0; c < 3; c++) { r.d[c] = h.d[c]; r.o[c] = ob.o[c]; }
    if (shader == RRT_SHADER_DEPTH) {  // shader.py:14-20
        float v = 1.0f - r.t / max_depth;
        rgb[0] = rgb[1] = rgb[2] = v;
        return;
    }
    if (!(ob.flags & 1)) {  // Sphere.normals shape.py:134-137 (object-space normal)
        float p0 = fmaf(r.t, r.d[0], r.o[0]), p1 = fmaf(r.t, r.d[1], r.o[1]), p2 = fmaf(r.t, r.d[2], r.o[2]);
        const float pn2 = p0 * p0 + p1 * p1 + p2 * p2;
        const float inv = rsqrtf(pn2);
        r.pn = pn2 * inv;
        r.nrm[0] = p0 * inv; r.nrm[1] = p1 * inv; r.nrm[2] = p2 * inv;
    } else {                // Square.normals shape.py:55-68
        r.nrm[0] = r.nrm[1] = 0.f;
        r.nrm[2] = (ob.o[2] > 0.0f) ? 1.0f : -1.0f;
        r.pn = 1.0f;
    }
    r.ndl = -(r.nrm[0] * g.Lh[0] + r.nrm[1] * g.Lh[1] + r.nrm[2] * g.Lh[2]);  // shader.py:40
    r.ph = mat[0] + mat[1] * r.ndl;
    r.rv = 0.f; r.pw = 0.f;
    if (shader == RRT_SHADER_PHONG) {  // shader.py:43-45
#pragma unroll
        for (int c = 0; c < 3; c++) r.rm[c] = 2.0f * r.ndl * r.nrm[c] + g.Lh[c];
        r.rv = r.rm[0] * g.look[0] + r.rm[1] * g.look[1] + r.rm[2] * g.look[2];
        r.pw = pow_shininess(r.rv, mat[3]);
        r.ph += mat[2] * r.pw;
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {      // shader.py:50-51
        float v = r.ph * mat[4 + c] * g.I[c];
        r.inside[c] = (v >= 0.0f && v <= 1.0f);
        rgb[c] = fminf(fmaxf(v, 0.0f), 1.0f);
    }
}

// ---------------------------------------------------------------- reverse pass, one winning ray
// Closed form of T.grad through the winner (masks constant).  og[19] receives
// [M = sum g_d' r_cam^T (9), g_b = sum g_o' (3), d/d(ka,kd,ks,sh,r,g,b)];
// gg[9] receives [d/d Lhat (3), d/d intensity (3), d/d look_at (3)].
// The chain M -> d/dA, d/d camera and Lhat -> L is applied by finalize_grads.
__device__ __forceinline__ void backward_ray(int shader, float max_depth, const Obj& ob, const float* mat,
                                             const Globals& g, const HitRec& h, const ShadeRec& r, const float rc[3],
                                             const float gc[3], float og[19], float gg[9]) {
    float g_t = 0.f;
    float g_o[3] = {0.f, 0.f, 0.f}, g_d[3] = {0.f, 0.f, 0.f};
    const bool sphere = !(ob.flags & 1);
    if (shader == RRT_SHADER_DEPTH) {
        g_t = -(gc[0] + gc[1] + gc[2]) / max_depth;
    } else {
        float g_ph = 0.f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (!r.inside[c]) continue;
            g_ph += gc[c] * mat[4 + c] * g.I[c];
            og[16 + c] += gc[c] * r.ph * g.I[c];
            gg[3 + c] += gc[c] * r.ph * mat[4 + c];
        }
        og[12] += g_ph;
        og[13] += g_ph * r.ndl;
        float g_ndl = g_ph * mat[1];
        float g_n[3] = {0.f, 0.f, 0.f}, g_Lh[3] = {0.f, 0.f, 0.f};
        if (shader == RRT_SHADER_PHONG) {
            og[14] += g_ph * r.pw;
            if (r.rv > 0.0f) og[15] += g_ph * mat[2] * r.pw * __logf(r.rv);
            float dpw = (r.rv != 0.0f) ? __fdividef(r.pw, r.rv) : pow_shininess(r.rv, mat[3] - 1.0f);  // rv^(sh-1)
            float g_rv = g_ph * mat[2] * mat[3] * dpw;
            float g_rm[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                g_rm[c] = g_rv * g.look[c];
                gg[6 + c] += g_rv * r.rm[c];
            }
            g_ndl += 2.0f * (g_rm[0] * r.nrm[0] + g_rm[1] * r.nrm[1] + g_rm[2] * r.nrm[2]);
#pragma unroll
            for (int c = 0; c < 3; c++) { g_n[c] += 2.0f * r.ndl * g_rm[c]; g_Lh[c] += g_rm[c]; }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) { g_n[c] -= g_ndl * g.Lh[c]; g_Lh[c] -= g_ndl * r.nrm[c]; }
#pragma unroll
        for (int c = 0; c < 3; c++) gg[c] += g_Lh[c];
        if (sphere) {
            float ndg = r.nrm[0] * g_n[0] + r.nrm[1] * g_n[1] + r.nrm[2] * g_n[2];
            float inv = __frcp_rn(r.pn);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float gp = (g_n[c] - r.nrm[c] * ndg) * inv;
                g_o[c] = gp;
                g_t += gp * r.d[c];
                g_d[c] = r.t * gp;
            }
        }
    }
    if (sphere) {
        float ivn = __frcp_rn(h.vn);
        float g_pd = -g_t * ivn, g_s = -g_t * ivn, g_vn = -g_t * r.t * ivn;
        float g_det = g_s * 0.5f * rsqrtf(h.det);
        g_pd += 2.0f * h.pd * g_det;
        g_vn += ob.ncc * g_det;            // -cc * g_det
        float g_cc = -h.vn * g_det;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            g_o[c] += 2.0f * r.o[c] * g_cc + r.d[c] * g_pd;
            g_d[c] += r.o[c] * g_pd + 2.0f * r.d[c] * g_vn;
        }
    } else {  // t = -o'_z / d'_z
        g_o[2] += -g_t / r.d[2];
        g_d[2] += -g_t * r.t / r.d[2];
    }
#pragma unroll
    for (int rr = 0; rr < 3; rr++) {
#pragma unroll
        for (int c = 0; c < 3; c++) og[rr * 3 + c] += g_d[rr] * rc[c];
        og[9 + rr] += g_o[rr];
    }
}

// ---------------------------------------------------------------- gradient reduction
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// find-or-insert a CTA slot for object `key` (called by one lane); -1 = table full
__device__ __forceinline__ int slot_for(int* slot_key, int key) {
    int h = key & (kSlots - 1);
#pragma unroll 1
    for (int probe = 0; probe < kSlots; probe++) {
        int old = atomicCAS(&slot_key[h], -1, key);
        if (old == -1 || old == key) return h;
        h = (h + 1) & (kSlots - 1);
    }
    return -1;
}

// Sum of 19 per-lane values over the warp as a TRANSPOSED butterfly: at every stage a lane
// keeps one half of its values and trades the other half with its partner, so the whole
// reduction is 10+5+3+2+1 = 21 shuffles (instead of 19 x 5) and ends with lane l holding the
// complete sum of value index `v` (returned; -1 for the lanes that hold padding).
__device__ __forceinline__ float xchg_add(float keep, float send, int offset) {
    return keep + __shfl_xor_sync(0xffffffffu, send, offset);
}
__device__ __forceinline__ float warp_reduce19(const float (&a)[19], bool mine, int lane, int& v) {
    float b[10], c[6], d[4], e[2];
    bool up = lane & 16;
#pragma unroll
    for (int j = 0; j < 10; j++) {
        const float lo = mine ? a[j] : 0.f;
        const float hi = (j < 9 && mine) ? a[j < 9 ? 10 + j : 18] : 0.f;   // value 19 is padding
        b[j] = xchg_add(up ? hi : lo, up ? lo : hi, 16);
    }
    up = lane & 8;
#pragma unroll
    for (int j = 0; j < 5; j++) c[j] = xchg_add(up ? b[5 + j] : b[j], up ? b[j] : b[5 + j], 8);
    c[5] = 0.f;
    up = lane & 4;
#pragma unroll
    for (int j = 0; j < 3; j++) d[j] = xchg_add(up ? c[3 + j] : c[j], up ? c[j] : c[3 + j], 4);
    d[3] = 0.f;
    up = lane & 2;
#pragma unroll
    for (int j = 0; j < 2; j++) e[j] = xchg_add(up ? d[2 + j] : d[j], up ? d[j] : d[2 + j], 2);
    up = lane & 1;
    const float total = xchg_add(up ? e[1] : e[0], up ? e[0] : e[1], 1);
    const int j3 = ((lane >> 1) & 1) * 2 + (lane & 1);
    const int ci = ((lane >> 2) & 1) * 3 + j3;
    const int idx = ((lane >> 4) & 1) * 10 + ((lane >> 3) & 1) * 5 + ci;
    v = (j3 < 3 && ci < 5 && idx < 19) ? idx : -1;
    return total;
}

// All 32 lanes call this together.  Each lane holds (key, acc[19]); lanes with the
// same key are summed (transposed butterfly) and 19 lanes add one sum each to the CTA slot.
__device__ __forceinline__ void warp_flush(int key, float (&acc)[19], int* slot_key, float* slots, float* gobj, int lane) {
    unsigned active = __ballot_sync(0xffffffffu, key >= 0);
    while (active) {
        int leader = __ffs(active) - 1;
        int k = __shfl_sync(0xffffffffu, key, leader);
        bool mine = (key == k);
        active &= ~__ballot_sync(0xffffffffu, mine);
        int slot = 0;
        if (lane == 0) slot = slot_for(slot_key, k);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        int v;
        const float x = warp_reduce19(acc, mine, lane, v);
        if (v >= 0 && x != 0.f) {
            if (slot >= 0) atomicAdd(&slots[slot * kSlotStride + v], x);
            else atomicAdd(&gobj[(size_t)k * RRT_OBJ_GRAD_STRIDE + v], x);
        }
    }
#pragma unroll
    for (int v = 0; v < 19; v++) acc[v] = 0.f;
}

// ---------------------------------------------------------------- conservative tile culling
// RRT_FLAG_CULL.  The CTA's rays (world directions, all through the camera origin) are
// bounded by a circular cone (axis u, half-angle theta).  In an object's space every ray
// direction lies within theta' of u' = A.u with sin(theta') <= |A|_F tan(theta) / |u'|.  The
// LINE through o' with such a direction passes the object's origin no closer than
// |o'| sin(phi - theta'), phi = angle(line u', -o') -- lines, not rays, because spheres have
// no t > 0 test (shape.py:109-126).  The object is skipped only if that distance exceeds its
// bounding radius (1 for the unit sphere, sqrt(1/2) for the unit square) inflated by far more
// than the float32 error of the canonical discriminant (delta(det/vn) <= ~6e-7 |o'|^2) and of
// this test itself.  Every comparison is written so that NaN keeps the object.
struct TileCone {
    float u[3];
    float tan_theta;
    int ok;
};

__device__ __forceinline__ bool cull_keep(const float4* __restrict__ rec, const TileCone& tc) {
    if (!tc.ok) return true;
    Obj ob;
    load_rec(rec, ob);
    const float ux = ob.a[0] * tc.u[0] + ob.a[1] * tc.u[1] + ob.a[2] * tc.u[2];
    const float uy = ob.a[3] * tc.u[0] + ob.a[4] * tc.u[1] + ob.a[5] * tc.u[2];
    const float uz = ob.a[6] * tc.u[0] + ob.a[7] * tc.u[1] + ob.a[8] * tc.u[2];
    const float lu = sqrtf(ux * ux + uy * uy + uz * uz);
    if (!(lu > 0.f)) return true;
    const float s = ob.afro * tc.tan_theta / lu * 1.001f;
    if (!(s < 0.99f)) return true;
    const float theta_o = asinf(s) + 1e-4f;
    const float lo2 = ob.o[0] * ob.o[0] + ob.o[1] * ob.o[1] + ob.o[2] * ob.o[2];
    const float r2 = (ob.flags & 1) ? 0.5f : 1.0f;
    const float rinfl = sqrtf(r2 + 1e-5f * (1.0f + lo2)) * 1.001f;
    const float lo = sqrtf(lo2);
    if (!(lo > rinfl)) return true;
    float cphi = fabsf(ob.o[0] * ux + ob.o[1] * uy + ob.o[2] * uz) / (lo * lu);
    cphi = fminf(cphi, 1.0f);
    const float phi = acosf(cphi);
    const float need = asinf(rinfl / lo) + 1e-4f;
    return !(phi - theta_o > need);   // keep unless provably out of reach
}

// ---------------------------------------------------------------- the render kernel
// grid = (ceil(n / (32*PIX)), ceil(rows / warps), B); block = 32 * warps.
// Thread (warp w, lane l) owns pixels (row = tile_row0 + w, cols = col0 + l*PIX .. +PIX-1),
// each with SPT samples: PIX*SPT = 8 rays.  S > SPT (generic path, PIX = 1) loops
// over chunks of SPT samples.
template <int PIX, int SPT, int MODE>
__global__ void __launch_bounds__(32 * RRT_MAX_WARPS, RRT_MIN_BLOCKS) render_kernel(const __grid_constant__ KParams P) {
    extern __shared__ float4 smem_tab[];  // kObjChunk (or N) sweep records
    __shared__ Globals g;
    __shared__ int slot_key[kSlots];
    __shared__ float slots[kSlots * kSlotStride];
    __shared__ float gglob[9];
    __shared__ float loss_warp[kMaxWarps];
    __shared__ __align__(16) float stage[kMaxWarps][32 * PIX * 3];   // per-warp tile-row staging (vector I/O)
    __shared__ int cam_identity_s;
    __shared__ float cone_red[kMaxWarps][6];
    __shared__ TileCone tcone;
    __shared__ unsigned keepmask[(kObjChunk + 31) / 32];
    __shared__ int chunk_class;  // sticky per CTA: bit0 squares, bit1 general spheres seen
    __shared__ __align__(8) unsigned long long tma_bar;   // mbarrier of the record-table bulk copies
    __shared__ unsigned tma_phase;

    const rrt_scene& sc = P.sc;
    // S is a compile-time constant except in the generic (PIX=1, SPT=8) instantiation, so the
    // sample-chunk loop below has exactly one trip and nothing reverse-pass related is live
    // across the sweep.
    const int n = sc.n, N = sc.num_objects;
    const int S = (PIX == 1) ? sc.samples : SPT;
    const int scene = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int al = blockIdx.y * nwarps + warp;      // slab-local row
    const int a = sc.row_begin + al;                // image row
    const int b0 = (blockIdx.x * 32 + lane) * PIX;  // first column of this thread
    const bool row_ok = al < P.rows;

    // ---- per-scene constants
    if (tid < 32) {
        const float* cam = sc.camera + (size_t)scene * sc.camera_scene_stride;
        const float* li = sc.light + (size_t)scene * sc.light_scene_stride;
        if (tid < 3) {
            g.C[tid * 3 + 0] = cam[tid * 4 + 0];
            g.C[tid * 3 + 1] = cam[tid * 4 + 1];
            g.C[tid * 3 + 2] = cam[tid * 4 + 2];
            g.ct[tid] = cam[tid * 4 + 3];
            g.look[tid] = cam[12 + tid];
            g.L[tid] = li[tid];
            g.I[tid] = li[3 + tid];
        }
        __syncwarp();
        if (tid == 0) {
            float ln = sqrtf(g.L[0] * g.L[0] + g.L[1] * g.L[1] + g.L[2] * g.L[2]);  // scene.py:83-86
            g.Ln = ln;
            g.Lh[0] = g.L[0] / ln; g.Lh[1] = g.L[1] / ln; g.Lh[2] = g.L[2] / ln;
            canon_to_light(g.L, g.U);
            chunk_class = 0;
            cam_identity_s = (g.C[0] == 1.f && g.C[4] == 1.f && g.C[8] == 1.f && g.C[1] == 0.f && g.C[2] == 0.f &&
                              g.C[3] == 0.f && g.C[5] == 0.f && g.C[6] == 0.f && g.C[7] == 0.f);
        }
        if (tid < kSlots) slot_key[tid] = -1;
        if (tid < 9) gglob[tid] = 0.f;
        if (tid == 0) {
            tma_phase = 0;
            if (sc.obj_records) mbar_init((uint32_t)__cvta_generic_to_shared(&tma_bar), 1);
        }
    }
    for (int q = tid; q < kSlots * kSlotStride; q += blockDim.x) slots[q] = 0.f;
    __syncthreads();

    const bool cam_identity = cam_identity_s != 0;
    const float* w2o = sc.w2o + (size_t)scene * sc.w2o_scene_stride;
    const float* mats = sc.material + (size_t)scene * sc.material_scene_stride;
    float* gobj = (MODE != MODE_FWD) ? P.grad + (size_t)scene * RRT_GRAD_SIZE(N) : nullptr;

    float pixsum[PIX][3];
#pragma unroll
    for (int px = 0; px < PIX; px++) pixsum[px][0] = pixsum[px][1] = pixsum[px][2] = 0.f;

    // upstream gradient per pixel (MODE_BWD known up front; MODE_FUSED after shading)
    float gpix[PIX][3];
#pragma unroll
    for (int px = 0; px < PIX; px++) {
        gpix[px][0] = gpix[px][1] = gpix[px][2] = 0.f;
        if (MODE == MODE_BWD) {
            int b = b0 + px;
            if (row_ok && b < n) {
                size_t po = (((size_t)scene * P.rows + al) * n + b) * 3;
                float inv = 1.0f / (float)S;
                gpix[px][0] = P.dl_dimage[po] * inv; gpix[px][1] = P.dl_dimage[po + 1] * inv; gpix[px][2] = P.dl_dimage[po + 2] * inv;
            }
        }
    }

    if (MODE == MODE_BWD) {
        // Sparse upstream gradients (optimize_brightness.py:51 touches two pixels): a CTA none
        // of whose pixels carries gradient contributes exactly zero -- leave before building rays.
        bool nz = false;
#pragma unroll
        for (int px = 0; px < PIX; px++) nz |= (gpix[px][0] != 0.f) | (gpix[px][1] != 0.f) | (gpix[px][2] != 0.f);
        if (!__syncthreads_or(nz)) return;
    }

    // base rays (float64 grid -> float32), one per owned pixel
    float bx[PIX], by[PIX], bz[PIX];
#pragma unroll
    for (int px = 0; px < PIX; px++) {
        int b = b0 + px;
        int i = sc.transpose ? b : a, j = sc.transpose ? a : b;
        if (row_ok && b < n) {
            if (sc.base_rays) {       // precomputed grid (rrt_primary_rays): same bits, no float64 chain
                const float* br = sc.base_rays + ((size_t)i * n + j) * 3;
                bx[px] = __ldg(br); by[px] = __ldg(br + 1); bz[px] = __ldg(br + 2);
            } else {
                base_ray(n, P.lin_step, i, j, bx[px], by[px], bz[px]);
            }
        }
        else { bx[px] = by[px] = bz[px] = 0.f; }
    }

    float acc[19];
#pragma unroll
    for (int v = 0; v < 19; v++) acc[v] = 0.f;
    int acc_key = -1;
    float gg[9];
#pragma unroll
    for (int v = 0; v < 9; v++) gg[v] = 0.f;
    float loss_part = 0.f;

    const int nchunks_s = (S + SPT - 1) / SPT;
#pragma unroll 1
    for (int sc0 = 0; sc0 < nchunks_s; sc0++) {
        // ---- build the 8 rays of this sample chunk.  Per-ray state lives in (L1-resident)
        // local memory: it is read by the rare path of the sweep and by the rolled shading /
        // reverse-pass loops below; only the packed world directions stay in registers.
        __align__(16) float l_rc[3 * kRays], l_dw[3 * kRays], l_tmin[kRays];  // SoA: [x0..x7 | y0..y7 | z0..z7]
        int l_idx[kRays];
        RayPack rp;
        // rolled on purpose (code size: this runs once per thread; instruction-cache misses
        // dominate small-scene workloads otherwise)
#pragma unroll 1
        for (int r = 0; r < kRays; r++) {
            const int px = r / SPT, sl = r % SPT;
            const int s = sc0 * SPT + sl;
            const int b = b0 + px;
            const bool ok = row_ok && b < n && s < S;
            float rcx = 0.f, rcy = 0.f, rcz = 0.f;
            float wx = 0.f, wy = 0.f, wz = 0.f;   // zero direction never hits (det == 0)
            if (ok) {
                float jx, jy;
                if (sc.jitter_x) {
                    size_t off = (size_t)scene * sc.jitter_scene_stride + ((size_t)al * n + b) * S + s;
                    jx = __ldg(sc.jitter_x + off);
                    jy = __ldg(sc.jitter_y + off);
                } else {
                    jx = rrt_rng(sc.seed, scene + sc.scene_begin, (uint32_t)(a * n + b), s, 0);
                    jy = rrt_rng(sc.seed, scene + sc.scene_begin, (uint32_t)(a * n + b), s, 1);
                }
                const float ox = P.pow2 ? jitter_offset_pow2(jx, s, P.inv_s, P.inv_n) : jitter_offset(jx, s, S, n);
                const float oy = P.pow2 ? jitter_offset_pow2(jy, s, P.inv_s, P.inv_n) : jitter_offset(jy, s, S, n);
                float bxv = 0.f, byv = 0.f, bzv = 0.f;
#pragma unroll
                for (int q = 0; q < PIX; q++)
                    if (q == px) { bxv = bx[q]; byv = by[q]; bzv = bz[q]; }
                rcx = __fadd_rn(bxv, ox);
                rcy = __fadd_rn(byv, oy);
                rcz = bzv;
                if (cam_identity) {     // root variant: C = I, the fma chain returns its input
                    wx = rcx; wy = rcy; wz = rcz;
                } else {                // camera.o2w, orbit_experiments/scene.py:80
                    wx = dot3_canon(g.C[0], g.C[1], g.C[2], rcx, rcy, rcz);
                    wy = dot3_canon(g.C[3], g.C[4], g.C[5], rcx, rcy, rcz);
                    wz = dot3_canon(g.C[6], g.C[7], g.C[8], rcx, rcy, rcz);
                }
            }
            l_rc[r] = rcx; l_rc[kRays + r] = rcy; l_rc[2 * kRays + r] = rcz;
            l_dw[r] = wx; l_dw[kRays + r] = wy; l_dw[2 * kRays + r] = wz;
            l_tmin[r] = __int_as_float(0x7f800000);
            l_idx[r] = -1;
        }
        {
            const u64* dp = reinterpret_cast<const u64*>(l_dw);   // (x0,x1) (x2,x3) ... pairs
#pragma unroll
            for (int p = 0; p < kRays / 2; p++) { rp.dx[p] = dp[p]; rp.dy[p] = dp[kRays / 2 + p]; rp.dz[p] = dp[kRays + p]; }
        }

        const bool use_stored = (MODE == MODE_BWD) && (P.hit_in != nullptr);
        const bool cull = (sc.flags & RRT_FLAG_CULL) && !use_stored;
        if (cull) {   // ---- bounding cone of this CTA's rays (exact min/max of the rays built above)
            const float big = 3.0e38f;
            float lo3[3] = {big, big, big}, hi3[3] = {-big, -big, -big};
#pragma unroll 1
            for (int r = 0; r < kRays; r++) {
                const float x = l_dw[r], y = l_dw[kRays + r], z = l_dw[2 * kRays + r];
                if (x != 0.f || y != 0.f || z != 0.f) {
                    lo3[0] = fminf(lo3[0], x); hi3[0] = fmaxf(hi3[0], x);
                    lo3[1] = fminf(lo3[1], y); hi3[1] = fmaxf(hi3[1], y);
                    lo3[2] = fminf(lo3[2], z); hi3[2] = fmaxf(hi3[2], z);
                }
            }
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    lo3[c] = fminf(lo3[c], __shfl_xor_sync(0xffffffffu, lo3[c], o));
                    hi3[c] = fmaxf(hi3[c], __shfl_xor_sync(0xffffffffu, hi3[c], o));
                }
            if (sc0 > 0) __syncthreads();          // previous use of cone_red / tcone is over
            if (lane == 0) {
#pragma unroll
                for (int c = 0; c < 3; c++) { cone_red[warp][c] = lo3[c]; cone_red[warp][3 + c] = hi3[c]; }
            }
            __syncthreads();
            if (tid == 0) {
                float l[3] = {big, big, big}, h[3] = {-big, -big, -big};
                for (int w = 0; w < nwarps; w++)
                    for (int c = 0; c < 3; c++) { l[c] = fminf(l[c], cone_red[w][c]); h[c] = fmaxf(h[c], cone_red[w][3 + c]); }
                TileCone tc;
                tc.ok = 0; tc.tan_theta = 0.f; tc.u[0] = tc.u[1] = tc.u[2] = 0.f;
                if (l[0] <= h[0]) {
                    const float cx = 0.5f * (l[0] + h[0]), cy = 0.5f * (l[1] + h[1]), cz = 0.5f * (l[2] + h[2]);
                    const float cn = sqrtf(cx * cx + cy * cy + cz * cz);
                    if (cn > 1e-20f) {
                        tc.u[0] = cx / cn; tc.u[1] = cy / cn; tc.u[2] = cz / cn;
                        float cosmin = 1.0f;
                        bool good = true;
                        for (int q = 0; q < 8; q++) {
                            const float vx = (q & 1) ? h[0] : l[0], vy = (q & 2) ? h[1] : l[1], vz = (q & 4) ? h[2] : l[2];
                            const float vn = sqrtf(vx * vx + vy * vy + vz * vz);
                            if (!(vn > 1e-20f)) { good = false; break; }
                            cosmin = fminf(cosmin, (vx * tc.u[0] + vy * tc.u[1] + vz * tc.u[2]) / vn);
                        }
                        if (good && cosmin > 0.2f) {
                            const float theta = acosf(fminf(cosmin, 1.0f)) * 1.001f + 1e-4f;
                            tc.tan_theta = tanf(theta);
                            tc.ok = 1;
                        }
                    }
                }
                tcone = tc;
            }
            __syncthreads();
        }

        // ---- nearest-hit sweep (or stored winners)
        if (use_stored) {
#pragma unroll 1
            for (int r = 0; r < kRays; r++) {
                const int px = r / SPT, s = sc0 * SPT + r % SPT, b = b0 + px;
                if (row_ok && b < n && s < S)
{
                    const int kk = P.hit_in[(((size_t)scene * S + s) * P.rows + al) * n + b];
                    // never trust an index buffer blindly; a winner stored with RRT_HIT_SHADOWED
                    // (>= N) shades to zero and carries no gradient
                    l_idx[r] = (kk >= 0 && kk < N) ? kk : -1;
                }
            }
        } else {
#pragma unroll 1
            for (int kb = 0; kb < N; kb += kObjChunk) {
                const int cnt = min(kObjChunk, N - kb);
                if (kb > 0 || sc0 > 0) __syncthreads();   // previous chunk fully consumed
                bool staged_by_tma = false;
                if (N > kObjChunk || sc0 == 0) {
                    if (sc.obj_records) {                  // precomputed records: one TMA bulk copy
                        stage_records_tma(smem_tab, sc.obj_records + ((size_t)scene * N + kb) * RRT_RECORD_FLOATS, cnt,
                                          &tma_bar, &tma_phase, &chunk_class, tid);
                        staged_by_tma = true;
                    } else {
                        int cls = 0;                       // bit0: squares present, bit1: general spheres present
                        for (int k = tid; k < cnt; k += blockDim.x) {
                            Obj ob;
                            make_obj(w2o + (size_t)(kb + k) * RRT_W2O_STRIDE, sc.obj_type[kb + k], g.ct, ob, cull);
                            store_rec(smem_tab + 4 * k, ob);
                            cls |= ob.flags;
                        }
                        if (cls) atomicOr(&chunk_class, cls);
                    }
                }
                __syncthreads();
                if (staged_by_tma && tid == 0) tma_phase ^= 1u;   // read again only after the next CTA-wide barrier
                const int cls = chunk_class;
                if (cull) {
                    // one ballot word per 32 objects keeps list order without a compaction pass
                    for (int k0 = warp * 32; k0 < cnt; k0 += 32 * nwarps) {
                        const int k = k0 + lane;
                        const bool keep = (k < cnt) && cull_keep(smem_tab + 4 * k, tcone);
                        const unsigned m = __ballot_sync(0xffffffffu, keep);
                        if (lane == 0) keepmask[k0 >> 5] = m;
                    }
                    __syncthreads();
#pragma unroll 1
                    for (int w = 0; w < (cnt + 31) / 32; w++) {
                        unsigned m = keepmask[w];
#pragma unroll 1
                        while (m) {
                            const int bit = __ffs(m) - 1;
                            m &= m - 1;
                            rare_group(smem_tab, w * 32 + bit, 1, kb, l_dw, l_tmin, l_idx);
                        }
                    }
                } else if (cls == 0) sweep_spheres<false>(smem_tab, cnt, kb, rp, l_dw, l_tmin, l_idx);
                else if (!(cls & 1)) sweep_spheres<true>(smem_tab, cnt, kb, rp, l_dw, l_tmin, l_idx);
                else sweep_mixed(smem_tab, cnt, kb, rp, l_dw, l_tmin, l_idx);
            }
        }

        // ---- hard shadows (opt-in): second pass over the object table for the winners
        unsigned shadowed = 0;
        if ((sc.flags & RRT_FLAG_SHADOWS) && !use_stored) {
#pragma unroll 1
            for (int kb = 0; kb < N; kb += kObjChunk) {
                const int cnt = min(kObjChunk, N - kb);
                if (N > kObjChunk) {                       // otherwise the whole table is still staged
                    __syncthreads();
                    if (sc.obj_records) {
                        stage_records_tma(smem_tab, sc.obj_records + ((size_t)scene * N + kb) * RRT_RECORD_FLOATS, cnt,
                                          &tma_bar, &tma_phase, nullptr, tid);
                    } else {
                        for (int k = tid; k < cnt; k += blockDim.x) {
                            Obj ob;
                            make_obj(w2o + (size_t)(kb + k) * RRT_W2O_STRIDE, sc.obj_type[kb + k], g.ct, ob);
                            store_rec(smem_tab + 4 * k, ob);
                        }
                    }
                    __syncthreads();
                    if (sc.obj_records && tid == 0) tma_phase ^= 1u;
                }
                shadowed = shadow_chunk(smem_tab, cnt, kb, l_dw, l_tmin, l_idx, g.U, shadowed);
            }
        }

        // ---- outputs of the sweep
        if (MODE != MODE_BWD && (P.hit_out || P.tmin_out)) {
#pragma unroll 1
            for (int r = 0; r < kRays; r++) {
                const int px = r / SPT, s = sc0 * SPT + r % SPT, b = b0 + px;
                if (row_ok && b < n && s < S) {
                    size_t ro = (((size_t)scene * S + s) * P.rows + al) * n + b;
                    if (P.hit_out) __stcs(P.hit_out + ro, l_idx[r] | ((shadowed >> r & 1u) ? RRT_HIT_SHADOWED : 0));
                    if (MODE == MODE_FWD && P.tmin_out) __stcs(P.tmin_out + ro, l_tmin[r]);
                }
            }
        }
        if (shadowed) {                                    // (0,0,0) and no gradient from here on
#pragma unroll 1
            for (int r = 0; r < kRays; r++)
                if (shadowed >> r & 1u) l_idx[r] = -1;
        }

        // ---- shade the winners (forward value)
        if (MODE != MODE_BWD) {
            // a thread's rays (samples of one pixel, neighbouring pixels) mostly share their
            // winner: the object record and material are re-fetched only when it changes
            Obj ob;
            float m7[7];
            int k_loaded = -1;
#pragma unroll 1
            for (int r = 0; r < kRays; r++) {
                const int k = l_idx[r];
                if (k < 0) continue;
                if (k != k_loaded) {
                    make_obj(w2o + (size_t)k * RRT_W2O_STRIDE, sc.obj_type[k], g.ct, ob);
                    const float* mat = mats + (size_t)k * RRT_MAT_STRIDE;
#pragma unroll
                    for (int q = 0; q < 7; q++) m7[q] = __ldg(mat + q);
                    k_loaded = k;
                }
                const float dwx = l_dw[r], dwy = l_dw[kRays + r], dwz = l_dw[2 * kRays + r];
                HitRec h;
                obj_test(ob, dwx, dwy, dwz, h);
                ShadeRec sr;
                float rgb[3];
                shade(sc.shader, sc.max_depth, ob, m7, g, h, sr, rgb);
                const int px = r / SPT;
#pragma unroll
                for (int q = 0; q < PIX; q++)
                    if (q == px) { pixsum[q][0] += rgb[0]; pixsum[q][1] += rgb[1]; pixsum[q][2] += rgb[2]; }
            }
        }

        const bool last_chunk = (sc0 == nchunks_s - 1);
        // ---- pixel value, image store, loss and upstream gradient.  A full, aligned tile row
        // (32*PIX pixels = 96*PIX contiguous floats per warp) moves through a per-warp
        // shared-memory stage so that global traffic is coalesced 16-byte vectors
        // (LDG.128 / STG.128, streaming); ragged tiles use scalar accesses.
        if (MODE != MODE_BWD && last_chunk) {
            const float inv = 1.0f / (float)S;
            const bool vec = P.vec_ok && row_ok && (blockIdx.x * 32 + 32) * PIX <= n;   // warp-uniform
            const size_t row_off = (((size_t)scene * P.rows + al) * n + (size_t)blockIdx.x * 32 * PIX) * 3;
            float* st = stage[warp];
            constexpr int kVec = 32 * PIX * 3 / 4;
            float tg[PIX][3];
            if (MODE == MODE_FUSED) {
                if (vec) {
                    const float4* g4 = reinterpret_cast<const float4*>(P.target + row_off);
                    for (int j = lane; j < kVec; j += 32) reinterpret_cast<float4*>(st)[j] = __ldcs(g4 + j);
                    __syncwarp();
#pragma unroll
                    for (int px = 0; px < PIX; px++)
#pragma unroll
                        for (int c = 0; c < 3; c++) tg[px][c] = st[(lane * PIX + px) * 3 + c];
                    __syncwarp();
                } else {
#pragma unroll
                    for (int px = 0; px < PIX; px++) {
                        const int b = b0 + px;
                        const bool ok = row_ok && b < n;
                        const size_t po = (((size_t)scene * P.rows + al) * n + b) * 3;
#pragma unroll
                        for (int c = 0; c < 3; c++) tg[px][c] = ok ? __ldcs(P.target + po + c) : 0.f;
                    }
                }
            }
            float v[PIX][3];
#pragma unroll
            for (int px = 0; px < PIX; px++) {
                const int b = b0 + px;
                const bool ok = row_ok && b < n;
#pragma unroll
                for (int c = 0; c < 3; c++) v[px][c] = pixsum[px][c] * inv;             // scene.py:49-50
                if (MODE == MODE_FUSED && ok) {
                    const float d0 = v[px][0] - tg[px][0], d1 = v[px][1] - tg[px][1], d2 = v[px][2] - tg[px][2];
                    loss_part += P.cw[0] * d0 * d0 + P.cw[1] * d1 * d1 + P.cw[2] * d2 * d2;
                    gpix[px][0] = 2.0f * P.cw[0] * d0 * inv;
                    gpix[px][1] = 2.0f * P.cw[1] * d1 * inv;
                    gpix[px][2] = 2.0f * P.cw[2] * d2 * inv;
                }
            }
            if (P.image) {
                if (vec) {
#pragma unroll
                    for (int px = 0; px < PIX; px++)
#pragma unroll
                        for (int c = 0; c < 3; c++) st[(lane * PIX + px) * 3 + c] = v[px][c];
                    __syncwarp();
                    float4* g4 = reinterpret_cast<float4*>(P.image + row_off);
                    for (int j = lane; j < kVec; j += 32) __stcs(g4 + j, reinterpret_cast<const float4*>(st)[j]);
                    __syncwarp();
                } else {
#pragma unroll
                    for (int px = 0; px < PIX; px++) {
                        const int b = b0 + px;
                        if (!(row_ok && b < n)) continue;
                        const size_t po = (((size_t)scene * P.rows + al) * n + b) * 3;
                        __stcs(P.image + po, v[px][0]); __stcs(P.image + po + 1, v[px][1]); __stcs(P.image + po + 2, v[px][2]);
                    }
                }
            }
        }

        // ---- reverse pass over the winners
        if (MODE != MODE_FWD) {
            // one extra (sentinel) trip after the last ray of the last sample chunk flushes the
            // running accumulator, so the warp-level flush code exists exactly once
            Obj ob;
            float m7[7];
            int k_loaded = -1;                       // see the shading loop
#pragma unroll 1
            for (int r = 0; r <= kRays; r++) {
                const bool fin = (r == kRays);
                if (fin && !last_chunk) break;
                int k = fin ? -1 : l_idx[r];
                HitRec h;
                float dwx = 0.f, dwy = 0.f, dwz = 0.f;
                float gc[3] = {0.f, 0.f, 0.f};
                if (!fin) {
                    const int px = r / SPT;
#pragma unroll
                    for (int q = 0; q < PIX; q++)
                        if (q == px) { gc[0] = gpix[q][0]; gc[1] = gpix[q][1]; gc[2] = gpix[q][2]; }
                }
                // a ray whose pixel has no upstream gradient contributes exactly zero (sparse
                // losses such as optimize_brightness.py:51 touch two pixels): skip it
                if (gc[0] == 0.f && gc[1] == 0.f && gc[2] == 0.f) k = -1;
                if (k >= 0) {
                    if (k != k_loaded) {
                        make_obj(w2o + (size_t)k * RRT_W2O_STRIDE, sc.obj_type[k], g.ct, ob);
                        const float* mat = mats + (size_t)k * RRT_MAT_STRIDE;
#pragma unroll
                        for (int q = 0; q < 7; q++) m7[q] = __ldg(mat + q);
                        k_loaded = k;
                    }
                    dwx = l_dw[r]; dwy = l_dw[kRays + r]; dwz = l_dw[2 * kRays + r];
                    obj_test(ob, dwx, dwy, dwz, h);
                    if (!(h.t < __int_as_float(0x7f800000))) k = -1;  // stale stored winner
                }
                // flush the running per-object accumulator when some lane changes object
                const bool change = fin ? (acc_key >= 0) : ((k >= 0) && (acc_key >= 0) && (k != acc_key));
                if (__any_sync(0xffffffffu, change)) {
                    warp_flush(acc_key, acc, slot_key, slots, gobj, lane);
                    acc_key = -1;
                }
                if (k >= 0) {
                    ShadeRec sr;
                    float rgb[3];
                    shade(sc.shader, sc.max_depth, ob, m7, g, h, sr, rgb);
                    acc_key = k;
                    {
                        const float rc3[3] = {l_rc[r], l_rc[kRays + r], l_rc[2 * kRays + r]};
                        backward_ray(sc.shader, sc.max_depth, ob, m7, g, h, sr, rc3, gc, acc, gg);
                    }
                }
            }
        }
    }  // sample chunks

    if (MODE != MODE_FWD) {
        // ---- warp -> CTA -> global reduction (per-object sums were flushed by the sentinel trip)
#pragma unroll
        for (int v = 0; v < 9; v++) {
            float x = warp_sum(gg[v]);
            if (lane == 0 && x != 0.f) atomicAdd(&gglob[v], x);
        }
        if (MODE == MODE_FUSED) {
            float x = warp_sum(loss_part);
            if (lane == 0) loss_warp[warp] = x;
        }
        __syncthreads();
        for (int q = tid; q < kSlots * 19; q += blockDim.x) {
            int s = q / 19, v = q - s * 19;
            int key = slot_key[s];
            float x = slots[s * kSlotStride + v];
            if (key >= 0 && x != 0.f) atomicAdd(&gobj[(size_t)key * RRT_OBJ_GRAD_STRIDE + v], x);
        }
        float* gglobal = gobj + (size_t)N * RRT_OBJ_GRAD_STRIDE;
        if (tid < 9) {
            // layout: Lhat -> slots 0..2, intensity 3..5, look_at 18..20
            int dst = tid < 6 ? tid : 12 + tid;
            if (gglob[tid] != 0.f) atomicAdd(&gglobal[dst], gglob[tid]);
        }
        if (MODE == MODE_FUSED && tid == 0) {
            double t = 0.0;
            for (int w = 0; w < nwarps; w++) t += (double)loss_warp[w];
            if (t != 0.0) atomicAdd(&P.loss[scene], t);
        }
    }
}

// ---------------------------------------------------------------- parameter -> matrix chain
// Affine 3x4 matrices [A|b] (bottom row 0 0 0 1 implied).  See include/rrt_b200.h.
struct Aff {
    float m[12];
};

__device__ __forceinline__ Aff aff_identity() {
    Aff r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.m[i] = 0.f;
    r.m[0] = r.m[5] = r.m[10] = 1.f;
    return r;
}

// C = A . B   (transform.py:35-38); products with exact zeros stay exact zeros
__device__ __forceinline__ Aff aff_mul(const Aff& A, const Aff& B) {
    Aff C;
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float v = A.m[r * 4 + 0] * B.m[0 * 4 + c] + A.m[r * 4 + 1] * B.m[1 * 4 + c] + A.m[r * 4 + 2] * B.m[2 * 4 + c];
            if (c == 3) v += A.m[r * 4 + 3];
            C.m[r * 4 + c] = v;
        }
    }
    return C;
}

// rotate(angle_deg, axis), transform.py:95-122 (Rodrigues form; axis assumed unit)
__device__ __forceinline__ void rot_entries(float angle, const float* a, float* R) {
    float s, c;
    sincosf(angle * 0.017453292519943295f, &s, &c);
    R[0] = a[0] * a[0] + (1.f - a[0] * a[0]) * c;
    R[1] = a[0] * a[1] * (1.f - c) - a[2] * s;
    R[2] = a[0] * a[2] * (1.f - c) + a[1] * s;
    R[3] = a[0] * a[1] * (1.f - c) + a[2] * s;
    R[4] = a[1] * a[1] + (1.f - a[1] * a[1]) * c;
    R[5] = a[1] * a[2] * (1.f - c) - a[0] * s;
    R[6] = a[0] * a[2] * (1.f - c) - a[1] * s;
    R[7] = a[1] * a[2] * (1.f - c) + a[0] * s;
    R[8] = a[2] * a[2] + (1.f - a[2] * a[2]) * c;
}

__device__ __forceinline__ Aff chain_op_matrix(const int32_t* op, const float* __restrict__ values) {
    const int kind = op[0] & 0xff;
    const bool inv = (op[0] & RRT_CHAIN_INVERT) != 0;
    Aff M = aff_identity();
    if (kind == RRT_CHAIN_TRANSLATE) {            // transform.py:60-75
        const float* v = values + op[1];
        M.m[3] = inv ? -v[0] : v[0]; M.m[7] = inv ? -v[1] : v[1]; M.m[11] = inv ? -v[2] : v[2];
    } else if (kind == RRT_CHAIN_SCALE) {         // transform.py:78-93 (inverse is 1/x)
        const float* v = values + op[1];
        M.m[0] = inv ? 1.f / v[0] : v[0]; M.m[5] = inv ? 1.f / v[1] : v[1]; M.m[10] = inv ? 1.f / v[2] : v[2];
    } else if (kind == RRT_CHAIN_ROTATE) {        // inverse = transpose
        float R[9];
        rot_entries(values[op[1]], values + op[2], R);
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) M.m[r * 4 + c] = inv ? R[c * 3 + r] : R[r * 3 + c];
    }
    return M;
}

__device__ __forceinline__ Aff chain_forward_one(const int32_t* __restrict__ ops, const int32_t* __restrict__ chain_begin,
                                                 int k, const float* __restrict__ values) {
    Aff M = aff_identity();
    for (int j = chain_begin[k]; j < chain_begin[k + 1]; j++) M = aff_mul(M, chain_op_matrix(ops + 4 * j, values));
    return M;
}

__global__ void chain_forward_kernel(const int32_t* __restrict__ ops, const int32_t* __restrict__ chain_begin,
                                     int num_chains, const float* __restrict__ values, float* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= num_chains) return;
    const Aff M = chain_forward_one(ops, chain_begin, k, values);
#pragma unroll
    for (int i = 0; i < 12; i++) out[(size_t)k * 12 + i] = M.m[i];
}

// dL/d(op j) = P_{j-1}^T . G . S_{j+1}^T with P = prefix product, S = suffix product
// (4x4 with the implied bottom row); then into the primitive's own parameters.
// G = dL/d(row k of the chain's output); accumulates into g_values with atomics (parameters may be
// shared between chains).
__device__ __noinline__ void chain_backward_one(const int32_t* __restrict__ ops, const int32_t* __restrict__ chain_begin,
                                                int k, const float* __restrict__ values, const float* G,
                                                float* __restrict__ g_values) {
    const int b = chain_begin[k], e = chain_begin[k + 1], n = e - b;
    if (n <= 0 || n > RRT_CHAIN_MAX_OPS) return;
    Aff mats[RRT_CHAIN_MAX_OPS], pre[RRT_CHAIN_MAX_OPS + 1];
    pre[0] = aff_identity();
    for (int j = 0; j < n; j++) {
        mats[j] = chain_op_matrix(ops + 4 * (b + j), values);
        pre[j + 1] = aff_mul(pre[j], mats[j]);
    }
    Aff suf = aff_identity();                     // product of ops j+1..n-1
    for (int j = n - 1; j >= 0; j--) {
        // out = P . M_j . S  (affine).  T = G . S^T restricted to what reaches M_j's 3x4 block:
        //   T[r][c] = sum_q G[r][q] S[c][q] (c<3: q over 0..3 with S[c][3]=b_c) ; T[r][3] = G[r][3]
        float T[12];
#pragma unroll
        for (int r = 0; r < 3; r++) {
#pragma unroll
            for (int c = 0; c < 3; c++)
                T[r * 4 + c] = G[r * 4 + 0] * suf.m[c * 4 + 0] + G[r * 4 + 1] * suf.m[c * 4 + 1] +
                               G[r * 4 + 2] * suf.m[c * 4 + 2] + G[r * 4 + 3] * suf.m[c * 4 + 3];
            T[r * 4 + 3] = G[r * 4 + 3];
        }
        // D = P_A^T . T   (gradient w.r.t. M_j's [A|b])
        const Aff& Pm = pre[j];
        float D[12];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 4; c++)
                D[r * 4 + c] = Pm.m[0 * 4 + r] * T[0 * 4 + c] + Pm.m[1 * 4 + r] * T[1 * 4 + c] + Pm.m[2 * 4 + r] * T[2 * 4 + c];
        const int32_t* op = ops + 4 * (b + j);
        const int kind = op[0] & 0xff;
        const bool inv = (op[0] & RRT_CHAIN_INVERT) != 0;
        if (kind == RRT_CHAIN_TRANSLATE) {
            const float sgn = inv ? -1.f : 1.f;
            atomicAdd(&g_values[op[1] + 0], sgn * D[3]);
            atomicAdd(&g_values[op[1] + 1], sgn * D[7]);
            atomicAdd(&g_values[op[1] + 2], sgn * D[11]);
        } else if (kind == RRT_CHAIN_SCALE) {
            const float* v = values + op[1];
#pragma unroll
            for (int i = 0; i < 3; i++) atomicAdd(&g_values[op[1] + i], inv ? -D[i * 5] / (v[i] * v[i]) : D[i * 5]);
        } else if (kind == RRT_CHAIN_ROTATE) {
            const float ang = values[op[1]];
            const float* a = values + op[2];
            float s, c;
            sincosf(ang * 0.017453292519943295f, &s, &c);
            float Gr[9];                         // gradient w.r.t. the (non-transposed) rotation entries
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int cc = 0; cc < 3; cc++) Gr[r * 3 + cc] = inv ? D[cc * 4 + r] : D[r * 4 + cc];
            const float dc = -s * 0.017453292519943295f, ds = c * 0.017453292519943295f, omc = 1.f - c;
            float g_ang = Gr[0] * (1.f - a[0] * a[0]) * dc + Gr[4] * (1.f - a[1] * a[1]) * dc + Gr[8] * (1.f - a[2] * a[2]) * dc +
                          Gr[1] * (-a[0] * a[1] * dc - a[2] * ds) + Gr[2] * (-a[0] * a[2] * dc + a[1] * ds) +
                          Gr[3] * (-a[0] * a[1] * dc + a[2] * ds) + Gr[5] * (-a[1] * a[2] * dc - a[0] * ds) +
                          Gr[6] * (-a[0] * a[2] * dc - a[1] * ds) + Gr[7] * (-a[1] * a[2] * dc + a[0] * ds);
            float g_a0 = Gr[0] * 2.f * a[0] * omc + (Gr[1] + Gr[3]) * a[1] * omc + (Gr[2] + Gr[6]) * a[2] * omc + (Gr[7] - Gr[5]) * s;
            float g_a1 = Gr[4] * 2.f * a[1] * omc + (Gr[1] + Gr[3]) * a[0] * omc + (Gr[5] + Gr[7]) * a[2] * omc + (Gr[2] - Gr[6]) * s;
            float g_a2 = Gr[8] * 2.f * a[2] * omc + (Gr[2] + Gr[6]) * a[0] * omc + (Gr[5] + Gr[7]) * a[1] * omc + (Gr[3] - Gr[1]) * s;
            atomicAdd(&g_values[op[1]], g_ang);
            atomicAdd(&g_values[op[2] + 0], g_a0);
            atomicAdd(&g_values[op[2] + 1], g_a1);
            atomicAdd(&g_values[op[2] + 2], g_a2);
        }
        suf = aff_mul(mats[j], suf);
    }
}

__global__ void chain_backward_kernel(const int32_t* __restrict__ ops, const int32_t* __restrict__ chain_begin,
                                      int num_chains, const float* __restrict__ values,
                                      const float* __restrict__ g_out, float* __restrict__ g_values) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= num_chains) return;
    float G[12];
#pragma unroll
    for (int i = 0; i < 12; i++) G[i] = g_out[(size_t)k * 12 + i];
    chain_backward_one(ops, chain_begin, k, values, G, g_values);
}

// ---------------------------------------------------------------- the small-scene kernel
// Latency-oriented variant for the reference's own workloads (optimize_brightness.py,
// match_mirror.py, test_balls.py, the orbit decoder: 32..128 pixels a side, 2..3 shapes).
// There the 8-rays-per-thread kernel above is a serial dependency chain on a grid that
// cannot fill the machine; here ONE thread owns ONE ray, the S samples of a pixel sit in
// adjacent lanes (S a power of two <= 32) and are combined by shuffles, and the whole
// object table (<= kSmallMaxN records) plus the materials live in shared memory.  Same
// device routines (obj_test / shade / backward_ray), same canonical order, same sample
// summation order as render_kernel, so the two kernels agree bit for bit on masks.
constexpr int kSmallMaxN = 32;
constexpr int kSmallThreads = 128;
#ifndef RRT_SMALL_MIN_BLOCKS
#define RRT_SMALL_MIN_BLOCKS 8   // 64 registers: measured 357 -> 276 us on the 512-scene orbit batch, C1/C3 unchanged
#endif
constexpr long long kSmallDefaultMaxRays = 16 << 20;  // total rays of a call (all scenes) up to which it is used
                                                      // (orbit batch, 8.4 M rays: 258 us here vs 295 us on the general kernel)

template <int MODE, bool STEP = false>
__global__ void __launch_bounds__(kSmallThreads, RRT_SMALL_MIN_BLOCKS) render_small_kernel(const __grid_constant__ KParams P) {
    __shared__ float4 tab[kSmallMaxN * 4];
    __shared__ float mat_s[kSmallMaxN * RRT_MAT_STRIDE];
    __shared__ Globals g;
    __shared__ float slots[kSmallMaxN * kSlotStride];
    __shared__ float gglob[9];
    __shared__ float loss_warp[kSmallThreads / 32];

    const rrt_scene& sc = P.sc;
    const int n = sc.n, N = sc.num_objects, S = sc.samples;
    const int scene = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    // 32-bit index arithmetic (the launcher guarantees rows*n*S < 2^31; S is a power of two)
    const unsigned rays_scene = (unsigned)P.rows * (unsigned)n * (unsigned)S;
    const unsigned gid = blockIdx.x * kSmallThreads + tid;
    const bool active = gid < rays_scene;
    const int s = (int)(gid & (unsigned)(S - 1));
    const unsigned pl = gid >> (31 - __clz(S));         // slab-local pixel index
    const int al = active ? (int)(pl / (unsigned)n) : 0, b = active ? (int)(pl - (unsigned)al * (unsigned)n) : 0;
    const int a = sc.row_begin + al;
    const size_t po = (((size_t)scene * P.rows + al) * n + b) * 3;
    const float inv = 1.0f / (float)S;

    float gc[3] = {0.f, 0.f, 0.f};
    if (MODE == MODE_BWD) {
        if (active) { gc[0] = P.dl_dimage[po] * inv; gc[1] = P.dl_dimage[po + 1] * inv; gc[2] = P.dl_dimage[po + 2] * inv; }
        // sparse upstream gradients: see render_kernel
        if (!__syncthreads_or((gc[0] != 0.f) | (gc[1] != 0.f) | (gc[2] != 0.f))) return;
    }

    // ---- per-scene constants, object records and materials -> shared memory (one barrier)
    const float* cam = sc.camera + (size_t)scene * sc.camera_scene_stride;
    const float* w2o = sc.w2o + (size_t)scene * sc.w2o_scene_stride;
    const float* mats = sc.material + (size_t)scene * sc.material_scene_stride;
    if (tid < N) {
        const float ct[3] = {__ldg(cam + 3), __ldg(cam + 7), __ldg(cam + 11)};
        Obj ob;
        if (STEP) {   // whole-step variant: the shape's w2o rows come straight from the parameter chain
            const Aff M = chain_forward_one(P.step.ops, P.step.chain_begin, tid, P.step.values);
            make_obj_rows(M.m, sc.obj_type[tid], ct, ob);
        } else {
            make_obj(w2o + (size_t)tid * RRT_W2O_STRIDE, sc.obj_type[tid], ct, ob);
        }
        store_rec(tab + 4 * tid, ob);
    }
    for (int q = tid; q < N * RRT_MAT_STRIDE; q += kSmallThreads) mat_s[q] = __ldg(mats + q);
    if (MODE != MODE_FWD) {
        for (int q = tid; q < N * kSlotStride; q += kSmallThreads) slots[q] = 0.f;
        if (tid < 9) gglob[tid] = 0.f;
    }
    if (warp == kSmallThreads / 32 - 1) {               // last warp: usually no object to build
        const float* li = sc.light + (size_t)scene * sc.light_scene_stride;
        if (lane < 3) {
            g.C[lane * 3 + 0] = cam[lane * 4 + 0];
            g.C[lane * 3 + 1] = cam[lane * 4 + 1];
            g.C[lane * 3 + 2] = cam[lane * 4 + 2];
            g.ct[lane] = cam[lane * 4 + 3];
            g.look[lane] = cam[12 + lane];
            g.I[lane] = li[3 + lane];
            const float l0 = li[0], l1 = li[1], l2 = li[2];
            const float ln = sqrtf(l0 * l0 + l1 * l1 + l2 * l2);          // scene.py:83-86
            g.L[lane] = li[lane];
            g.Lh[lane] = li[lane] / ln;
            if (lane == 0) {
                g.Ln = ln;
                const float L3[3] = {l0, l1, l2};
                canon_to_light(L3, g.U);
            }
        }
    }

    // ---- this thread's ray (independent of the shared tables up to the camera matrix)
    float rcx = 0.f, rcy = 0.f, rcz = 0.f;
    if (active) {
        const int i = sc.transpose ? b : a, j = sc.transpose ? a : b;
        float bx, by, bz;
        if (sc.base_rays) {
            const float* br = sc.base_rays + ((size_t)i * n + j) * 3;
            bx = __ldg(br); by = __ldg(br + 1); bz = __ldg(br + 2);
        } else {
            base_ray(n, P.lin_step, i, j, bx, by, bz);
        }
        float jx, jy;
        if (sc.jitter_x) {
            const size_t off = (size_t)scene * sc.jitter_scene_stride + (size_t)gid;   // [rows][n][S]
            jx = __ldg(sc.jitter_x + off);
            jy = __ldg(sc.jitter_y + off);
        } else {
            jx = rrt_rng(sc.seed, scene + sc.scene_begin, (uint32_t)(a * n + b), s, 0);
            jy = rrt_rng(sc.seed, scene + sc.scene_begin, (uint32_t)(a * n + b), s, 1);
        }
        const float ox = P.pow2 ? jitter_offset_pow2(jx, s, P.inv_s, P.inv_n) : jitter_offset(jx, s, S, n);
        const float oy = P.pow2 ? jitter_offset_pow2(jy, s, P.inv_s, P.inv_n) : jitter_offset(jy, s, S, n);
        rcx = __fadd_rn(bx, ox);
        rcy = __fadd_rn(by, oy);
        rcz = bz;
    }
    float tgt[3] = {0.f, 0.f, 0.f};
    if (MODE == MODE_FUSED && active) { tgt[0] = __ldg(P.target + po); tgt[1] = __ldg(P.target + po + 1); tgt[2] = __ldg(P.target + po + 2); }
    int stored = -1;
    const bool use_stored = (MODE == MODE_BWD) && (P.hit_in != nullptr);
    if (use_stored && active) {
        const int kk = P.hit_in[(((size_t)scene * S + s) * P.rows + al) * n + b];
        stored = (kk >= 0 && kk < N) ? kk : -1;          // never trust an index buffer blindly; a winner
    }                                                    // flagged RRT_HIT_SHADOWED (>= N) carries no gradient
    __syncthreads();

    // camera.o2w (orbit_experiments/scene.py:80); the fma chain returns its input for C = I
    const float wx = dot3_canon(g.C[0], g.C[1], g.C[2], rcx, rcy, rcz);
    const float wy = dot3_canon(g.C[3], g.C[4], g.C[5], rcx, rcy, rcz);
    const float wz = dot3_canon(g.C[6], g.C[7], g.C[8], rcx, rcy, rcz);

    // ---- nearest hit: list order, strict '<' (scene.py:46-47)
    const float inf = __int_as_float(0x7f800000);
    float tmin = inf;
    int idx = -1;
    if (use_stored) {
        idx = stored;
    } else if (active) {
#pragma unroll 1
        for (int k = 0; k < N; k++) {
            Obj ob;
            load_rec(tab + 4 * k, ob);
            HitRec h;
            const float t = obj_test<true>(ob, wx, wy, wz, h);
            if (t < tmin) { tmin = t; idx = k; }
        }
    }
    bool in_shadow = false;
    if ((sc.flags & RRT_FLAG_SHADOWS) && !use_stored && idx >= 0) {   // hard shadows (opt-in)
#pragma unroll 1
        for (int k = 0; k < N && !in_shadow; k++)
            if (k != idx) in_shadow = shadow_test(tab + 4 * k, wx, wy, wz, tmin, g.U);
    }
    if (MODE != MODE_BWD && active) {
        const size_t ro = (((size_t)scene * S + s) * P.rows + al) * n + b;
        if (P.hit_out) P.hit_out[ro] = idx | (in_shadow ? RRT_HIT_SHADOWED : 0);
        if (MODE == MODE_FWD && P.tmin_out) P.tmin_out[ro] = tmin;
    }
    if (in_shadow) idx = -1;                             // (0,0,0) and no gradient from here on

    // ---- winner: hit record, shading
    Obj ob;
    HitRec h;
    ShadeRec sr;
    float m7[7];
    float rgb[3] = {0.f, 0.f, 0.f};
    if (MODE == MODE_BWD && gc[0] == 0.f && gc[1] == 0.f && gc[2] == 0.f) idx = -1;
    if (idx >= 0) {
        load_rec(tab + 4 * idx, ob);
        obj_test<true>(ob, wx, wy, wz, h);
        if (!(h.t < inf)) idx = -1;                      // stale stored winner
    }
    if (idx >= 0) {
#pragma unroll
        for (int q = 0; q < 7; q++) m7[q] = mat_s[idx * RRT_MAT_STRIDE + q];
        shade(sc.shader, sc.max_depth, ob, m7, g, h, sr, rgb);
    }

    float loss_part = 0.f;
    if (MODE != MODE_BWD) {
        // pixel = mean over the S lanes of this pixel, summed in sample order like render_kernel
        const int base = lane & ~(S - 1);
        float sum[3] = {0.f, 0.f, 0.f};
        if (S == 4) {                                    // the reference's default (scene.py:18)
#pragma unroll
            for (int j = 0; j < 4; j++) {
#pragma unroll
                for (int c = 0; c < 3; c++) sum[c] += __shfl_sync(full, rgb[c], base + j);
            }
        } else {
            for (int j = 0; j < S; j++) {
#pragma unroll
                for (int c = 0; c < 3; c++) sum[c] += __shfl_sync(full, rgb[c], base + j);
            }
        }
        const float v0 = sum[0] * inv, v1 = sum[1] * inv, v2 = sum[2] * inv;      // scene.py:49-50
        if (active && s == 0 && P.image) { P.image[po] = v0; P.image[po + 1] = v1; P.image[po + 2] = v2; }
        if (MODE == MODE_FUSED && active) {
            const float d0 = v0 - tgt[0], d1 = v1 - tgt[1], d2 = v2 - tgt[2];
            if (s == 0) loss_part = P.cw[0] * d0 * d0 + P.cw[1] * d1 * d1 + P.cw[2] * d2 * d2;
            gc[0] = 2.0f * P.cw[0] * d0 * inv;
            gc[1] = 2.0f * P.cw[1] * d1 * inv;
            gc[2] = 2.0f * P.cw[2] * d2 * inv;
        }
    }

    if (MODE != MODE_FWD) {
        // ---- reverse pass through the winner, then thread -> warp -> CTA -> global
        float acc[19], gg[9];
#pragma unroll
        for (int v = 0; v < 19; v++) acc[v] = 0.f;
#pragma unroll
        for (int v = 0; v < 9; v++) gg[v] = 0.f;
        int key = idx;
        if (gc[0] == 0.f && gc[1] == 0.f && gc[2] == 0.f) key = -1;
        if (key >= 0) {
            const float rc3[3] = {rcx, rcy, rcz};
            backward_ray(sc.shader, sc.max_depth, ob, m7, g, h, sr, rc3, gc, acc, gg);
        }
        unsigned todo = __ballot_sync(full, key >= 0);
        if (todo) {
            while (todo) {
                const int leader = __ffs(todo) - 1;
                const int k = __shfl_sync(full, key, leader);
                const bool mine = (key == k);
                todo &= ~__ballot_sync(full, mine);
                int v;
                const float x = warp_reduce19(acc, mine, lane, v);
                if (v >= 0 && x != 0.f) atomicAdd(&slots[k * kSlotStride + v], x);
            }
#pragma unroll
            for (int v = 0; v < 9; v++) {
                const float x = warp_sum(gg[v]);
                if (lane == 0 && x != 0.f) atomicAdd(&gglob[v], x);
            }
        }
        if (MODE == MODE_FUSED) {
            const float x = warp_sum(loss_part);
            if (lane == 0) loss_warp[warp] = x;
        }
        __syncthreads();
        float* gobj = P.grad + (size_t)scene * RRT_GRAD_SIZE(N);
        for (int q = tid; q < N * 19; q += kSmallThreads) {
            const int k = q / 19, v = q - k * 19;
            const float x = slots[k * kSlotStride + v];
            if (x != 0.f) atomicAdd(&gobj[(size_t)k * RRT_OBJ_GRAD_STRIDE + v], x);
        }
        float* gglobal = gobj + (size_t)N * RRT_OBJ_GRAD_STRIDE;
        if (tid < 9) {
            const int dst = tid < 6 ? tid : 12 + tid;    // Lhat 0..2, intensity 3..5, look_at 18..20
            if (gglob[tid] != 0.f) atomicAdd(&gglobal[dst], gglob[tid]);
        }
        if (MODE == MODE_FUSED && tid == 0) {
            double t = 0.0;
            for (int w = 0; w < kSmallThreads / 32; w++) t += (double)loss_warp[w];
            if (t != 0.0) atomicAdd(&P.loss[scene], t);
        }
        if (STEP) {
            // ---- the rest of the optimise step, by the LAST CTA to get here (ticket): finalize the
            // raw sums into d/d w2o, chain them back to the parameters (T.grad through
            // translate/scale/rotate, transform.py:56-122), apply var <- var - lr*grad
            // (optimize.py:26-27), publish the loss and re-zero the scratch for the next step.
            __shared__ int is_last;
            __threadfence();
            __syncthreads();
            if (tid == 0) is_last = (atomicAdd(P.step.ticket, 1u) == gridDim.x - 1);
            __syncthreads();
            if (is_last) {
                __threadfence();
                const rrt_step& q = P.step;
                if (tid < N) {
                    float Mm[9], gb[3], G[12];
#pragma unroll
                    for (int v = 0; v < 9; v++) Mm[v] = __ldcg(gobj + (size_t)tid * RRT_OBJ_GRAD_STRIDE + v);
#pragma unroll
                    for (int v = 0; v < 3; v++) gb[v] = __ldcg(gobj + (size_t)tid * RRT_OBJ_GRAD_STRIDE + 9 + v);
#pragma unroll
                    for (int r = 0; r < 3; r++) {          // d/dA = M C^T + g_b ct^T ; d/db = g_b (finalize_grads)
#pragma unroll
                        for (int c = 0; c < 3; c++)
                            G[r * 4 + c] = Mm[r * 3] * g.C[c * 3] + Mm[r * 3 + 1] * g.C[c * 3 + 1] + Mm[r * 3 + 2] * g.C[c * 3 + 2] +
                                           gb[r] * g.ct[c];
                        G[r * 4 + 3] = gb[r];
                    }
                    chain_backward_one(q.ops, q.chain_begin, tid, q.values, G, q.g_values);
                }
                __threadfence();
                __syncthreads();
                for (int p = tid; p < q.num_values; p += kSmallThreads) {
                    const float gv = __ldcg(q.g_values + p);
                    if (p >= q.param_begin) q.values[p] -= q.lr * gv;
                    q.g_values[p] = 0.f;
                }
                for (int v = tid; v < (int)RRT_GRAD_SIZE(N); v += kSmallThreads) gobj[v] = 0.f;
                if (tid == 0) {
                    *q.loss_out = (float)__ldcg(P.loss);
                    *P.loss = 0.0;
                    *q.ticket = 0u;
                }
            }
        }
    }
}

// ---------------------------------------------------------------- sweep-record table (for TMA staging)
// grid = (chunks of kObjChunk objects, scenes).  Writes the 64-byte records the render kernels
// bulk-copy into shared memory; the chunk's class bits (squares / general spheres present) go
// into the spare slot of its first record.
__global__ void __launch_bounds__(128) build_records_kernel(const rrt_scene sc, float* __restrict__ records) {
    __shared__ int cls_s;
    const int scene = blockIdx.y, kb = blockIdx.x * kObjChunk, N = sc.num_objects;
    const int cnt = min(kObjChunk, N - kb);
    if (threadIdx.x == 0) cls_s = 0;
    __syncthreads();
    const float* cam = sc.camera + (size_t)scene * sc.camera_scene_stride;
    const float ct[3] = {__ldg(cam + 3), __ldg(cam + 7), __ldg(cam + 11)};
    const float* w2o = sc.w2o + (size_t)scene * sc.w2o_scene_stride;
    float4* out = reinterpret_cast<float4*>(records + ((size_t)scene * N + kb) * RRT_RECORD_FLOATS);
    int cls = 0;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
        Obj ob;
        make_obj(w2o + (size_t)(kb + k) * RRT_W2O_STRIDE, sc.obj_type[kb + k], ct, ob, true);
        store_rec(out + 4 * k, ob);
        cls |= ob.flags;
    }
    if (cls) atomicOr(&cls_s, cls);
    __syncthreads();
    if (threadIdx.x == 0 && cnt > 0) reinterpret_cast<float*>(out)[15] = __int_as_float(cls_s);
}

// ---------------------------------------------------------------- primary-ray grid table
__global__ void primary_rays_kernel(int n, double step, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n) return;
    float x, y, z;
    base_ray(n, step, i, j, x, y, z);
    float* o = out + ((size_t)i * n + j) * 3;
    o[0] = x; o[1] = y; o[2] = z;
}

// ---------------------------------------------------------------- gradient finalisation
// One CTA per scene.  Converts the raw per-object sums [M, g_b] into d/d w2o and
// folds the camera and light chains (see backward_ray).
__global__ void finalize_grads(const __grid_constant__ KParams P) {
    const rrt_scene& sc = P.sc;
    const int N = sc.num_objects, scene = blockIdx.x, tid = threadIdx.x;
    float* gobj = P.grad + (size_t)scene * RRT_GRAD_SIZE(N);
    float* gglobal = gobj + (size_t)N * RRT_OBJ_GRAD_STRIDE;
    const float* cam = sc.camera + (size_t)scene * sc.camera_scene_stride;
    const float* w2o = sc.w2o + (size_t)scene * sc.w2o_scene_stride;
    __shared__ float camg[12];
    if (tid < 12) camg[tid] = 0.f;
    __syncthreads();
    float C[9], ct[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        C[r * 3] = cam[r * 4]; C[r * 3 + 1] = cam[r * 4 + 1]; C[r * 3 + 2] = cam[r * 4 + 2];
        ct[r] = cam[r * 4 + 3];
    }
    float cg[12];
#pragma unroll
    for (int q = 0; q < 12; q++) cg[q] = 0.f;
    for (int k = tid; k < N; k += blockDim.x) {
        float* og = gobj + (size_t)k * RRT_OBJ_GRAD_STRIDE;
        float M[9], gb[3], A[9];
#pragma unroll
        for (int q = 0; q < 9; q++) M[q] = og[q];
#pragma unroll
        for (int q = 0; q < 3; q++) gb[q] = og[9 + q];
        const float* w = w2o + (size_t)k * RRT_W2O_STRIDE;
#pragma unroll
        for (int r = 0; r < 3; r++) { A[r * 3] = w[r * 4]; A[r * 3 + 1] = w[r * 4 + 1]; A[r * 3 + 2] = w[r * 4 + 2]; }
        // d/dA = M C^T + g_b ct^T ;  d/db = g_b        (d' = A C r, o' = A ct + b)
        float out[12];
#pragma unroll
        for (int r = 0; r < 3; r++) {
#pragma unroll
            for (int c = 0; c < 3; c++)
                out[r * 4 + c] = M[r * 3] * C[c * 3] + M[r * 3 + 1] * C[c * 3 + 1] + M[r * 3 + 2] * C[c * 3 + 2] + gb[r] * ct[c];
            out[r * 4 + 3] = gb[r];
        }
#pragma unroll
        for (int q = 0; q < 12; q++) og[q] = out[q];
        if (sc.camera_grad) {
            // d/dC = sum_k A_k^T M_k ; d/dct = sum_k A_k^T g_b,k
#pragma unroll
            for (int r = 0; r < 3; r++) {
#pragma unroll
                for (int c = 0; c < 3; c++)
                    cg[r * 4 + c] += A[0 * 3 + r] * M[0 * 3 + c] + A[1 * 3 + r] * M[1 * 3 + c] + A[2 * 3 + r] * M[2 * 3 + c];
                cg[r * 4 + 3] += A[0 * 3 + r] * gb[0] + A[1 * 3 + r] * gb[1] + A[2 * 3 + r] * gb[2];
            }
        }
    }
    if (sc.camera_grad) {
#pragma unroll
        for (int q = 0; q < 12; q++) {
            float x = warp_sum(cg[q]);
            if ((tid & 31) == 0 && x != 0.f) atomicAdd(&camg[q], x);
        }
    }
    __syncthreads();
    if (tid < 12) gglobal[6 + tid] = camg[tid];
    if (tid == 0) {
        // Lhat = L/|L|  =>  g_L = (g_Lhat - Lhat (Lhat . g_Lhat)) / |L|    scene.py:83-86
        const float* li = sc.light + (size_t)scene * sc.light_scene_stride;
        float L0 = li[0], L1 = li[1], L2 = li[2];
        float ln = sqrtf(L0 * L0 + L1 * L1 + L2 * L2);
        float h0 = L0 / ln, h1 = L1 / ln, h2 = L2 / ln;
        float g0 = gglobal[0], g1 = gglobal[1], g2 = gglobal[2];
        float dot = h0 * g0 + h1 * g1 + h2 * g2;
        gglobal[0] = (g0 - h0 * dot) / ln;
        gglobal[1] = (g1 - h1 * dot) / ln;
        gglobal[2] = (g2 - h2 * dot) / ln;
    }
}

// ---------------------------------------------------------------- gradient exchange over peer memory
// The one exchange step of the sharded path (row slabs / scene ranges per GPU): every rank holds
// a small vector [gradient (float32) | loss (float64)] and all ranks need the sum.  Instead of
// an NCCL allreduce (plus the two copy kernels that pack its buffer) ONE kernel per rank
//   1. PUSHES its values, converted to float64, into slot[rank] of every peer's buffer with
//      plain stores through the NVLink peer mapping,
//   2. publishes a per-(CTA, source) flag on every peer (fence + release store) and waits for
//      the same flags from all peers (acquire loads) -- CTA c only depends on CTA c of the
//      peers, so no grid-wide barrier is needed,
//   3. sums the `world` slots in RANK ORDER (every rank gets the same bits; deterministic).
// Buffers alternate by epoch parity: a rank can only start epoch e+2 after every peer has
// finished reading epoch e (it needs their epoch e+1 flags first), so two copies suffice.
// Flags carry the epoch number, one per source, so a fast peer's next epoch cannot be
// mistaken for a slow peer's current one.
constexpr int kPeerCtas = 16;
constexpr int kPeerMaxWorld = 16;
constexpr int kPeerEpochOffset = kPeerCtas * kPeerMaxWorld;   // per-CTA epoch counters (local use only)

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) peer_allreduce_kernel(const float* __restrict__ grad, const double* __restrict__ loss,
                                                             int n, int nloss, void* const* __restrict__ peer_buf,
                                                             void* const* __restrict__ peer_sig, int rank, int world,
                                                             double* __restrict__ out) {
    const int cta = blockIdx.x, tid = threadIdx.x;
    unsigned* sig_local = reinterpret_cast<unsigned*>(peer_sig[rank]);
    __shared__ unsigned epoch_s;
    __shared__ int timed_out;
    if (tid == 0) {
        timed_out = 0;
        epoch_s = sig_local[kPeerEpochOffset + cta] + 1u;
        sig_local[kPeerEpochOffset + cta] = epoch_s;
    }
    __syncthreads();
    const unsigned epoch = epoch_s;
    const int total = n + nloss;
    const int per = (total + gridDim.x - 1) / gridDim.x;
    const int lo = cta * per, hi = min(total, lo + per);
    const size_t slot = ((size_t)(epoch & 1u) * world + rank) * (size_t)total;
    // 1. push
    for (int i = lo + tid; i < hi; i += blockDim.x) {
        const double v = i < n ? (double)grad[i] : loss[i - n];
        for (int p = 0; p < world; p++) reinterpret_cast<double*>(peer_buf[p])[slot + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish + wait
    if (tid < world) {
        st_release_sys(reinterpret_cast<unsigned*>(peer_sig[tid]) + cta * kPeerMaxWorld + rank, epoch);
        const unsigned* mine = sig_local + cta * kPeerMaxWorld + tid;
        // bounded wait (~10 s): a peer that died must not hang this GPU; the sums become NaN
        long long spins = 0;
        while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
            if (++spins > (1LL << 26)) { timed_out = 1; break; }
            if (spins > 1024) __nanosleep(128);
        }
    }
    __syncthreads();
    // 3. sum in rank order
    const double* local = reinterpret_cast<const double*>(peer_buf[rank]) + (size_t)(epoch & 1u) * world * (size_t)total;
    for (int i = lo + tid; i < hi; i += blockDim.x) {
        double sum = 0.0;
        for (int p = 0; p < world; p++) sum += __ldcg(local + (size_t)p * total + i);
        out[i] = timed_out ? __longlong_as_double(0x7ff8000000000000LL) : sum;
    }
}

// ---------------------------------------------------------------- FP32 peak micro-benchmarks
template <int MODE>
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float seed) {
    if (MODE == 0) {
        float a[16];
#pragma unroll
        for (int q = 0; q < 16; q++) a[q] = seed + (float)(threadIdx.x + q);
        float b = 1.0000001f, c = 1e-7f;
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int rep = 0; rep < 8; rep++)
#pragma unroll
                for (int q = 0; q < 16; q++) a[q] = __fmaf_rn(a[q], b, c);
        }
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 16; q++) s += a[q];
        if (s == 123.456f) out[0] = s;
    } else {
        // MODE 1: packed FFMA2 only.  MODE 2: + one ALU-pipe FMNMX3 per 4 FFMA2 (diagnostic:
        // does a non-FMA instruction issue in the shadow of an FFMA2 or cost its own cycle?)
        u64 a[16];
#pragma unroll
        for (int q = 0; q < 16; q++) a[q] = pk(seed + (float)(threadIdx.x + q), seed - (float)q);
        u64 b = pk(1.0000001f, 0.9999999f), c = pk(1e-7f, -1e-7f);
        float m = 0.f;
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int rep = 0; rep < 8; rep++)
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    a[q] = fma2(a[q], b, c);
                    if (MODE == 2 && (q & 3) == 3) {
                        float lo, hi;
                        upk(a[q - 3], lo, hi);
                        m = fmaxf(m, fmaxf(lo, hi));
                    }
                }
        }
        float s = m;
#pragma unroll
        for (int q = 0; q < 16; q++) { float lo, hi; upk(a[q], lo, hi); s += lo + hi; }
        if (s == 123.456f) out[0] = s;
    }
}

// ---------------------------------------------------------------- host side
thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* a = "") {
    snprintf(g_err, sizeof g_err, fmt, a);
    return code;
}

int check_scene(const rrt_scene* sc, int* rows_out) {
    if (!sc) return fail(RRT_ERR_INVALID, "scene is NULL");
    if (sc->n <= 0) return fail(RRT_ERR_INVALID, "n must be > 0");
    if (sc->samples <= 0 || sc->samples > 128) return fail(RRT_ERR_INVALID, "samples must be in 1..128");
    if (sc->num_objects < 0) return fail(RRT_ERR_INVALID, "num_objects must be >= 0");
    if (sc->num_scenes <= 0 || sc->num_scenes > 65535) return fail(RRT_ERR_INVALID, "num_scenes must be in 1..65535");
    if (sc->shader < 0 || sc->shader > 2) return fail(RRT_ERR_INVALID, "unknown shader");
    if (sc->row_begin < 0 || sc->row_begin >= sc->n) return fail(RRT_ERR_INVALID, "row_begin out of range");
    int rows = sc->row_count > 0 ? sc->row_count : sc->n - sc->row_begin;
    if (sc->row_begin + rows > sc->n) return fail(RRT_ERR_INVALID, "row slab exceeds the image");
    if (sc->num_objects > 0 && (!sc->obj_type || !sc->w2o || !sc->material)) return fail(RRT_ERR_INVALID, "object tables are NULL");
    if (!sc->light || !sc->camera) return fail(RRT_ERR_INVALID, "light/camera tables are NULL");
    if ((sc->jitter_x == nullptr) != (sc->jitter_y == nullptr)) return fail(RRT_ERR_INVALID, "jitter_x and jitter_y must both be set or both NULL");
    if (((uintptr_t)sc->w2o & 15) || (sc->w2o_scene_stride & 3)) return fail(RRT_ERR_INVALID, "w2o must be 16-byte aligned");
    if (sc->shader == RRT_SHADER_DEPTH && !(sc->max_depth != 0.0f)) return fail(RRT_ERR_INVALID, "max_depth must be non-zero");
    if ((uintptr_t)sc->obj_records & 15) return fail(RRT_ERR_INVALID, "obj_records must be 16-byte aligned");
    *rows_out = rows;
    return RRT_OK;
}

// Small scenes (few objects, few rays) take the one-ray-per-thread kernel.  The ray limit
// can be overridden for A/B measurements: RRT_SMALL_MAX_RAYS=<count> (0 disables the kernel).
bool use_small_kernel(const KParams& P) {
    const rrt_scene& sc = P.sc;
    const int S = sc.samples;
    if (sc.num_objects > kSmallMaxN || S > 32 || (S & (S - 1)) || (sc.flags & (RRT_FLAG_CULL | RRT_FLAG_NO_SMALL)))
        return false;
    static long long limit = -1;
    if (limit < 0) {
        const char* e = getenv("RRT_SMALL_MAX_RAYS");
        limit = e ? atoll(e) : kSmallDefaultMaxRays;
    }
    const long long rays_scene = (long long)P.rows * sc.n * S;
    if (rays_scene >= 0x7fffffffLL - kSmallThreads) return false;    // 32-bit ray index in the kernel
    return rays_scene * sc.num_scenes <= limit;
}

template <int MODE>
int launch(KParams& P, cudaStream_t st) {
    const rrt_scene& sc = P.sc;
    P.lin_step = sc.n > 1 ? 1.0 / (double)(sc.n - 1) : 0.0;
    P.inv_s = 1.0f / (float)sc.samples;
    P.inv_n = 1.0f / (float)sc.n;
    P.pow2 = ((sc.samples & (sc.samples - 1)) == 0) && ((sc.n & (sc.n - 1)) == 0);
    P.vec_ok = (sc.n % 4 == 0) && (((uintptr_t)P.image & 15) == 0) && (((uintptr_t)P.target & 15) == 0);
    const int S = sc.samples;
    if (use_small_kernel(P)) {
        const long long rays_scene = (long long)P.rows * sc.n * S;
        dim3 grid((unsigned)((rays_scene + kSmallThreads - 1) / kSmallThreads), sc.num_scenes);
        render_small_kernel<MODE><<<grid, kSmallThreads, 0, st>>>(P);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "small-scene kernel launch: %s", cudaGetErrorString(e));
        return RRT_OK;
    }
    if (MODE == MODE_FUSED && S > kRays)   // one thread must own all samples of its pixels (single sample chunk)
        return fail(RRT_ERR_UNSUPPORTED, "fused kernel supports samples <= 8 beyond the small-scene limits; use forward + backward");
    int pix;
    if (S == 1) pix = kRays; else if (S == 2) pix = kRays / 2; else if (S == 4) pix = kRays / 4; else pix = 1;
    // block height: keep the grid >= ~2 waves of 148 SMs when the image is small
    const int cols = (sc.n + 32 * pix - 1) / (32 * pix);
    int warps = kMaxWarps;
    while (warps > 1 && (long long)cols * ((P.rows + warps - 1) / warps) * sc.num_scenes < 2 * 148) warps >>= 1;
    dim3 grid(cols, (P.rows + warps - 1) / warps, sc.num_scenes), block(32 * warps);
    if (grid.y > 65535) return fail(RRT_ERR_UNSUPPORTED, "row slab too tall for one launch");
    const int staged = sc.num_objects < kObjChunk ? sc.num_objects : kObjChunk;
    size_t smem = (size_t)(staged > 0 ? staged : 1) * 64;
    void (*kern)(const KParams) = nullptr;
    if (S == 1) kern = render_kernel<kRays, 1, MODE>;
    else if (S == 2) kern = render_kernel<kRays / 2, 2, MODE>;
    else if (S == 4) kern = render_kernel<kRays / 4, 4, MODE>;
    else kern = render_kernel<1, kRays, MODE>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    kern<<<grid, block, smem, st>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "render kernel launch: %s", cudaGetErrorString(e));
    return RRT_OK;
}

int launch_finalize(const KParams& P, cudaStream_t st) {
    finalize_grads<<<P.sc.num_scenes, 128, 0, st>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "finalize launch: %s", cudaGetErrorString(e));
    return RRT_OK;
}

}  // namespace

extern "C" {

int rrt_version(void) { return RRT_VERSION; }

const char* rrt_last_error(void) { return g_err; }

int rrt_render_forward(const rrt_scene* scene, float* image, int32_t* hit_index, float* tmin, void* stream) {
    KParams P;
    memset(&P, 0, sizeof P);
    int rc = check_scene(scene, &P.rows);
    if (rc) return rc;
    if (!image && !hit_index && !tmin) return fail(RRT_ERR_INVALID, "no output buffer given");
    P.sc = *scene;
    P.image = image;
    P.hit_out = hit_index;
    P.tmin_out = tmin;
    return launch<MODE_FWD>(P, (cudaStream_t)stream);
}

int rrt_render_backward(const rrt_scene* scene, const float* dl_dimage, const int32_t* hit_index, float* grad, void* stream) {
    KParams P;
    memset(&P, 0, sizeof P);
    int rc = check_scene(scene, &P.rows);
    if (rc) return rc;
    if (!dl_dimage || !grad) return fail(RRT_ERR_INVALID, "dl_dimage and grad are required");
    P.sc = *scene;
    P.dl_dimage = dl_dimage;
    P.hit_in = hit_index;
    P.grad = grad;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(grad, 0, sizeof(float) * RRT_GRAD_SIZE(scene->num_objects) * scene->num_scenes, st);
    if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "memset grad: %s", cudaGetErrorString(e));
    rc = launch<MODE_BWD>(P, st);
    if (rc) return rc;
    return launch_finalize(P, st);
}

int rrt_render_fused_mse(const rrt_scene* scene, const float* target, const float* channel_weight, float* image,
                         int32_t* hit_index, double* loss, float* grad, void* stream) {
    KParams P;
    memset(&P, 0, sizeof P);
    int rc = check_scene(scene, &P.rows);
    if (rc) return rc;
    if (!target || !loss || !grad) return fail(RRT_ERR_INVALID, "target, loss and grad are required");
    P.sc = *scene;
    P.target = target;
    P.cw[0] = channel_weight ? channel_weight[0] : 1.f;
    P.cw[1] = channel_weight ? channel_weight[1] : 1.f;
    P.cw[2] = channel_weight ? channel_weight[2] : 1.f;
    P.image = image;
    P.hit_out = hit_index;
    P.loss = loss;
    P.grad = grad;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(grad, 0, sizeof(float) * RRT_GRAD_SIZE(scene->num_objects) * scene->num_scenes, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(loss, 0, sizeof(double) * scene->num_scenes, st);
    if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "memset grad/loss: %s", cudaGetErrorString(e));
    rc = launch<MODE_FUSED>(P, st);
    if (rc) return rc;
    return launch_finalize(P, st);
}

int rrt_build_records(const rrt_scene* scene, float* records, void* stream) {
    int rows = 0;
    int rc = check_scene(scene, &rows);
    if (rc) return rc;
    if (scene->num_objects == 0) return RRT_OK;
    if (!records || ((uintptr_t)records & 15)) return fail(RRT_ERR_INVALID, "records must be a 16-byte aligned device buffer");
    dim3 grid((scene->num_objects + kObjChunk - 1) / kObjChunk, scene->num_scenes);
    build_records_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*scene, records);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "build records launch: %s", cudaGetErrorString(e));
    return RRT_OK;
}

int rrt_primary_rays(int n, float* out, void* stream) {
    if (n <= 0 || n > 65535 || !out) return fail(RRT_ERR_INVALID, "rrt_primary_rays: bad arguments");
    primary_rays_kernel<<<dim3((n + 127) / 128, n), 128, 0, (cudaStream_t)stream>>>(n, n > 1 ? 1.0 / (double)(n - 1) : 0.0, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "primary rays launch: %s", cudaGetErrorString(e));
    return RRT_OK;
}

int rrt_chain_forward(const int32_t* ops, const int32_t* chain_begin, int num_chains, const float* values, float* out,
                      void* stream) {
    if (num_chains < 0) return fail(RRT_ERR_INVALID, "num_chains must be >= 0");
    if (num_chains == 0) return RRT_OK;
    if (!ops || !chain_begin || !values || !out) return fail(RRT_ERR_INVALID, "chain tables are NULL");
    chain_forward_kernel<<<(num_chains + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ops, chain_begin, num_chains, values, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "chain forward launch: %s", cudaGetErrorString(e));
    return RRT_OK;
}

int rrt_chain_backward(const int32_t* ops, const int32_t* chain_begin, int num_chains, const float* values,
                       const float* g_out, float* g_values, int num_values, void* stream) {
    if (num_chains < 0 || num_values < 0) return fail(RRT_ERR_INVALID, "negative size");
    if (!g_values && num_values > 0) return fail(RRT_ERR_INVALID, "g_values is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (num_values > 0) {
        cudaError_t e = cudaMemsetAsync(g_values, 0, sizeof(float) * (size_t)num_values, st);
        if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "memset g_values: %s", cudaGetErrorString(e));
    }
    if (num_chains == 0) return RRT_OK;
    if (!ops || !chain_begin || !values || !g_out) return fail(RRT_ERR_INVALID, "chain tables are NULL");
    chain_backward_kernel<<<(num_chains + 63) / 64, 64, 0, st>>>(ops, chain_begin, num_chains, values, g_out, g_values);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "chain backward launch: %s", cudaGetErrorString(e));
    return RRT_OK;
}

int rrt_small_step_mse(const rrt_scene* scene, const rrt_step* step, const float* target, const float* channel_weight,
                       float* image, void* stream) {
    KParams P;
    memset(&P, 0, sizeof P);
    int rc = check_scene(scene, &P.rows);
    if (rc) return rc;
    if (!step || !target) return fail(RRT_ERR_INVALID, "step and target are required");
    if (!step->ops || !step->chain_begin || !step->values || !step->grad || !step->g_values || !step->loss_acc ||
        !step->loss_out || !step->ticket)
        return fail(RRT_ERR_INVALID, "rrt_step has a NULL field");
    if (step->num_values <= 0 || step->param_begin < 0 || step->param_begin > step->num_values)
        return fail(RRT_ERR_INVALID, "rrt_step: bad parameter range");
    P.sc = *scene;
    if (scene->num_scenes != 1 || scene->num_objects < 1) return fail(RRT_ERR_UNSUPPORTED, "whole-step kernel: one scene, >= 1 shape");
    P.step = *step;
    P.target = target;
    P.cw[0] = channel_weight ? channel_weight[0] : 1.f;
    P.cw[1] = channel_weight ? channel_weight[1] : 1.f;
    P.cw[2] = channel_weight ? channel_weight[2] : 1.f;
    P.image = image;
    P.loss = step->loss_acc;
    P.grad = step->grad;
    if (!use_small_kernel(P))
        return fail(RRT_ERR_UNSUPPORTED, "whole-step kernel needs a small scene (<= 32 shapes, power-of-two samples)");
    const rrt_scene& sc = P.sc;
    P.lin_step = sc.n > 1 ? 1.0 / (double)(sc.n - 1) : 0.0;
    P.inv_s = 1.0f / (float)sc.samples;
    P.inv_n = 1.0f / (float)sc.n;
    P.pow2 = ((sc.samples & (sc.samples - 1)) == 0) && ((sc.n & (sc.n - 1)) == 0);
    const long long rays_scene = (long long)P.rows * sc.n * sc.samples;
    dim3 grid((unsigned)((rays_scene + kSmallThreads - 1) / kSmallThreads), 1);
    render_small_kernel<MODE_FUSED, true><<<grid, kSmallThreads, 0, (cudaStream_t)stream>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "whole-step kernel launch: %s", cudaGetErrorString(e));
    return RRT_OK;
}

size_t rrt_peer_buffer_bytes(int n, int nloss, int world) {
    if (n < 0 || nloss < 0 || world < 1) return 0;
    return (size_t)2 * (size_t)world * (size_t)(n + nloss) * sizeof(double);
}

size_t rrt_peer_signal_bytes(void) { return (size_t)(kPeerEpochOffset + kPeerCtas) * sizeof(unsigned); }

int rrt_peer_allreduce(const float* grad, const double* loss, int n, int nloss, void* const* peer_buf,
                       void* const* peer_sig, int rank, int world, double* out, void* stream) {
    if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world) return fail(RRT_ERR_INVALID, "bad rank/world (world <= 16)");
    if (n < 0 || nloss < 0 || n + nloss <= 0) return fail(RRT_ERR_INVALID, "nothing to reduce");
    if ((n > 0 && !grad) || (nloss > 0 && !loss) || !peer_buf || !peer_sig || !out)
        return fail(RRT_ERR_INVALID, "rrt_peer_allreduce: NULL argument");
    peer_allreduce_kernel<<<kPeerCtas, 256, 0, (cudaStream_t)stream>>>(grad, loss, n, nloss, peer_buf, peer_sig, rank, world, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "peer allreduce launch: %s", cudaGetErrorString(e));
    return RRT_OK;
}

int rrt_measure_fp32_peak(int mode, int iters, double* tflops, double* ms, void* stream) {
    if (!tflops || iters <= 0 || mode < 0 || mode > 2) return fail(RRT_ERR_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    float* out = nullptr;
    if (cudaMalloc(&out, 4) != cudaSuccess) return fail(RRT_ERR_CUDA, "cudaMalloc failed");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, st);
        if (mode == 0) fp32_peak_kernel<0><<<blocks, threads, 0, st>>>(out, iters, 0.5f);
        else if (mode == 1) fp32_peak_kernel<1><<<blocks, threads, 0, st>>>(out, iters, 0.5f);
        else fp32_peak_kernel<2><<<blocks, threads, 0, st>>>(out, iters, 0.5f);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float t = 0.f;
        cudaEventElapsedTime(&t, e0, e1);
        if (rep > 0 && t < best) best = t;
    }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "peak kernel: %s", cudaGetErrorString(e));
    // per thread per iteration: 8*16 FMA instructions, x2 lanes when packed, 2 flops each
    double flops = (double)blocks * threads * (double)iters * 8.0 * 16.0 * (mode >= 1 ? 2.0 : 1.0) * 2.0;
    *tflops = flops / ((double)best * 1e-3) / 1e12;
    if (ms) *ms = best;
    return RRT_OK;
}

}  // extern "C"
