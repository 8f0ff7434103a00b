// rrt_sweep.cuh -- part of rrt_kernels.cu (one translation unit; included inside its anonymous namespace).
// The packed nearest-hit sweep (hot loop), its rare path, and TMA staging of prebuilt records.
#pragma once

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

// same arithmetic, also returning vn and pd (the rare path finishes t from them)
template <bool GENERAL>
__device__ __forceinline__ u64 pair_det_full(const float4& q0, const float4& q1, const float4& q2, const float4& q3,
                                             u64 dx, u64 dy, u64 dz, u64& vn, u64& pd) {
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
    vn = fma2(ez, ez, fma2(ey, ey, mul2(ex, ex)));
    pd = fma2(ez, bc(q1.y), fma2(ey, bc(q1.x), mul2(ex, bc(q0.w))));
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
        const float4 q0 = rec[0], q1 = rec[1];
        const int flags = __float_as_int(q1.w);
        if (flags & 1) {                                   // square: the scalar routine for every ray
            Obj ob;
            load_rec(rec, ob);
#pragma unroll 1
            for (int r = 0; r < kRays; r++) {
                HitRec h;
                const float t = obj_test(ob, dw[r], dw[kRays + r], dw[2 * kRays + r], h);
                if (t < tmin[r]) { tmin[r] = t; idx[r] = kbase + k0 + j; }
            }
            continue;
        }
        // sphere: the packed pass yields det, vn and pd with the bits of the scalar routine
        // (IEEE fma per lane; the diagonal form differs only in the sign of a zero, which neither
        // vn, pd nor -pd - sqrt(det) can see), so a hit only needs the sqrt and the divide
        float det[kRays], vn[kRays], pd[kRays];
        if (flags & 2) {
            const float4 q2 = rec[2], q3 = rec[3];
#pragma unroll
            for (int p = 0; p < kRays / 2; p++) {
                u64 v2, p2;
                upk(pair_det_full<true>(q0, q1, q2, q3, dx[p], dy[p], dz[p], v2, p2), det[2 * p], det[2 * p + 1]);
                upk(v2, vn[2 * p], vn[2 * p + 1]);
                upk(p2, pd[2 * p], pd[2 * p + 1]);
            }
        } else {
#pragma unroll
            for (int p = 0; p < kRays / 2; p++) {
                u64 v2, p2;
                upk(pair_det_full<false>(q0, q1, q0, q0, dx[p], dy[p], dz[p], v2, p2), det[2 * p], det[2 * p + 1]);
                upk(v2, vn[2 * p], vn[2 * p + 1]);
                upk(p2, pd[2 * p], pd[2 * p + 1]);
            }
        }
#pragma unroll
        for (int r = 0; r < kRays; r++) {
            if (det[r] > 0.0f) {                           // shape.py:121-125, first root
                const float t = __fdiv_rn(__fsub_rn(-pd[r], __fsqrt_rn(det[r])), vn[r]);
                if (t < tmin[r]) { tmin[r] = t; idx[r] = kbase + k0 + j; }
            }
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
__device__ __noinline__ void stage_bytes_tma(float4* smem_tab, const float* src, unsigned bytes, unsigned long long* bar,
                                             const unsigned* phase_s, int tid) {
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(bar);
    const unsigned phase = *phase_s;
    if (tid == 0) tma_bulk_load((uint32_t)__cvta_generic_to_shared(smem_tab), src, bytes, mbar);
    mbar_wait(mbar, phase);
}
__device__ __forceinline__ void stage_records_tma(float4* smem_tab, const float* src, int cnt, unsigned long long* bar,
                                                  const unsigned* phase_s, int* chunk_class, int tid) {
    stage_bytes_tma(smem_tab, src, (unsigned)cnt * 64u, bar, phase_s, tid);
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

// ---------------------------------------------------------------- conservative pre-filter sweep
// The default hot loop when a prebuilt table is available (RRT_FLAG_CANONICAL_SWEEP switches it
// off).  Per (ray, sphere) pair ONE float32 quadratic form in (u, v) = (d_x/d_z, d_y/d_z), in
// Horner form
//     F = ((c00 u + c02) u + c22) + v ((c01 u + c12) + c11 v)          5 fused multiply-adds,
// i.e. 20 FFMA2 per object for the thread's 8 rays with only u and v (16 registers) held per
// thread -- whose sign conservatively bounds the sign of the canonical discriminant (quadric_row
// in rrt_aux_kernels.cuh has the error budget).  Groups of kQGroup objects are branch-free (6
// floats = 1.5 LDS.128 per object, coefficients as scalar-broadcast .F32 operands in all three
// operand positions, 4 FMNMX3 per object, one vote per group); a group with any F > 0 goes to
// rare_group, which evaluates the CANONICAL arithmetic from the full records (read from the
// global table: the shared-memory chunk holds only the 24-byte pre-filter rows) and alone
// decides hits.  So results are bit-identical to the canonical sweep.
#ifndef RRT_QGROUP
#define RRT_QGROUP 8   // measured on C5 (final round-2 kernel, 4 CTAs/SM): groups of 8 objects 15.28 ms, of 4 15.41 ms
#endif
constexpr int kQGroup = RRT_QGROUP;      // objects per branch; the table is padded to a multiple of it... of 4 (see below)
static_assert(kQGroup == 4 || kQGroup == 8, "pre-filter rows are padded to multiples of 4 objects; groups of 4 or 8");

struct RayQ {
    u64 u[kRays / 2], v[kRays / 2];
};

// row = (c00, c02, c22, c01, c12, c11)
__device__ __forceinline__ float quad_obj_max(const float* g, const RayQ& rq, float gmax) {
#pragma unroll
    for (int p = 0; p < kRays / 2; p++) {
        u64 t1 = fma2(bc(g[0]), rq.u[p], bc(g[1]));
        u64 t2 = fma2(bc(g[3]), rq.u[p], bc(g[4]));
        t1 = fma2(t1, rq.u[p], bc(g[2]));
        t2 = fma2(bc(g[5]), rq.v[p], t2);
        const u64 f = fma2(t2, rq.v[p], t1);
        float lo, hi;
        upk(f, lo, hi);
        gmax = fmaxf(gmax, fmaxf(lo, hi));        // NaN (padding rays) is dropped: never a candidate
    }
    return gmax;
}

// `quad`: staged pre-filter rows of `count_pad` objects (a multiple of 4; padding rows never pass);
// `rec_g`: the same chunk's full records in the global table; `count`: real objects in the chunk.
__device__ __forceinline__ void sweep_quadric(const float4* __restrict__ quad, int count_pad, int count,
                                              const float4* __restrict__ rec_g, int kbase, const RayQ& rq,
                                              const float* dw, float* tmin, int* idx) {
    uint32_t q0 = (uint32_t)__cvta_generic_to_shared(quad);
    const int full = count_pad - count_pad % kQGroup;
    uint32_t q_end = q0 + 24u * (uint32_t)full;
    asm volatile("mov.u32 %0, %0;" : "+r"(q0));
    asm volatile("mov.u32 %0, %0;" : "+r"(q_end));
#pragma unroll 1
    for (uint32_t q = q0; q != q_end; q += 24 * kQGroup) {
        float g[6 * kQGroup];
#pragma unroll
        for (int i = 0; i < 6 * kQGroup / 4; i++) {
            const float4 x = lds128(q + 16 * i);
            g[4 * i] = x.x; g[4 * i + 1] = x.y; g[4 * i + 2] = x.z; g[4 * i + 3] = x.w;
        }
        float gmax = 0.0f;
#pragma unroll
        for (int j = 0; j < kQGroup; j++) gmax = quad_obj_max(g + 6 * j, rq, gmax);
        if (__builtin_expect(__any_sync(0xffffffffu, gmax > 0.0f), 0)) {
            const int k = (int)((q - q0) / 24u);
            rare_group(rec_g, k, min(kQGroup, count - k), kbase, dw, tmin, idx);
        }
    }
    if (full < count) rare_group(rec_g, full, count - full, kbase, dw, tmin, idx);   // kQGroup == 8: a last half group
}
