// rrt_reduce.cuh -- part of rrt_kernels.cu (one translation unit; included inside its anonymous namespace).
// Warp / CTA gradient reductions and the conservative tile-culling test.
#pragma once

// ---------------------------------------------------------------- gradient reduction
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- RRT_FLAG_DETERMINISTIC
// 128-bit fixed point: value = hi * 2^-20 + lo * 2^-60, two int64 limbs combined with integer
// atomics (associative => the sum does not depend on the order in which warps / CTAs arrive).
// The split is exact for |x| < 2^43 down to 2^-60; smaller bits are rounded once, per contribution.
__device__ __forceinline__ void det_add(long long* limbs, double x) {
    const long long hi = __double2ll_rn(x * 1048576.0);                              // 2^20
    const double r = x - (double)hi * 9.5367431640625e-07;                          // exact
    const long long lo = __double2ll_rn(r * 1152921504606846976.0);                 // 2^60
    atomicAdd(reinterpret_cast<unsigned long long*>(limbs), (unsigned long long)hi);
    atomicAdd(reinterpret_cast<unsigned long long*>(limbs) + 1, (unsigned long long)lo);
}
__device__ __forceinline__ double det_value(const long long* limbs) {
    const long long hi = __ldcg(limbs), lo = __ldcg(limbs + 1);
    return (double)hi * 9.5367431640625e-07 + (double)lo * 8.673617379884035e-19;    // 2^-20, 2^-60
}
__device__ __forceinline__ long long* det_scene(const rrt_scene& sc, int scene) {
    return reinterpret_cast<long long*>(sc.det_workspace) + (size_t)scene * (RRT_GRAD_SIZE(sc.num_objects) + 1) * 2;
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

// Sum of NV per-lane values over the warp as a TRANSPOSED butterfly: at every stage a lane
// keeps one half of its values and trades the other half with its partner, so the whole
// reduction is ceil(NV/2) + ceil(NV/4) + ... shuffles (21 for NV = 19, 13 for NV = 12, instead
// of NV x 5) and ends with lane l holding the complete sum of value index `v` (returned; -1
// for the lanes that hold padding).  Fixed lane pairing => the result does not depend on
// scheduling (it is one of the deterministic stages of RRT_FLAG_DETERMINISTIC).
__device__ __forceinline__ float xchg_add(float keep, float send, int offset) {
    return keep + __shfl_xor_sync(0xffffffffu, send, offset);
}
template <int NI, int NO>
__device__ __forceinline__ void reduce_stage(const float (&in)[NI], float (&out)[NO], bool up, int offset) {
#pragma unroll
    for (int j = 0; j < NO; j++) {
        const float lo = in[j];
        const float hi = (NO + j < NI) ? in[(NO + j < NI) ? NO + j : 0] : 0.f;
        out[j] = xchg_add(up ? hi : lo, up ? lo : hi, offset);
    }
}
template <int NV>
__device__ __forceinline__ float warp_reduce_n(const float (&a)[NV], bool mine, int lane, int& v) {
    constexpr int n1 = (NV + 1) / 2, n2 = (n1 + 1) / 2, n3 = (n2 + 1) / 2, n4 = (n3 + 1) / 2;
    static_assert(n4 <= 2, "warp_reduce_n: at most 32 values");
    float m[NV], b[n1], c[n2], d[n3], e[n4], f[1];
#pragma unroll
    for (int j = 0; j < NV; j++) m[j] = mine ? a[j] : 0.f;
    reduce_stage<NV, n1>(m, b, lane & 16, 16);
    reduce_stage<n1, n2>(b, c, lane & 8, 8);
    reduce_stage<n2, n3>(c, d, lane & 4, 4);
    reduce_stage<n3, n4>(d, e, lane & 2, 2);
    reduce_stage<n4, 1>(e, f, lane & 1, 1);
    // which value this lane ended up with: at every stage the upper partner kept the upper half
    const int j1 = lane & 1;
    const int j2 = ((lane >> 1) & 1) * n4 + j1;
    const int j3 = ((lane >> 2) & 1) * n3 + j2;
    const int j4 = ((lane >> 3) & 1) * n2 + j3;
    const int idx = ((lane >> 4) & 1) * n1 + j4;
    v = (j1 < n4 && j2 < n3 && j3 < n2 && j4 < n1 && idx < NV) ? idx : -1;
    return f[0];
}
__device__ __forceinline__ float warp_reduce19(const float (&a)[19], bool mine, int lane, int& v) {
    return warp_reduce_n<19>(a, mine, lane, v);
}

// All 32 lanes call this together.  Each lane holds (key, acc[19]); lanes with the
// same key are summed (transposed butterfly) and 19 lanes add one sum each to the CTA slot.
__device__ __forceinline__ void warp_flush(int key, float (&acc)[19], int* slot_key, float* slots, float* gobj, int lane,
                                           long long* det_ws = nullptr) {
    unsigned active = __ballot_sync(0xffffffffu, key >= 0);
    while (active) {
        int leader = __ffs(active) - 1;
        int k = __shfl_sync(0xffffffffu, key, leader);
        bool mine = (key == k);
        active &= ~__ballot_sync(0xffffffffu, mine);
        int v;
        if (det_ws) {                 // deterministic: the warp's sums go straight to the fixed-point workspace
            const float x = warp_reduce19(acc, mine, lane, v);
            if (v >= 0 && x != 0.f) det_add(det_ws + ((size_t)k * RRT_OBJ_GRAD_STRIDE + v) * 2, (double)x);
            continue;
        }
        int slot = 0;
        if (lane == 0) slot = slot_for(slot_key, k);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        const float x = warp_reduce19(acc, mine, lane, v);
        if (v >= 0 && x != 0.f) {
            if (slot >= 0) atomicAdd(&slots[slot * kSlotStride + v], x);
            else atomicAdd(&gobj[(size_t)k * RRT_OBJ_GRAD_STRIDE + v], x);
        }
    }
#pragma unroll
    for (int v = 0; v < 19; v++) acc[v] = 0.f;
}

// ---------------------------------------------------------------- gradient finalisation of one scene
// Converts the raw per-object sums [M = sum g_d' r_cam^T (9), g_b (3)] into d/d w2o and folds the
// camera and light chains (see backward_ray).  Called by all `nthreads` threads of ONE CTA after
// every contribution to the scene has landed (separate launch, or last CTA + __threadfence);
// reads through L2 (__ldcg) because the sums were produced by other CTAs' atomics.
__device__ __forceinline__ void finalize_scene(const KParams& P, int scene, int tid, int nthreads, float* camg /* shared [48] */) {
    const rrt_scene& sc = P.sc;
    const int N = sc.num_objects;
    float* gobj = P.grad + (size_t)scene * RRT_GRAD_SIZE(N);
    float* gglobal = gobj + (size_t)N * RRT_OBJ_GRAD_STRIDE;
    const float* cam = sc.camera + (size_t)scene * sc.camera_scene_stride;
    const float* w2o = sc.w2o + (size_t)scene * sc.w2o_scene_stride;
    const bool geom_only = (sc.flags & RRT_FLAG_NO_MATERIAL_GRAD) != 0;
    if ((sc.flags & RRT_FLAG_DETERMINISTIC) && sc.det_workspace) {
        // fixed-point sums -> the raw float sums the rest of this routine expects (and the loss)
        const long long* ws = det_scene(sc, scene);
        const int G = (int)RRT_GRAD_SIZE(N);
        for (int i = tid; i < G; i += nthreads) gobj[i] = (float)det_value(ws + 2 * (size_t)i);
        if (tid == 0 && P.loss) P.loss[scene] = det_value(ws + 2 * (size_t)G);
        __threadfence();
        __syncthreads();
    }
    __syncthreads();                 // (callers reuse camg)
    float C[9], ct[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        C[r * 3] = cam[r * 4]; C[r * 3 + 1] = cam[r * 4 + 1]; C[r * 3 + 2] = cam[r * 4 + 2];
        ct[r] = cam[r * 4 + 3];
    }
    float cg[12];
#pragma unroll
    for (int q = 0; q < 12; q++) cg[q] = 0.f;
    for (int k = tid; k < N; k += nthreads) {
        float* og = gobj + (size_t)k * RRT_OBJ_GRAD_STRIDE;
        float M[9], gb[3], A[9];
#pragma unroll
        for (int q = 0; q < 9; q++) M[q] = __ldcg(og + q);
#pragma unroll
        for (int q = 0; q < 3; q++) gb[q] = __ldcg(og + 9 + q);
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
        if (geom_only) {                 // RRT_FLAG_NO_MATERIAL_GRAD: material entries are defined to be zero
#pragma unroll
            for (int q = 12; q < 19; q++) og[q] = 0.f;
        }
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
    // camera sums: warp butterfly, then the (<= 4) warps' partials in warp order -- a fixed order, so the
    // finalisation never adds run-to-run noise (RRT_FLAG_DETERMINISTIC relies on it)
    if (sc.camera_grad) {
#pragma unroll
        for (int q = 0; q < 12; q++) {
            const float x = warp_sum(cg[q]);
            if ((tid & 31) == 0) camg[(tid >> 5) * 12 + q] = x;
        }
    }
    __syncthreads();
    if (tid < 12) {
        float x = 0.f;
        if (sc.camera_grad)
            for (int w = 0; w < (nthreads + 31) / 32; w++) x += camg[w * 12 + tid];
        gglobal[6 + tid] = x;
    }
    if (tid == 0) {
        if (geom_only) {
#pragma unroll
            for (int q = 0; q < 6; q++) gglobal[q] = 0.f;
            gglobal[18] = gglobal[19] = gglobal[20] = 0.f;
        } else {
            // Lhat = L/|L|  =>  g_L = (g_Lhat - Lhat (Lhat . g_Lhat)) / |L|    scene.py:83-86
            const float* li = sc.light + (size_t)scene * sc.light_scene_stride;
            float L0 = li[0], L1 = li[1], L2 = li[2];
            float ln = sqrtf(L0 * L0 + L1 * L1 + L2 * L2);
            float h0 = L0 / ln, h1 = L1 / ln, h2 = L2 / ln;
            float g0 = __ldcg(gglobal), g1 = __ldcg(gglobal + 1), g2 = __ldcg(gglobal + 2);
            float dot = h0 * g0 + h1 * g1 + h2 * g2;
            gglobal[0] = (g0 - h0 * dot) / ln;
            gglobal[1] = (g1 - h1 * dot) / ln;
            gglobal[2] = (g2 - h2 * dot) / ln;
        }
    }
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
