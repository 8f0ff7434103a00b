// rrt_reduce.cuh -- part of rrt_kernels.cu (one translation unit; included inside its anonymous namespace).
// Warp / CTA gradient reductions and the conservative tile-culling test.
#pragma once

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
