// rrt_objects.cuh -- part of rrt_kernels.cu (one translation unit; included inside its anonymous namespace).
// Object records (the per-object constants of the hit test) and the canonical scalar ray-object test.
#pragma once

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
    int cam_identity;  // camera rotation part is exactly I (small-scene kernel)
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

// The hit record of a ray whose winner AND ray parameter t are already known from the sweep:
// same d', vn, pd, det as obj_test (same operations), but no second sqrt / divide.
template <bool NEED_QUADRATIC>
__device__ __forceinline__ void hit_record(const Obj& ob, float dwx, float dwy, float dwz, float t, HitRec& h) {
    if (!(ob.flags & 2)) {
        h.d[0] = __fmul_rn(ob.a[0], dwx);
        h.d[1] = __fmul_rn(ob.a[4], dwy);
        h.d[2] = __fmul_rn(ob.a[8], dwz);
    } else {
        h.d[0] = dot3_canon(ob.a[0], ob.a[1], ob.a[2], dwx, dwy, dwz);
        h.d[1] = dot3_canon(ob.a[3], ob.a[4], ob.a[5], dwx, dwy, dwz);
        h.d[2] = dot3_canon(ob.a[6], ob.a[7], ob.a[8], dwx, dwy, dwz);
    }
    h.t = t;
    h.vn = h.pd = h.det = 0.f;
    if (NEED_QUADRATIC && !(ob.flags & 1)) {
        h.vn = dot3_canon(h.d[0], h.d[1], h.d[2], h.d[0], h.d[1], h.d[2]);
        h.pd = dot3_canon(h.d[0], h.d[1], h.d[2], ob.o[0], ob.o[1], ob.o[2]);
        h.det = __fmaf_rn(h.pd, h.pd, __fmul_rn(h.vn, ob.ncc));
    }
}
