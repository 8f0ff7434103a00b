// rrt_shade.cuh -- part of rrt_kernels.cu (one translation unit; included inside its anonymous namespace).
// Hard shadows, shading (forward value) and the closed-form reverse pass of one winning ray.
#pragma once

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

// Packed form of the same pass for chunks of spheres: the decider of 8 rays against one object
// as FFMA2/FMUL2 (bit-identical to shadow_test's scalar arithmetic, so it is an EXACT filter),
// folded into a running max; only (ray, object) pairs with dec > 0 -- the ray's own winner and
// real occluder candidates -- take the scalar routine, which makes the decision.
__device__ __forceinline__ u64 neg2(u64 v) { return v ^ 0x8000000080000000ULL; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

template <bool GENERAL>
__device__ __forceinline__ u64 pair_dec(const float4& q0, const float4& q1, const float4& q2, const float4& q3,
                                        u64 dx, u64 dy, u64 dz, u64 t2, float U0, float U1, float U2) {
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
    const u64 y0 = fma2(t2, ex, bc(q0.w)), y1 = fma2(t2, ey, bc(q1.x)), y2 = fma2(t2, ez, bc(q1.y));
    const u64 x = fma2(y2, bc(U2), fma2(y1, bc(U1), mul2(y0, bc(U0))));
    const u64 yy = fma2(y2, y2, fma2(y1, y1, mul2(y0, y0)));
    return add2(fma2(x, x, neg2(yy)), bc(1.0f));
}

__device__ __noinline__ unsigned shadow_chunk_packed(const float4* __restrict__ tab, int cnt, int kbase, const float* dw,
                                                     const float* tmin, const int* idx, const float* U, unsigned shadowed) {
    const u64* dp = reinterpret_cast<const u64*>(dw);
    u64 dx[kRays / 2], dy[kRays / 2], dz[kRays / 2], tp[kRays / 2];
    const float nan = __int_as_float(0x7fc00000);
#pragma unroll
    for (int p = 0; p < kRays / 2; p++) {
        dx[p] = dp[p]; dy[p] = dp[kRays / 2 + p]; dz[p] = dp[kRays + p];
        // rays without a (still lit) winner carry t = NaN: their decider is NaN, never > 0
        const float t0 = (idx[2 * p] >= 0 && !(shadowed >> (2 * p) & 1u)) ? tmin[2 * p] : nan;
        const float t1 = (idx[2 * p + 1] >= 0 && !(shadowed >> (2 * p + 1) & 1u)) ? tmin[2 * p + 1] : nan;
        tp[p] = pk(t0, t1);
    }
    const float U0 = U[0], U1 = U[1], U2 = U[2];
#pragma unroll 1
    for (int k = 0; k < cnt; k++) {
        const float4* rec = tab + 4 * k;
        const float4 q0 = rec[0], q1 = rec[1];
        const int flags = __float_as_int(q1.w);
        if (flags & 1) continue;                           // squares cast no shadow
        float dec[kRays];
        float gmax = 0.0f;
        if (flags & 2) {
            const float4 q2 = rec[2], q3 = rec[3];
#pragma unroll
            for (int p = 0; p < kRays / 2; p++) {
                upk(pair_dec<true>(q0, q1, q2, q3, dx[p], dy[p], dz[p], tp[p], U0, U1, U2), dec[2 * p], dec[2 * p + 1]);
                gmax = fmaxf(gmax, fmaxf(dec[2 * p], dec[2 * p + 1]));
            }
        } else {
#pragma unroll
            for (int p = 0; p < kRays / 2; p++) {
                upk(pair_dec<false>(q0, q1, q0, q0, dx[p], dy[p], dz[p], tp[p], U0, U1, U2), dec[2 * p], dec[2 * p + 1]);
                gmax = fmaxf(gmax, fmaxf(dec[2 * p], dec[2 * p + 1]));
            }
        }
        if (!(gmax > 0.0f)) continue;
#pragma unroll
        for (int r = 0; r < kRays; r++) {
            if (dec[r] > 0.0f && kbase + k != idx[r] &&
                shadow_test(rec, dw[r], dw[kRays + r], dw[2 * kRays + r], tmin[r], U)) {
                shadowed |= 1u << r;
                float lo, hi;
                upk(tp[r / 2], lo, hi);
                tp[r / 2] = (r & 1) ? pk(lo, nan) : pk(nan, hi);
            }
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
        // T.clip = switch(x < 0, 0, switch(x > 1, 1, x)): a NaN (negative base ** non-integer
        // shininess, shader.py:45) stays NaN like in the reference -- fminf/fmaxf would drop it
        rgb[c] = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    }
}

// ---------------------------------------------------------------- the fused entry points' cost of one pixel
// Default: squared error against the target, sum_c w_c (v_c - t_c)^2 (match_mirror.py:45).  With
// RRT_FLAG_LINEAR_COST the `target` buffer holds a WEIGHT image W and the cost is the linear functional
// sum_c w_c W_c v_c -- optimize_brightness.py:51, -image[90,85].sum() - image[50,90].sum(), is W = -1 at two
// pixels; pixels with W = 0 carry no upstream gradient, so their rays skip the reverse pass.
// `inv` = 1/S: d pixel / d sample.
__device__ __forceinline__ void pixel_cost(bool linear, const float* cw, float inv, float v0, float v1, float v2,
                                           const float* t, float& loss, float* gc) {
    if (linear) {
        loss += cw[0] * t[0] * v0 + cw[1] * t[1] * v1 + cw[2] * t[2] * v2;
        gc[0] = cw[0] * t[0] * inv;
        gc[1] = cw[1] * t[1] * inv;
        gc[2] = cw[2] * t[2] * inv;
    } else {
        const float d0 = v0 - t[0], d1 = v1 - t[1], d2 = v2 - t[2];
        loss += cw[0] * d0 * d0 + cw[1] * d1 * d1 + cw[2] * d2 * d2;
        gc[0] = 2.0f * cw[0] * d0 * inv;
        gc[1] = 2.0f * cw[1] * d1 * inv;
        gc[2] = 2.0f * cw[2] * d2 * inv;
    }
}

// ---------------------------------------------------------------- reverse pass, one winning ray
// Closed form of T.grad through the winner (masks constant).  og[19] receives
// [M = sum g_d' r_cam^T (9), g_b = sum g_o' (3), d/d(ka,kd,ks,sh,r,g,b)];
// gg[9] receives [d/d Lhat (3), d/d intensity (3), d/d look_at (3)].
// The chain M -> d/dA, d/d camera and Lhat -> L is applied by finalize_grads.
// GEOM_ONLY (RRT_FLAG_NO_MATERIAL_GRAD): only og[0..11] is produced -- the material, light and
// look_at sums are neither computed nor touched (og may then be a 12-float array).
// Mirror bounce (RRT_FLAG_MIRROR): `extra` = {dL/d(object-space normal)[3], dL/dt} arriving through the
// reflection (primary rays), `god` receives {g_o'[3], g_d'[3]}, `origin` is the world origin of a
// secondary ray (o'' = A P + b also depends on A: M += g_o' P^T; valid because the camera is the identity).
template <bool GEOM_ONLY = false, int NOG = 19>
__device__ __forceinline__ void backward_ray(int shader, float max_depth, const Obj& ob, const float* mat,
                                             const Globals& g, const HitRec& h, const ShadeRec& r, const float rc[3],
                                             const float gc[3], float (&og)[NOG], float* gg,
                                             const float* extra = nullptr, float* god = nullptr, const float* origin = nullptr) {
    static_assert(NOG == (GEOM_ONLY ? 12 : 19), "backward_ray: accumulator size");
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
            if (!GEOM_ONLY) {
                og[(GEOM_ONLY ? 0 : 16) + c] += gc[c] * r.ph * g.I[c];
                gg[3 + c] += gc[c] * r.ph * mat[4 + c];
            }
        }
        if (!GEOM_ONLY) {
            og[GEOM_ONLY ? 0 : 12] += g_ph;
            og[GEOM_ONLY ? 0 : 13] += g_ph * r.ndl;
        }
        float g_ndl = g_ph * mat[1];
        float g_n[3] = {0.f, 0.f, 0.f}, g_Lh[3] = {0.f, 0.f, 0.f};
        if (shader == RRT_SHADER_PHONG) {
            if (!GEOM_ONLY) {
                og[GEOM_ONLY ? 0 : 14] += g_ph * r.pw;
                if (r.rv > 0.0f) og[GEOM_ONLY ? 0 : 15] += g_ph * mat[2] * r.pw * __logf(r.rv);
            }
            float dpw = (r.rv != 0.0f) ? __fdividef(r.pw, r.rv) : pow_shininess(r.rv, mat[3] - 1.0f);  // rv^(sh-1)
            float g_rv = g_ph * mat[2] * mat[3] * dpw;
            float g_rm[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                g_rm[c] = g_rv * g.look[c];
                if (!GEOM_ONLY) gg[6 + c] += g_rv * r.rm[c];
            }
            g_ndl += 2.0f * (g_rm[0] * r.nrm[0] + g_rm[1] * r.nrm[1] + g_rm[2] * r.nrm[2]);
#pragma unroll
            for (int c = 0; c < 3; c++) { g_n[c] += 2.0f * r.ndl * g_rm[c]; g_Lh[c] += g_rm[c]; }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) { g_n[c] -= g_ndl * g.Lh[c]; g_Lh[c] -= g_ndl * r.nrm[c]; }
        if (!GEOM_ONLY) {
#pragma unroll
            for (int c = 0; c < 3; c++) gg[c] += g_Lh[c];
        }
        if (extra) {
#pragma unroll
            for (int c = 0; c < 3; c++) g_n[c] += extra[c];
            g_t += extra[3];
        }
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
        for (int c = 0; c < 3; c++) og[rr * 3 + c] += g_d[rr] * rc[c] + (origin ? g_o[rr] * origin[c] : 0.f);
        og[9 + rr] += g_o[rr];
    }
    if (god) {
#pragma unroll
        for (int c = 0; c < 3; c++) { god[c] = g_o[c]; god[3 + c] = g_d[c]; }
    }
}

// ---------------------------------------------------------------- mirror bounce (RRT_FLAG_MIRROR)
// Semantics in include/rrt_b200.h; canonical float32 order shared with orc_bounce_geom /
// orc_secondary in oracle/oracle_c.c (this defines which object a reflected ray sees), shading and
// the reverse pass at float32 tolerance.  An opt-in extension outside the roofline-accountable
// path: scalar, per ray, out of line.
struct Bounce {
    float P[3], r[3], nw[3], no[3], mn, dn;
};

__device__ __forceinline__ void bounce_geom(const Obj& ob, const HitRec& h, float dwx, float dwy, float dwz, Bounce& b) {
    if (!(ob.flags & 1)) {
        const float p0 = __fmaf_rn(h.t, h.d[0], ob.o[0]), p1 = __fmaf_rn(h.t, h.d[1], ob.o[1]), p2 = __fmaf_rn(h.t, h.d[2], ob.o[2]);
        const float pn = __fsqrt_rn(dot3_canon(p0, p1, p2, p0, p1, p2));
        b.no[0] = __fdiv_rn(p0, pn); b.no[1] = __fdiv_rn(p1, pn); b.no[2] = __fdiv_rn(p2, pn);
    } else {
        b.no[0] = b.no[1] = 0.f;
        b.no[2] = (ob.o[2] > 0.0f) ? 1.0f : -1.0f;
    }
    float m[3];
#pragma unroll
    for (int i = 0; i < 3; i++) m[i] = dot3_canon(ob.a[i], ob.a[3 + i], ob.a[6 + i], b.no[0], b.no[1], b.no[2]);   // A^T n_o
    b.mn = __fsqrt_rn(dot3_canon(m[0], m[1], m[2], m[0], m[1], m[2]));
#pragma unroll
    for (int c = 0; c < 3; c++) b.nw[c] = __fdiv_rn(m[c], b.mn);
    b.dn = dot3_canon(dwx, dwy, dwz, b.nw[0], b.nw[1], b.nw[2]);
    const float k2 = __fmul_rn(-2.0f, b.dn);
    b.r[0] = __fmaf_rn(k2, b.nw[0], dwx); b.r[1] = __fmaf_rn(k2, b.nw[1], dwy); b.r[2] = __fmaf_rn(k2, b.nw[2], dwz);
    b.P[0] = __fmul_rn(h.t, dwx); b.P[1] = __fmul_rn(h.t, dwy); b.P[2] = __fmul_rn(h.t, dwz);
}

// nearest OTHER object along the reflected ray (list order, strict '<', t2 > 0); fills ob2 (object j2
// re-based at origin P) and its hit record
__device__ __noinline__ int mirror_secondary(const float* __restrict__ w2o, const int* __restrict__ obj_type, int N, int k,
                                             const Bounce& b, Obj& ob2, HitRec& h2) {
    float tmin = __int_as_float(0x7f800000);
    int j2 = -1;
#pragma unroll 1
    for (int j = 0; j < N; j++) {
        if (j == k) continue;
        Obj t;
        make_obj(w2o + (size_t)j * RRT_W2O_STRIDE, obj_type[j], b.P, t);
        HitRec hh;
        const float t2 = obj_test<false>(t, b.r[0], b.r[1], b.r[2], hh);
        if (t2 > 0.0f && t2 < tmin) { tmin = t2; j2 = j; ob2 = t; h2 = hh; }
    }
    return j2;
}

// forward: what the reflected ray of a winning primary ray sees (0 if nothing)
__device__ __noinline__ void mirror_shade(int shader, float max_depth, const float* __restrict__ w2o,
                                          const float* __restrict__ mats, const int* __restrict__ obj_type, int N,
                                          const Globals& g, int k, const Obj& ob, const HitRec& h,
                                          float dwx, float dwy, float dwz, float rgb2[3]) {
    Bounce b;
    bounce_geom(ob, h, dwx, dwy, dwz, b);
    Obj ob2;
    HitRec h2;
    const int j2 = mirror_secondary(w2o, obj_type, N, k, b, ob2, h2);
    rgb2[0] = rgb2[1] = rgb2[2] = 0.f;
    if (j2 >= 0) {
        float m7[7];
#pragma unroll
        for (int q = 0; q < 7; q++) m7[q] = __ldg(mats + (size_t)j2 * RRT_MAT_STRIDE + q);
        ShadeRec sr2;
        shade(shader, max_depth, ob2, m7, g, h2, sr2, rgb2);
    }
}

// reverse: the secondary object's sums og2 (for object j2 = return value, -1: none), and what flows
// back into the primary object through the reflection: extra = {dL/d n_o [3], dL/dt}, dA[9] (direct
// term of n_w = A^T n_o / |.|; added to the primary's M sums, which ARE d/dA for the identity camera)
template <bool GEOM_ONLY, int NOG>
__device__ __noinline__ int mirror_backward(int shader, float max_depth, const float* __restrict__ w2o,
                                            const float* __restrict__ mats, const int* __restrict__ obj_type, int N,
                                            const Globals& g, int k, const Obj& ob, const HitRec& h,
                                            float dwx, float dwy, float dwz, const float gc2[3],
                                            float (&og2)[NOG], float* gg, float extra[4], float dA[9]) {
    Bounce b;
    bounce_geom(ob, h, dwx, dwy, dwz, b);
    Obj ob2;
    HitRec h2;
    const int j2 = mirror_secondary(w2o, obj_type, N, k, b, ob2, h2);
#pragma unroll
    for (int v = 0; v < NOG; v++) og2[v] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; q++) extra[q] = 0.f;
#pragma unroll
    for (int q = 0; q < 9; q++) dA[q] = 0.f;
    if (j2 < 0) return -1;
    float m7[7];
#pragma unroll
    for (int q = 0; q < 7; q++) m7[q] = __ldg(mats + (size_t)j2 * RRT_MAT_STRIDE + q);
    ShadeRec sr2;
    float rgb2[3], god[6];
    shade(shader, max_depth, ob2, m7, g, h2, sr2, rgb2);
    backward_ray<GEOM_ONLY, NOG>(shader, max_depth, ob2, m7, g, h2, sr2, b.r, gc2, og2, gg, nullptr, god, b.P);
    float GP[3], Gr[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {                      // A_j2^T g_o'', A_j2^T g_d''
        GP[c] = ob2.a[c] * god[0] + ob2.a[3 + c] * god[1] + ob2.a[6 + c] * god[2];
        Gr[c] = ob2.a[c] * god[3] + ob2.a[3 + c] * god[4] + ob2.a[6 + c] * god[5];
    }
    const float dw[3] = {dwx, dwy, dwz};
    extra[3] = GP[0] * dwx + GP[1] * dwy + GP[2] * dwz;                        // P = t d
    const float grn = Gr[0] * b.nw[0] + Gr[1] * b.nw[1] + Gr[2] * b.nw[2];
    float g_nw[3], dot = 0.f;
#pragma unroll
    for (int c = 0; c < 3; c++) { g_nw[c] = -2.0f * (b.dn * Gr[c] + grn * dw[c]); dot += b.nw[c] * g_nw[c]; }   // r = d - 2 (d.n) n
    float g_m[3];
    const float imn = 1.0f / b.mn;
#pragma unroll
    for (int c = 0; c < 3; c++) g_m[c] = (g_nw[c] - b.nw[c] * dot) * imn;      // n_w = m/|m|
#pragma unroll
    for (int rr = 0; rr < 3; rr++)
#pragma unroll
        for (int i = 0; i < 3; i++) dA[rr * 3 + i] = b.no[rr] * g_m[i];         // m = A^T n_o
    if (!(ob.flags & 1)) {
#pragma unroll
        for (int rr = 0; rr < 3; rr++) extra[rr] = ob.a[rr * 3] * g_m[0] + ob.a[rr * 3 + 1] * g_m[1] + ob.a[rr * 3 + 2] * g_m[2];
    }
    return j2;
}
