/*
 * oracle_c.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, loaded by or called
 * from the product path; only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs use it (through oracle/oracle_c.py).
 *
 * Per-pixel CPU restatement (plain C, OpenMP over rows) of the hot path of
 * lebek/reversible-raytracer, in the CANONICAL EVALUATION ORDER that defines the
 * bit-exact hit masks (DESIGN.md "Canonical order").  The reference itself is
 * Python 2 + Theano and cannot run here; this file restates, per ray, what the
 * reference's dense graph computes:
 *     Camera.make_rays          scene.py:61-75  (orbit variant: orbit_experiments/scene.py:55-80)
 *     Transform.__call__        transform.py:40-47   (spatial transpose -> `transpose` flag)
 *     Sphere.distance/normals   shape.py:78-83, 109-138
 *     Square._hit/normals       shape.py:25-69
 *     PhongShader.shade         shader.py:28-53 (orbit: specular dropped, orbit shader.py:45,48)
 *     DepthMapShader.shade      shader.py:14-20
 *     Scene.build               scene.py:18-52  (strict '<' nearest hit, list order, mean over S)
 *     T.grad(loss, params)      optimize.py:25,73  (closed form, masks constant)
 * It is pinned against oracle_numpy.py (dense, reference-structured), which is
 * pinned against the reference's golden renders (tests/test_oracle_*.py).
 *
 * Precision: every mask-determining quantity (rays, d', vn, pd, det, t, compares)
 * is float32 with explicit fmaf exactly where the CUDA kernels use FMA; compile
 * with -ffp-contract=off.  Ray generation is non-contracted float64 like NumPy.
 * Shading and the reverse pass are evaluated in DOUBLE on the float32 hit record,
 * and gradients are accumulated in double, so that comparing the float32 kernels
 * against this file measures the kernels' arithmetic, not this file's.
 *
 * Uses the product's descriptor struct (include/rrt_b200.h) with HOST pointers.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/rrt_b200.h"

#define FMAF(a, b, c) __builtin_fmaf((a), (b), (c))

typedef struct {
    float A[9], b[3]; /* w2o rows 0..2: [A|b]                          */
    float o[3];       /* o' = A.c + b   (transform.py:44)              */
    float cc;         /* o'.o' - 1      (shape.py:79,82)               */
    int type;
    float ka, kd, ks, sh, col[3];
    int has64;        /* secondary (mirror) objects in the f64-record validation mode: origin kept in double */
    double o64[3];
} orc_obj;

typedef struct {
    float C[9], ct[3]; /* camera.o2w rows 0..2 */
    float look[3];
    float L[3], I[3];
    double Lh[3], Ln; /* normalised light direction, |L| (scene.py:83-86) */
    float U[3];       /* -Lhat in float32, canonical order (shadow test only) */
} orc_glob;

/* ---- jitter RNG shared (by specification) with the CUDA kernels ------------- */
static inline float orc_rng(uint64_t seed, uint32_t scene, uint32_t pix, uint32_t s, uint32_t axis) {
    /* counter-based 32-bit hash (lowbias32 finaliser) -> 24-bit uniform in [0,1) */
    uint32_t x = (pix * 0x9E3779B1u) ^ (scene * 0x85EBCA77u) ^ ((s * 2u + axis) * 0xC2B2AE3Du) ^
                 (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x27D4EB2Fu);
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return (float)(x >> 8) * 5.9604644775390625e-08f; /* 2^-24 */
}

/* np.linspace(start, stop, n)[i] : fl(fl(i*step)+start), last element = stop */
static inline double orc_lin(int i, int n, double start, double stop) {
    if (n == 1) return start;
    if (i == n - 1) return stop;
    double step = (stop - start) / (double)(n - 1);
    double m = (double)i * step; /* -ffp-contract=off: two roundings like NumPy */
    return m + start;
}

/* Camera.make_rays, scene.py:66-74.  (i,j) index the reference's ray array. */
static inline void orc_gen_ray(int n, int i, int j, float jx, float jy, int s, int S, float r[3]) {
    double x = orc_lin(i, n, 0.5, -0.5);
    double y = orc_lin(j, n, -0.5, 0.5);
    double xx = x * x, yy = y * y;
    double sxy = xx + yy; /* np.linalg.norm: sqrt((x^2 + y^2) + 1) */
    double nrm = sqrt(sxy + 1.0);
    r[0] = (float)(x / nrm);
    r[1] = (float)(y / nrm);
    r[2] = (float)(1.0 / nrm);
    /* (sampleDist + sample)/antialias_samples, scene.py:31-32; then / x_dims, :73-74 */
    float sdx = ((jx + (float)s) / (float)S) / (float)n;
    float sdy = ((jy + (float)s) / (float)S) / (float)n;
    r[0] = r[0] + sdx;
    r[1] = r[1] + sdy;
}

static inline void orc_mat3(const float* M, const float* v, float* out) {
    for (int r = 0; r < 3; r++)
        out[r] = FMAF(M[r * 3 + 2], v[2], FMAF(M[r * 3 + 1], v[1], M[r * 3 + 0] * v[0]));
}

static void orc_prep(const rrt_scene* sc, int scene, orc_obj* objs, orc_glob* g) {
    const float* cam = sc->camera + (size_t)scene * sc->camera_scene_stride;
    const float* li = sc->light + (size_t)scene * sc->light_scene_stride;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) g->C[r * 3 + c] = cam[r * 4 + c];
        g->ct[r] = cam[r * 4 + 3];
        g->look[r] = cam[12 + r];
        g->L[r] = li[r];
        g->I[r] = li[3 + r];
    }
    g->Ln = sqrt((double)g->L[0] * g->L[0] + (double)g->L[1] * g->L[1] + (double)g->L[2] * g->L[2]);
    for (int r = 0; r < 3; r++) g->Lh[r] = g->L[r] / g->Ln;
    {   /* canonical float32 -Lhat for the shadow mask: sqrt and div round-to-nearest */
        float ln = sqrtf(FMAF(g->L[2], g->L[2], FMAF(g->L[1], g->L[1], g->L[0] * g->L[0])));
        for (int r = 0; r < 3; r++) g->U[r] = -(g->L[r] / ln);
    }
    for (int k = 0; k < sc->num_objects; k++) {
        const float* w = sc->w2o + (size_t)scene * sc->w2o_scene_stride + (size_t)k * RRT_W2O_STRIDE;
        const float* m = sc->material + (size_t)scene * sc->material_scene_stride + (size_t)k * RRT_MAT_STRIDE;
        orc_obj* o = &objs[k];
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++) o->A[r * 3 + c] = w[r * 4 + c];
            o->b[r] = w[r * 4 + 3];
        }
        for (int r = 0; r < 3; r++) /* o' = A.c + b */
            o->o[r] = FMAF(o->A[r * 3 + 2], g->ct[2],
                           FMAF(o->A[r * 3 + 1], g->ct[1], FMAF(o->A[r * 3 + 0], g->ct[0], o->b[r])));
        o->cc = FMAF(o->o[2], o->o[2], FMAF(o->o[1], o->o[1], o->o[0] * o->o[0])) - 1.0f;
        o->type = sc->obj_type[k];
        o->has64 = 0;
        o->ka = m[0]; o->kd = m[1]; o->ks = m[2]; o->sh = m[3];
        o->col[0] = m[4]; o->col[1] = m[5]; o->col[2] = m[6];
    }
}

/* one ray-object test in canonical order; returns t (+inf on miss) */
typedef struct { float d[3], vn, pd, det, t; double d64[3], vn64, pd64, det64, t64; } orc_hit;

/* TEST SWITCH: 1 => the hit record used by shading/backward is recomputed in
 * double from the float32 inputs (winners still come from the float32 sweep or
 * hit_in).  Used only to validate the closed-form reverse pass against float64
 * autograd to ~1e-12; the canonical oracle runs with 0. */
static int orc_f64_record = 0;
void orc_set_f64_record(int on) { orc_f64_record = on; }

static inline void orc_promote_d(const orc_obj* o, const double dw[3], orc_hit* h);
static inline void orc_promote(const orc_obj* o, const float dw[3], orc_hit* h) {
    if (!orc_f64_record) {
        for (int c = 0; c < 3; c++) h->d64[c] = h->d[c];
        h->vn64 = h->vn; h->pd64 = h->pd; h->det64 = h->det; h->t64 = h->t;
        return;
    }
    for (int r = 0; r < 3; r++)
        h->d64[r] = (double)o->A[r * 3] * dw[0] + (double)o->A[r * 3 + 1] * dw[1] + (double)o->A[r * 3 + 2] * dw[2];
    const double o0 = o->has64 ? o->o64[0] : (double)o->o[0], o1 = o->has64 ? o->o64[1] : (double)o->o[1],
                 o2 = o->has64 ? o->o64[2] : (double)o->o[2];
    if (o->type == RRT_OBJ_SPHERE) {
        double oo = o0 * o0 + o1 * o1 + o2 * o2;
        h->vn64 = h->d64[0] * h->d64[0] + h->d64[1] * h->d64[1] + h->d64[2] * h->d64[2];
        h->pd64 = h->d64[0] * o0 + h->d64[1] * o1 + h->d64[2] * o2;
        h->det64 = h->pd64 * h->pd64 - h->vn64 * (oo - 1.0);
        h->t64 = (-h->pd64 - sqrt(h->det64)) / h->vn64;
    } else {
        h->vn64 = h->pd64 = h->det64 = 0.0;
        h->t64 = -o2 / h->d64[2];
    }
}

/* same, world direction given in double (the reflected ray of the mirror bounce in validation mode) */
static inline void orc_promote_d(const orc_obj* o, const double dw[3], orc_hit* h) {
    for (int r = 0; r < 3; r++)
        h->d64[r] = (double)o->A[r * 3] * dw[0] + (double)o->A[r * 3 + 1] * dw[1] + (double)o->A[r * 3 + 2] * dw[2];
    const double o0 = o->has64 ? o->o64[0] : (double)o->o[0], o1 = o->has64 ? o->o64[1] : (double)o->o[1],
                 o2 = o->has64 ? o->o64[2] : (double)o->o[2];
    if (o->type == RRT_OBJ_SPHERE) {
        double oo = o0 * o0 + o1 * o1 + o2 * o2;
        h->vn64 = h->d64[0] * h->d64[0] + h->d64[1] * h->d64[1] + h->d64[2] * h->d64[2];
        h->pd64 = h->d64[0] * o0 + h->d64[1] * o1 + h->d64[2] * o2;
        h->det64 = h->pd64 * h->pd64 - h->vn64 * (oo - 1.0);
        h->t64 = (-h->pd64 - sqrt(h->det64)) / h->vn64;
    } else {
        h->vn64 = h->pd64 = h->det64 = 0.0;
        h->t64 = -o2 / h->d64[2];
    }
}

static inline float orc_test(const orc_obj* o, const float dw[3], orc_hit* h) {
    orc_mat3(o->A, dw, h->d); /* d' = A.dw */
    if (o->type == RRT_OBJ_SPHERE) {
        h->vn = FMAF(h->d[2], h->d[2], FMAF(h->d[1], h->d[1], h->d[0] * h->d[0]));
        h->pd = FMAF(h->d[2], o->o[2], FMAF(h->d[1], o->o[1], h->d[0] * o->o[0]));
        float vc = h->vn * o->cc;
        h->det = FMAF(h->pd, h->pd, -vc);
        if (!(h->det > 0.0f)) return h->t = INFINITY;   /* det<=0 or NaN -> inf, shape.py:124-125 */
        float sq = sqrtf(h->det);
        return h->t = (-h->pd - sq) / h->vn;            /* first root, shape.py:121-123 */
    } else {
        float t = (-o->o[2]) / h->d[2];                  /* shape.py:27 */
        float px = FMAF(t, h->d[0], o->o[0]);
        float py = FMAF(t, h->d[1], o->o[1]);
        int m = (h->d[2] != 0.0f) && (t > 0.0f) && (px > -0.5f) && (px < 0.5f) && (py > -0.5f) && (py < 0.5f);
        h->vn = h->pd = h->det = 0.0f;
        return h->t = m ? t : INFINITY;
    }
}

/* nearest-hit sweep, scene.py:38-47: strict '<', list order => first wins ties */
static inline int orc_sweep(const orc_obj* objs, int N, const float dw[3], float* tmin_out) {
    float tmin = INFINITY;
    int idx = -1;
    orc_hit h;
    for (int k = 0; k < N; k++) {
        float t = orc_test(&objs[k], dw, &h);
        if (t < tmin) { tmin = t; idx = k; }
    }
    *tmin_out = tmin;
    return idx;
}

/* Hard shadows, RRT_FLAG_SHADOWS (include/rrt_b200.h): the formula of Sphere.shadow,
 * shape.py:85-97, at the call site scene.py:41-45 (commented out in the reference), in
 * canonical float32 order.  `t` is the winner's ray parameter; the surface point is taken
 * in the CASTER's object space (y = o'_k + t d'_k = w2o_k (c + t d)). */
static inline int orc_shadowed(const orc_obj* objs, int N, int winner, const float dw[3], float t, const float U[3]) {
    for (int k = 0; k < N; k++) {
        if (k == winner || objs[k].type != RRT_OBJ_SPHERE) continue;
        const orc_obj* o = &objs[k];
        float d[3], y[3];
        orc_mat3(o->A, dw, d);
        for (int c = 0; c < 3; c++) y[c] = FMAF(t, d[c], o->o[c]);
        float x = FMAF(y[2], U[2], FMAF(y[1], U[1], y[0] * U[0]));
        float yy = FMAF(y[2], y[2], FMAF(y[1], y[1], y[0] * y[0]));
        float dec = FMAF(x, x, -yy) + 1.0f;
        if (dec > 0.0f && (-x - sqrtf(dec)) >= 0.0f) return 1;
    }
    return 0;
}

/* vectorisable sphere-only sweep used when every object is a sphere (bit-identical:
 * same operations per test; only the control flow differs) */
typedef struct { float *a[9], *o[3], *cc; } orc_soa;

static inline int orc_sweep_spheres(const orc_obj* objs, const orc_soa* S, int N, const float dw[3], float* tmin_out) {
    float tmin = INFINITY;
    int idx = -1;
    enum { CH = 64 };
    float det[CH];
    for (int k0 = 0; k0 < N; k0 += CH) {
        int m = N - k0 < CH ? N - k0 : CH;
        int any = 0;
#pragma omp simd reduction(| : any)
        for (int q = 0; q < m; q++) {
            int k = k0 + q;
            float dx = FMAF(S->a[2][k], dw[2], FMAF(S->a[1][k], dw[1], S->a[0][k] * dw[0]));
            float dy = FMAF(S->a[5][k], dw[2], FMAF(S->a[4][k], dw[1], S->a[3][k] * dw[0]));
            float dz = FMAF(S->a[8][k], dw[2], FMAF(S->a[7][k], dw[1], S->a[6][k] * dw[0]));
            float vn = FMAF(dz, dz, FMAF(dy, dy, dx * dx));
            float pd = FMAF(dz, S->o[2][k], FMAF(dy, S->o[1][k], dx * S->o[0][k]));
            float vc = vn * S->cc[k];
            float dt = FMAF(pd, pd, -vc);
            det[q] = dt;
            any |= (dt > 0.0f);
        }
        if (!any) continue;
        for (int q = 0; q < m; q++) {
            if (det[q] > 0.0f) {
                orc_hit h;
                float t = orc_test(&objs[k0 + q], dw, &h);
                if (t < tmin) { tmin = t; idx = k0 + q; }
            }
        }
    }
    *tmin_out = tmin;
    return idx;
}

/* ---- shading (double, on the float32 hit record) ---------------------------- */
typedef struct {
    double t, d[3], o[3], p[3], pn, nrm[3], ndl, rm[3], rv, pw, ph, col[3];
    int inside[3];
} orc_shade_rec;

static inline void orc_shade(const rrt_scene* sc, const orc_obj* o, const orc_glob* g, const orc_hit* h,
                             orc_shade_rec* r, double rgb[3]) {
    r->t = h->t64;
    for (int c = 0; c < 3; c++) { r->d[c] = h->d64[c]; r->o[c] = o->has64 ? o->o64[c] : (double)o->o[c]; }
    if (sc->shader == RRT_SHADER_DEPTH) { /* shader.py:14-20 */
        double v = 1.0 - r->t / (double)sc->max_depth;
        rgb[0] = rgb[1] = rgb[2] = v;
        return;
    }
    if (o->type == RRT_OBJ_SPHERE) { /* shape.py:134-137 */
        for (int c = 0; c < 3; c++) r->p[c] = r->o[c] + r->t * r->d[c];
        r->pn = sqrt(r->p[0] * r->p[0] + r->p[1] * r->p[1] + r->p[2] * r->p[2]);
        for (int c = 0; c < 3; c++) r->nrm[c] = r->p[c] / r->pn;
    } else { /* shape.py:55-68 */
        r->nrm[0] = r->nrm[1] = 0.0;
        r->nrm[2] = (o->o[2] > 0.0f) ? 1.0 : -1.0;
        r->pn = 1.0;
    }
    r->ndl = -(r->nrm[0] * g->Lh[0] + r->nrm[1] * g->Lh[1] + r->nrm[2] * g->Lh[2]); /* shader.py:40 */
    r->ph = (double)o->ka + (double)o->kd * r->ndl;
    r->rv = 0.0; r->pw = 0.0;
    if (sc->shader == RRT_SHADER_PHONG) { /* shader.py:43-45 */
        for (int c = 0; c < 3; c++) r->rm[c] = 2.0 * r->ndl * r->nrm[c] + g->Lh[c];
        r->rv = r->rm[0] * g->look[0] + r->rm[1] * g->look[1] + r->rm[2] * g->look[2];
        r->pw = pow(r->rv, (double)o->sh);
        r->ph += (double)o->ks * r->pw;
    }
    for (int c = 0; c < 3; c++) { /* shader.py:50-51 */
        double v = r->ph * (double)o->col[c] * (double)g->I[c];
        r->col[c] = v;
        r->inside[c] = (v >= 0.0 && v <= 1.0);
        rgb[c] = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
    }
}

/* ---- mirror bounce (RRT_FLAG_MIRROR; include/rrt_b200.h) ---------------------------------------
 * An EXTENSION: the reference has no secondary ray (match_mirror.py:40,45 matches an image to its
 * left-right flip; the hook would be scene.py:41-45 / shader.py:43-45).  PARITY UNPINNED by the
 * reference; pinned instead by the dense NumPy restatement, float64 autograd and finite differences.
 * Root camera variant only (camera.o2w = identity, origin 0).  For a winning primary ray (object k,
 * parameter t, world direction d), all in float32 canonical order (this defines the secondary mask):
 *     n_o  = p'/|p'| (sphere, p' = o' + t d')   or  (0,0,+-1) (square)          object-space normal
 *     m    = A_k^T n_o ,  n_w = m/|m|                                           world normal
 *     r    = d - 2 (d.n_w) n_w ,   P = t d                                       reflected ray
 * The secondary ray (P, r) is tested against every OTHER object in list order with the canonical
 * ray-object test (o'' = A_j P + b_j), nearest hit with t2 > 0, strict '<'; the hit is shaded by the
 * scene's shader as seen along r;  rgb = (1 - k_k) rgb_primary + k_k rgb_secondary  (0 if nothing is
 * hit), k_k = reflectivity of the primary object (a constant: no gradient).  One bounce. */
typedef struct {
    float P[3], r[3], nw[3], no[3], mn, dn;
    double P64[3], r64[3], nw64[3], no64[3], mn64, dn64;
} orc_bounce;

static inline void orc_bounce_geom(const orc_obj* o, const orc_hit* h, const float dw[3], orc_bounce* b) {
    float m[3];
    if (o->type == RRT_OBJ_SPHERE) {
        float p[3];
        for (int c = 0; c < 3; c++) p[c] = FMAF(h->t, h->d[c], o->o[c]);
        float pn = sqrtf(FMAF(p[2], p[2], FMAF(p[1], p[1], p[0] * p[0])));
        for (int c = 0; c < 3; c++) b->no[c] = p[c] / pn;
    } else {
        b->no[0] = b->no[1] = 0.0f;
        b->no[2] = (o->o[2] > 0.0f) ? 1.0f : -1.0f;
    }
    for (int i = 0; i < 3; i++) m[i] = FMAF(o->A[6 + i], b->no[2], FMAF(o->A[3 + i], b->no[1], o->A[i] * b->no[0]));
    b->mn = sqrtf(FMAF(m[2], m[2], FMAF(m[1], m[1], m[0] * m[0])));
    for (int c = 0; c < 3; c++) b->nw[c] = m[c] / b->mn;
    b->dn = FMAF(dw[2], b->nw[2], FMAF(dw[1], b->nw[1], dw[0] * b->nw[0]));
    const float k2 = -2.0f * b->dn;
    for (int c = 0; c < 3; c++) { b->r[c] = FMAF(k2, b->nw[c], dw[c]); b->P[c] = h->t * dw[c]; }
    if (!orc_f64_record) {
        for (int c = 0; c < 3; c++) { b->P64[c] = b->P[c]; b->r64[c] = b->r[c]; b->nw64[c] = b->nw[c]; b->no64[c] = b->no[c]; }
        b->mn64 = b->mn; b->dn64 = b->dn;
        return;
    }
    double m64[3];
    if (o->type == RRT_OBJ_SPHERE) {
        double p[3], pn = 0.0;
        for (int c = 0; c < 3; c++) { p[c] = (double)o->o[c] + h->t64 * h->d64[c]; pn += p[c] * p[c]; }
        pn = sqrt(pn);
        for (int c = 0; c < 3; c++) b->no64[c] = p[c] / pn;
    } else {
        for (int c = 0; c < 3; c++) b->no64[c] = b->no[c];
    }
    b->mn64 = 0.0;
    for (int i = 0; i < 3; i++) {
        m64[i] = (double)o->A[i] * b->no64[0] + (double)o->A[3 + i] * b->no64[1] + (double)o->A[6 + i] * b->no64[2];
        b->mn64 += m64[i] * m64[i];
    }
    b->mn64 = sqrt(b->mn64);
    b->dn64 = 0.0;
    for (int c = 0; c < 3; c++) { b->nw64[c] = m64[c] / b->mn64; b->dn64 += (double)dw[c] * b->nw64[c]; }
    for (int c = 0; c < 3; c++) { b->r64[c] = (double)dw[c] - 2.0 * b->dn64 * b->nw64[c]; b->P64[c] = h->t64 * (double)dw[c]; }
}

/* nearest OTHER object along the reflected ray; fills o2 (object j2 re-based at origin P) and its hit record */
static inline int orc_secondary(const orc_obj* objs, int N, int k, const orc_bounce* b, orc_obj* o2, orc_hit* h2) {
    float tmin = INFINITY;
    int j2 = -1;
    for (int j = 0; j < N; j++) {
        if (j == k) continue;
        orc_obj t = objs[j];
        for (int r = 0; r < 3; r++)
            t.o[r] = FMAF(t.A[r * 3 + 2], b->P[2], FMAF(t.A[r * 3 + 1], b->P[1], FMAF(t.A[r * 3 + 0], b->P[0], t.b[r])));
        t.cc = FMAF(t.o[2], t.o[2], FMAF(t.o[1], t.o[1], t.o[0] * t.o[0])) - 1.0f;
        orc_hit hh;
        float t2 = orc_test(&t, b->r, &hh);
        if (t2 > 0.0f && t2 < tmin) { tmin = t2; j2 = j; *o2 = t; *h2 = hh; }
    }
    if (j2 >= 0) {
        if (orc_f64_record) {
            o2->has64 = 1;
            for (int r = 0; r < 3; r++)
                o2->o64[r] = (double)o2->A[r * 3] * b->P64[0] + (double)o2->A[r * 3 + 1] * b->P64[1] +
                             (double)o2->A[r * 3 + 2] * b->P64[2] + (double)o2->b[r];
            orc_promote_d(o2, b->r64, h2);
        } else {
            orc_promote(o2, b->r, h2);
        }
    }
    return j2;
}

static void orc_backward_ray(const rrt_scene* sc, const orc_obj* o, const orc_glob* g, const orc_hit* h,
                             const orc_shade_rec* r, const float rcam[3], const double origin[3], const double dwd[3],
                             int camera_grad, const double* extra, double* god,
                             const double gc[3], double* og, double* gg);

/* reverse pass of one winning primary ray with the bounce: secondary object first (its own parameters,
 * and the gradients w.r.t. P and r), then the primary object with the chain through the reflection */
static void orc_mirror_backward(const rrt_scene* sc, const orc_obj* objs, int N, const orc_glob* g, int k,
                                const orc_hit* h, const orc_shade_rec* r, const float dw[3], const double gc[3],
                                double kr, double* gl, double* ggl) {
    const orc_obj* o = &objs[k];
    orc_bounce b;
    orc_obj o2;
    orc_hit h2;
    memset(&o2, 0, sizeof o2);
    memset(&h2, 0, sizeof h2);
    orc_bounce_geom(o, h, dw, &b);
    const int j2 = orc_secondary(objs, N, k, &b, &o2, &h2);
    double gc1[3], gc2[3], extra[4] = {0, 0, 0, 0}, dA[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < 3; c++) { gc1[c] = (1.0 - kr) * gc[c]; gc2[c] = kr * gc[c]; }
    const double dw64[3] = {dw[0], dw[1], dw[2]}, zero3[3] = {0, 0, 0};
    if (j2 >= 0) {
        orc_shade_rec r2;
        double rgb2[3], god[6], GP[3], Gr[3];
        orc_shade(sc, &o2, g, &h2, &r2, rgb2);
        orc_backward_ray(sc, &o2, g, &h2, &r2, NULL, b.P64, b.r64, 0, NULL, god, gc2,
                         gl + (size_t)j2 * RRT_OBJ_GRAD_STRIDE, ggl);
        for (int c = 0; c < 3; c++) {
            GP[c] = o2.A[c] * god[0] + o2.A[3 + c] * god[1] + o2.A[6 + c] * god[2];
            Gr[c] = o2.A[c] * god[3] + o2.A[3 + c] * god[4] + o2.A[6 + c] * god[5];
        }
        extra[3] = GP[0] * dw64[0] + GP[1] * dw64[1] + GP[2] * dw64[2];              /* P = t d */
        const double grn = Gr[0] * b.nw64[0] + Gr[1] * b.nw64[1] + Gr[2] * b.nw64[2];
        double g_nw[3], g_m[3], dot = 0.0;
        for (int c = 0; c < 3; c++) { g_nw[c] = -2.0 * (b.dn64 * Gr[c] + grn * dw64[c]); dot += b.nw64[c] * g_nw[c]; }
        for (int c = 0; c < 3; c++) g_m[c] = (g_nw[c] - b.nw64[c] * dot) / b.mn64;   /* n_w = m/|m| */
        for (int rr = 0; rr < 3; rr++)
            for (int i = 0; i < 3; i++) dA[rr * 3 + i] = b.no64[rr] * g_m[i];          /* m = A^T n_o */
        if (o->type == RRT_OBJ_SPHERE)
            for (int rr = 0; rr < 3; rr++)
                extra[rr] = (double)o->A[rr * 3] * g_m[0] + (double)o->A[rr * 3 + 1] * g_m[1] + (double)o->A[rr * 3 + 2] * g_m[2];
    }
    double* og1 = gl + (size_t)k * RRT_OBJ_GRAD_STRIDE;
    orc_backward_ray(sc, o, g, h, r, NULL, zero3, dw64, 0, extra, NULL, gc1, og1, ggl);
    for (int rr = 0; rr < 3; rr++)
        for (int i = 0; i < 3; i++) og1[rr * 4 + i] += dA[rr * 3 + i];
}

/* ---- reverse pass for one winning ray (closed form) ---------
 * gc[3] = dL/d(image[a,b,:]) / S.  Accumulates into og (this object's 19 slots)
 * and gg (21 global slots).  Light-direction slots hold d/d(Lhat) here; the
 * normalisation chain is applied once at the end (orc_finish_grads). */
/* Mirror bounce (RRT_FLAG_MIRROR): `origin` / `dwd` are the ray's world origin and direction (camera
 * origin and primary direction, or hit point P and reflected direction r of a secondary ray);
 * `extra` = {dL/d(object-space normal)[3], dL/dt} fed in from the secondary ray through the reflection
 * (primary rays only), `god` receives {g_o'[3], g_d'[3]} (object-space origin / direction gradients). */
static void orc_backward_ray(const rrt_scene* sc, const orc_obj* o, const orc_glob* g, const orc_hit* h,
                             const orc_shade_rec* r, const float rcam[3], const double origin[3], const double dwd[3],
                             int camera_grad, const double* extra, double* god,
                             const double gc[3], double* og, double* gg) {
    double g_t = 0.0, g_p[3] = {0, 0, 0};
    double g_o[3] = {0, 0, 0}, g_d[3] = {0, 0, 0};
    if (sc->shader == RRT_SHADER_DEPTH) {
        g_t = -(gc[0] + gc[1] + gc[2]) / (double)sc->max_depth;
    } else {
        double g_ph = 0.0;
        for (int c = 0; c < 3; c++) {
            if (!r->inside[c]) continue;
            g_ph += gc[c] * (double)o->col[c] * (double)g->I[c];
            og[16 + c] += gc[c] * r->ph * (double)g->I[c];       /* d/d color_c     */
            gg[3 + c] += gc[c] * r->ph * (double)o->col[c];      /* d/d intensity_c */
        }
        og[12] += g_ph;               /* ka */
        og[13] += g_ph * r->ndl;      /* kd */
        double g_ndl = g_ph * (double)o->kd;
        double g_n[3] = {0, 0, 0}, g_Lh[3] = {0, 0, 0};
        if (sc->shader == RRT_SHADER_PHONG) {
            og[14] += g_ph * r->pw;   /* ks */
            if (r->rv > 0.0) og[15] += g_ph * (double)o->ks * r->pw * log(r->rv); /* shininess */
            double g_rv = g_ph * (double)o->ks * (double)o->sh * pow(r->rv, (double)o->sh - 1.0);
            double g_rm[3];
            for (int c = 0; c < 3; c++) {
                g_rm[c] = g_rv * (double)g->look[c];
                gg[18 + c] += g_rv * r->rm[c];                   /* d/d look_at */
            }
            g_ndl += 2.0 * (g_rm[0] * r->nrm[0] + g_rm[1] * r->nrm[1] + g_rm[2] * r->nrm[2]);
            for (int c = 0; c < 3; c++) { g_n[c] += 2.0 * r->ndl * g_rm[c]; g_Lh[c] += g_rm[c]; }
        }
        for (int c = 0; c < 3; c++) { g_n[c] -= g_ndl * g->Lh[c]; g_Lh[c] -= g_ndl * r->nrm[c]; }
        for (int c = 0; c < 3; c++) gg[c] += g_Lh[c];
        if (extra) { for (int c = 0; c < 3; c++) g_n[c] += extra[c]; g_t += extra[3]; }
        if (o->type == RRT_OBJ_SPHERE) {
            double ndg = r->nrm[0] * g_n[0] + r->nrm[1] * g_n[1] + r->nrm[2] * g_n[2];
            for (int c = 0; c < 3; c++) g_p[c] = (g_n[c] - r->nrm[c] * ndg) / r->pn;
            for (int c = 0; c < 3; c++) { g_o[c] = g_p[c]; g_t += g_p[c] * r->d[c]; g_d[c] = r->t * g_p[c]; }
        }
    }
    if (o->type == RRT_OBJ_SPHERE) {
        double vn = h->vn64, pd = h->pd64, det = h->det64;
        double cc = orc_f64_record ? (r->o[0] * r->o[0] + r->o[1] * r->o[1] + r->o[2] * r->o[2] - 1.0) : (double)o->cc;
        double g_pd = -g_t / vn, g_s = -g_t / vn, g_vn = -g_t * r->t / vn;
        double g_det = g_s / (2.0 * sqrt(det));
        g_pd += 2.0 * pd * g_det;
        g_vn -= cc * g_det;
        double g_cc = -vn * g_det;
        for (int c = 0; c < 3; c++) {
            g_o[c] += 2.0 * r->o[c] * g_cc + r->d[c] * g_pd;
            g_d[c] += r->o[c] * g_pd + 2.0 * r->d[c] * g_vn;
        }
    } else { /* t = -o'_z / d'_z */
        g_o[2] += -g_t / r->d[2];
        g_d[2] += -g_t * r->t / r->d[2];
    }
    /* d' = A.dw ; o' = A.origin + b */
    for (int rr = 0; rr < 3; rr++) {
        for (int c = 0; c < 3; c++) og[rr * 4 + c] += g_d[rr] * dwd[c] + g_o[rr] * origin[c];
        og[rr * 4 + 3] += g_o[rr];
    }
    if (god) for (int c = 0; c < 3; c++) { god[c] = g_o[c]; god[3 + c] = g_d[c]; }
    if (camera_grad) {
        double g_dw[3], g_ct[3];
        for (int c = 0; c < 3; c++) {
            g_dw[c] = o->A[0 * 3 + c] * g_d[0] + o->A[1 * 3 + c] * g_d[1] + o->A[2 * 3 + c] * g_d[2];
            g_ct[c] = o->A[0 * 3 + c] * g_o[0] + o->A[1 * 3 + c] * g_o[1] + o->A[2 * 3 + c] * g_o[2];
        }
        for (int rr = 0; rr < 3; rr++) {
            for (int c = 0; c < 3; c++) gg[6 + rr * 4 + c] += g_dw[rr] * (double)rcam[c];
            gg[6 + rr * 4 + 3] += g_ct[rr];
        }
    }
}

/* light direction: Lhat = L/|L|  =>  g_L = (g_Lhat - Lhat (Lhat.g_Lhat)) / |L| */
static void orc_finish_grads(const orc_glob* g, double* gg) {
    double dot = g->Lh[0] * gg[0] + g->Lh[1] * gg[1] + g->Lh[2] * gg[2];
    for (int c = 0; c < 3; c++) gg[c] = (gg[c] - g->Lh[c] * dot) / g->Ln;
}

/* ---- drivers ---------------------------------------------------------------- */
static int orc_rows(const rrt_scene* sc) { return sc->row_count > 0 ? sc->row_count : sc->n - sc->row_begin; }

static int orc_all_spheres(const orc_obj* objs, int N) {
    for (int k = 0; k < N; k++) if (objs[k].type != RRT_OBJ_SPHERE) return 0;
    return 1;
}

static void orc_make_soa(const orc_obj* objs, int N, orc_soa* S, float* store) {
    for (int q = 0; q < 9; q++) S->a[q] = store + (size_t)q * N;
    for (int q = 0; q < 3; q++) S->o[q] = store + (size_t)(9 + q) * N;
    S->cc = store + (size_t)12 * N;
    for (int k = 0; k < N; k++) {
        for (int q = 0; q < 9; q++) S->a[q][k] = objs[k].A[q];
        for (int q = 0; q < 3; q++) S->o[q][k] = objs[k].o[q];
        S->cc[k] = objs[k].cc;
    }
}

static inline void orc_pixel_ray(const rrt_scene* sc, const orc_glob* g, int scene, int a, int b, int s,
                                 float rcam[3], float dw[3]) {
    int n = sc->n, S = sc->samples, rows = orc_rows(sc);
    float jx, jy;
    if (sc->jitter_x) {
        size_t off = (size_t)scene * sc->jitter_scene_stride + ((size_t)(a - sc->row_begin) * n + b) * S + s;
        jx = sc->jitter_x[off];
        jy = sc->jitter_y[off];
    } else {
        jx = orc_rng(sc->seed, (uint32_t)(scene + sc->scene_begin), (uint32_t)(a * n + b), (uint32_t)s, 0);
        jy = orc_rng(sc->seed, (uint32_t)(scene + sc->scene_begin), (uint32_t)(a * n + b), (uint32_t)s, 1);
    }
    (void)rows;
    int i = sc->transpose ? b : a, j = sc->transpose ? a : b;
    orc_gen_ray(n, i, j, jx, jy, s, S, rcam);
    orc_mat3(g->C, rcam, dw); /* camera.o2w, orbit_experiments/scene.py:80 (identity in the root variant) */
}

/*
 * mode bits: 1 = write image/hit_index/tmin, 2 = backward from dl_dimage,
 *            4 = fused mse (loss + backward from 2 w (image - target))
 * hit_in: optional winners to use instead of sweeping (backward only).
 */
static int orc_run(const rrt_scene* sc, int mode, float* image, int32_t* hit_out, float* tmin_out,
                   const float* dl_dimage, const int32_t* hit_in, const float* target, const float* cw,
                   double* loss, double* grad, int32_t* hit2_out) {
    int n = sc->n, S = sc->samples, N = sc->num_objects, B = sc->num_scenes, rows = orc_rows(sc);
    const int mirror = (sc->flags & RRT_FLAG_MIRROR) != 0;
    if (mirror && (!sc->reflectivity || sc->shader == RRT_SHADER_DEPTH || sc->camera_grad)) return RRT_ERR_UNSUPPORTED;
    size_t gsz = RRT_GRAD_SIZE(N);
    float w3[3] = {1.f, 1.f, 1.f};
    if (cw) { w3[0] = cw[0]; w3[1] = cw[1]; w3[2] = cw[2]; }
    if (S > 64) return RRT_ERR_UNSUPPORTED;
    for (int scene = 0; scene < B; scene++) {
        orc_obj* objs = (orc_obj*)malloc(sizeof(orc_obj) * (size_t)(N > 0 ? N : 1));
        float* soa_store = (float*)malloc(sizeof(float) * 13 * (size_t)(N > 0 ? N : 1));
        orc_glob g;
        orc_soa soa;
        orc_prep(sc, scene, objs, &g);
        const float* refl = mirror ? sc->reflectivity + (size_t)scene * sc->reflectivity_scene_stride : NULL;
        int all_sph = orc_all_spheres(objs, N);
        if (all_sph) orc_make_soa(objs, N, &soa, soa_store);
        double* gscene = grad ? grad + (size_t)scene * gsz : NULL;
        double loss_scene = 0.0;
        if (gscene) memset(gscene, 0, sizeof(double) * gsz);
#pragma omp parallel
        {
            double* gl = gscene ? (double*)calloc(gsz, sizeof(double)) : NULL;
            double loss_l = 0.0;
#pragma omp for schedule(dynamic, 1)
            for (int al = 0; al < rows; al++) {
                int a = sc->row_begin + al;
                for (int b = 0; b < n; b++) {
                    float rc[64][3], dws[64][3], tm[64];
                    int idx[64];
                    double pix[3] = {0, 0, 0};
                    double rgbs[64][3];
                    for (int s = 0; s < S; s++) {
                        orc_pixel_ray(sc, &g, scene, a, b, s, rc[s], dws[s]);
                        size_t ro = (((size_t)scene * S + s) * rows + al) * n + b;
                        int shadowed = 0;
                        if (hit_in) {
                            idx[s] = hit_in[ro];
                            if (idx[s] >= 0 && (idx[s] & RRT_HIT_SHADOWED)) { shadowed = 1; idx[s] &= ~RRT_HIT_SHADOWED; }
                            if (idx[s] >= N) idx[s] = -1;
                            tm[s] = INFINITY;
                        } else {
                            idx[s] = all_sph ? orc_sweep_spheres(objs, &soa, N, dws[s], &tm[s])
                                             : orc_sweep(objs, N, dws[s], &tm[s]);
                        }
                        rgbs[s][0] = rgbs[s][1] = rgbs[s][2] = 0.0;
                        if (idx[s] >= 0) {
                            orc_hit h;
                            orc_shade_rec r;
                            tm[s] = orc_test(&objs[idx[s]], dws[s], &h);
                            if (!hit_in && (sc->flags & RRT_FLAG_SHADOWS) && isfinite(tm[s]))
                                shadowed = orc_shadowed(objs, N, idx[s], dws[s], tm[s], g.U);
                            if (!shadowed) {
                                orc_promote(&objs[idx[s]], dws[s], &h);
                                orc_shade(sc, &objs[idx[s]], &g, &h, &r, rgbs[s]);
                                if (mirror && isfinite(tm[s])) {
                                    orc_bounce bn;
                                    orc_obj o2;
                                    orc_hit h2;
                                    orc_shade_rec r2;
                                    double rgb2[3] = {0, 0, 0};
                                    memset(&o2, 0, sizeof o2);
                                    memset(&h2, 0, sizeof h2);
                                    orc_bounce_geom(&objs[idx[s]], &h, dws[s], &bn);
                                    int j2 = orc_secondary(objs, N, idx[s], &bn, &o2, &h2);
                                    if (j2 >= 0) orc_shade(sc, &o2, &g, &h2, &r2, rgb2);
                                    const double kr = refl[idx[s]];
                                    for (int c = 0; c < 3; c++) rgbs[s][c] = (1.0 - kr) * rgbs[s][c] + kr * rgb2[c];
                                    if (hit2_out) hit2_out[ro] = j2;
                                }
                            }
                        }
                        if (hit2_out && !(mirror && idx[s] >= 0 && !shadowed && isfinite(tm[s]))) hit2_out[ro] = -1;
                        if (hit_out) hit_out[ro] = (shadowed && idx[s] >= 0) ? (idx[s] | RRT_HIT_SHADOWED) : idx[s];
                        if (shadowed) idx[s] = -1;   /* (0,0,0) and no gradient from here on */
                        if (tmin_out) tmin_out[ro] = tm[s];
                        for (int c = 0; c < 3; c++) pix[c] += rgbs[s][c]; /* scene.py:49 */
                    }
                    size_t po = (((size_t)scene * rows + al) * n + b) * 3;
                    for (int c = 0; c < 3; c++) pix[c] /= (double)S;      /* scene.py:50 */
                    if (image) for (int c = 0; c < 3; c++) image[po + c] = (float)pix[c];
                    if (!(mode & 6)) continue;
                    double gc[3];
                    if (mode & 4) {
                        for (int c = 0; c < 3; c++) {
                            double df = (double)(float)pix[c] - (double)target[po + c];
                            loss_l += (double)w3[c] * df * df;
                            gc[c] = 2.0 * (double)w3[c] * df / (double)S;
                        }
                    } else {
                        for (int c = 0; c < 3; c++) gc[c] = (double)dl_dimage[po + c] / (double)S;
                    }
                    for (int s = 0; s < S; s++) {
                        if (idx[s] < 0) continue;
                        orc_hit h;
                        orc_shade_rec r;
                        double rgb[3];
                        orc_test(&objs[idx[s]], dws[s], &h);
                        if (!isfinite(h.t)) continue; /* stale hit_in */
                        orc_promote(&objs[idx[s]], dws[s], &h);
                        orc_shade(sc, &objs[idx[s]], &g, &h, &r, rgb);
                        const double ct64[3] = {g.ct[0], g.ct[1], g.ct[2]}, dw64[3] = {dws[s][0], dws[s][1], dws[s][2]};
                        double* og1 = gl + (size_t)idx[s] * RRT_OBJ_GRAD_STRIDE;
                        double* ggl = gl + (size_t)N * RRT_OBJ_GRAD_STRIDE;
                        if (mirror) {
                            orc_mirror_backward(sc, objs, N, &g, idx[s], &h, &r, dws[s], gc, refl[idx[s]], gl, ggl);
                            continue;
                        }
                        orc_backward_ray(sc, &objs[idx[s]], &g, &h, &r, rc[s], ct64, dw64, sc->camera_grad, NULL, NULL, gc,
                                         og1, ggl);
                    }
                }
            }
#pragma omp critical
            {
                if (gscene) for (size_t q = 0; q < gsz; q++) gscene[q] += gl[q];
                loss_scene += loss_l;
            }
            free(gl);
        }
        if (gscene) orc_finish_grads(&g, gscene + (size_t)N * RRT_OBJ_GRAD_STRIDE);
        if (loss) loss[scene] = loss_scene;
        free(objs);
        free(soa_store);
    }
    return RRT_OK;
}

int orc_render_forward(const rrt_scene* sc, float* image, int32_t* hit_index, float* tmin) {
    return orc_run(sc, 1, image, hit_index, tmin, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
}

/* forward + the secondary (mirror) hit index per ray, [B][S][rows][n], -1 = none (test infrastructure) */
int orc_render_forward_secondary(const rrt_scene* sc, float* image, int32_t* hit_index, int32_t* hit2) {
    return orc_run(sc, 1, image, hit_index, NULL, NULL, NULL, NULL, NULL, NULL, NULL, hit2);
}

int orc_render_backward(const rrt_scene* sc, const float* dl_dimage, const int32_t* hit_index, double* grad) {
    return orc_run(sc, 2, NULL, NULL, NULL, dl_dimage, hit_index, NULL, NULL, NULL, grad, NULL);
}

int orc_render_fused_mse(const rrt_scene* sc, const float* target, const float* channel_weight, float* image,
                         int32_t* hit_index, double* loss, double* grad) {
    return orc_run(sc, 4 | 1, image, hit_index, NULL, NULL, NULL, target, channel_weight, loss, grad, NULL);
}

/* primary rays only (for pinning against numpy make_rays): out[rows][n][S][3] in
 * IMAGE index space, camera-space ray before camera.o2w */
int orc_primary_rays(const rrt_scene* sc, float* out) {
    int n = sc->n, S = sc->samples, rows = orc_rows(sc);
    orc_glob g;
    memset(&g, 0, sizeof g);
    g.C[0] = g.C[4] = g.C[8] = 1.0f;
    for (int al = 0; al < rows; al++)
        for (int b = 0; b < n; b++)
            for (int s = 0; s < S; s++) {
                float rc[3], dw[3];
                orc_pixel_ray(sc, &g, 0, sc->row_begin + al, b, s, rc, dw);
                memcpy(out + (((size_t)al * n + b) * S + s) * 3, rc, sizeof rc);
            }
    return RRT_OK;
}

float orc_rng_value(uint64_t seed, uint32_t scene, uint32_t pix, uint32_t s, uint32_t axis) {
    return orc_rng(seed, scene, pix, s, axis);
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
