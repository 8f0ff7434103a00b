/*
 * rrt_b200.h -- C ABI of the B200-native differentiable ray-tracer hot path.
 *
 * Drop-in boundary for ONE path of lebek/reversible-raytracer (Python 2 + Theano):
 * primary rays -> ray/shape intersection -> nearest hit -> Phong / depth shading
 * -> reverse pass to scene-parameter gradients.  The reference has no FFI of its
 * own; the interface replaced is the Python class API
 *     Scene(shapes, lights, camera, shader).build(antialias_samples)   scene.py:11-52
 *     T.grad(loss, params) over that graph                             optimize.py:25,73
 * and each entry point below cites the reference code it stands in for
 * (paths relative to the reference checkout).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, int return: 0 = OK, <0 = error
 *     (RRT_ERR_*); rrt_last_error() gives a thread-local message.  No exceptions
 *     cross the ABI.
 *   - Every buffer is DEVICE memory owned by the caller (except where a parameter
 *     is documented as host).  The library never allocates, frees or synchronises
 *     (measurement helpers that do live in a separate library, include/rrt_b200_bench.h).
 *   - All work is enqueued on the caller's CUDA stream (`stream` is a
 *     cudaStream_t passed as void*; NULL = legacy default stream).
 *   - No global mutable state: re-entrant from several host threads, one process
 *     per GPU.
 *   - Gradient / loss outputs are zero-initialised by the callee (on the stream).
 *
 * Index conventions: image[a][b][c] is the reference's
 * (x_dims, y_dims, 3) array.  With transpose=1 (root variant, scene.py:55-75)
 * pixel (a,b) is shaded with camera ray [b][a] -- the spatial transpose that
 * Transform.__call__ performs once (transform.py:46).  With transpose=0 (orbit
 * variant, orbit_experiments/scene.py:55-80: camera.o2w then shape.w2o, two
 * transposes) pixel (a,b) is shaded with camera ray [a][b].
 *
 * Row slabs (multi-GPU sharding): a call renders image rows
 * [row_begin, row_begin+row_count).  All per-pixel buffers passed to that call
 * (image, hit_index, tmin, target, dl_dimage, jitter) are SLAB-LOCAL: their row
 * index is (a - row_begin) and they hold row_count rows.
 */
#ifndef RRT_B200_H
#define RRT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RRT_VERSION 100 /* 0.1.0 */

/* shape kinds: shape.py:72 (Sphere), shape.py:16 (Square) */
#define RRT_OBJ_SPHERE 0
#define RRT_OBJ_SQUARE 1

/* shaders: shader.py:23-53 (Phong), orbit_experiments/shader.py:45,48 (Phong with
 * the specular term commented out), shader.py:9-20 (DepthMapShader) */
/* Shininess: x ** shininess follows C pow() like Theano's T.pow (shader.py:45): a negative base with a
 * non-integer exponent gives NaN, which the clip passes on (the pixel becomes NaN, as in the reference).
 * Deviation, stated: d/d shininess is accumulated only where the base is > 0 (the reference's T.grad would
 * put NaN there; every reference script keeps shininess constant and never asks for this gradient), and the
 * reverse pass uses rsqrt / fast-divide / fast-log approximations -- gradients are float32-tolerance
 * quantities (1e-3 per block), only hit masks and tmin are bit-exact. */
#define RRT_SHADER_PHONG 0
#define RRT_SHADER_PHONG_NOSPEC 1
#define RRT_SHADER_DEPTH 2

/* packed strides (floats) */
#define RRT_W2O_STRIDE 12      /* rows 0..2 of shape.w2o.m, row-major 3x4 = [A | b]       */
#define RRT_MAT_STRIDE 7       /* ka, kd, ks, shininess, color r, g, b   scene.py:89-101   */
#define RRT_LIGHT_STRIDE 6     /* direction[3], intensity[3]             scene.py:78-86    */
#define RRT_CAMERA_STRIDE 15   /* rows 0..2 of camera.o2w.m (3x4), look_at[3]              */
#define RRT_OBJ_GRAD_STRIDE 19 /* d/d w2o[12] then d/d material[7]                         */
#define RRT_GLOBAL_GRAD 21     /* d/d light dir[3], intensity[3], camera o2w[12], look_at[3] */

/* flat gradient vector per scene: [N][RRT_OBJ_GRAD_STRIDE] then [RRT_GLOBAL_GRAD] */
#define RRT_GRAD_SIZE(num_objects) ((size_t)(num_objects) * RRT_OBJ_GRAD_STRIDE + RRT_GLOBAL_GRAD)

/* rrt_scene.flags */
/* Conservative per-tile object culling before the sweep: every CTA bounds its rays by a
 * cone, rejects objects whose (inflated) bounding ball the cone cannot touch, and runs the
 * canonical hit test only on the survivors, in list order.  Results are BIT-IDENTICAL to
 * the exhaustive sweep (tested); only the work changes, so throughput measured with this
 * flag is never reported as a roofline fraction. */
#define RRT_FLAG_CULL 1
/* Force the general 8-rays-per-thread kernel even where the small-scene (one ray per thread)
 * kernel would be chosen.  Results agree; this exists for A/B measurements and tests. */
#define RRT_FLAG_NO_SMALL 2
/* Hard shadows (the reference's call site is commented out, scene.py:41-45,
 * and its helper is broken, shape.py:100-106, so PARITY IS WEAKLY PINNED: the formula of
 * Sphere.shadow, shape.py:85-97, is followed, evaluated in the frame of the shadow caster).
 * A winning ray with parameter t is in shadow if for some OTHER SPHERE k (list order)
 *     y = o'_k + t d'_k ;  x = y . (-Lhat) ;  dec = (x^2 - y.y) + 1 ;
 *     dec > 0  and  (-x - sqrt(dec)) >= 0
 * (unit sphere of the caster's object space, WORLD-space light direction like the shading,
 * squares cast no shadow: Square has no shadow method).  A shadowed ray shades to (0,0,0)
 * and carries no gradient (masks are constants).  hit_index keeps the winner and gets
 * RRT_HIT_SHADOWED or-ed in, so the shadow mask is observable and testable bit for bit. */
#define RRT_FLAG_SHADOWS 4
#define RRT_HIT_SHADOWED 0x40000000
/* Diagnostic: run the shadow pass of the general kernel with the scalar routine only (the
 * packed FFMA2 filter in front of it is exact, so results are identical; A/B and tests). */
#define RRT_FLAG_SCALAR_SHADOWS 8
/* Reverse pass produces d/d w2o (and, with camera_grad, d/d camera.o2w) only: the material, light
 * and look_at entries of the gradient vector are left ZERO and their sums are not computed.  For
 * callers that chain only through the transforms -- every autoencoder decoder of the reference
 * (autoencoder.py:57-71, orbit_experiments/autoencoder_2ly.py:82-91: materials, light and camera
 * direction are constants there). */
#define RRT_FLAG_NO_MATERIAL_GRAD 16
/* Nearest-hit sweep without the conservative pre-filter.  By default, when a prebuilt record table
 * is given (rrt_scene.obj_records), the hot loop of the sweep evaluates for every (ray, sphere)
 * pair a float32 QUADRATIC FORM in (u, v) = (d_x/d_z, d_y/d_z) whose sign conservatively bounds the
 * sign of the reference's discriminant det = pd^2 - vn*(o'.o' - 1) (shape.py:78-83): 5 fused
 * multiply-adds per pair instead of 11 operations.  A pair it cannot exclude (every true hit, plus a
 * shell of relative width 2^-18 around silhouettes) is then decided by the CANONICAL arithmetic, so
 * hit masks, tmin and everything downstream are bit-identical with and without the filter (tested).
 * Every (ray, object) pair is still tested individually -- this is not spatial culling.  This flag
 * forces the canonical packed sweep for all pairs (A/B measurements, roofline accounting). */
#define RRT_FLAG_CANONICAL_SWEEP 32
/* Deterministic reverse pass: gradients and loss are bit-identical from run to run (the reference's
 * T.grad is, optimize.py:25; float atomics are not).  Everything up to a warp's per-object sums is
 * evaluated in a fixed order anyway (thread: ray order; warp: fixed butterfly); with this flag the
 * combination ACROSS warps and CTAs -- normally float atomics in shared and global memory -- is
 * done in 128-bit fixed point (two int64 limbs per value, 2^-20 and 2^-60 units, integer atomics:
 * integer addition is associative, so the result does not depend on the order of arrival).
 * Range +-2^43, absolute resolution 2^-60 per contribution.  Needs rrt_scene.det_workspace. */
#define RRT_FLAG_DETERMINISTIC 64
/* One mirror-reflection bounce (BASELINE config 3 names it; the reference itself has NO secondary
 * ray -- match_mirror.py:40,45 matches an image to its left-right flip, the hook would be
 * scene.py:41-45 / shader.py:43-45 -- so this is an extension and PARITY IS UNPINNED by the
 * reference; it is pinned by a dense NumPy restatement, float64 autograd and finite differences).
 * Root camera variant only (camera.o2w rotation = identity, origin 0), Phong shaders only.
 * For a winning primary ray (object k, parameter t, world direction d), float32 canonical order:
 *     n_o = p'/|p'| (sphere; p' = o' + t d')  or (0,0,+-1) (square)      object-space normal
 *     n_w = A_k^T n_o / |A_k^T n_o|                                      world normal
 *     r   = d - 2 (d.n_w) n_w ,  P = t d                                  reflected ray
 * (P, r) is tested against every OTHER object in list order (canonical test with o'' = A_j P + b_j),
 * nearest hit with t2 > 0, strict '<'; that hit is shaded by the scene's shader as seen along r;
 *     rgb = (1 - k_k) rgb_primary + k_k rgb_secondary        (rgb_secondary = 0 if nothing is hit)
 * with k_k = rrt_scene.reflectivity[k] of the PRIMARY object (a constant: it gets no gradient).
 * The reverse pass differentiates through the secondary shading, the secondary object's transform,
 * and -- through P, r, n_w -- the primary object's transform (masks constant, like everywhere). */
#define RRT_FLAG_MIRROR 128
/* The small-scene kernel has two thread mappings with identical results (same device routines, same
 * canonical order, same sample summation order): one RAY per thread (the S samples of a pixel in
 * adjacent lanes, combined by shuffles) and one PIXEL per thread (its S samples in registers, a warp
 * covering a compact 8 x 4 pixel tile; S in {1, 2, 4}, no shadows / mirror).
 * By default the library takes pixel threads when the call has enough pixels to fill the GPU with them
 * (batches of scenes, e.g. the orbit decoder batch) and ray threads for a single small image.  These
 * two flags force one or the other where it applies (A/B measurements, tests). */
#define RRT_FLAG_PIXEL_THREADS 256
#define RRT_FLAG_RAY_THREADS 512
/* The fused entry points (rrt_render_fused_mse, rrt_small_step_mse) evaluate a LINEAR cost instead of the
 * squared error: their `target` argument is read as a weight image W [B][rows][n][3] and
 *     cost = sum_c w_c * sum(W * image)        d cost / d image = w_c * W
 * -- the loss of optimize_brightness.py:51, -image[90,85].sum() - image[50,90].sum(), is W = -1 at two
 * pixels.  Pixels with W = 0 carry no upstream gradient: their rays skip the reverse pass. */
#define RRT_FLAG_LINEAR_COST 1024

#define RRT_OK 0
#define RRT_ERR_INVALID (-1)   /* bad argument (message in rrt_last_error)            */
#define RRT_ERR_CUDA (-2)      /* CUDA runtime error at launch                        */
#define RRT_ERR_UNSUPPORTED (-3)

/*
 * Scene descriptor: everything Scene.build() reads (scene.py:11-52).
 * Passed by pointer from host memory; the pointers inside are device pointers.
 * Batches of scenes (num_scenes = B > 1) use a per-scene stride in floats for each
 * table; a stride of 0 shares the table across the batch.
 */
typedef struct rrt_scene {
    int32_t n;            /* image side: camera.x_dims == camera.y_dims (scene.py:21,24) */
    int32_t samples;      /* antialias_samples S (scene.py:18)                           */
    int32_t num_objects;  /* N = len(scene.shapes)                                        */
    int32_t num_scenes;   /* B >= 1                                                       */
    int32_t shader;       /* RRT_SHADER_*                                                 */
    int32_t transpose;    /* 1 = root camera variant, 0 = orbit variant (see above)       */
    int32_t row_begin;    /* first image row rendered by this call                        */
    int32_t row_count;    /* rows rendered; 0 means n - row_begin                         */
    float max_depth;      /* DepthMapShader.maxDepth (shader.py:11)                       */
    int32_t camera_grad;  /* 1: also produce d/d camera.o2w and d/d look_at               */
    uint64_t seed;        /* jitter seed, used only when jitter_x == NULL                 */

    const int32_t* obj_type; /* [N] RRT_OBJ_*, list order = scene.shapes order           */
    const float* w2o;        /* [B][N][RRT_W2O_STRIDE]                                    */
    const float* material;   /* [B][N][RRT_MAT_STRIDE]                                    */
    const float* light;      /* [B][RRT_LIGHT_STRIDE]   only lights[0] is used, shader.py:33 */
    const float* camera;     /* [B][RRT_CAMERA_STRIDE]                                    */
    /* Anti-alias jitter in [0,1), IMAGE index space, slab-local:
     * jitter_x[scene][a - row_begin][b][s] is the value the reference adds to ray
     * channel 0 of the ray that shades pixel (a,b) (scene.py:24-25,31-32,73-74).
     * NULL => in-kernel counter RNG keyed by (seed, scene_begin+scene, a*n+b, s).                   */
    const float* jitter_x;
    const float* jitter_y;

    int64_t w2o_scene_stride;      /* floats between scenes; 0 = shared */
    int64_t material_scene_stride;
    int64_t light_scene_stride;
    int64_t camera_scene_stride;
    int64_t jitter_scene_stride;

    /* Optional [n][n][3] float32 table of the camera-space ray grid BEFORE jitter, in
     * the reference's ray index space (value of Camera.make_rays without sampleDist,
     * scene.py:66-72), as filled by rrt_primary_rays().  NULL => the kernels evaluate
     * the float64 grid themselves.  Same bits either way; the table only saves the
     * float64 divide/sqrt chains per pixel (small images, batches of scenes).          */
    const float* base_rays;

    /* Global index of this call's scene 0 (scene-batch sharding across GPUs): only keys the
     * in-kernel jitter RNG, so that a sharded batch draws the same jitter as the whole one. */
    int32_t scene_begin;
    int32_t flags;        /* RRT_FLAG_* */

    /* Optional float32 table of RRT_RECORD_TABLE_FLOATS(B, N) floats, 16-byte aligned, as filled
     * by rrt_build_records() for THIS scene's current w2o / camera tables: first
     * [B][N][RRT_RECORD_FLOATS] sweep records (the per-object constants of the ray-object test:
     * A's diagonal, o' = A.c + b, 1 - o'.o', the off-diagonals), then [B][Npad][RRT_QUADRIC_FLOATS]
     * pre-filter rows (see RRT_FLAG_CANONICAL_SWEEP; Npad = N rounded up to a multiple of 4).
     * With it the render kernels stage the object table in shared memory by one TMA bulk copy
     * (cp.async.bulk + mbarrier) per 512-object chunk instead of rebuilding the records in
     * every CTA, and sphere-only chunks are swept with the pre-filter.  NULL => the kernels
     * build the records themselves.  Same bits either way.  Must be rebuilt whenever w2o or the
     * camera changes.                                                                          */
    const float* obj_records;

    /* Optional scratch, uint32 [num_scenes] in device memory: ZERO before its first use, left zero
     * by every call (the kernels reset what they used), never touched by the host afterwards.
     * With it the reverse-pass entry points are ONE launch: the last CTA to finish a scene
     * (device-side ticket + __threadfence) finalises that scene's gradient (d/dA = M C^T + g_b ct^T,
     * camera and light chains) inside the render kernel instead of a second launch.  Calls that
     * may run concurrently (different streams) need different scratch.  NULL => separate launch. */
    uint32_t* ticket;

    /* RRT_FLAG_DETERMINISTIC only: int64 [B][RRT_GRAD_SIZE(N) + 1][2] (RRT_DET_WORKSPACE_BYTES),
     * 16-byte aligned device scratch, zeroed by the callee. */
    int64_t* det_workspace;

    /* RRT_FLAG_MIRROR only: float32 [B][N] (or [N] with stride 0) mirror coefficient k in [0,1] per object */
    const float* reflectivity;
    int64_t reflectivity_scene_stride;
} rrt_scene;

#define RRT_DET_WORKSPACE_BYTES(num_scenes, num_objects) ((size_t)(num_scenes) * (RRT_GRAD_SIZE(num_objects) + 1) * 16)

#define RRT_RECORD_FLOATS 16
#define RRT_QUADRIC_FLOATS 6
#define RRT_RECORD_TABLE_FLOATS(num_scenes, num_objects) \
    ((size_t)(num_scenes) * ((size_t)(num_objects) * RRT_RECORD_FLOATS + (((size_t)(num_objects) + 3) / 4 * 4) * RRT_QUADRIC_FLOATS))

int rrt_version(void);
const char* rrt_last_error(void);

/*
 * Forward render.  Replaces Scene.build() + theano.function([], image)()
 * (scene.py:18-52; optimize_brightness.py:43-46).
 *   image      [B][rows][n][3] float32   mean over S samples, background 0
 *   hit_index  [B][S][rows][n] int32 or NULL: winning shape per ray, -1 = none
 *   tmin       [B][S][rows][n] float32 or NULL: min_dists (scene.py:47), +inf = none
 */
int rrt_render_forward(const rrt_scene* scene, float* image, int32_t* hit_index,
                       float* tmin, void* stream);

/*
 * Reverse pass.  Replaces T.grad(loss, params) through the render graph
 * (optimize.py:25,73; orbit_experiments/optimize.py:76) for an arbitrary upstream
 * gradient.
 *   dl_dimage  [B][rows][n][3] float32
 *   hit_index  as written by rrt_render_forward, or NULL to re-run the nearest-hit
 *              sweep (hit records themselves are always recomputed, never stored)
 *   grad       [B][RRT_GRAD_SIZE(N)] float32, zeroed by the callee then accumulated
 */
int rrt_render_backward(const rrt_scene* scene, const float* dl_dimage,
                        const int32_t* hit_index, float* grad, void* stream);

/*
 * Fused forward + squared-error loss + reverse pass in one kernel: hit records
 * stay in registers.  Covers cost = sum_c w_c * sum((image - target)^2)
 * (match_mirror.py:45; autoencoder.py:76; autoencoder_2ly.py:91; test_balls.py via
 * channel weights (1,0,0)).
 *   target          [B][rows][n][3] float32
 *   channel_weight  HOST float[3] or NULL (= 1,1,1)
 *   image, hit_index  optional outputs (NULL to skip the stores)
 *   loss            [B] float64, zeroed by the callee
 *   grad            [B][RRT_GRAD_SIZE(N)] float32, zeroed by the callee (ONE memset node instead of two when
 *                   loss lies directly in front of grad in one allocation: (char*)loss + 8*B == (char*)grad)
 */
int rrt_render_fused_mse(const rrt_scene* scene, const float* target,
                         const float* channel_weight, float* image, int32_t* hit_index,
                         double* loss, float* grad, void* stream);

/*
 * Fills rrt_scene.obj_records (RRT_RECORD_TABLE_FLOATS(B, N) floats) from scene->w2o, obj_type and
 * the camera translation (o' = A.c + b, transform.py:44; cc = o'.o' - 1, shape.py:79,82), and the
 * pre-filter rows derived from them in float64 (Q = A^T (o' o'^T - cc I) A, inflated by 2^-18 of
 * its scale; objects outside the range in which that bound is proven -- non-finite entries,
 * |A|_F or |o'| beyond 2^+-16 -- and squares get an always-pass row).
 * One tiny kernel; call it before the render entry points whenever the tables changed.
 */
int rrt_build_records(const rrt_scene* scene, float* records, void* stream);

/*
 * Camera.make_rays grid (scene.py:66-72: float64 linspace / normalise, cast to float32),
 * without jitter: out[i][j][0..2] for the reference's ray indices (i, j).  Fills the
 * optional rrt_scene.base_rays table.
 */
int rrt_primary_rays(int n, float* out, void* stream);

/*
 * Parameter -> matrix chain.  Replaces the symbolic transform algebra that feeds the
 * renderer: translate / scale / rotate (transform.py:60-122), Transform.__mul__ and
 * .inverse (transform.py:32-38) and, in rrt_chain_backward, T.grad through them.
 * One thread evaluates one chain = product of up to RRT_CHAIN_MAX_OPS primitive
 * matrices, e.g. shape.w2o = (translate(c) * rotate(a, axis) * scale(s)).inverse()
 * is the chain [scale^-1, rotate^-1, translate^-1].
 *   ops        [num_ops][4] int32 (device): kind | RRT_CHAIN_INVERT, arg0, arg1, unused
 *                kind RRT_CHAIN_TRANSLATE / _SCALE: 3 floats at values[arg0]
 *                kind RRT_CHAIN_ROTATE: angle (degrees) at values[arg0], axis[3] at values[arg1]
 *   chain_begin [num_chains+1] int32 (device): ops of chain k are [chain_begin[k], chain_begin[k+1])
 *   values     [num_values] float32 (device): constants and live parameters, gathered by the host
 *   out        [num_chains][12] float32: rows 0..2 of each product (row-major 3x4)
 *   g_out      [num_chains][12] float32: dL/d out;  g_values [num_values] float32, zeroed by the callee
 */
#define RRT_CHAIN_TRANSLATE 1
#define RRT_CHAIN_SCALE 2
#define RRT_CHAIN_ROTATE 3
#define RRT_CHAIN_INVERT 0x100
#define RRT_CHAIN_MAX_OPS 8
int rrt_chain_forward(const int32_t* ops, const int32_t* chain_begin, int num_chains, const float* values,
                      float* out, void* stream);
int rrt_chain_backward(const int32_t* ops, const int32_t* chain_begin, int num_chains, const float* values,
                       const float* g_out, float* g_values, int num_values, void* stream);

/*
 * A WHOLE optimise step of a small scene with a squared-error cost in ONE kernel launch -- the
 * counterpart of the single compiled Theano function the reference builds from
 * T.grad + updates (optimize.py:19-29) for the cost of match_mirror.py:45:
 *     w2o rows = chains(values)                 transform.py:32-38, 56-122 (as rrt_chain_forward)
 *     image, loss, d/d w2o                      as rrt_render_fused_mse (small-scene kernel)
 *     d/d values = chains^T(d/d w2o)            as rrt_chain_backward
 *     values[p] -= lr * d/d values[p]           for p >= param_begin   (optimize.py:26-27)
 * The last CTA to finish (device-side ticket) runs everything after the render and re-zeroes the
 * scratch buffers, so consecutive steps need no memsets and no other launches.  scene->w2o is not
 * read (the chains replace it); one scene, <= 32 shapes, power-of-two samples.
 */
typedef struct rrt_step {
    const int32_t* ops;          /* chain program of the scene's shapes, see rrt_chain_forward       */
    const int32_t* chain_begin;  /* [num_objects + 1]                                                */
    float* values;               /* [num_values] constants, then trainable parameters (updated)     */
    int32_t num_values;
    int32_t param_begin;         /* values[param_begin .. num_values) are trainable                  */
    float lr;
    int32_t reserved;
    float* grad;                 /* scratch [RRT_GRAD_SIZE(N)], ZERO before the first step           */
    float* g_values;             /* scratch [num_values], ZERO before the first step                 */
    double* loss_acc;            /* scratch [1], ZERO before the first step                          */
    float* loss_out;             /* [1]: the step's loss (evaluated before the update)               */
    uint32_t* ticket;            /* scratch [1], ZERO before the first step                          */
} rrt_step;
int rrt_small_step_mse(const rrt_scene* scene, const rrt_step* step, const float* target,
                       const float* channel_weight, float* image, void* stream);

/*
 * The one exchange step of the sharded path (the reference is single-process,
 * so there is no reference interface to cite): sum of the per-rank vector
 * [grad (n float32) | loss (nloss float64)] over the GPUs of one box, done by ONE kernel per
 * rank over NVLink peer memory (push to every peer's slot, per-source epoch flags, sum in
 * rank order => identical bits on every rank), instead of an NCCL allreduce.
 *   peer_buf   DEVICE array [world] of pointers: rank p's exchange buffer as mapped into THIS
 *              process (rrt_peer_buffer_bytes(n, nloss, world) bytes each, symmetric allocation,
 *              e.g. torch.distributed._symmetric_memory / cuMem + IPC)
 *   peer_sig   DEVICE array [world] of pointers: rank p's flag area (rrt_peer_signal_bytes()
 *              bytes, ZEROED once before first use, never touched by the host afterwards)
 *   out        [n + nloss] float64 (local): the sums
 * All ranks must call it the same number of times with the same sizes; the kernels of
 * different ranks wait for each other on the device (no host synchronisation).
 */
size_t rrt_peer_buffer_bytes(int n, int nloss, int world);
size_t rrt_peer_signal_bytes(void);
int rrt_peer_allreduce(const float* grad, const double* loss, int n, int nloss, void* const* peer_buf,
                       void* const* peer_sig, int rank, int world, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RRT_B200_H */
