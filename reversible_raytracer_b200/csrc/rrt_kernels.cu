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
#define RRT_MIN_BLOCKS 4   // 128 registers, no spills.  5 CTAs/SM (96 registers) run the same speed but their extra
                           // resident threads push local memory (per-ray arrays) out of L2: DRAM writes 377 MB vs
                           // the algorithmic 201 MB per C5 launch (profiles/r2_traffic_probe.txt)
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
    int small_per;     // render_small_kernel: work items (blocks of kSmallThreads rays) per persistent CTA
    int n_shift;       // log2(n) when n is a power of two, else -1
};

// The device code lives in the .cuh parts below, in dependency order (single translation unit).
#include "rrt_math.cuh"
#include "rrt_objects.cuh"
#include "rrt_sweep.cuh"
#include "rrt_shade.cuh"
#include "rrt_reduce.cuh"
#include "rrt_render_kernel.cuh"
#include "rrt_chain.cuh"
#include "rrt_small_kernel.cuh"
#include "rrt_aux_kernels.cuh"

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
    if (sc->flags & RRT_FLAG_MIRROR) {
        if (!sc->reflectivity) return fail(RRT_ERR_INVALID, "RRT_FLAG_MIRROR needs the reflectivity table");
        if (sc->shader == RRT_SHADER_DEPTH) return fail(RRT_ERR_UNSUPPORTED, "RRT_FLAG_MIRROR: Phong shaders only");
        if (sc->camera_grad || !sc->transpose)
            return fail(RRT_ERR_UNSUPPORTED, "RRT_FLAG_MIRROR: root camera variant only (identity camera.o2w, no camera gradient)");
    }
    if ((sc->flags & RRT_FLAG_DETERMINISTIC) && (!sc->det_workspace || ((uintptr_t)sc->det_workspace & 15)))
        return fail(RRT_ERR_INVALID, "RRT_FLAG_DETERMINISTIC needs a 16-byte aligned det_workspace");
    *rows_out = rows;
    return RRT_OK;
}

// Small scenes (few objects, few rays) take the one-ray-per-thread kernel (RRT_FLAG_NO_SMALL
// forces the general one).  Pure function of the descriptor: no environment, no static state.
bool use_small_kernel(const KParams& P) {
    const rrt_scene& sc = P.sc;
    const int S = sc.samples;
    if (sc.num_objects > kSmallMaxN || S > 32 || (S & (S - 1)) || (sc.flags & (RRT_FLAG_CULL | RRT_FLAG_NO_SMALL)))
        return false;
    const long long limit = kSmallDefaultMaxRays;
    const long long rays_scene = (long long)P.rows * sc.n * S;
    if (rays_scene >= 0x7fffffffLL - kSmallThreads) return false;    // 32-bit ray index in the kernel
    return rays_scene * sc.num_scenes <= limit;
}

int sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        sms = 148;      // B200
    return sms;
}

// The pixel-per-thread form of the small-scene kernel (one thread = one pixel and its S samples):
// all three modes, S in {1, 2, 4}, no shadows / mirror bounce.  Taken by default when the
// call has enough pixels to fill the machine with pixel threads (batches of scenes); a single small
// image keeps one ray per thread (4 x the threads, shorter dependency chains).
constexpr long long kPixelMinPixels = 192 * 1024;   // measured crossover on the orbit batch: 16 scene pairs (131 k pixels) ray threads, 32 pixel threads
bool use_pixel_threads(const KParams& P, int mode) {
    const rrt_scene& sc = P.sc;
    const int S = sc.samples;
    if (!(S == 1 || S == 2 || S == 4)) return false;
    if (sc.flags & (RRT_FLAG_SHADOWS | RRT_FLAG_MIRROR | RRT_FLAG_RAY_THREADS)) return false;
    if (sc.flags & RRT_FLAG_PIXEL_THREADS) return true;
    return (long long)P.rows * sc.n * sc.num_scenes >= kPixelMinPixels;
}

// Persistent grid of the small-scene kernel: at most one resident wave, every CTA >= 1 work item.
unsigned small_grid(KParams& P, bool pixel = false, int pixel_blocks = RRT_PIXEL_MIN_BLOCKS) {
    const rrt_scene& sc = P.sc;
    const long long rays_scene = (long long)P.rows * sc.n * sc.samples;
    long long items_scene = (rays_scene + kSmallThreads - 1) / kSmallThreads;
    if (pixel)          // tiles of 8 x 4 pixels, one per warp; a work item is a 2 x 2 block of tiles
        items_scene = (long long)((sc.n + 2 * kPixTileW - 1) / (2 * kPixTileW)) * ((P.rows + 2 * kPixTileH - 1) / (2 * kPixTileH));
    const long long total = items_scene * sc.num_scenes;
    const long long cap = (long long)sm_count() * (pixel ? pixel_blocks : RRT_SMALL_MIN_BLOCKS);
    long long per = (total + cap - 1) / cap;
    if (per < 1) per = 1;
    P.small_per = (int)per;
    P.n_shift = -1;
    if ((sc.n & (sc.n - 1)) == 0)
        for (int q = 0; q < 31; q++)
            if ((1 << q) == sc.n) P.n_shift = q;
    return (unsigned)((total + per - 1) / per);
}

// *finalized is set when the kernel itself finalised the gradients (rrt_scene.ticket given and
// supported by the kernel taken); otherwise the caller launches finalize_grads.
template <int MODE>
int launch(KParams& P, cudaStream_t st, bool* finalized = nullptr) {
    const rrt_scene& sc = P.sc;
    P.lin_step = sc.n > 1 ? 1.0 / (double)(sc.n - 1) : 0.0;
    P.inv_s = 1.0f / (float)sc.samples;
    P.inv_n = 1.0f / (float)sc.n;
    P.pow2 = ((sc.samples & (sc.samples - 1)) == 0) && ((sc.n & (sc.n - 1)) == 0);
    P.vec_ok = (sc.n % 4 == 0) && (((uintptr_t)P.image & 15) == 0) && (((uintptr_t)P.target & 15) == 0);
    const int S = sc.samples;
    if (finalized) *finalized = false;
    if (use_small_kernel(P)) {
        const bool geom = MODE != MODE_FWD && (sc.flags & RRT_FLAG_NO_MATERIAL_GRAD);
        const bool pixel = use_pixel_threads(P, MODE);
        const unsigned grid = small_grid(P, pixel, pixel_min_blocks(MODE, geom));
        if (pixel) {
            constexpr int PM = MODE;
            void (*kern)(const KParams) = nullptr;
            constexpr bool G = MODE != MODE_FWD;                         // (forward: one instantiation)
            if (geom) kern = S == 1 ? render_small_kernel<PM, false, G, false, 1> :
                             S == 2 ? render_small_kernel<PM, false, G, false, 2> : render_small_kernel<PM, false, G, false, 4>;
            else kern = S == 1 ? render_small_kernel<PM, false, false, false, 1> :
                        S == 2 ? render_small_kernel<PM, false, false, false, 2> : render_small_kernel<PM, false, false, false, 4>;
            kern<<<grid, kSmallThreads, 0, st>>>(P);
        } else if (sc.flags & RRT_FLAG_MIRROR) {
            if (geom) render_small_kernel<MODE, false, true, true><<<grid, kSmallThreads, 0, st>>>(P);
            else render_small_kernel<MODE, false, false, true><<<grid, kSmallThreads, 0, st>>>(P);
        } else if (geom) {
            render_small_kernel<MODE, false, true><<<grid, kSmallThreads, 0, st>>>(P);
        } else {
            render_small_kernel<MODE><<<grid, kSmallThreads, 0, st>>>(P);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "small-scene kernel launch: %s", cudaGetErrorString(e));
        if (finalized) *finalized = (MODE != MODE_FWD) && sc.ticket != nullptr;
        return RRT_OK;
    }
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
    if (sc.flags & RRT_FLAG_MIRROR) {
        if (S == 1) kern = render_kernel<kRays, 1, MODE, true>;
        else if (S == 2) kern = render_kernel<kRays / 2, 2, MODE, true>;
        else if (S == 4) kern = render_kernel<kRays / 4, 4, MODE, true>;
        else kern = render_kernel<1, kRays, MODE, true>;
    } else if (S == 1) kern = render_kernel<kRays, 1, MODE>;
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
    if (finalized) *finalized = (MODE != MODE_FWD) && sc.ticket != nullptr;
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
    if (e == cudaSuccess && (scene->flags & RRT_FLAG_DETERMINISTIC))
        e = cudaMemsetAsync(scene->det_workspace, 0, RRT_DET_WORKSPACE_BYTES(scene->num_scenes, scene->num_objects), st);
    if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "memset grad: %s", cudaGetErrorString(e));
    bool finalized = false;
    rc = launch<MODE_BWD>(P, st, &finalized);
    if (rc) return rc;
    return finalized ? RRT_OK : launch_finalize(P, st);
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
    // zero-initialise the outputs: ONE memset when the caller carved loss and grad out of one allocation
    // (loss directly in front of grad), else two
    const size_t gbytes = sizeof(float) * RRT_GRAD_SIZE(scene->num_objects) * scene->num_scenes;
    const size_t lbytes = sizeof(double) * scene->num_scenes;
    cudaError_t e;
    if ((const char*)loss + lbytes == (const char*)grad) {
        e = cudaMemsetAsync(loss, 0, lbytes + gbytes, st);
    } else {
        e = cudaMemsetAsync(grad, 0, gbytes, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(loss, 0, lbytes, st);
    }
    if (e == cudaSuccess && (scene->flags & RRT_FLAG_DETERMINISTIC))
        e = cudaMemsetAsync(scene->det_workspace, 0, RRT_DET_WORKSPACE_BYTES(scene->num_scenes, scene->num_objects), st);
    if (e != cudaSuccess) return fail(RRT_ERR_CUDA, "memset grad/loss: %s", cudaGetErrorString(e));
    bool finalized = false;
    rc = launch<MODE_FUSED>(P, st, &finalized);
    if (rc) return rc;
    return finalized ? RRT_OK : launch_finalize(P, st);
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
    if (scene->flags & (RRT_FLAG_DETERMINISTIC | RRT_FLAG_MIRROR))
        return fail(RRT_ERR_UNSUPPORTED, "whole-step kernel: RRT_FLAG_DETERMINISTIC / RRT_FLAG_MIRROR are not supported");
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
    render_small_kernel<MODE_FUSED, true><<<small_grid(P), kSmallThreads, 0, (cudaStream_t)stream>>>(P);
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

}  // extern "C"
