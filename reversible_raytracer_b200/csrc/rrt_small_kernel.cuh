// rrt_small_kernel.cuh -- part of rrt_kernels.cu (one translation unit; included inside its anonymous namespace).
// render_small_kernel<MODE,STEP,GEOM,MIRROR,SPP>: small scenes, persistent CTAs; one ray per thread (SPP = 0) or
// one pixel and its S = SPP samples per thread (SPP in {1, 2, 4}); STEP = the whole optimise step in one launch;
// GEOM = RRT_FLAG_NO_MATERIAL_GRAD (only the 12 transform sums per object are reduced); MIRROR = one bounce.
#pragma once

// ---------------------------------------------------------------- the small-scene kernel
// Latency-oriented variant for the reference's own workloads (optimize_brightness.py,
// match_mirror.py, test_balls.py, the orbit decoder: 32..128 pixels a side, 2..3 shapes).
// There the 8-rays-per-thread kernel is a serial dependency chain on a grid that cannot fill
// the machine; here ONE thread owns ONE ray, the S samples of a pixel sit in adjacent lanes
// (S a power of two <= 32) and are combined by shuffles, and the whole object table
// (<= kSmallMaxN records) plus the materials live in shared memory.  Same device routines
// (obj_test / shade / backward_ray), same canonical order, same sample summation order as
// render_kernel, so the two kernels agree bit for bit on masks.
//
// Work items are blocks of kSmallThreads consecutive rays of one scene (pixel-per-thread form: blocks of 16 x 8
// pixels, see below), numbered scene-major.
// The grid is PERSISTENT: at most (SM count x resident CTAs) CTAs, each walking a contiguous
// range of `small_per` items.  Everything that used to be paid per 128 rays is paid once per
// CTA and scene: the object-table build and its barrier, the CTA-level reduction barrier and
// the global atomics (the orbit batch: 65 536 items -> ~1 200 CTAs touching <= 2 scenes each).
// Per-object gradient sums stay in the thread's registers across items, keyed by the winning
// object, and go through the warp butterfly only when a lane's winner changes.
constexpr int kSmallMaxN = 32;
constexpr int kSmallThreads = 128;
#ifndef RRT_SMALL_MIN_BLOCKS
#define RRT_SMALL_MIN_BLOCKS 6   // 80 registers; measured on the orbit batch (geom-only gradients): 8 CTAs/SM
                                 // (64 registers, spills) 161 us, 6 CTAs/SM 148 us
#endif
constexpr long long kSmallDefaultMaxRays = 16 << 20;  // total rays of a call (all scenes) up to which it is used

// The running sums of thread `tid` live in shared memory (column tid of acc_s[NACC][kSmallThreads],
// conflict-free) instead of registers: they are touched only by the few rays that hit something,
// and 12..19 registers held across the whole item loop cost a fifth of the resident warps.
template <int NACC>
__device__ __forceinline__ void small_warp_flush(int key, float* acc_col, float* slots, int lane, long long* det_ws = nullptr) {
    const unsigned full = 0xffffffffu;
    float acc[NACC];
#pragma unroll
    for (int v = 0; v < NACC; v++) { acc[v] = acc_col[v * kSmallThreads]; acc_col[v * kSmallThreads] = 0.f; }
    unsigned todo = __ballot_sync(full, key >= 0);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const int k = __shfl_sync(full, key, leader);
        const bool mine = (key == k);
        todo &= ~__ballot_sync(full, mine);
        int v;
        const float x = warp_reduce_n<NACC>(acc, mine, lane, v);
        if (v >= 0 && x != 0.f) {
            if (det_ws) det_add(det_ws + ((size_t)k * RRT_OBJ_GRAD_STRIDE + v) * 2, (double)x);   // RRT_FLAG_DETERMINISTIC
            else atomicAdd(&slots[k * kSlotStride + v], x);
        }
    }
}

// SPP > 0 selects the PIXEL-PER-THREAD form (batches of small scenes: the orbit decoder batch is 2.1 M
// pixels): one thread owns one pixel and its SPP = S anti-alias samples (S in {1, 2, 4}), a warp
// covers a compact tile of 8 x 4 pixels (a warp meets an object's silhouette about half as often as
// with a 32 x 1 strip, and the shading / reverse-pass code is only entered by warps that hit
// something), a work item is a 2 x 2 block of tiles (one per warp).  What the ray-per-thread form pays
// per RAY is paid per PIXEL: index arithmetic, the camera-space grid ray, the target load, the loss,
// the image store; the pixel mean is a register sum in sample order (same bits as the shuffles of
// the ray-per-thread form); the S samples of a pixel give the sweep S independent dependency
// chains.  Same device routines, same canonical order => same masks.  All three modes; no shadows /
// mirror / whole-step (those keep SPP = 0).
constexpr int kPixTileW = 8, kPixTileH = 4;                 // pixels per warp: 8 x 4
// a[q] for a run-time q out of a register array (select chain instead of local memory): lets the
// per-sample shading / reverse-pass loops stay ROLLED -- one copy of that code instead of S, which is
// what keeps the kernel inside the instruction cache (S unrolled copies: 56 % I-cache hit rate, 7
// "no instruction" stall cycles per issue, measured)
template <typename T, int NQ>
__device__ __forceinline__ void put(T (&a)[NQ], int q, T v) {
#pragma unroll
    for (int i = 0; i < NQ; i++) a[i] = (q == i) ? v : a[i];
}
__device__ __noinline__ void jitter_offset_pair(float ux, float uy, int s, int S, int n, float& ox, float& oy) {
    ox = jitter_offset(ux, s, S, n);
    oy = jitter_offset(uy, s, S, n);
}
template <typename T, int NQ>
__device__ __forceinline__ T pick(const T (&a)[NQ], int q) {
    T v = a[0];
#pragma unroll
    for (int i = 1; i < NQ; i++) v = (q == i) ? a[i] : v;
    return v;
}
// resident CTAs per SM of the pixel-per-thread form (measured on the orbit batch): 5 (<= 102 registers) for the
// forward and the d/d w2o-only kernels, 4 (<= 128) when all 19 + 9 gradient sums are live (98 / 119 us; 5: 98 / 130 us)
#ifndef RRT_PIXEL_MIN_BLOCKS
#define RRT_PIXEL_MIN_BLOCKS 5
#endif
#ifndef RRT_PIXEL_MIN_BLOCKS_ALL
#define RRT_PIXEL_MIN_BLOCKS_ALL 4
#endif
constexpr int pixel_min_blocks(int mode, bool geom) { return (mode == MODE_FWD || geom) ? RRT_PIXEL_MIN_BLOCKS : RRT_PIXEL_MIN_BLOCKS_ALL; }

template <int MODE, bool STEP = false, bool GEOM = false, bool MIRROR = false, int SPP = 0>
__global__ void __launch_bounds__(kSmallThreads, SPP > 0 ? pixel_min_blocks(MODE, GEOM) : RRT_SMALL_MIN_BLOCKS)
render_small_kernel(const __grid_constant__ KParams P) {
    static_assert(SPP == 0 || (!STEP && !MIRROR), "pixel-per-thread form: no whole-step / mirror variants");
    constexpr int NACC = GEOM ? 12 : 19;
    __shared__ float4 tab[kSmallMaxN * 4];
    __shared__ float mat_s[kSmallMaxN * RRT_MAT_STRIDE];
    __shared__ Globals g;
    __shared__ float slots[kSmallMaxN * kSlotStride];
    __shared__ float gglob[9];
    __shared__ float loss_warp[kSmallThreads / 32];
    __shared__ float camg_s[48];
    __shared__ int flag_s;
    __shared__ float acc_s[NACC * kSmallThreads];                 // per-thread running sums, [value][thread]
    __shared__ float gg_s[(GEOM ? 1 : 9) * kSmallThreads];        // light / look_at sums, [value][thread]

    const rrt_scene& sc = P.sc;
    const int n = sc.n, N = sc.num_objects, S = sc.samples;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    // 32-bit index arithmetic (the launcher guarantees rows*n*S < 2^31 and a total below 2^31; S is a power of two)
    const unsigned rays_scene = (unsigned)P.rows * (unsigned)n * (unsigned)S;
    // pixel-per-thread form: tiles of 8 x 4 pixels, 4 tiles (one per warp) per work item
    // (a work item = the 2 x 2 block of tiles of its 4 warps, 16 x 8 pixels: compact, so the warps of a CTA
    // meet an object together and reach the scene barriers together)
    const unsigned items_x = ((unsigned)n + 2 * kPixTileW - 1) / (2 * kPixTileW);
    const unsigned items_y = ((unsigned)P.rows + 2 * kPixTileH - 1) / (2 * kPixTileH);
    const unsigned bps = SPP > 0 ? items_x * items_y
                                 : (rays_scene + kSmallThreads - 1) / kSmallThreads;      // work items per scene
    const unsigned total = bps * (unsigned)sc.num_scenes;
    const unsigned item0 = blockIdx.x * (unsigned)P.small_per;
    const unsigned item1 = min(total, item0 + (unsigned)P.small_per);
    const float inv = 1.0f / (float)S;
    const float inf = __int_as_float(0x7f800000);
    const int s = (int)((unsigned)tid & (unsigned)(S - 1));
    const int sshift = 31 - __clz(S);
    const bool use_stored = (MODE == MODE_BWD) && (P.hit_in != nullptr);

    // running per-thread sums of the reverse pass, keyed by the winning object
    float* const acc_col = acc_s + tid;
    float* const gg_col = gg_s + tid;
    if (MODE != MODE_FWD) {
#pragma unroll
        for (int v = 0; v < NACC; v++) acc_col[v * kSmallThreads] = 0.f;
        if (!GEOM) {
#pragma unroll
            for (int v = 0; v < 9; v++) gg_col[v * kSmallThreads] = 0.f;
        }
    }
    int acc_key = -1;
    float loss_part = 0.f;
    int cur_scene = -1;
    unsigned scene_items = 0;                          // items of cur_scene processed by this CTA
    float* gobj = nullptr;
    long long* det_ws = nullptr;                       // RRT_FLAG_DETERMINISTIC: fixed-point sums of cur_scene
    // RRT_FLAG_MIRROR (opt-in extension): scene tables the secondary rays read; requires the identity camera
    const bool mirror_on = MIRROR && !STEP && (sc.flags & RRT_FLAG_MIRROR) && sc.reflectivity && sc.shader != RRT_SHADER_DEPTH;
    const float* w2o_s = nullptr;
    const float* mats_g = nullptr;
    const float* refl_s = nullptr;

    // ---- end of a scene's items on this CTA: thread -> warp -> CTA -> global, ticket, finalisation
    auto flush_scene = [&](bool reduced) {
        if (MODE != MODE_FWD && reduced) {
            if (__any_sync(full, acc_key >= 0)) small_warp_flush<NACC>(acc_key, acc_col, slots, lane, det_ws);
            acc_key = -1;
            if (!GEOM) {
#pragma unroll
                for (int v = 0; v < 9; v++) {
                    const float x = warp_sum(gg_col[v * kSmallThreads]);
                    if (lane == 0 && x != 0.f) {
                        if (det_ws) det_add(det_ws + ((size_t)N * RRT_OBJ_GRAD_STRIDE + (v < 6 ? v : 12 + v)) * 2, (double)x);
                        else atomicAdd(&gglob[v], x);
                    }
                    gg_col[v * kSmallThreads] = 0.f;
                }
            }
            if (MODE == MODE_FUSED) {
                const float x = warp_sum(loss_part);
                if (lane == 0) {
                    loss_warp[warp] = x;
                    if (det_ws && x != 0.f) det_add(det_ws + (size_t)RRT_GRAD_SIZE(N) * 2, (double)x);
                }
                loss_part = 0.f;
            }
            __syncthreads();
            if (!det_ws) {
            for (int q = tid; q < N * NACC; q += kSmallThreads) {
                const int k = q / NACC, v = q - k * NACC;
                const float x = slots[k * kSlotStride + v];
                if (x != 0.f) atomicAdd(&gobj[(size_t)k * RRT_OBJ_GRAD_STRIDE + v], x);
            }
            float* gglobal = gobj + (size_t)N * RRT_OBJ_GRAD_STRIDE;
            if (!GEOM && tid < 9) {
                const int dst = tid < 6 ? tid : 12 + tid;    // Lhat 0..2, intensity 3..5, look_at 18..20
                if (gglob[tid] != 0.f) atomicAdd(&gglobal[dst], gglob[tid]);
            }
            if (MODE == MODE_FUSED && tid == 0) {
                double t = 0.0;
                for (int w = 0; w < kSmallThreads / 32; w++) t += (double)loss_warp[w];
                if (t != 0.0) atomicAdd(&P.loss[cur_scene], t);
            }
            }
        }
        if (MODE != MODE_FWD && !STEP && sc.ticket) {
            // the last CTA to finish this scene turns the raw sums into d/d w2o, camera and light
            // gradients (finalize_scene): one launch per reverse pass
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const unsigned old = atomicAdd(sc.ticket + cur_scene, scene_items);
                flag_s = (old + scene_items == bps);
            }
            __syncthreads();
            if (flag_s) {
                __threadfence();
                finalize_scene(P, cur_scene, tid, kSmallThreads, camg_s);
                if (tid == 0) sc.ticket[cur_scene] = 0u;
            }
        }
    };

    // BWD with one item per CTA (single small image): a CTA none of whose pixels carries upstream
    // gradient (optimize_brightness.py:51 touches two pixels) skips the table build altogether
    bool skip_cta = false;
    if (MODE == MODE_BWD && SPP == 0 && item1 - item0 == 1) {
        const unsigned scene0 = item0 / bps, gid0 = (item0 - scene0 * bps) * kSmallThreads + tid;
        bool nz = false;
        if (gid0 < rays_scene) {
            const size_t po0 = ((size_t)scene0 * P.rows * n + (gid0 >> sshift)) * 3;
            nz = (P.dl_dimage[po0] != 0.f) | (P.dl_dimage[po0 + 1] != 0.f) | (P.dl_dimage[po0 + 2] != 0.f);
        }
        skip_cta = !__syncthreads_or(nz);
        if (skip_cta) {
            cur_scene = (int)scene0;
            scene_items = 1;
            flush_scene(false);
            return;
        }
    }

    // (scene, blk) of the current item, advanced incrementally (no division in the loop)
    int scene = (int)(item0 / bps);
    unsigned blk = item0 - (unsigned)scene * bps;
    const float* tgt_s = nullptr;      // per-scene bases: 32-bit offsets inside a scene
    const float* dl_s = nullptr;
    float* img_s = nullptr;
#pragma unroll 1
    for (unsigned item = item0;; item++, blk++) {
        // (one call site for the scene flush -- the loop also runs once past the last item: the flush
        // inlines the whole gradient finalisation, and this kernel lives or dies by its code size)
        const bool done = item >= item1;
        if (!done && blk == bps) { blk = 0; scene++; }
        if (done || scene != cur_scene) {                  // CTA-uniform
            if (cur_scene >= 0) {
                flush_scene(true);
                if (!done) __syncthreads();                // everybody is done with the old tables / slots
            }
            if (done) break;
            cur_scene = scene;
            scene_items = 0;
            gobj = (MODE != MODE_FWD) ? P.grad + (size_t)scene * RRT_GRAD_SIZE(N) : nullptr;
            det_ws = (MODE != MODE_FWD && !STEP && (sc.flags & RRT_FLAG_DETERMINISTIC)) ? det_scene(sc, scene) : nullptr;
            {
                const size_t so = (size_t)scene * P.rows * n * 3;
                tgt_s = (MODE == MODE_FUSED) ? P.target + so : nullptr;
                dl_s = (MODE == MODE_BWD) ? P.dl_dimage + so : nullptr;
                img_s = (MODE != MODE_BWD && P.image) ? P.image + so : nullptr;
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
            w2o_s = w2o;
            mats_g = mats;
            refl_s = mirror_on ? sc.reflectivity + (size_t)scene * sc.reflectivity_scene_stride : nullptr;
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
                        g.cam_identity = (cam[0] == 1.f && cam[5] == 1.f && cam[10] == 1.f && cam[1] == 0.f && cam[2] == 0.f &&
                                          cam[4] == 0.f && cam[6] == 0.f && cam[8] == 0.f && cam[9] == 0.f);
                    }
                }
            }
            __syncthreads();
        }
        scene_items++;

        if (SPP > 0) {
            // ================= pixel-per-thread form: this thread's pixel and its S = SPP samples
            constexpr int SP = SPP > 0 ? SPP : 1;
            const unsigned iy = blk / items_x, ix = blk - iy * items_x;
            const int al = (int)((2u * iy + ((unsigned)warp >> 1)) * kPixTileH) + (lane >> 3);
            const int b = (int)((2u * ix + ((unsigned)warp & 1u)) * kPixTileW) + (lane & 7);
            const bool active = (al < P.rows) && (b < n);
            const int a = sc.row_begin + al;
            const unsigned pl = active ? (unsigned)al * (unsigned)n + (unsigned)b : 0u;
            const unsigned po = pl * 3u;

            float gc[3] = {0.f, 0.f, 0.f};                       // upstream gradient of the pixel, per sample
            if (MODE == MODE_BWD) {
                if (active) { gc[0] = dl_s[po] * inv; gc[1] = dl_s[po + 1] * inv; gc[2] = dl_s[po + 2] * inv; }
                // sparse upstream gradients: a warp without any contributes exactly zero
                if (!__any_sync(full, (gc[0] != 0.f) | (gc[1] != 0.f) | (gc[2] != 0.f))) continue;
            }

            // ---- rays: the grid ray once per pixel, jitter per sample (scene.py:24-32,66-74)
            float rcx[SP], rcy[SP], rcz = 0.f;
#pragma unroll
            for (int q = 0; q < SP; q++) rcx[q] = rcy[q] = 0.f;
            if (active) {
                const int i = sc.transpose ? b : a, j = sc.transpose ? a : b;
                float bx, by;
                if (sc.base_rays) {
                    const float* br = sc.base_rays + ((size_t)i * n + j) * 3;
                    bx = __ldg(br); by = __ldg(br + 1); rcz = __ldg(br + 2);
                } else {
                    base_ray(n, P.lin_step, i, j, bx, by, rcz);
                }
#pragma unroll
                for (int q = 0; q < SP; q++) {
                    float jx, jy;
                    if (sc.jitter_x) {
                        const size_t off = (size_t)scene * sc.jitter_scene_stride + (size_t)pl * SP + q;   // [rows][n][S]
                        jx = __ldg(sc.jitter_x + off);
                        jy = __ldg(sc.jitter_y + off);
                    } else {
                        jx = rrt_rng(sc.seed, scene + sc.scene_begin, (uint32_t)(a * n + b), q, 0);
                        jy = rrt_rng(sc.seed, scene + sc.scene_begin, (uint32_t)(a * n + b), q, 1);
                    }
                    float ox, oy;
                    if (P.pow2) { ox = jitter_offset_pow2(jx, q, P.inv_s, P.inv_n); oy = jitter_offset_pow2(jy, q, P.inv_s, P.inv_n); }
                    else jitter_offset_pair(jx, jy, q, SP, n, ox, oy);      // (out of line: four IEEE divisions)
                    rcx[q] = __fadd_rn(bx, ox);
                    rcy[q] = __fadd_rn(by, oy);
                }
            }
            float tgt[3] = {0.f, 0.f, 0.f};
            if (MODE == MODE_FUSED && active) { tgt[0] = __ldg(tgt_s + po); tgt[1] = __ldg(tgt_s + po + 1); tgt[2] = __ldg(tgt_s + po + 2); }
            const bool cam_id = g.cam_identity;
            // camera.o2w (orbit_experiments/scene.py:80); identity => the fma chain returns its input
            auto world = [&](float rx, float ry, float& wx, float& wy, float& wz) {
                wx = rx; wy = ry; wz = rcz;
                if (!cam_id) {
                    wx = dot3_canon(g.C[0], g.C[1], g.C[2], rx, ry, rcz);
                    wy = dot3_canon(g.C[3], g.C[4], g.C[5], rx, ry, rcz);
                    wz = dot3_canon(g.C[6], g.C[7], g.C[8], rx, ry, rcz);
                }
            };

            // ---- nearest hit per sample: list order, strict '<' (scene.py:46-47); objects outer, samples inner
            float tmin[SP];
            int idx[SP];
#pragma unroll
            for (int q = 0; q < SP; q++) { tmin[q] = inf; idx[q] = -1; }
            if (use_stored) {
                if (active) {
#pragma unroll
                    for (int q = 0; q < SP; q++) {
                        const int kk = P.hit_in[(((size_t)scene * SP + q) * P.rows + al) * n + b];
                        idx[q] = (kk >= 0 && kk < N) ? kk : -1;   // never trust an index buffer blindly; a winner flagged
                    }                                             // RRT_HIT_SHADOWED (>= N) carries no gradient
                }
            } else if (active) {
                float wx[SP], wy[SP], wz[SP];
#pragma unroll
                for (int q = 0; q < SP; q++) world(rcx[q], rcy[q], wx[q], wy[q], wz[q]);
#pragma unroll 1
                for (int k = 0; k < N; k++) {
                    Obj ob;
                    load_rec(tab + 4 * k, ob);
                    if (ob.flags & 1) {                    // Square: the scalar routine, one sample at a time
#pragma unroll 1
                        for (int q = 0; q < SP; q++) {
                            HitRec h;
                            const float t = obj_test<true>(ob, pick(wx, q), pick(wy, q), pick(wz, q), h);
                            if (t < pick(tmin, q)) put(tmin, q, t), put(idx, q, k);
                        }
                        continue;
                    }
                    // Sphere: the S discriminants side by side (independent chains; the operations and their
                    // order are obj_test's: d' = A d, vn = d'.d', pd = d'.o', det = fma(pd, pd, vn * -cc)) ...
                    float vn[SP], pd[SP], det[SP];
                    bool cand = false;
#pragma unroll
                    for (int q = 0; q < SP; q++) {
                        float d0, d1, d2;
                        if (!(ob.flags & 2)) {
                            d0 = __fmul_rn(ob.a[0], wx[q]); d1 = __fmul_rn(ob.a[4], wy[q]); d2 = __fmul_rn(ob.a[8], wz[q]);
                        } else {
                            d0 = dot3_canon(ob.a[0], ob.a[1], ob.a[2], wx[q], wy[q], wz[q]);
                            d1 = dot3_canon(ob.a[3], ob.a[4], ob.a[5], wx[q], wy[q], wz[q]);
                            d2 = dot3_canon(ob.a[6], ob.a[7], ob.a[8], wx[q], wy[q], wz[q]);
                        }
                        vn[q] = dot3_canon(d0, d1, d2, d0, d1, d2);
                        pd[q] = dot3_canon(d0, d1, d2, ob.o[0], ob.o[1], ob.o[2]);
                        det[q] = __fmaf_rn(pd[q], pd[q], __fmul_rn(vn[q], ob.ncc));
                        cand |= det[q] > 0.0f;
                    }
                    // ... and the first root only where det > 0 (shape.py:121-125), one copy of the sqrt / divide
                    if (cand) {
#pragma unroll 1
                        for (int q = 0; q < SP; q++) {
                            const float dq = pick(det, q);
                            if (dq > 0.0f) {
                                const float t = __fdiv_rn(__fsub_rn(-pick(pd, q), __fsqrt_rn(dq)), pick(vn, q));
                                if (t < pick(tmin, q)) put(tmin, q, t), put(idx, q, k);
                            }
                        }
                    }
                }
                if (MODE != MODE_BWD && (P.hit_out || (MODE == MODE_FWD && P.tmin_out))) {
#pragma unroll
                    for (int q = 0; q < SP; q++) {
                        const size_t ro = (((size_t)scene * SP + q) * P.rows + al) * n + b;
                        if (P.hit_out) P.hit_out[ro] = idx[q];
                        if (MODE == MODE_FWD && P.tmin_out) P.tmin_out[ro] = tmin[q];
                    }
                }
            }

            // ---- shade the winners; pixel = mean over the samples, summed in sample order (scene.py:49-50)
            float sum[3] = {0.f, 0.f, 0.f};
            bool any_hit = false;
#pragma unroll
            for (int q = 0; q < SP; q++) any_hit |= (idx[q] >= 0);
            if (MODE != MODE_BWD && __any_sync(full, any_hit)) {
#pragma unroll 1
                for (int q = 0; q < SP; q++) {
                    float rgb[3] = {0.f, 0.f, 0.f};
                    const int kq = pick(idx, q);
                    if (kq >= 0) {
                        Obj ob;
                        HitRec h;
                        ShadeRec sr;
                        float m7[7], wx, wy, wz;
                        load_rec(tab + 4 * kq, ob);
                        world(pick(rcx, q), pick(rcy, q), wx, wy, wz);
                        hit_record<false>(ob, wx, wy, wz, pick(tmin, q), h);
#pragma unroll
                        for (int v = 0; v < 7; v++) m7[v] = mat_s[kq * RRT_MAT_STRIDE + v];
                        shade(sc.shader, sc.max_depth, ob, m7, g, h, sr, rgb);
                    }
#pragma unroll
                    for (int c = 0; c < 3; c++) sum[c] += rgb[c];
                }
            }
            const float v0 = sum[0] * inv, v1 = sum[1] * inv, v2 = sum[2] * inv;
            if (MODE != MODE_BWD && active && img_s) { img_s[po] = v0; img_s[po + 1] = v1; img_s[po + 2] = v2; }
            if (MODE != MODE_FWD) {
                if (MODE == MODE_FUSED && active)
                    pixel_cost(sc.flags & RRT_FLAG_LINEAR_COST, P.cw, inv, v0, v1, v2, tgt, loss_part, gc);
                // ---- reverse pass through the winners: the samples of a pixel usually share their
                // winner, so their sums meet in registers and reach the thread's column once per pixel
                const bool gnz = (gc[0] != 0.f) | (gc[1] != 0.f) | (gc[2] != 0.f);
                if (__any_sync(full, gnz && any_hit)) {
                    float da[NACC], dg[9];
#pragma unroll
                    for (int v = 0; v < NACC; v++) da[v] = 0.f;
#pragma unroll
                    for (int v = 0; v < 9; v++) dg[v] = 0.f;
                    bool pend = false, touched = false;
#pragma unroll 1
                    for (int q = 0; q < SP; q++) {
                        int key = gnz ? pick(idx, q) : -1;
                        if (use_stored && key >= 0) {
                            // reverse-only entry point with stored winners: the ray parameter comes from the
                            // canonical test of the stored object; a stale winner (no hit any more) carries nothing
                            Obj ob0;
                            HitRec h0;
                            float wx, wy, wz;
                            load_rec(tab + 4 * key, ob0);
                            world(pick(rcx, q), pick(rcy, q), wx, wy, wz);
                            const float t0 = obj_test<true>(ob0, wx, wy, wz, h0);
                            if (t0 < inf) put(tmin, q, t0);
                            else key = -1;
                        }
                        const bool change = (key >= 0) && (acc_key >= 0) && (key != acc_key);
                        if (__any_sync(full, change)) {          // some lane's winner changed: flush the warp's sums
                            if (pend) {
#pragma unroll
                                for (int v = 0; v < NACC; v++) { acc_col[v * kSmallThreads] += da[v]; da[v] = 0.f; }
                                pend = false;
                            }
                            small_warp_flush<NACC>(acc_key, acc_col, slots, lane, det_ws);
                            acc_key = -1;
                        }
                        if (key >= 0) {
                            Obj ob;
                            HitRec h;
                            ShadeRec sr;
                            float m7[7], rgb[3], wx, wy, wz;
                            const float rx = pick(rcx, q), ry = pick(rcy, q);
                            load_rec(tab + 4 * key, ob);
                            world(rx, ry, wx, wy, wz);
                            hit_record<true>(ob, wx, wy, wz, pick(tmin, q), h);   // t is known (sweep, or the test above)
#pragma unroll
                            for (int v = 0; v < 7; v++) m7[v] = mat_s[key * RRT_MAT_STRIDE + v];
                            shade(sc.shader, sc.max_depth, ob, m7, g, h, sr, rgb);
                            const float rc3[3] = {rx, ry, rcz};
                            backward_ray<GEOM, NACC>(sc.shader, sc.max_depth, ob, m7, g, h, sr, rc3, gc, da, dg);
                            acc_key = key;
                            pend = touched = true;
                        }
                    }
                    if (pend) {
#pragma unroll
                        for (int v = 0; v < NACC; v++) acc_col[v * kSmallThreads] += da[v];
                    }
                    if (!GEOM && touched) {            // light / look_at sums are not keyed by object
#pragma unroll
                        for (int v = 0; v < 9; v++) gg_col[v * kSmallThreads] += dg[v];
                    }
                }
            }
            continue;
        }
        const bool mir = MIRROR && mirror_on && g.cam_identity && g.ct[0] == 0.f && g.ct[1] == 0.f && g.ct[2] == 0.f;

        const unsigned gid = blk * kSmallThreads + tid;      // ray index within the scene: [rows][n][S]
        const bool active = gid < rays_scene;
        const unsigned pl = gid >> sshift;                   // slab-local pixel index
        const int al = active ? (P.n_shift >= 0 ? (int)(pl >> P.n_shift) : (int)(pl / (unsigned)n)) : 0;
        const int b = active ? (int)(pl - (unsigned)al * (unsigned)n) : 0;
        const int a = sc.row_begin + al;
        const unsigned po = pl * 3u;                         // offset of the pixel inside its scene

        float gc[3] = {0.f, 0.f, 0.f};
        if (MODE == MODE_BWD) {
            if (active) { gc[0] = dl_s[po] * inv; gc[1] = dl_s[po + 1] * inv; gc[2] = dl_s[po + 2] * inv; }
            // sparse upstream gradients: a warp without any contributes exactly zero
            if (!__any_sync(full, (gc[0] != 0.f) | (gc[1] != 0.f) | (gc[2] != 0.f))) continue;
        }

        // ---- this thread's ray
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
        if (MODE == MODE_FUSED && active) { tgt[0] = __ldg(tgt_s + po); tgt[1] = __ldg(tgt_s + po + 1); tgt[2] = __ldg(tgt_s + po + 2); }
        int stored = -1;
        if (use_stored && active) {
            const int kk = P.hit_in[(((size_t)scene * S + s) * P.rows + al) * n + b];
            stored = (kk >= 0 && kk < N) ? kk : -1;          // never trust an index buffer blindly; a winner
        }                                                    // flagged RRT_HIT_SHADOWED (>= N) carries no gradient

        // camera.o2w (orbit_experiments/scene.py:80); for C = I (root variant, translated orbit cameras)
        // the fma chain returns its input
        float wx = rcx, wy = rcy, wz = rcz;
        if (!g.cam_identity) {
            wx = dot3_canon(g.C[0], g.C[1], g.C[2], rcx, rcy, rcz);
            wy = dot3_canon(g.C[3], g.C[4], g.C[5], rcx, rcy, rcz);
            wz = dot3_canon(g.C[6], g.C[7], g.C[8], rcx, rcy, rcz);
        }

        // ---- nearest hit: list order, strict '<' (scene.py:46-47)
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
            if (use_stored) {
                obj_test<true>(ob, wx, wy, wz, h);
                if (!(h.t < inf)) idx = -1;                  // stale stored winner
            } else {
                hit_record<MODE != MODE_FWD>(ob, wx, wy, wz, tmin, h);   // t is known from the sweep
            }
        }
        if (idx >= 0) {
#pragma unroll
            for (int q = 0; q < 7; q++) m7[q] = mat_s[idx * RRT_MAT_STRIDE + q];
            shade(sc.shader, sc.max_depth, ob, m7, g, h, sr, rgb);
            if (MIRROR && MODE != MODE_BWD && mir) {       // one mirror bounce (extension)
                float rgb2[3];
                mirror_shade(sc.shader, sc.max_depth, w2o_s, mats_g, sc.obj_type, N, g, idx, ob, h, wx, wy, wz, rgb2);
                const float kr = __ldg(refl_s + idx);
#pragma unroll
                for (int c = 0; c < 3; c++) rgb[c] = (1.0f - kr) * rgb[c] + kr * rgb2[c];
            }
        }

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
            if (active && s == 0 && img_s) { img_s[po] = v0; img_s[po + 1] = v1; img_s[po + 2] = v2; }
            if (MODE == MODE_FUSED && active) {
                float lp = 0.f;
                pixel_cost(sc.flags & RRT_FLAG_LINEAR_COST, P.cw, inv, v0, v1, v2, tgt, lp, gc);
                if (s == 0) loss_part += lp;                 // once per pixel (its S lanes hold the same value)
            }
        }

        if (MODE != MODE_FWD) {
            // ---- reverse pass through the winner into the running per-thread sums
            int key = idx;
            if (gc[0] == 0.f && gc[1] == 0.f && gc[2] == 0.f) key = -1;
            const bool change = (key >= 0) && (acc_key >= 0) && (key != acc_key);
            if (__any_sync(full, change)) {                  // some lane's winner changed: flush the warp's sums
                small_warp_flush<NACC>(acc_key, acc_col, slots, lane, det_ws);
                acc_key = -1;
            }
            if (key >= 0) {
                const float rc3[3] = {rcx, rcy, rcz};
                float da[NACC], dg[9];
#pragma unroll
                for (int v = 0; v < NACC; v++) da[v] = 0.f;
#pragma unroll
                for (int v = 0; v < 9; v++) dg[v] = 0.f;
                if (MIRROR && mir) {
                    // secondary object first (its sums go straight to its CTA slot / the fixed-point
                    // workspace), then the primary one with the chain through the reflection
                    const float kr = __ldg(refl_s + key);
                    const float gc2[3] = {kr * gc[0], kr * gc[1], kr * gc[2]};
                    const float gc1[3] = {(1.0f - kr) * gc[0], (1.0f - kr) * gc[1], (1.0f - kr) * gc[2]};
                    float og2[NACC], extra[4], dA[9];
                    const int j2 = mirror_backward<GEOM, NACC>(sc.shader, sc.max_depth, w2o_s, mats_g, sc.obj_type, N, g, key, ob, h,
                                                               wx, wy, wz, gc2, og2, dg, extra, dA);
                    if (j2 >= 0) {
#pragma unroll 1
                        for (int v = 0; v < NACC; v++) {
                            if (og2[v] == 0.f) continue;
                            if (det_ws) det_add(det_ws + ((size_t)j2 * RRT_OBJ_GRAD_STRIDE + v) * 2, (double)og2[v]);
                            else atomicAdd(&slots[j2 * kSlotStride + v], og2[v]);
                        }
                    }
                    backward_ray<GEOM, NACC>(sc.shader, sc.max_depth, ob, m7, g, h, sr, rc3, gc1, da, dg, j2 >= 0 ? extra : nullptr);
#pragma unroll
                    for (int q = 0; q < 9; q++) da[q] += dA[q];
                } else
                backward_ray<GEOM, NACC>(sc.shader, sc.max_depth, ob, m7, g, h, sr, rc3, gc, da, dg);
#pragma unroll
                for (int v = 0; v < NACC; v++) acc_col[v * kSmallThreads] += da[v];
                if (!GEOM) {
#pragma unroll
                    for (int v = 0; v < 9; v++) gg_col[v * kSmallThreads] += dg[v];
                }
                acc_key = key;
            }
        }
    }

    if (MODE != MODE_FWD && STEP) {
        // ---- the rest of the optimise step, by the LAST CTA to get here (ticket): finalize the
        // raw sums into d/d w2o, chain them back to the parameters (T.grad through
        // translate/scale/rotate, transform.py:56-122), apply var <- var - lr*grad
        // (optimize.py:26-27), publish the loss and re-zero the scratch for the next step.
        __threadfence();
        __syncthreads();
        if (tid == 0) flag_s = (atomicAdd(P.step.ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (flag_s) {
            __threadfence();
            const rrt_step& q = P.step;
            if (tid < N) {
                float Mm[9], gb[3], G[12];
#pragma unroll
                for (int v = 0; v < 9; v++) Mm[v] = __ldcg(gobj + (size_t)tid * RRT_OBJ_GRAD_STRIDE + v);
#pragma unroll
                for (int v = 0; v < 3; v++) gb[v] = __ldcg(gobj + (size_t)tid * RRT_OBJ_GRAD_STRIDE + 9 + v);
#pragma unroll
                for (int r = 0; r < 3; r++) {          // d/dA = M C^T + g_b ct^T ; d/db = g_b (finalize_scene)
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
