// rrt_render_kernel.cuh -- part of rrt_kernels.cu (one translation unit; included inside its anonymous namespace).
// render_kernel<PIX,SPT,MODE>: the general 8-rays-per-thread kernel.
#pragma once

// ---------------------------------------------------------------- the render kernel
// grid = (ceil(n / (32*PIX)), ceil(rows / warps), B); block = 32 * warps.
// Thread (warp w, lane l) owns pixels (row = tile_row0 + w, cols = col0 + l*PIX .. +PIX-1),
// each with SPT samples: PIX*SPT = 8 rays.  S > SPT (generic path, PIX = 1) loops
// over chunks of SPT samples.
// MIRROR instantiations carry the opt-in reflection bounce (RRT_FLAG_MIRROR); the default ones contain none
// of its code (it would triple the stack frame, and local memory of all resident threads competes for L2).
template <int PIX, int SPT, int MODE, bool MIRROR = false>
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
    __shared__ float camg_s[48];
    __shared__ int last_s;

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
    long long* const det_ws = (MODE != MODE_FWD && (sc.flags & RRT_FLAG_DETERMINISTIC)) ? det_scene(sc, scene) : nullptr;
    // RRT_FLAG_MIRROR (opt-in extension): needs the identity camera, Phong shaders
    const bool mirror_on = MIRROR && (sc.flags & RRT_FLAG_MIRROR) && sc.reflectivity && sc.shader != RRT_SHADER_DEPTH && cam_identity &&
                           g.ct[0] == 0.f && g.ct[1] == 0.f && g.ct[2] == 0.f;
    const float* refl = mirror_on ? sc.reflectivity + (size_t)scene * sc.reflectivity_scene_stride : nullptr;
    // the last CTA of a scene (rrt_scene.ticket) finalises its gradients: one launch per reverse pass
    auto take_ticket = [&]() {
        if (MODE == MODE_FWD || !sc.ticket) return;
        __threadfence();
        __syncthreads();
        if (tid == 0) last_s = (atomicAdd(sc.ticket + scene, 1u) == gridDim.x * gridDim.y - 1);
        __syncthreads();
        if (last_s) {
            __threadfence();
            finalize_scene(P, scene, tid, blockDim.x, camg_s);
            if (tid == 0) sc.ticket[scene] = 0u;
        }
    };

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
        if (!__syncthreads_or(nz)) { take_ticket(); return; }
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
    // Fused mode with more samples than one thread holds (generic instantiation only, S > 8): the
    // upstream gradient of a pixel is known only after ALL its samples are shaded, so the chunk
    // loop runs twice -- pass 0 sweeps + shades, pass 1 sweeps again + runs the reverse pass
    // (hit records are recomputed, never stored).  Compile-time false wherever S == SPT.
    const bool two_pass = (MODE == MODE_FUSED) && (PIX == 1) && nchunks_s > 1;
    const int ntrips = two_pass ? 2 * nchunks_s : nchunks_s;
#pragma unroll 1
    for (int trip = 0; trip < ntrips; trip++) {
        const int sc0 = two_pass ? (trip >= nchunks_s ? trip - nchunks_s : trip) : trip;
        const bool do_fwd = (MODE != MODE_BWD) && (!two_pass || trip < nchunks_s);
        const bool do_bwd = (MODE != MODE_FWD) && (!two_pass || trip >= nchunks_s);
        // ---- build the 8 rays of this sample chunk.  Per-ray state lives in (L1-resident)
        // local memory: it is read by the rare path of the sweep and by the rolled shading /
        // reverse-pass loops below; only the packed world directions stay in registers.
        __align__(16) float l_rc[3 * kRays], l_dw[3 * kRays], l_tmin[kRays];  // SoA: [x0..x7 | y0..y7 | z0..z7]
        int l_idx[kRays];
        __align__(16) float l_uv[2 * kRays];          // (u, v) = (d_x/d_z, d_y/d_z) of the pre-filter: [u0..u7 | v0..v7]
        bool rays_safe = true;                        // every ray inside the range the pre-filter's bound is proven for
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
            {
                // padding rays get NaN (never a candidate); a real ray must have |d_z| within 2^+-10
                // and |u|, |v| <= 2^10, else the whole CTA takes the canonical sweep
                float u = __int_as_float(0x7fc00000), v = u;
                if (ok) {
                    u = __fdiv_rn(wx, wz);
                    v = __fdiv_rn(wy, wz);
                    const float az = fabsf(wz);
                    rays_safe &= (az >= 9.765625e-4f) && (az <= 1024.0f) && (fabsf(u) <= 1024.0f) && (fabsf(v) <= 1024.0f);
                }
                l_uv[r] = u; l_uv[kRays + r] = v;
            }
            l_rc[r] = rcx; l_rc[kRays + r] = rcy; l_rc[2 * kRays + r] = rcz;
            l_dw[r] = wx; l_dw[kRays + r] = wy; l_dw[2 * kRays + r] = wz;
            l_tmin[r] = __int_as_float(0x7f800000);
            l_idx[r] = -1;
        }
        const bool use_stored = (MODE == MODE_BWD) && (P.hit_in != nullptr);
        const bool cull = (sc.flags & RRT_FLAG_CULL) && !use_stored;
        // conservative pre-filter sweep (default with a prebuilt table): CTA-wide decision, because the
        // shared-memory chunk then holds pre-filter rows instead of records
        bool use_q = false;
        if (sc.obj_records && !cull && !use_stored && !(sc.flags & RRT_FLAG_CANONICAL_SWEEP))
            use_q = __syncthreads_and(rays_safe);
        bool staged_quadrics = false;                 // the staged chunk does not hold records (shadow pass restages)
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
            if (trip > 0) __syncthreads();         // previous use of cone_red / tcone is over
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
                if (kb > 0 || trip > 0) __syncthreads();  // previous chunk fully consumed
                bool staged_by_tma = false;
                const float* rec_g = sc.obj_records ? sc.obj_records + ((size_t)scene * N + kb) * RRT_RECORD_FLOATS : nullptr;
                // chunk class from the prebuilt table: sphere-only chunks take the pre-filter
                const bool qchunk = use_q && !(__float_as_int(__ldg(rec_g + 15)) & 1);
                const int cnt_pad = (cnt + 3) & ~3;
                if (qchunk) {
                    const int npad = (N + 3) & ~3;
                    stage_bytes_tma(smem_tab, sc.obj_records + (size_t)sc.num_scenes * N * RRT_RECORD_FLOATS +
                                                  ((size_t)scene * npad + kb) * RRT_QUADRIC_FLOATS,
                                    (unsigned)cnt_pad * 24u, &tma_bar, &tma_phase, tid);
                    staged_by_tma = true;
                    staged_quadrics = true;
                } else if (N > kObjChunk || trip == 0 || staged_quadrics) {
                    staged_quadrics = false;
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
                // the packed rays of the sweep that is actually taken are (re)loaded from local memory per
                // chunk, so that only ONE of the two packs is ever live in registers
                if (qchunk) {
                    RayQ rq;
                    const u64* up = reinterpret_cast<const u64*>(l_uv);
#pragma unroll
                    for (int p = 0; p < kRays / 2; p++) { rq.u[p] = up[p]; rq.v[p] = up[kRays / 2 + p]; }
                    sweep_quadric(smem_tab, cnt_pad, cnt, reinterpret_cast<const float4*>(rec_g), kb, rq, l_dw, l_tmin, l_idx);
                    continue;
                }
                RayPack rp;
                {
                    const u64* dp = reinterpret_cast<const u64*>(l_dw);   // (x0,x1) (x2,x3) ... pairs
#pragma unroll
                    for (int p = 0; p < kRays / 2; p++) { rp.dx[p] = dp[p]; rp.dy[p] = dp[kRays / 2 + p]; rp.dz[p] = dp[kRays + p]; }
                }
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
                if (N > kObjChunk || staged_quadrics) {    // otherwise the whole table is still staged
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
                shadowed = (sc.flags & RRT_FLAG_SCALAR_SHADOWS)
                               ? shadow_chunk(smem_tab, cnt, kb, l_dw, l_tmin, l_idx, g.U, shadowed)
                               : shadow_chunk_packed(smem_tab, cnt, kb, l_dw, l_tmin, l_idx, g.U, shadowed);
            }
        }

        // ---- outputs of the sweep
        if (do_fwd && (P.hit_out || P.tmin_out)) {
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
        if (do_fwd) {
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
                hit_record<false>(ob, dwx, dwy, dwz, l_tmin[r], h);   // t is known from the sweep
                ShadeRec sr;
                float rgb[3];
                shade(sc.shader, sc.max_depth, ob, m7, g, h, sr, rgb);
                if (MIRROR && mirror_on) {                // one mirror bounce (extension)
                    float rgb2[3];
                    mirror_shade(sc.shader, sc.max_depth, w2o, mats, sc.obj_type, N, g, k, ob, h, dwx, dwy, dwz, rgb2);
                    const float kr = __ldg(refl + k);
#pragma unroll
                    for (int c = 0; c < 3; c++) rgb[c] = (1.0f - kr) * rgb[c] + kr * rgb2[c];
                }
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
        if (do_fwd && last_chunk) {
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
                if (MODE == MODE_FUSED && ok)
                    pixel_cost(sc.flags & RRT_FLAG_LINEAR_COST, P.cw, inv, v[px][0], v[px][1], v[px][2], tg[px], loss_part, gpix[px]);
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
        if (do_bwd) {
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
                    if (use_stored) {
                        obj_test(ob, dwx, dwy, dwz, h);
                        if (!(h.t < __int_as_float(0x7f800000))) k = -1;  // stale stored winner
                    } else {
                        hit_record<true>(ob, dwx, dwy, dwz, l_tmin[r], h);   // t is known from the sweep
                    }
                }
                // flush the running per-object accumulator when some lane changes object
                const bool change = fin ? (acc_key >= 0) : ((k >= 0) && (acc_key >= 0) && (k != acc_key));
                if (__any_sync(0xffffffffu, change)) {
                    warp_flush(acc_key, acc, slot_key, slots, gobj, lane, det_ws);
                    acc_key = -1;
                }
                if (k >= 0) {
                    ShadeRec sr;
                    float rgb[3];
                    shade(sc.shader, sc.max_depth, ob, m7, g, h, sr, rgb);
                    acc_key = k;
                    {
                        const float rc3[3] = {l_rc[r], l_rc[kRays + r], l_rc[2 * kRays + r]};
                        if (MIRROR && mirror_on) {
                            const float kr = __ldg(refl + k);
                            const float gc2[3] = {kr * gc[0], kr * gc[1], kr * gc[2]};
                            const float gc1[3] = {(1.0f - kr) * gc[0], (1.0f - kr) * gc[1], (1.0f - kr) * gc[2]};
                            float og2[19], extra[4], dA[9];
                            const int j2 = mirror_backward<false, 19>(sc.shader, sc.max_depth, w2o, mats, sc.obj_type, N, g, k, ob, h,
                                                                      dwx, dwy, dwz, gc2, og2, gg, extra, dA);
                            if (j2 >= 0) {
#pragma unroll 1
                                for (int v = 0; v < 19; v++) {
                                    if (og2[v] == 0.f) continue;
                                    if (det_ws) det_add(det_ws + ((size_t)j2 * RRT_OBJ_GRAD_STRIDE + v) * 2, (double)og2[v]);
                                    else atomicAdd(&gobj[(size_t)j2 * RRT_OBJ_GRAD_STRIDE + v], og2[v]);
                                }
                            }
                            backward_ray(sc.shader, sc.max_depth, ob, m7, g, h, sr, rc3, gc1, acc, gg, j2 >= 0 ? extra : nullptr);
#pragma unroll
                            for (int q = 0; q < 9; q++) acc[q] += dA[q];
                        } else
                        backward_ray(sc.shader, sc.max_depth, ob, m7, g, h, sr, rc3, gc, acc, gg);
                    }
                }
            }
        }
    }  // sample chunks

    if (MODE != MODE_FWD) {
        // ---- warp -> CTA -> global reduction (per-object sums were flushed by the sentinel trip)
        if (det_ws) {
            // deterministic: every warp adds its (fixed-order) sums to the fixed-point workspace
#pragma unroll
            for (int v = 0; v < 9; v++) {
                const float x = warp_sum(gg[v]);
                const int dst = v < 6 ? v : 12 + v;
                if (lane == 0 && x != 0.f) det_add(det_ws + ((size_t)N * RRT_OBJ_GRAD_STRIDE + dst) * 2, (double)x);
            }
            if (MODE == MODE_FUSED) {
                const float x = warp_sum(loss_part);
                if (lane == 0 && x != 0.f) det_add(det_ws + (size_t)RRT_GRAD_SIZE(N) * 2, (double)x);
            }
            take_ticket();
            return;
        }
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
        take_ticket();
    }
}
