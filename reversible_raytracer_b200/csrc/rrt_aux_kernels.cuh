// rrt_aux_kernels.cuh -- part of rrt_kernels.cu (one translation unit; included inside its anonymous namespace).
// Small kernels: record table, primary-ray table, gradient finalisation, peer-memory exchange.
#pragma once

// ---------------------------------------------------------------- sweep-record table (for TMA staging)
// Pre-filter row of one object (see RRT_FLAG_CANONICAL_SWEEP in include/rrt_b200.h and
// sweep_quadric in rrt_sweep.cuh).  With the STORED float32 record values A, o', cc the
// reference's discriminant is the quadratic form det(d) = d^T Q d, Q = A^T (o' o'^T - cc I) A.
// For d = d_z (u, v, 1):  det / d_z^2 = F(u,v) = Q22 + 2 Q02 u + 2 Q12 v + Q00 u^2 + Q11 v^2 + 2 Q01 uv.
// Error budget, in units of eps = 2^-24 times B |(u,v,1)|^2 with B = |A|_F^2 (|o'|^2 + |cc|):
//   canonical float32 det (11 roundings in the fma chains)      <= 13.3
//   u, v = fl(d_x/d_z), fl(d_y/d_z) instead of the exact ratios  <=  3.5
//   float32 Horner evaluation of F (5 fma; every term passes <= 3 roundings + its coefficient's) <= 7
// so with mu = 64 eps B added to the three diagonal coefficients, canonical det > 0 implies the
// evaluated F > 0 (margin 2.7x) as long as nothing over/underflows: |A|_F and |o'| within
// 2^+-16 here, |d_z| within 2^+-10 and |u|, |v| <= 2^10 on the ray side (checked per CTA).
// Everything else gets an always-pass row, i.e. is decided by the canonical arithmetic alone.
// Row layout (6 floats):  c00 c02 c22 c01 c12 c11   (F = ((c00 u + c02) u + c22) + v ((c01 u + c12) + c11 v))
__device__ __forceinline__ void quadric_row(const Obj& ob, float* __restrict__ row) {
    const double a[9] = {ob.a[0], ob.a[1], ob.a[2], ob.a[3], ob.a[4], ob.a[5], ob.a[6], ob.a[7], ob.a[8]};
    const double o[3] = {ob.o[0], ob.o[1], ob.o[2]};
    const double cc = -(double)ob.ncc;
    double af2 = 0.0;
#pragma unroll
    for (int q = 0; q < 9; q++) af2 += a[q] * a[q];
    const double oo = o[0] * o[0] + o[1] * o[1] + o[2] * o[2];
    const double lo = 1.52587890625e-05, hi = 65536.0;            // 2^-16, 2^16
    bool safe = !(ob.flags & 1) && (af2 >= lo * lo) && (af2 <= hi * hi) && (oo <= hi * hi) && (fabs(cc) <= 2.0 * hi * hi);
    // (comparisons are false for NaN; infinities fail the upper bounds)
    float c22 = 1.0e30f, c02 = 0.f, c12 = 0.f, c00 = 0.f, c11 = 0.f, c01 = 0.f;   // always-pass row
    if (safe) {
        // M = o o^T - cc I ;  Q = A^T M A  (A row-major: a[r*3+c])
        double M[9], T[9], Q[9];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) M[r * 3 + c] = o[r] * o[c] - (r == c ? cc : 0.0);
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) T[r * 3 + c] = M[r * 3] * a[c] + M[r * 3 + 1] * a[3 + c] + M[r * 3 + 2] * a[6 + c];   // M A
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) Q[r * 3 + c] = a[r] * T[c] + a[3 + r] * T[3 + c] + a[6 + r] * T[6 + c];             // A^T (M A)
        const double mu = 3.814697265625e-06 * af2 * (oo + fabs(cc));      // 2^-18 B
        c22 = (float)(Q[8] + mu); c00 = (float)(Q[0] + mu); c11 = (float)(Q[4] + mu);
        c02 = (float)(Q[2] + Q[6]); c12 = (float)(Q[5] + Q[7]); c01 = (float)(Q[1] + Q[3]);
    }
    row[0] = c00; row[1] = c02; row[2] = c22; row[3] = c01; row[4] = c12; row[5] = c11;
}

// grid = (chunks of kObjChunk objects, scenes).  Writes the 64-byte records the render kernels
// bulk-copy into shared memory; the chunk's class bits (squares / general spheres present) go
// into the spare slot of its first record.  Then the pre-filter rows (second plane of the table,
// padded to a multiple of 4 objects with never-pass rows so that the hot loop has no tail).
__global__ void __launch_bounds__(128) build_records_kernel(const rrt_scene sc, float* __restrict__ records) {
    __shared__ int cls_s;
    const int scene = blockIdx.y, kb = blockIdx.x * kObjChunk, N = sc.num_objects;
    const int cnt = min(kObjChunk, N - kb);
    const int npad = (N + 3) / 4 * 4;
    if (threadIdx.x == 0) cls_s = 0;
    __syncthreads();
    const float* cam = sc.camera + (size_t)scene * sc.camera_scene_stride;
    const float ct[3] = {__ldg(cam + 3), __ldg(cam + 7), __ldg(cam + 11)};
    const float* w2o = sc.w2o + (size_t)scene * sc.w2o_scene_stride;
    float4* out = reinterpret_cast<float4*>(records + ((size_t)scene * N + kb) * RRT_RECORD_FLOATS);
    float* quad = records + (size_t)sc.num_scenes * N * RRT_RECORD_FLOATS + ((size_t)scene * npad + kb) * RRT_QUADRIC_FLOATS;
    int cls = 0;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
        Obj ob;
        make_obj(w2o + (size_t)(kb + k) * RRT_W2O_STRIDE, sc.obj_type[kb + k], ct, ob, true);
        store_rec(out + 4 * k, ob);
        quadric_row(ob, quad + (size_t)k * RRT_QUADRIC_FLOATS);
        cls |= ob.flags;
    }
    const int cnt_pad = min(kObjChunk, npad - kb);
    for (int k = cnt + threadIdx.x; k < cnt_pad; k += blockDim.x) {       // never-pass rows: F = -1
        float* row = quad + (size_t)k * RRT_QUADRIC_FLOATS;
        row[0] = row[1] = row[3] = row[4] = row[5] = 0.f;
        row[2] = -1.f;
    }
    if (cls) atomicOr(&cls_s, cls);
    __syncthreads();
    if (threadIdx.x == 0 && cnt > 0) reinterpret_cast<float*>(out)[15] = __int_as_float(cls_s);
}

// ---------------------------------------------------------------- primary-ray grid table
__global__ void primary_rays_kernel(int n, double step, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n) return;
    float x, y, z;
    base_ray(n, step, i, j, x, y, z);
    float* o = out + ((size_t)i * n + j) * 3;
    o[0] = x; o[1] = y; o[2] = z;
}

// ---------------------------------------------------------------- gradient finalisation
// One CTA per scene (the separate-launch form; with rrt_scene.ticket the render kernels call
// finalize_scene themselves from the last CTA of each scene).
__global__ void __launch_bounds__(128) finalize_grads(const __grid_constant__ KParams P) {
    __shared__ float camg[48];
    finalize_scene(P, blockIdx.x, threadIdx.x, blockDim.x, camg);
}

// ---------------------------------------------------------------- gradient exchange over peer memory
// The one exchange step of the sharded path (row slabs / scene ranges per GPU): every rank holds
// a small vector [gradient (float32) | loss (float64)] and all ranks need the sum.  Instead of
// an NCCL allreduce (plus the two copy kernels that pack its buffer) ONE kernel per rank
//   1. PUSHES its values, converted to float64, into slot[rank] of every peer's buffer with
//      plain stores through the NVLink peer mapping,
//   2. publishes a per-(CTA, source) flag on every peer (fence + release store) and waits for
//      the same flags from all peers (acquire loads) -- CTA c only depends on CTA c of the
//      peers, so no grid-wide barrier is needed,
//   3. sums the `world` slots in RANK ORDER (every rank gets the same bits; deterministic).
// Buffers alternate by epoch parity: a rank can only start epoch e+2 after every peer has
// finished reading epoch e (it needs their epoch e+1 flags first), so two copies suffice.
// Flags carry the epoch number, one per source, so a fast peer's next epoch cannot be
// mistaken for a slow peer's current one.
constexpr int kPeerCtas = 16;
constexpr int kPeerMaxWorld = 16;
constexpr int kPeerEpochOffset = kPeerCtas * kPeerMaxWorld;   // per-CTA epoch counters (local use only)

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) peer_allreduce_kernel(const float* __restrict__ grad, const double* __restrict__ loss,
                                                             int n, int nloss, void* const* __restrict__ peer_buf,
                                                             void* const* __restrict__ peer_sig, int rank, int world,
                                                             double* __restrict__ out) {
    const int cta = blockIdx.x, tid = threadIdx.x;
    unsigned* sig_local = reinterpret_cast<unsigned*>(peer_sig[rank]);
    __shared__ unsigned epoch_s;
    __shared__ int timed_out;
    if (tid == 0) {
        timed_out = 0;
        epoch_s = sig_local[kPeerEpochOffset + cta] + 1u;
        sig_local[kPeerEpochOffset + cta] = epoch_s;
    }
    __syncthreads();
    const unsigned epoch = epoch_s;
    const int total = n + nloss;
    const int per = (total + gridDim.x - 1) / gridDim.x;
    const int lo = cta * per, hi = min(total, lo + per);
    const size_t slot = ((size_t)(epoch & 1u) * world + rank) * (size_t)total;
    // 1. push
    for (int i = lo + tid; i < hi; i += blockDim.x) {
        const double v = i < n ? (double)grad[i] : loss[i - n];
        for (int p = 0; p < world; p++) reinterpret_cast<double*>(peer_buf[p])[slot + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish + wait
    if (tid < world) {
        st_release_sys(reinterpret_cast<unsigned*>(peer_sig[tid]) + cta * kPeerMaxWorld + rank, epoch);
        const unsigned* mine = sig_local + cta * kPeerMaxWorld + tid;
        // bounded wait (~10 s): a peer that died must not hang this GPU; the sums become NaN
        long long spins = 0;
        while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
            if (++spins > (1LL << 26)) { timed_out = 1; break; }
            if (spins > 1024) __nanosleep(128);
        }
    }
    __syncthreads();
    // 3. sum in rank order
    const double* local = reinterpret_cast<const double*>(peer_buf[rank]) + (size_t)(epoch & 1u) * world * (size_t)total;
    for (int i = lo + tid; i < hi; i += blockDim.x) {
        double sum = 0.0;
        for (int p = 0; p < world; p++) sum += __ldcg(local + (size_t)p * total + i);
        out[i] = timed_out ? __longlong_as_double(0x7ff8000000000000LL) : sum;
    }
}
