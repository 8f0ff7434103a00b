// rrt_aux_kernels.cuh -- part of rrt_kernels.cu (one translation unit; included inside its anonymous namespace).
// Small kernels: record table, primary-ray table, gradient finalisation, peer-memory exchange.
#pragma once

// ---------------------------------------------------------------- sweep-record table (for TMA staging)
// grid = (chunks of kObjChunk objects, scenes).  Writes the 64-byte records the render kernels
// bulk-copy into shared memory; the chunk's class bits (squares / general spheres present) go
// into the spare slot of its first record.
__global__ void __launch_bounds__(128) build_records_kernel(const rrt_scene sc, float* __restrict__ records) {
    __shared__ int cls_s;
    const int scene = blockIdx.y, kb = blockIdx.x * kObjChunk, N = sc.num_objects;
    const int cnt = min(kObjChunk, N - kb);
    if (threadIdx.x == 0) cls_s = 0;
    __syncthreads();
    const float* cam = sc.camera + (size_t)scene * sc.camera_scene_stride;
    const float ct[3] = {__ldg(cam + 3), __ldg(cam + 7), __ldg(cam + 11)};
    const float* w2o = sc.w2o + (size_t)scene * sc.w2o_scene_stride;
    float4* out = reinterpret_cast<float4*>(records + ((size_t)scene * N + kb) * RRT_RECORD_FLOATS);
    int cls = 0;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
        Obj ob;
        make_obj(w2o + (size_t)(kb + k) * RRT_W2O_STRIDE, sc.obj_type[kb + k], ct, ob, true);
        store_rec(out + 4 * k, ob);
        cls |= ob.flags;
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
    __shared__ float camg[12];
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
