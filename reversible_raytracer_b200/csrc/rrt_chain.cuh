// rrt_chain.cuh -- part of rrt_kernels.cu (one translation unit; included inside its anonymous namespace).
// Parameter -> matrix chains (translate / scale / rotate products) and their reverse pass.
#pragma once

// ---------------------------------------------------------------- parameter -> matrix chain
// Affine 3x4 matrices [A|b] (bottom row 0 0 0 1 implied).  See include/rrt_b200.h.
struct Aff {
    float m[12];
};

__device__ __forceinline__ Aff aff_identity() {
    Aff r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.m[i] = 0.f;
    r.m[0] = r.m[5] = r.m[10] = 1.f;
    return r;
}

// C = A . B   (transform.py:35-38); products with exact zeros stay exact zeros
__device__ __forceinline__ Aff aff_mul(const Aff& A, const Aff& B) {
    Aff C;
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float v = A.m[r * 4 + 0] * B.m[0 * 4 + c] + A.m[r * 4 + 1] * B.m[1 * 4 + c] + A.m[r * 4 + 2] * B.m[2 * 4 + c];
            if (c == 3) v += A.m[r * 4 + 3];
            C.m[r * 4 + c] = v;
        }
    }
    return C;
}

// rotate(angle_deg, axis), transform.py:95-122 (Rodrigues form; axis assumed unit)
__device__ __forceinline__ void rot_entries(float angle, const float* a, float* R) {
    float s, c;
    sincosf(angle * 0.017453292519943295f, &s, &c);
    R[0] = a[0] * a[0] + (1.f - a[0] * a[0]) * c;
    R[1] = a[0] * a[1] * (1.f - c) - a[2] * s;
    R[2] = a[0] * a[2] * (1.f - c) + a[1] * s;
    R[3] = a[0] * a[1] * (1.f - c) + a[2] * s;
    R[4] = a[1] * a[1] + (1.f - a[1] * a[1]) * c;
    R[5] = a[1] * a[2] * (1.f - c) - a[0] * s;
    R[6] = a[0] * a[2] * (1.f - c) - a[1] * s;
    R[7] = a[1] * a[2] * (1.f - c) + a[0] * s;
    R[8] = a[2] * a[2] + (1.f - a[2] * a[2]) * c;
}

__device__ __forceinline__ Aff chain_op_matrix(const int32_t* op, const float* __restrict__ values) {
    const int kind = op[0] & 0xff;
    const bool inv = (op[0] & RRT_CHAIN_INVERT) != 0;
    Aff M = aff_identity();
    if (kind == RRT_CHAIN_TRANSLATE) {            // transform.py:60-75
        const float* v = values + op[1];
        M.m[3] = inv ? -v[0] : v[0]; M.m[7] = inv ? -v[1] : v[1]; M.m[11] = inv ? -v[2] : v[2];
    } else if (kind == RRT_CHAIN_SCALE) {         // transform.py:78-93 (inverse is 1/x)
        const float* v = values + op[1];
        M.m[0] = inv ? 1.f / v[0] : v[0]; M.m[5] = inv ? 1.f / v[1] : v[1]; M.m[10] = inv ? 1.f / v[2] : v[2];
    } else if (kind == RRT_CHAIN_ROTATE) {        // inverse = transpose
        float R[9];
        rot_entries(values[op[1]], values + op[2], R);
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) M.m[r * 4 + c] = inv ? R[c * 3 + r] : R[r * 3 + c];
    }
    return M;
}

__device__ __forceinline__ Aff chain_forward_one(const int32_t* __restrict__ ops, const int32_t* __restrict__ chain_begin,
                                                 int k, const float* __restrict__ values) {
    Aff M = aff_identity();
    for (int j = chain_begin[k]; j < chain_begin[k + 1]; j++) M = aff_mul(M, chain_op_matrix(ops + 4 * j, values));
    return M;
}

__global__ void chain_forward_kernel(const int32_t* __restrict__ ops, const int32_t* __restrict__ chain_begin,
                                     int num_chains, const float* __restrict__ values, float* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= num_chains) return;
    const Aff M = chain_forward_one(ops, chain_begin, k, values);
#pragma unroll
    for (int i = 0; i < 12; i++) out[(size_t)k * 12 + i] = M.m[i];
}

// dL/d(op j) = P_{j-1}^T . G . S_{j+1}^T with P = prefix product, S = suffix product
// (4x4 with the implied bottom row); then into the primitive's own parameters.
// G = dL/d(row k of the chain's output); accumulates into g_values with atomics (parameters may be
// shared between chains).
__device__ __noinline__ void chain_backward_one(const int32_t* __restrict__ ops, const int32_t* __restrict__ chain_begin,
                                                int k, const float* __restrict__ values, const float* G,
                                                float* __restrict__ g_values) {
    const int b = chain_begin[k], e = chain_begin[k + 1], n = e - b;
    if (n <= 0 || n > RRT_CHAIN_MAX_OPS) return;
    Aff mats[RRT_CHAIN_MAX_OPS], pre[RRT_CHAIN_MAX_OPS + 1];
    pre[0] = aff_identity();
    for (int j = 0; j < n; j++) {
        mats[j] = chain_op_matrix(ops + 4 * (b + j), values);
        pre[j + 1] = aff_mul(pre[j], mats[j]);
    }
    Aff suf = aff_identity();                     // product of ops j+1..n-1
    for (int j = n - 1; j >= 0; j--) {
        // out = P . M_j . S  (affine).  T = G . S^T restricted to what reaches M_j's 3x4 block:
        //   T[r][c] = sum_q G[r][q] S[c][q] (c<3: q over 0..3 with S[c][3]=b_c) ; T[r][3] = G[r][3]
        float T[12];
#pragma unroll
        for (int r = 0; r < 3; r++) {
#pragma unroll
            for (int c = 0; c < 3; c++)
                T[r * 4 + c] = G[r * 4 + 0] * suf.m[c * 4 + 0] + G[r * 4 + 1] * suf.m[c * 4 + 1] +
                               G[r * 4 + 2] * suf.m[c * 4 + 2] + G[r * 4 + 3] * suf.m[c * 4 + 3];
            T[r * 4 + 3] = G[r * 4 + 3];
        }
        // D = P_A^T . T   (gradient w.r.t. M_j's [A|b])
        const Aff& Pm = pre[j];
        float D[12];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 4; c++)
                D[r * 4 + c] = Pm.m[0 * 4 + r] * T[0 * 4 + c] + Pm.m[1 * 4 + r] * T[1 * 4 + c] + Pm.m[2 * 4 + r] * T[2 * 4 + c];
        const int32_t* op = ops + 4 * (b + j);
        const int kind = op[0] & 0xff;
        const bool inv = (op[0] & RRT_CHAIN_INVERT) != 0;
        if (kind == RRT_CHAIN_TRANSLATE) {
            const float sgn = inv ? -1.f : 1.f;
            atomicAdd(&g_values[op[1] + 0], sgn * D[3]);
            atomicAdd(&g_values[op[1] + 1], sgn * D[7]);
            atomicAdd(&g_values[op[1] + 2], sgn * D[11]);
        } else if (kind == RRT_CHAIN_SCALE) {
            const float* v = values + op[1];
#pragma unroll
            for (int i = 0; i < 3; i++) atomicAdd(&g_values[op[1] + i], inv ? -D[i * 5] / (v[i] * v[i]) : D[i * 5]);
        } else if (kind == RRT_CHAIN_ROTATE) {
            const float ang = values[op[1]];
            const float* a = values + op[2];
            float s, c;
            sincosf(ang * 0.017453292519943295f, &s, &c);
            float Gr[9];                         // gradient w.r.t. the (non-transposed) rotation entries
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int cc = 0; cc < 3; cc++) Gr[r * 3 + cc] = inv ? D[cc * 4 + r] : D[r * 4 + cc];
            const float dc = -s * 0.017453292519943295f, ds = c * 0.017453292519943295f, omc = 1.f - c;
            float g_ang = Gr[0] * (1.f - a[0] * a[0]) * dc + Gr[4] * (1.f - a[1] * a[1]) * dc + Gr[8] * (1.f - a[2] * a[2]) * dc +
                          Gr[1] * (-a[0] * a[1] * dc - a[2] * ds) + Gr[2] * (-a[0] * a[2] * dc + a[1] * ds) +
                          Gr[3] * (-a[0] * a[1] * dc + a[2] * ds) + Gr[5] * (-a[1] * a[2] * dc - a[0] * ds) +
                          Gr[6] * (-a[0] * a[2] * dc - a[1] * ds) + Gr[7] * (-a[1] * a[2] * dc + a[0] * ds);
            float g_a0 = Gr[0] * 2.f * a[0] * omc + (Gr[1] + Gr[3]) * a[1] * omc + (Gr[2] + Gr[6]) * a[2] * omc + (Gr[7] - Gr[5]) * s;
            float g_a1 = Gr[4] * 2.f * a[1] * omc + (Gr[1] + Gr[3]) * a[0] * omc + (Gr[5] + Gr[7]) * a[2] * omc + (Gr[2] - Gr[6]) * s;
            float g_a2 = Gr[8] * 2.f * a[2] * omc + (Gr[2] + Gr[6]) * a[0] * omc + (Gr[5] + Gr[7]) * a[1] * omc + (Gr[3] - Gr[1]) * s;
            atomicAdd(&g_values[op[1]], g_ang);
            atomicAdd(&g_values[op[2] + 0], g_a0);
            atomicAdd(&g_values[op[2] + 1], g_a1);
            atomicAdd(&g_values[op[2] + 2], g_a2);
        }
        suf = aff_mul(mats[j], suf);
    }
}

__global__ void chain_backward_kernel(const int32_t* __restrict__ ops, const int32_t* __restrict__ chain_begin,
                                      int num_chains, const float* __restrict__ values,
                                      const float* __restrict__ g_out, float* __restrict__ g_values) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= num_chains) return;
    float G[12];
#pragma unroll
    for (int i = 0; i < 12; i++) G[i] = g_out[(size_t)k * 12 + i];
    chain_backward_one(ops, chain_begin, k, values, G, g_values);
}
