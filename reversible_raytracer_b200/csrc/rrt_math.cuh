// rrt_math.cuh -- part of rrt_kernels.cu (one translation unit; included inside its anonymous namespace).
// Packed f32x2 arithmetic (FFMA2 / FMUL2), the jitter RNG and primary-ray generation.
#pragma once

// ---------------------------------------------------------------- packed f32x2
__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 bc(float v) { return pk(v, v); }  // ptxas folds this into the .F32 broadcast operand
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// ---------------------------------------------------------------- jitter RNG
// Counter-based 32-bit hash (lowbias32 finaliser) -> 24-bit uniform in [0,1).
// Same function (by specification) as orc_rng in oracle/oracle_c.c.
__device__ __forceinline__ float rrt_rng(u64 seed, uint32_t scene, uint32_t pix, uint32_t s, uint32_t axis) {
    uint32_t x = (pix * 0x9E3779B1u) ^ (scene * 0x85EBCA77u) ^ ((s * 2u + axis) * 0xC2B2AE3Du) ^
                 (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x27D4EB2Fu);
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return (float)(x >> 8) * 5.9604644775390625e-08f;
}

// ---------------------------------------------------------------- primary rays
// np.linspace(start, stop, n)[i] = fl(fl(i*step) + start), last element = stop; the
// step (stop-start)/(n-1) = +-1/(n-1) is computed once on the host in IEEE double.
__device__ __forceinline__ double lin(int i, int n, double start, double stop, double step) {
    if (n == 1) return start;
    if (i == n - 1) return stop;
    return __dadd_rn(__dmul_rn((double)i, step), start);
}

// Camera.make_rays scene.py:66-72: float64 grid, normalise, cast to float32.
__device__ __forceinline__ void base_ray(int n, double step, int i, int j, float& rx, float& ry, float& rz) {
    double x = lin(i, n, 0.5, -0.5, -step);
    double y = lin(j, n, -0.5, 0.5, step);
    double s = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), 1.0);
    double nrm = __dsqrt_rn(s);
    rx = __double2float_rn(__ddiv_rn(x, nrm));
    ry = __double2float_rn(__ddiv_rn(y, nrm));
    rz = __double2float_rn(__ddiv_rn(1.0, nrm));
}

// scene.py:31-32 then :73-74, all float32 round-to-nearest.
__device__ __forceinline__ float jitter_offset(float u, int s, int S, int n) {
    return __fdiv_rn(__fdiv_rn(__fadd_rn(u, (float)s), (float)S), (float)n);
}
// same value when S and n are powers of two (division by 2^k == multiplication by 2^-k, exact)
__device__ __forceinline__ float jitter_offset_pow2(float u, int s, float inv_s, float inv_n) {
    return __fmul_rn(__fmul_rn(__fadd_rn(u, (float)s), inv_s), inv_n);
}

__device__ __forceinline__ float dot3_canon(float a0, float a1, float a2, float v0, float v1, float v2) {
    return __fmaf_rn(a2, v2, __fmaf_rn(a1, v1, __fmul_rn(a0, v0)));
}
