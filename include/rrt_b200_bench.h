/*
 * rrt_b200_bench.h -- C ABI of librrt_b200_bench.so: MEASUREMENT helpers only.
 *
 * Not part of the product library (librrt_b200.so, include/rrt_b200.h): these entry points
 * allocate, synchronise and time, which the product ABI never does.  bench.py loads this
 * library to measure the FP32-pipe roofline denominator on the box it runs on, because the
 * driver-written MEASURED_PEAKS.json has no FP32 entry.  The reference has no counterpart.
 */
#ifndef RRT_B200_BENCH_H
#define RRT_B200_BENCH_H

#ifdef __cplusplus
extern "C" {
#endif

/*
 * FP32 pipe micro-benchmarks.  One persistent grid of (SM count x 8) CTAs x 256 threads, each
 * thread 16 independent accumulator chains, `iters` trips of 128 FMA instructions.
 *   mode 0  scalar FFMA                                   (fma.rn.f32)
 *   mode 1  packed FFMA2                                  (fma.rn.f32x2)  <- roofline denominator
 *   mode 2  FFMA2 + one ALU-pipe FMNMX3 per 4 FFMA2       (does anything issue in an FFMA2's shadow?)
 *   mode 3  FFMA2 + one broadcast LDS.128 per 8 FFMA2     (object constants from shared memory)
 *   mode 4  FFMA2 + one LDC (constant bank, uniform dynamic address) per 8 FFMA2
 *   mode 5  a pre-filter-like mix: 24 FFMA2 : 4 FMNMX3 : 1.5 LDS.128
 *   mode 6  FFMA2 with multiplier and addend as scalar-broadcast (.F32) operands (the pre-filter's Horner steps)
 *   mode 7  FFMA2 with a broadcast multiplier and a pre-duplicated packed addend
 * tflops counts 2 flops per FMA lane (the other instructions are overhead, not credited).
 * Synchronises the stream.  tflops / ms are HOST pointers.  Returns 0, or -1 (bad argument),
 * -2 (CUDA error).
 */
int rrt_bench_fp32_peak(int mode, int iters, double* tflops, double* ms, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RRT_B200_BENCH_H */
