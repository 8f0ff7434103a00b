"""FP32-pipe instruction-mix micro-benchmarks (include/rrt_b200_bench.h), one line per mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reversible_raytracer_b200 import render as R
names = ['scalar FFMA', 'FFMA2', 'FFMA2 + 1 FMNMX3 per 4', 'FFMA2 + 1 LDS.128 per 8', 'FFMA2 + 1 LDC per 8',
         'quadric mix 24 FFMA2 : 4 FMNMX3 : 1.5 LDS.128', 'FFMA2 with two .F32 broadcast operands', 'FFMA2 with .F32 multiplier + packed-duplicate addend']
torch.zeros(1, device='cuda')
for m, nm in enumerate(names):
    tf, ms = R.measure_fp32_peak(m, 4096)
    print('mode %d  %-48s %7.2f TFLOP/s  %.3f ms' % (m, nm, tf, ms))
