import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.latency import c4, timeit
fn, rays = c4(256)
us = timeit(fn, warm=3, iters=10)
print('C4 us', us, rays / us, 'Mrays/s')
