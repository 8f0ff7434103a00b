import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.latency import c3, timeit
train, sc = c3('graph')
print('C3 fused step us', timeit(train, warm=5, iters=20))
