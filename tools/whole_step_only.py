import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.latency import c3, timeit
train, sc = c3(True)
assert train.state.get('whole_step') is not None
print('C3 whole-step kernel step us', timeit(train, warm=5, iters=20))
