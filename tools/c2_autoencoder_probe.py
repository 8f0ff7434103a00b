"""examples/test_balls.py (the reference's test_balls.py: MLP encoder + depth-map renderer, MGDAutoOptimizer):
time per epoch with the step captured into a CUDA graph vs eager stepping."""
import os, sys, time, importlib.util
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import torch
spec = importlib.util.spec_from_file_location('ex_test_balls', os.path.join(root, 'examples', 'test_balls.py'))
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
for graph in ('auto', False):
    mod.main(num_epoch=10, verbose=False, graph=graph)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    losses = mod.main(num_epoch=200, verbose=False, graph=graph)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('test_balls graph=%s: %.1f us per epoch (200 epochs incl. set-up), captured=%s, loss %.4f -> %.4f' %
          (graph, dt / 200 * 1e6, mod.main.last_state['graph'] is not None, losses[0], losses[-1]))
