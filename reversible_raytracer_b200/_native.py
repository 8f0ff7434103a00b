"""ctypes binding of the C ABI in include/rrt_b200.h (librrt_b200.so).

The product path has NO fallback: if the CUDA library is missing or a call
fails, an exception is raised.  Nothing under oracle/ is ever imported here.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
LIB_PATH = os.environ.get('RRT_B200_LIB') or os.path.join(HERE, 'librrt_b200.so')   # env override: A/B kernel variants
SRC = os.path.join(HERE, 'csrc', 'rrt_kernels.cu')
BENCH_LIB_PATH = os.path.join(HERE, 'librrt_b200_bench.so')       # measurement helpers (include/rrt_b200_bench.h)
BENCH_SRC = os.path.join(HERE, 'csrc', 'rrt_bench_kernels.cu')
INCLUDE = os.path.join(REPO, 'include')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']

OBJ_SPHERE, OBJ_SQUARE = 0, 1
SHADER_PHONG, SHADER_PHONG_NOSPEC, SHADER_DEPTH = 0, 1, 2
W2O_STRIDE, MAT_STRIDE, LIGHT_STRIDE, CAMERA_STRIDE = 12, 7, 6, 15
OBJ_GRAD_STRIDE, GLOBAL_GRAD = 19, 21


def grad_size(num_objects):
    return num_objects * OBJ_GRAD_STRIDE + GLOBAL_GRAD


def record_table_floats(num_scenes, num_objects):
    """RRT_RECORD_TABLE_FLOATS: sweep records [B][N][16] + pre-filter rows [B][N padded to 4][6]."""
    return num_scenes * (num_objects * 16 + (num_objects + 3) // 4 * 4 * 6)


class RrtScene(C.Structure):
    """`struct rrt_scene` of include/rrt_b200.h."""
    _fields_ = [
        ('n', C.c_int32), ('samples', C.c_int32), ('num_objects', C.c_int32),
        ('num_scenes', C.c_int32), ('shader', C.c_int32), ('transpose', C.c_int32),
        ('row_begin', C.c_int32), ('row_count', C.c_int32), ('max_depth', C.c_float),
        ('camera_grad', C.c_int32), ('seed', C.c_uint64),
        ('obj_type', C.c_void_p), ('w2o', C.c_void_p), ('material', C.c_void_p),
        ('light', C.c_void_p), ('camera', C.c_void_p),
        ('jitter_x', C.c_void_p), ('jitter_y', C.c_void_p),
        ('w2o_scene_stride', C.c_int64), ('material_scene_stride', C.c_int64),
        ('light_scene_stride', C.c_int64), ('camera_scene_stride', C.c_int64),
        ('jitter_scene_stride', C.c_int64), ('base_rays', C.c_void_p),
        ('scene_begin', C.c_int32), ('flags', C.c_int32), ('obj_records', C.c_void_p),
        ('ticket', C.c_void_p), ('det_workspace', C.c_void_p),
        ('reflectivity', C.c_void_p), ('reflectivity_scene_stride', C.c_int64),
    ]


class NativeError(RuntimeError):
    pass


def build(force=False, verbose=False):
    """Compile csrc/rrt_kernels.cu (+ its .cuh parts, one translation unit) for sm_100a into
    librrt_b200.so (in-tree)."""
    srcs = [os.path.join(os.path.dirname(SRC), f) for f in os.listdir(os.path.dirname(SRC))] + \
           [os.path.join(INCLUDE, 'rrt_b200.h'), os.path.join(INCLUDE, 'rrt_b200_bench.h')]
    newest = max(os.path.getmtime(f) for f in srcs)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    for out, src in ((LIB_PATH, SRC), (BENCH_LIB_PATH, BENCH_SRC)):
        if not force and os.path.exists(out) and os.path.getmtime(out) >= newest:
            continue
        cmd = [nvcc] + NVCC_FLAGS + ['-I', INCLUDE, '-o', out, src]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        subprocess.check_call(cmd)
    return LIB_PATH


class RrtStep(C.Structure):
    """`struct rrt_step` of include/rrt_b200.h (whole optimise step in one launch)."""
    _fields_ = [
        ('ops', C.c_void_p), ('chain_begin', C.c_void_p), ('values', C.c_void_p),
        ('num_values', C.c_int32), ('param_begin', C.c_int32), ('lr', C.c_float), ('reserved', C.c_int32),
        ('grad', C.c_void_p), ('g_values', C.c_void_p), ('loss_acc', C.c_void_p), ('loss_out', C.c_void_p),
        ('ticket', C.c_void_p),
    ]


_lib = None


def lib():
    """Load the CUDA library; raises NativeError when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError('librrt_b200.so is not built (run `python -c "import __graft_entry__ as g; g.build()"`); '
                              'there is no CPU fallback')
        L = C.CDLL(LIB_PATH)
        L.rrt_version.restype = C.c_int
        L.rrt_last_error.restype = C.c_char_p
        P = C.c_void_p
        L.rrt_render_forward.argtypes = [C.POINTER(RrtScene), P, P, P, P]
        L.rrt_render_backward.argtypes = [C.POINTER(RrtScene), P, P, P, P]
        L.rrt_render_fused_mse.argtypes = [C.POINTER(RrtScene), P, C.POINTER(C.c_float), P, P, P, P, P]
        L.rrt_primary_rays.argtypes = [C.c_int, P, P]
        L.rrt_chain_forward.argtypes = [P, P, C.c_int, P, P, P]
        L.rrt_chain_backward.argtypes = [P, P, C.c_int, P, P, P, C.c_int, P]
        if hasattr(L, 'rrt_small_step_mse'):
            L.rrt_small_step_mse.argtypes = [C.POINTER(RrtScene), C.POINTER(RrtStep), P, C.POINTER(C.c_float), P, P]
            L.rrt_small_step_mse.restype = C.c_int
        if hasattr(L, 'rrt_build_records'):      # absent only in older A/B builds loaded through RRT_B200_LIB
            L.rrt_build_records.argtypes = [C.POINTER(RrtScene), P, P]
            L.rrt_build_records.restype = C.c_int
        L.rrt_peer_buffer_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
        L.rrt_peer_buffer_bytes.restype = C.c_size_t
        L.rrt_peer_signal_bytes.argtypes = []
        L.rrt_peer_signal_bytes.restype = C.c_size_t
        L.rrt_peer_allreduce.argtypes = [P, P, C.c_int, C.c_int, P, P, C.c_int, C.c_int, P, P]
        L.rrt_peer_allreduce.restype = C.c_int
        for f in (L.rrt_render_forward, L.rrt_render_backward, L.rrt_render_fused_mse,
                  L.rrt_chain_forward, L.rrt_chain_backward, L.rrt_primary_rays):
            f.restype = C.c_int
        _lib = L
    return _lib


_bench = None


def bench_lib():
    """librrt_b200_bench.so: measurement helpers (FP32 peak micro-benchmarks), used by bench.py only."""
    global _bench
    if _bench is None:
        if not os.path.exists(BENCH_LIB_PATH):
            raise NativeError('librrt_b200_bench.so is not built (run __graft_entry__.build())')
        L = C.CDLL(BENCH_LIB_PATH)
        L.rrt_bench_fp32_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]
        L.rrt_bench_fp32_peak.restype = C.c_int
        _bench = L
    return _bench


BENCH_EXPORTS = ['rrt_bench_fp32_peak']

EXPORTS = ['rrt_version', 'rrt_last_error', 'rrt_render_forward', 'rrt_render_backward',
           'rrt_render_fused_mse', 'rrt_chain_forward', 'rrt_chain_backward', 'rrt_primary_rays',
           'rrt_peer_allreduce', 'rrt_peer_buffer_bytes', 'rrt_peer_signal_bytes', 'rrt_build_records',
           'rrt_small_step_mse']

FLAG_CULL, FLAG_NO_SMALL, FLAG_SHADOWS, FLAG_SCALAR_SHADOWS, FLAG_NO_MATERIAL_GRAD, FLAG_CANONICAL_SWEEP, FLAG_DETERMINISTIC, FLAG_MIRROR = 1, 2, 4, 8, 16, 32, 64, 128
FLAG_PIXEL_THREADS, FLAG_RAY_THREADS, FLAG_LINEAR_COST = 256, 512, 1024
HIT_SHADOWED = 0x40000000
CHAIN_TRANSLATE, CHAIN_SCALE, CHAIN_ROTATE, CHAIN_INVERT, CHAIN_MAX_OPS = 1, 2, 3, 0x100, 8


def check(rc, what):
    if rc != 0:
        raise NativeError('%s failed (%d): %s' % (what, rc, lib().rrt_last_error().decode()))
