"""ctypes binding of libvfnerf_b200.so (the C ABI declared in include/vfnerf_b200.h).

The library is built in-tree by ``build()`` (plain ``nvcc -shared`` for sm_100a, no torch headers)
and loaded lazily.  There is NO fallback: if the shared object is missing or a call returns a
non-zero status, a RuntimeError is raised -- the product never routes around the CUDA path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libvfnerf_b200.so")
SOURCES = ["api.cu", "geometry_sampler.cu", "density_composite.cu", "render_fused.cu", "mlp_simt.cu", "mlp_train.cu", "mlp_tc.cu", "mlp_tc_bwd.cu", "loss.cu", "optim.cu", "mc_preprocess.cu", "host_rng.cu", "supervision.cu"]
HEADERS = ["common.cuh", "host_plan.cuh", "ray_ops.cuh", "mlp_tc.cuh", "tc_common.cuh", os.path.join("..", "..", "include", "vfnerf_b200.h")]
# test-only library (UMMA probes, micro-benchmarks, stash read-back): the product sources + tc_debug.cu, compiled with
# -DVFNERF_DEBUG_EXPORTS; built on demand by the tests (tests/conftest.py: debug_lib), never loaded by the product
DEBUG_LIB_PATH = os.path.join(HERE, "libvfnerf_b200_debug.so")
DEBUG_SOURCES = SOURCES + ["tc_debug.cu"]
DEBUG_HEADERS = HEADERS + [os.path.join("..", "..", "include", "vfnerf_b200_debug.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--threads", "0"]
SQNORM_SCRATCH_FLOATS = 1024

MAX_LAYERS = 16
MAX_SAMPLES = 256
PREC_FP32, PREC_BF16, PREC_BF16X3, PREC_FP16F8 = 0, 1, 2, 3
FLAG_RECOMPUTE_COARSE, FLAG_WHITE_BG, FLAG_NERF_WEIGHTS, FLAG_WEIGHTS_PACKED = 1, 2, 4, 8
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16x3": PREC_BF16X3, "fp16f8": PREC_FP16F8}


class MlpDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int32),
                ("in_dim", C.c_int32 * MAX_LAYERS), ("out_dim", C.c_int32 * MAX_LAYERS),
                ("w_off", C.c_int64 * MAX_LAYERS), ("b_off", C.c_int64 * MAX_LAYERS),
                ("gamma_off", C.c_int64 * MAX_LAYERS), ("beta_off", C.c_int64 * MAX_LAYERS),
                ("mean_off", C.c_int64 * MAX_LAYERS), ("var_off", C.c_int64 * MAX_LAYERS),
                ("arena_floats", C.c_int64)]


class RenderCfg(C.Structure):
    _fields_ = [("n_rays", C.c_int32), ("n_coarse", C.c_int32), ("n_fine", C.c_int32),
                ("perturb", C.c_int32), ("pose_is_quat", C.c_int32), ("window", C.c_int32),
                ("normalize", C.c_int32), ("multires", C.c_int32), ("multires_view", C.c_int32),
                ("skip_layer", C.c_int32), ("precision", C.c_int32), ("flags", C.c_int32),
                ("near_", C.c_double), ("far_", C.c_double), ("fine_range", C.c_double),
                ("dir_to_normal_th", C.c_float),
                ("beta_lo", C.c_float), ("beta_hi", C.c_float), ("mean_lo", C.c_float),
                ("mean_hi", C.c_float), ("scale_min", C.c_float), ("bn_eps", C.c_float),
                ("fine_near_", C.c_double), ("fine_far_", C.c_double)]


class RenderOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("points", "normals", "rgb", "depth", "z_vals", "ray_dirs_rep",
                                           "colors", "weights", "z_coarse", "weights_coarse")]


# every symbol include/vfnerf_b200.h declares, with its ctypes prototype
_P, _I, _L, _D, _F = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_float
_DESC, _CFG, _OUT = C.POINTER(MlpDesc), C.POINTER(RenderCfg), C.POINTER(RenderOut)
PROTOTYPES = {
    "vfnerf_abi_version": (C.c_int, []),
    "vfnerf_last_error": (C.c_char_p, []),
    "vfnerf_launch_count": (C.c_longlong, []),
    "vfnerf_render_workspace_bytes": (_L, [_CFG, _DESC, _DESC, _I]),
    "vfnerf_render_fwd": (_I, [_CFG, _DESC, _P, _DESC, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _OUT, _P, _L, _I, _P]),
    "vfnerf_render_bwd": (_I, [_CFG, _DESC, _P, _DESC, _P, _P, _OUT, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P]),
    "vfnerf_vf_workspace_bytes": (_L, [_DESC, _L, _I, _I, _I]),
    "vfnerf_vf_fwd": (_I, [_DESC, _P, _I, _I, _F, _I, _P, _L, _P, _L, _I, _P, _L, _I, _P]),
    "vfnerf_vf_bwd": (_I, [_DESC, _P, _I, _I, _F, _I, _L, _P, _L, _P, _L, _I, _P, _I, _P, _L, _P]),
    "vfnerf_vf_grid_query": (_I, [_DESC, _P, _I, _I, _F, _I, _I, _L, _L, C.POINTER(C.c_float),
                                  C.POINTER(C.c_float), C.POINTER(C.c_float), _F, _P, _P, _L, _P]),
    "vfnerf_vf_loss_fwd": (_I, [_L, _L, _L, _L, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(C.c_float), _F, _I, _P, _P]),
    "vfnerf_vf_loss_bwd": (_I, [_L, _L, _L, _P, _P, _P, _P, _P, _P, _P, C.POINTER(C.c_float), _F, _I, _P, _P, _P, _P, _P, _P]),
    "vfnerf_sqnorm_accumulate": (_I, [_P, _L, _F, _P, _P, _P]),
    "vfnerf_adam_step": (_I, [_P, _P, _P, _P, _P, _L, _P, _P, _F, _F, _F, _F, _F, _P, _F, _P]),
    "vfnerf_mlp_points_workspace_bytes": (_L, [_DESC, _DESC, _I, _I, _I]),
    "vfnerf_mlp_points_fwd": (_I, [_DESC, _P, _DESC, _P, _I, _I, _I, _F, _I, _P, _P, _I, _L, _P, _P, _P, _L, _I, _P]),
    "vfnerf_ray_geometry": (_I, [_I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "vfnerf_coarse_sample": (_I, [_I, _I, _D, _D, _I, _P, _P, _P, _P, _P, _P, _P]),
    "vfnerf_fine_sample": (_I, [_I, _I, _I, _D, _D, _D, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vfnerf_mt19937_uniform": (_I, [_P, _P, _P, _L, _P]),
    "vfnerf_smooth_vf": (_I, [_P, _P, _P, _I, _I, C.POINTER(C.c_float), _P]),
    "vfnerf_mc_count": (_I, [_P, _I, _P, _P, _P, _P, _P, _P]),
    "vfnerf_mc_emit": (_I, [_P, _I, _P, _P, _P, _P, _P, _P]),
    "vfnerf_sample_pdf": (_I, [_I, _I, _I, _P, _P, _P, _I, _P, _P]),
    "vfnerf_pdf_fine_sample": (_I, [_I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P]),
    "vfnerf_density_weights": (_I, [_CFG, _I, _P, _P, _L, _P, _P, _P, _P, _P, _P]),
    "vfnerf_volume_weights": (_I, [_I, _I, _I, _I, _P, _P, _P, _P]),
    "vfnerf_composite": (_I, [_I, _I, _P, _P, _P, _P, _P, _P]),
    "vfnerf_composite_white": (_I, [_I, _I, _P, _P, _P, _P, _P, _P]),
    "vfnerf_sphere_points": (_I, [_L, _P, _P, _P, _D, _D, C.POINTER(C.c_float), _I, _P, _P, _P]),
    "vfnerf_select_supervised": (_I, [_L, _P, C.POINTER(C.c_float), _F, _I, _P, _P, _P]),
    "vfnerf_render_train_workspace_bytes": (_L, [_CFG, _DESC, _DESC]),
    "vfnerf_render_train_fwd": (_I, [_CFG, _DESC, _P, _DESC, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _OUT, _P, _F, _P, _L, _P]),
    "vfnerf_render_train_bwd": (_I, [_CFG, _DESC, _P, _DESC, _P, _P, _OUT, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P]),
    "vfnerf_vf_train_workspace_bytes": (_L, [_DESC, _L, _I]),
    "vfnerf_vf_train_fwd": (_I, [_DESC, _P, _I, _I, _F, _F, _P, _L, _P, _L, _I, _P, _L, _P, _L, _P]),
    "vfnerf_vf_train_bwd": (_I, [_DESC, _P, _I, _I, _L, _P, _L, _P, _L, _I, _P, _I, _P, _L, _P]),
}

DEBUG_PROTOTYPES = {
    "vfnerf_debug_last_error": (C.c_char_p, []),
    "vfnerf_debug_umma_gemm": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "vfnerf_debug_umma_mn_gemm": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "vfnerf_debug_umma2_gemm": (_I, [_P, _P, _P, _I, _I, _P]),
    "vfnerf_debug_umma2_alt_gemm": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "vfnerf_debug_umma2_m128_probe": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "vfnerf_debug_umma2_bench": (_I, [_I, _I, _I, _P, _P]),
    "vfnerf_debug_stash_read": (_I, [_P, _P, _P, _P, _I, _P, _P, _P]),
    "vfnerf_debug_umma_bench": (_I, [_I, _I, _I, _I, _P, _P]),
    "vfnerf_debug_chunk_table": (_I, [_CFG, _DESC, _DESC, _I, _I, _P, _P, _P, _P]),
}

_lib: Optional[C.CDLL] = None
_dbg: Optional[C.CDLL] = None
_lock = threading.Lock()


def _stale(path=None, sources=None, headers=None) -> bool:
    path, sources, headers = path or LIB_PATH, sources or SOURCES, headers or HEADERS
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    deps = [os.path.join(CSRC, s) for s in sources] + [os.path.normpath(os.path.join(CSRC, h)) for h in headers]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _compile(path: str, sources, defines, verbose: bool) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    tmp = path + ".building.so"      # swapped in atomically: a concurrent snapshot/load never sees a partial file
    extra = os.environ.get("NVCC_EXTRA", "").split()     # e.g. -DVFNERF_TC_PROFILE for the in-kernel cycle counters
    cmd = [nvcc] + NVCC_FLAGS + list(defines) + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + \
          [os.path.join(CSRC, s) for s in sources]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(tmp, path)
    if verbose:
        print(res.stderr)
    return path


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into vfnerf_b200/libvfnerf_b200.so (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    return _compile(LIB_PATH, SOURCES, [], verbose)


def build_debug(force: bool = False, verbose: bool = False) -> str:
    """The test-only library (include/vfnerf_b200_debug.h) -> vfnerf_b200/libvfnerf_b200_debug.so."""
    if not force and not _stale(DEBUG_LIB_PATH, DEBUG_SOURCES, DEBUG_HEADERS):
        return DEBUG_LIB_PATH
    return _compile(DEBUG_LIB_PATH, DEBUG_SOURCES, ["-DVFNERF_DEBUG_EXPORTS"], verbose)


def debug_lib() -> C.CDLL:
    """The loaded test-only library (tests and profiling scripts only)."""
    global _dbg
    if _dbg is None:
        with _lock:
            if _dbg is None:
                if not os.path.exists(DEBUG_LIB_PATH):
                    raise RuntimeError(f"{DEBUG_LIB_PATH} is missing: call vfnerf_b200._lib.build_debug() first (test-only library)")
                L = C.CDLL(DEBUG_LIB_PATH)
                for name, (res, args) in DEBUG_PROTOTYPES.items():
                    fn = getattr(L, name)
                    fn.restype = res
                    fn.argtypes = args
                _dbg = L
    return _dbg


def check_debug(status: int, what: str) -> None:
    if status != 0:
        msg = debug_lib().vfnerf_debug_last_error()
        raise RuntimeError(f"{what} failed (status {status}): {msg.decode() if msg else '?'}")


def lib() -> C.CDLL:
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: the CUDA extension has not been built (run "
                        "`python -c 'import __graft_entry__ as g; g.build()'`). vfnerf_b200 has no CPU fallback.")
                L = C.CDLL(LIB_PATH)
                for name, (res, args) in PROTOTYPES.items():
                    fn = getattr(L, name)      # AttributeError if a declared symbol is not exported
                    fn.restype = res
                    fn.argtypes = args
                _lib = L
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().vfnerf_last_error()
        raise RuntimeError(f"{what} failed (status {status}): {msg.decode() if msg else '?'}")


def ptr(t) -> Optional[int]:
    """Device pointer of a (contiguous, fp32) torch tensor, or None."""
    return None if t is None else t.data_ptr()
