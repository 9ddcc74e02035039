"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

import vfn_testutil as U   # noqa: F401  (sys.path setup)
from vfnerf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="vfnerf_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vfnerf_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol(built_lib):
    syms = declared_symbols()
    assert len(syms) >= 14
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/vfnerf_b200.h but not exported"
    assert set(syms) == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert built_lib.vfnerf_abi_version() == 1


def test_product_library_has_no_debug_surface(built_lib):
    """Probe kernels, micro-benchmarks and the stash read-back live in the test-only library."""
    raw = ctypes.CDLL(_lib.LIB_PATH)
    assert not any(s.startswith("vfnerf_debug") for s in declared_symbols())
    for s in declared_symbols("vfnerf_b200_debug.h"):
        assert not hasattr(raw, s), f"{s} leaked into the product library"
    import subprocess
    strings = subprocess.run(["strings", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "VFNERF_TC_DBG" not in strings            # the product build reads no environment variable


def test_debug_library_exports_its_header(debug_lib):
    syms = declared_symbols("vfnerf_b200_debug.h")
    assert set(syms) == set(_lib.DEBUG_PROTOTYPES)
    raw = ctypes.CDLL(_lib.DEBUG_LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), s


def test_struct_layouts_match_header():
    """Exact sizes: the header static_asserts the same numbers (include/vfnerf_b200.h), so the C side and the ctypes
    mirror cannot drift apart silently."""
    header = open(os.path.join(ROOT, "include", "vfnerf_b200.h")).read()
    want = {name: int(n) for name, n in re.findall(r"VFNERF_STATIC_ASSERT\(sizeof\((\w+)\) == (\d+)", header)}
    assert want == {"vfnerf_mlp_desc": 912, "vfnerf_render_cfg": 120, "vfnerf_render_out": 80}
    assert ctypes.sizeof(_lib.MlpDesc) == want["vfnerf_mlp_desc"]
    assert ctypes.sizeof(_lib.RenderCfg) == want["vfnerf_render_cfg"]
    assert ctypes.sizeof(_lib.RenderOut) == want["vfnerf_render_out"]
    # field offsets that matter for alignment: the first double of render_cfg and the first int64 of mlp_desc
    assert _lib.RenderCfg.near_.offset == 48 and _lib.MlpDesc.w_off.offset == 136


def test_sass_has_no_legacy_tensor_path(built_lib):
    """The library is compiled for sm_100a only (no PTX for other archs, no multi-backend dispatch)."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(8|9)\d", out)


def test_workspace_query_errors_are_reported_not_raised(built_lib):
    cfg = _lib.RenderCfg()
    cfg.n_rays, cfg.n_coarse, cfg.n_fine, cfg.multires, cfg.multires_view, cfg.skip_layer = 8, 200, 100, 6, 4, 4
    vf, rn = _lib.MlpDesc(), _lib.MlpDesc()
    n = built_lib.vfnerf_render_workspace_bytes(ctypes.byref(cfg), ctypes.byref(vf), ctypes.byref(rn), 0)
    assert n == -1
    assert b"samples per ray" in built_lib.vfnerf_last_error() or b"VF net" in built_lib.vfnerf_last_error()
