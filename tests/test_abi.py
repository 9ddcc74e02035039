"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

import vfn_testutil as U   # noqa: F401  (sys.path setup)
from vfnerf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vfnerf_b200.h")).read()
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


def test_struct_layouts_match_header():
    # sizes the C side static-asserts implicitly through the ABI: 4 + 2*16*4 + 6*16*8 + 8 (+pad)
    assert ctypes.sizeof(_lib.MlpDesc) == 8 + 2 * 16 * 4 + 6 * 16 * 8 + 8 - 4 + 4 or ctypes.sizeof(_lib.MlpDesc) == 912
    assert ctypes.sizeof(_lib.RenderOut) == 10 * 8
    assert ctypes.sizeof(_lib.RenderCfg) == 12 * 4 + 3 * 8 + 7 * 4 + 4


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
