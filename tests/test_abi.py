"""CPU-side checks of the drop-in boundary: the CUDA library builds for sm_100a, loads, and exports every
symbol include/tmg_b200.h declares (no compute calls -- there is no GPU here and no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "tmg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tmg_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from tile_match_gym_b200 import _native
    path = _native.build()
    L = C.CDLL(path)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/tmg_b200.h but not exported"
    assert set(_native.EXPORTS) == set(names)
    lib = _native.lib()
    assert lib.tmg_abi_version() == 2
    assert lib.tmg_num_actions(10, 10) == 180 and lib.tmg_num_actions(9, 9) == 144 and lib.tmg_num_actions(32, 32) == 1984
    assert lib.tmg_onehot_planes(6, 15) == 10 and lib.tmg_onehot_planes(5, 0) == 5
    assert lib.tmg_status_string(1 | 8).decode() == "bad_action|reset_cap"
    out = (C.c_int32 * 4)()
    assert lib.tmg_action_to_coords(3, 5, 6, C.byref(out)) == 0 and list(out) == [1, 1, 2, 1]   # tests/test_env.py: action 6
    assert lib.tmg_action_to_coords(3, 5, 16, C.byref(out)) == 0 and list(out) == [1, 2, 1, 3]
    assert lib.tmg_action_to_coords(3, 5, 22, C.byref(out)) != 0


def test_sass_is_sm_100a_and_has_the_kernels():
    from tile_match_gym_b200 import _native
    path = _native.build()
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    for k in ("k_gate", "k_work", "k_pregen", "k_reset", "k_mask", "k_onehot", "k_debug"):
        assert k in res


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tile_match_gym_b200 import TileMatchVecEnv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TileMatchVecEnv(4, 10, 10, 4, 30, ["cookie"], ["bomb"])
    # and the C ABI itself refuses: tmg_create returns TMG_ERR_NO_DEVICE
    from tile_match_gym_b200 import _native
    lib = _native.lib()
    cfg = _native.Config(C.sizeof(_native.Config), 0, 4, 10, 10, 4, 30, 15, 0, 0, 0, 0, 1, 0)
    h = C.c_void_p()
    assert lib.tmg_create(C.byref(cfg), C.byref(h)) == 4
    assert b"no CPU fallback" in lib.tmg_error_string(4)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "tile_match_gym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions the oracle"
                assert "tests/emu" not in txt or f == "tmg_device.cuh"
