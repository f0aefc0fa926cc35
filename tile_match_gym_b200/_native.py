"""ctypes binding of libtmg_b200.so (include/tmg_b200.h).  There is no CPU fallback: if the CUDA
library is missing or no B200 is present, constructing an env raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("TMG_B200_LIB", os.path.join(_HERE, "libtmg_b200.so"))  # override: A/B builds only
SRC = os.path.join(_HERE, "csrc", "tmg_b200.cu")
DEVICE_HDR = os.path.join(_HERE, "csrc", "tmg_device.cuh")
RB_HDR = os.path.join(_HERE, "csrc", "tmg_rb.cuh")
ABI_HDR = os.path.join(ROOT, "include", "tmg_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

SP_COOKIE, SP_VERTICAL_LASER, SP_HORIZONTAL_LASER, SP_BOMB = 1, 2, 4, 8
SPECIAL_BITS = {"cookie": SP_COOKIE, "vertical_laser": SP_VERTICAL_LASER,
                "horizontal_laser": SP_HORIZONTAL_LASER, "bomb": SP_BOMB}
AUTORESET = {"disabled": 0, "next_step": 1, "same_step": 2}
REFILL = {"philox": 0, "injected": 1}
POLICY = {"uniform": 1, "mask": 2}
FLAG_NO_MASK = 1
FLAG_NO_PREGEN = 2
FLAG_BYTE_PLANES = 4
FLAG_CONSTRUCTIVE_RESET = 8
OP_BYTE_PLANES = 0x100
LINES_WORDS = 65

ST_BAD_ACTION, ST_NEEDS_RESET, ST_DRAWS_EXHAUSTED, ST_RESET_CAP = 1, 2, 4, 8
ST_LINE_OVERFLOW, ST_DFS_OVERFLOW, ST_INVALID_BOARD, ST_INTERNAL = 16, 32, 64, 128

OPS = {"gravity": 1, "refill": 2, "resolve_round": 3, "activate": 4, "combine": 5, "move": 6, "effective": 7,
       "generate": 8, "shuffle": 9, "count_lines": 10}


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("num_envs", C.c_int32),
                ("num_rows", C.c_int32), ("num_cols", C.c_int32), ("num_colours", C.c_int32),
                ("num_moves", C.c_int32), ("specials", C.c_uint32), ("autoreset", C.c_int32),
                ("refill_mode", C.c_int32), ("flags", C.c_uint32), ("max_reset_iters", C.c_int32),
                ("seed", C.c_uint64), ("env_id_offset", C.c_uint64)]


BUFFER_FIELDS = ["board", "timer", "draw_cursor", "shuffle_cursor", "reward", "terminated",
                 "is_combination_match", "num_new_specials", "num_specials_activated", "shuffled", "mask",
                 "num_moves_left", "status", "episode"]


class Buffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in BUFFER_FIELDS]


HOST_IO_FIELDS = ["actions", "board", "reward", "terminated", "mask", "mask_bits", "num_moves_left", "is_combination_match",
                  "num_new_specials", "num_specials_activated", "shuffled", "status", "board_packed"]


class HostIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in HOST_IO_FIELDS]


EXPORTS = {
    "tmg_abi_version": (C.c_int, []),
    "tmg_build_id": (C.c_char_p, []),
    "tmg_error_string": (C.c_char_p, [C.c_int]),
    "tmg_status_string": (C.c_char_p, [C.c_uint32]),
    "tmg_num_actions": (C.c_int, [C.c_int32, C.c_int32]),
    "tmg_onehot_planes": (C.c_int, [C.c_int32, C.c_uint32]),
    "tmg_action_to_coords": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32 * 4)]),
    "tmg_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "tmg_destroy": (C.c_int, [C.c_void_p]),
    "tmg_get_buffers": (C.c_int, [C.c_void_p, C.POINTER(Buffers)]),
    "tmg_set_injected_draws": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "tmg_reset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tmg_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "tmg_step_many": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tmg_rollout_policy": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tmg_legal_mask": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tmg_encode_onehot": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "tmg_encode_onehot_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "tmg_encode_onehot_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "tmg_clear_status": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tmg_join": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tmg_set_seed": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "tmg_step_host": (C.c_int, [C.c_void_p, C.POINTER(HostIO), C.c_void_p]),
    "tmg_host_bind": (C.c_int, [C.c_void_p, C.POINTER(HostIO), C.c_void_p]),
    "tmg_set_profile_buffer": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tmg_debug_op": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "tmg_debug_lines": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
}


def source_build_id() -> str:
    """sha256 over the CUDA sources, the ABI header and the compiler flags: what the library must have been built from."""
    import hashlib
    h = hashlib.sha256()
    for path in (SRC, DEVICE_HDR, RB_HDR, ABI_HDR):
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:32]


def built_build_id(path: str = None) -> str:
    """The build id embedded in a built library (read from the file, the library is not loaded)."""
    path = path or LIB_PATH
    try:
        with open(path, "rb") as f:
            blob = f.read()
    except OSError:
        return ""
    i = blob.find(b"TMG_BUILD_ID=")
    return blob[i + 13:i + 45].decode("ascii", "replace") if i >= 0 else ""


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU).  The library embeds a hash of
    the sources it was built from; it is rebuilt whenever that differs from the sources in the tree (not mtime-gated)."""
    want = source_build_id()
    if force or built_build_id() != want:
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        tmp = LIB_PATH + ".tmp"
        cmd = [nvcc] + NVCC_FLAGS + [f"-DTMG_BUILD_ID_STR=\"{want}\"", "-I", os.path.join(ROOT, "include"), "-o", tmp, SRC]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
        os.replace(tmp, LIB_PATH)
        if built_build_id() != want:
            raise RuntimeError("libtmg_b200.so does not carry the build id of its sources")
    return LIB_PATH


_lib = None


def lib():
    """Loads libtmg_b200.so.  Raises if it has not been built -- the product has no other compute path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(tile_match_gym_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            if not hasattr(L, name) and "TMG_B200_LIB" in os.environ:
                continue
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        if L.tmg_abi_version() != 2 and "TMG_B200_LIB" not in os.environ:   # (A/B builds of older sources are loaded as they are)
            raise RuntimeError("libtmg_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"tmg_b200 {what}: {lib().tmg_error_string(rc).decode()} (code {rc})")


def specials_mask(colourless_specials, colour_specials) -> int:
    m = 0
    for s in list(colourless_specials) + list(colour_specials):
        if s not in SPECIAL_BITS:
            raise ValueError(f"unknown special {s!r}")
        m |= SPECIAL_BITS[s]
    return m
