"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU lane emulator (tests/emu/emu_main.cpp), which
compiles the product's device code (tile_match_gym_b200/csrc/tmg_device.cuh) for the host so its logic can
be fuzzed against the oracle in a container without a GPU.  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libtmg_emu.so")
_DEV = os.path.join(_HERE, "..", "..", "tile_match_gym_b200", "csrc", "tmg_device.cuh")

SPECIAL_BITS = {"cookie": 1, "vertical_laser": 2, "horizontal_laser": 4, "bomb": 8}


class _Cfg(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("num_rows", C.c_int32), ("num_cols", C.c_int32), ("num_colours", C.c_int32),
                ("num_moves", C.c_int32), ("specials", C.c_uint32), ("autoreset", C.c_int32), ("refill_mode", C.c_int32),
                ("flags", C.c_uint32), ("max_reset_iters", C.c_int32), ("seed", C.c_uint64), ("env_id_offset", C.c_uint64)]


_FIELDS = ["board", "timer", "draw_cursor", "shuffle_cursor", "reward", "terminated", "is_combination_match",
           "num_new_specials", "num_specials_activated", "shuffled", "mask", "num_moves_left", "status", "episode"]
_NP = {"board": np.int8, "timer": np.int32, "draw_cursor": np.uint64, "shuffle_cursor": np.uint64, "reward": np.int32,
       "terminated": np.uint8, "is_combination_match": np.uint8, "num_new_specials": np.int32,
       "num_specials_activated": np.int32, "shuffled": np.uint8, "mask": np.uint8, "num_moves_left": np.int32,
       "status": np.uint32, "episode": np.int32}


class _Bufs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _FIELDS]


_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, "emu_main.cpp"), os.path.join(_HERE, "emu_shim.h"), _DEV, _DEV.replace("tmg_device.cuh", "tmg_rb.cuh")]
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(s) for s in srcs):
            subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-I", _HERE, "-o", _LIB,
                                   os.path.join(_HERE, "emu_main.cpp")])
        L = C.CDLL(_LIB)
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.POINTER(_Cfg)]
        L.emu_destroy.argtypes = [C.c_void_p]
        L.emu_get_buffers.argtypes = [C.c_void_p, C.POINTER(_Bufs)]
        L.emu_host_bind.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.emu_host_bind_packed.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_set_injected_draws.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.emu_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_step.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_step_many.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.emu_legal_mask.argtypes = [C.c_void_p]
        L.emu_debug_op.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.emu_debug_lines.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class EmuVecEnv:
    def __init__(self, num_envs, num_rows, num_cols, num_colours, num_moves, colourless_specials=(), colour_specials=(),
                 seed=1, autoreset="disabled", refill="philox", env_id_offset=0, max_reset_iters=16384, flags=0):
        self.L = lib()
        sp = 0
        for s in list(colourless_specials) + list(colour_specials):
            sp |= SPECIAL_BITS[s]
        cfg = _Cfg(num_envs, num_rows, num_cols, num_colours, num_moves, sp,
                   {"disabled": 0, "next_step": 1, "same_step": 2}[autoreset], {"philox": 0, "injected": 1}[refill],
                   flags, max_reset_iters, seed, env_id_offset)
        self.h = self.L.emu_create(C.byref(cfg))
        self.N, self.R, self.Cc, self.K = num_envs, num_rows, num_cols, num_colours
        self.A = 2 * num_rows * num_cols - num_rows - num_cols
        bufs = _Bufs()
        self.L.emu_get_buffers(self.h, C.byref(bufs))
        shapes = {"board": (self.N, 2, self.R, self.Cc), "mask": (self.N, self.A)}
        for name in _FIELDS:
            shape = shapes.get(name, (self.N,))
            dt = np.dtype(_NP[name])
            raw = (C.c_char * (int(np.prod(shape)) * dt.itemsize)).from_address(getattr(bufs, name))
            setattr(self, name, np.frombuffer(raw, dtype=dt).reshape(shape))
        self._keep = []

    def __del__(self):
        try:
            self.L.emu_destroy(self.h)
        except Exception:
            pass

    def set_injected_draws(self, draws):
        d = np.ascontiguousarray(draws, dtype=np.uint8)
        self._keep.append(d)
        self.L.emu_set_injected_draws(self.h, _ptr(d), d.shape[1])

    def host_bind(self):
        """Mirror arrays (plain memory here) initialised from the current state, as tmg_host_bind does."""
        bpe = (self.A + 7) // 8
        self.h_board = self.board.copy()
        self.h_mask = self.mask.copy()
        self.h_mask_bits = np.packbits(self.mask, axis=1, bitorder="little")[:, :bpe].copy()
        self.h_reward = self.reward.copy(); self.h_terminated = self.terminated.copy()
        self.h_moves_left = self.num_moves_left.copy()
        self.L.emu_host_bind(self.h, _ptr(self.h_board), _ptr(self.h_mask), _ptr(self.h_mask_bits), _ptr(self.h_reward),
                             _ptr(self.h_terminated), _ptr(self.h_moves_left))
        self.h_board_packed = ((self.board[:, 0] & 15) | ((self.board[:, 1] & 7) << 4)).astype(np.uint8).copy()
        self.L.emu_host_bind_packed(self.h, _ptr(self.h_board_packed))

    def reset(self, reset_mask=None, init_boards=None):
        m = None if reset_mask is None else np.ascontiguousarray(reset_mask, dtype=np.uint8)
        b = None if init_boards is None else np.ascontiguousarray(init_boards, dtype=np.int8)
        self.L.emu_reset(self.h, None if m is None else _ptr(m), None if b is None else _ptr(b))

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        self.L.emu_step(self.h, _ptr(a))

    def step_many(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        T = a.shape[0]
        rew = np.zeros((T, self.N), np.int32); term = np.zeros((T, self.N), np.uint8)
        self.L.emu_step_many(self.h, _ptr(a), T, _ptr(rew), _ptr(term), 0, None)
        return rew, term

    def rollout(self, T, policy):
        act = np.zeros((T, self.N), np.int32); rew = np.zeros((T, self.N), np.int32); term = np.zeros((T, self.N), np.uint8)
        self.L.emu_step_many(self.h, None, T, _ptr(rew), _ptr(term), {"uniform": 1, "mask": 2}[policy], _ptr(act))
        return act, rew, term

    def legal_mask(self):
        self.L.emu_legal_mask(self.h)

    def debug_op(self, op, args=None):
        a = None if args is None else np.ascontiguousarray(args, dtype=np.int32)
        self.L.emu_debug_op(self.h, op, None if a is None else _ptr(a))

    def debug_lines(self, byte_planes=False):
        out = np.zeros((self.N, 65), np.uint32)
        self.L.emu_debug_lines(self.h, _ptr(out), int(byte_planes))
        return out
