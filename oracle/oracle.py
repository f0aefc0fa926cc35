"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the C oracle (oracle/tmg_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  `OracleBoard` mirrors the reference's `Board` (board.py:41-726) method
for method so that parity tests read like the reference's own tests; `OracleVecEnv` mirrors
the product's vectorised semantics (include/tmg_b200.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtmg_oracle.so")

SPECIAL_BITS = {"cookie": 1, "vertical_laser": 2, "horizontal_laser": 4, "bomb": 8}
NAMES = ["normal", "vertical_laser", "horizontal_laser", "bomb", "cookie"]

ST_BAD_ACTION, ST_NEEDS_RESET, ST_DRAWS_EXHAUSTED, ST_RESET_CAP = 1, 2, 4, 8
ST_LINE_OVERFLOW, ST_DFS_OVERFLOW, ST_INVALID_BOARD, ST_INTERNAL = 16, 32, 64, 128


def specials_mask(colourless_specials, colour_specials) -> int:
    m = 0
    for s in list(colourless_specials) + list(colour_specials):
        m |= SPECIAL_BITS[s]
    return m


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tmg_oracle.c")
    hdr = os.path.join(_HERE, "tmg_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libtmg_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _VecConfig(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("num_rows", C.c_int32), ("num_cols", C.c_int32),
                ("num_colours", C.c_int32), ("num_moves", C.c_int32), ("specials", C.c_uint32),
                ("autoreset", C.c_int32), ("refill_mode", C.c_int32), ("seed", C.c_uint64),
                ("env_id_offset", C.c_uint64), ("max_reset_iters", C.c_int64), ("num_threads", C.c_int32),
                ("flags", C.c_uint32)]


class _VecBuffers(C.Structure):
    _fields_ = [("board", C.c_void_p), ("timer", C.c_void_p), ("draw_cursor", C.c_void_p),
                ("shuffle_cursor", C.c_void_p), ("reward", C.c_void_p), ("terminated", C.c_void_p),
                ("is_combination_match", C.c_void_p), ("num_new_specials", C.c_void_p),
                ("num_specials_activated", C.c_void_p), ("shuffled", C.c_void_p), ("mask", C.c_void_p),
                ("num_moves_left", C.c_void_p), ("status", C.c_void_p), ("episode", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp, i32, u32, u64, i64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_int64
    sig = {
        "tmgo_board_create": (vp, [i32, i32, i32, u32]),
        "tmgo_board_destroy": (None, [vp]),
        "tmgo_board_set_stream": (None, [vp, u64, u32, u64, u64]),
        "tmgo_board_set_injected": (None, [vp, vp, i64, i64]),
        "tmgo_board_get_cursors": (None, [vp, C.POINTER(u64), C.POINTER(u64)]),
        "tmgo_board_status": (u32, [vp]),
        "tmgo_board_set_episode": (None, [vp, i32]),
        "tmgo_board_get_episode": (i32, [vp]),
        "tmgo_board_set": (None, [vp, vp]),
        "tmgo_board_get": (None, [vp, vp]),
        "tmgo_board_set_counters": (None, [vp, i32, i32]),
        "tmgo_board_get_counters": (None, [vp, C.POINTER(i32), C.POINTER(i32)]),
        "tmgo_board_set_iter_cap": (None, [vp, i64]),
        "tmgo_board_diag": (None, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]),
        "tmgo_num_actions": (i32, [vp]),
        "tmgo_generate_board": (None, [vp]),
        "tmgo_shuffle": (None, [vp]),
        "tmgo_gravity": (None, [vp]),
        "tmgo_refill": (None, [vp]),
        "tmgo_is_move_legal": (i32, [vp, i32, i32, i32, i32]),
        "tmgo_is_move_effective": (i32, [vp, i32, i32, i32, i32]),
        "tmgo_possible_move": (i32, [vp]),
        "tmgo_get_colour_lines": (i32, [vp, vp, vp, i32, i32]),
        "tmgo_detect_colour_matches": (i32, [vp, vp, vp, vp, vp, i32, i32]),
        "tmgo_resolve_round": (i32, [vp]),
        "tmgo_special_creation_pos": (i32, [vp, vp, i32, vp, i32, i32]),
        "tmgo_activate_special": (None, [vp, i32, i32, i32, i32]),
        "tmgo_combination_match": (None, [vp, i32, i32, i32, i32]),
        "tmgo_move": (i32, [vp, i32, i32, i32, i32, vp]),
        "tmgo_effective_mask": (None, [vp, vp]),
        "tmgo_onehot_planes": (i32, [vp]),
        "tmgo_onehot": (None, [vp, vp]),
        "tmgo_vec_create": (vp, [C.POINTER(_VecConfig)]),
        "tmgo_vec_destroy": (None, [vp]),
        "tmgo_vec_get_buffers": (None, [vp, C.POINTER(_VecBuffers)]),
        "tmgo_vec_set_injected_draws": (None, [vp, vp, i64]),
        "tmgo_vec_reset": (None, [vp, vp, vp]),
        "tmgo_vec_step": (None, [vp, vp]),
        "tmgo_vec_rollout": (i64, [vp, i32, u64, u64]),
        "tmgo_vec_onehot": (None, [vp, vp]),
        "tmgo_vec_diag": (None, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]),
        "tmgo_philox4x32_10": (None, [vp, vp, vp]),
        "tmgo_stream_word": (u32, [u64, u32, u32, u64]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleBoard:
    """Mirror of reference `Board` (board.py:41).  `board` is an int32 (2,R,C) array that is
    pushed to / pulled from the C state around every call."""

    def __init__(self, num_rows, num_cols, num_colours, colourless_specials=("cookie",),
                 colour_specials=("vertical_laser", "horizontal_laser", "bomb"), seed=0, env_id=0,
                 board=None, injected_draws=None):
        self.L = lib()
        self.num_rows, self.num_cols, self.num_colours = int(num_rows), int(num_cols), int(num_colours)
        self.specials = specials_mask(colourless_specials, colour_specials)
        self.h = self.L.tmgo_board_create(self.num_rows, self.num_cols, self.num_colours, self.specials)
        if not self.h:
            raise ValueError("bad board shape")
        self.L.tmgo_board_set_stream(self.h, int(seed), int(env_id), 0, 0)
        self._inj = None
        if injected_draws is not None:
            self.set_injected(injected_draws)
        self.num_actions = self.L.tmgo_num_actions(self.h)
        self.flat_size = self.num_rows * self.num_cols
        self.board = np.zeros((2, self.num_rows, self.num_cols), dtype=np.int32)
        if board is not None:
            b = np.asarray(board, dtype=np.int32)
            if b.ndim == 2:
                b = np.stack([b, np.ones_like(b)])
            self.board = b.copy()
        self.action_to_coords = tuple(self._a2c(i) for i in range(self.num_actions))

    def __del__(self):
        try:
            self.L.tmgo_board_destroy(self.h)
        except Exception:
            pass

    def _a2c(self, i):
        R, Cc = self.num_rows, self.num_cols
        if i < Cc * (R - 1):
            return ((i // Cc, i % Cc), (i // Cc + 1, i % Cc))
        j = i - Cc * (R - 1)
        return ((j // (Cc - 1), j % (Cc - 1)), (j // (Cc - 1), j % (Cc - 1) + 1))

    # -- state sync ------------------------------------------------------------------------
    def _push(self):
        self._buf = np.ascontiguousarray(self.board, dtype=np.int32)
        self.L.tmgo_board_set(self.h, _ptr(self._buf))

    def _pull(self):
        out = np.empty((2, self.num_rows, self.num_cols), dtype=np.int32)
        self.L.tmgo_board_get(self.h, _ptr(out))
        self.board = out

    def set_injected(self, draws, cursor=0):
        self._inj = np.ascontiguousarray(draws, dtype=np.uint8)
        self.L.tmgo_board_set_injected(self.h, _ptr(self._inj), len(self._inj), int(cursor))

    def set_stream(self, seed, env_id, draw_cursor=0, shuffle_cursor=0):
        self.L.tmgo_board_set_stream(self.h, int(seed), int(env_id), int(draw_cursor), int(shuffle_cursor))

    @property
    def cursors(self):
        a, b = C.c_uint64(), C.c_uint64()
        self.L.tmgo_board_get_cursors(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    @property
    def status(self):
        return self.L.tmgo_board_status(self.h)

    @property
    def episode(self):
        return self.L.tmgo_board_get_episode(self.h)

    def set_episode(self, episode):
        self.L.tmgo_board_set_episode(self.h, int(episode))

    @property
    def counters(self):
        a, b = C.c_int(), C.c_int()
        self.L.tmgo_board_get_counters(self.h, C.byref(a), C.byref(b))
        return a.value, b.value  # (num_new_specials, num_specials_activated)

    def set_counters(self, num_new_specials=0, num_specials_activated=0):
        self.L.tmgo_board_set_counters(self.h, num_new_specials, num_specials_activated)

    def set_iter_cap(self, cap):
        self.L.tmgo_board_set_iter_cap(self.h, int(cap))

    def diag(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int64()
        self.L.tmgo_board_diag(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"max_lines": a.value, "max_dfs_depth": b.value, "reset_iters": c.value}

    # -- reference API ---------------------------------------------------------------------
    def generate_board(self):
        self.L.tmgo_generate_board(self.h)
        self._pull()

    def shuffle(self):
        self._push(); self.L.tmgo_shuffle(self.h); self._pull()

    def gravity(self):
        self._push(); self.L.tmgo_gravity(self.h); self._pull()

    def refill(self):
        self._push(); self.L.tmgo_refill(self.h); self._pull()

    def is_move_legal(self, c1, c2):
        return bool(self.L.tmgo_is_move_legal(self.h, c1[0], c1[1], c2[0], c2[1]))

    def is_move_effective(self, c1, c2):
        self._push()
        return bool(self.L.tmgo_is_move_effective(self.h, c1[0], c1[1], c2[0], c2[1]))

    def possible_move(self):
        self._push()
        return bool(self.L.tmgo_possible_move(self.h))

    def _lists(self, n, cells, offs):
        Cc = self.num_cols
        return [[(int(x) // Cc, int(x) % Cc) for x in cells[offs[i]:offs[i + 1]]] for i in range(n)]

    def get_colour_lines(self):
        self._push()
        cap = 4 * self.flat_size + 64
        cells = np.zeros(cap * 8, dtype=np.int32)
        offs = np.zeros(cap + 1, dtype=np.int32)
        n = self.L.tmgo_get_colour_lines(self.h, _ptr(cells), _ptr(offs), len(cells), cap)
        assert n >= 0
        return self._lists(n, cells, offs)

    def detect_colour_matches(self):
        self._push()
        cap = 4 * self.flat_size + 64
        cells = np.zeros(cap * 8, dtype=np.int32)
        offs = np.zeros(cap + 1, dtype=np.int32)
        names = np.zeros(cap, dtype=np.int32)
        cols = np.zeros(cap, dtype=np.int32)
        n = self.L.tmgo_detect_colour_matches(self.h, _ptr(cells), _ptr(offs), _ptr(names), _ptr(cols), len(cells), cap)
        assert n >= 0
        return self._lists(n, cells, offs), [NAMES[i] for i in names[:n]], [int(c) for c in cols[:n]]

    def resolve_round(self):
        """detect_colour_matches + resolve_colour_matches (board.py:369-373), no gravity/refill."""
        self._push(); n = self.L.tmgo_resolve_round(self.h); self._pull()
        return n

    def get_special_creation_pos(self, coords, taken_pos=(), straight_match=True):
        Cc = self.num_cols
        cells = np.array([r * Cc + c for r, c in coords], dtype=np.int32)
        taken = np.array([r * Cc + c for r, c in taken_pos] or [0], dtype=np.int32)
        pos = self.L.tmgo_special_creation_pos(self.h, _ptr(cells), len(cells), _ptr(taken), len(list(taken_pos)),
                                               int(bool(straight_match)))
        return (pos // Cc, pos % Cc)

    def activate_special(self, coord, tile_type, tile_colour=0, is_combination_match=False):
        self._push()
        self.L.tmgo_activate_special(self.h, coord[0], coord[1], int(tile_type), int(bool(is_combination_match)))
        self._pull()

    def combination_match(self, c1, c2):
        self._push(); self.L.tmgo_combination_match(self.h, c1[0], c1[1], c2[0], c2[1]); self._pull()

    def move(self, c1, c2):
        self._push()
        out = np.zeros(5, dtype=np.int32)
        rc = self.L.tmgo_move(self.h, c1[0], c1[1], c2[0], c2[1], _ptr(out))
        if rc != 0:
            raise ValueError(f"Invalid move: {c1}, {c2}")
        self._pull()
        return int(out[0]), bool(out[1]), int(out[2]), int(out[3]), bool(out[4])

    def effective_mask(self):
        self._push()
        out = np.zeros(self.num_actions, dtype=np.uint8)
        self.L.tmgo_effective_mask(self.h, _ptr(out))
        return out

    def onehot(self):
        self._push()
        planes = self.L.tmgo_onehot_planes(self.h)
        out = np.zeros((planes, self.num_rows, self.num_cols), dtype=np.uint8)
        self.L.tmgo_onehot(self.h, _ptr(out))
        return out


_NP = {"board": np.int8, "timer": np.int32, "draw_cursor": np.uint64, "shuffle_cursor": np.uint64,
       "reward": np.int32, "terminated": np.uint8, "is_combination_match": np.uint8,
       "num_new_specials": np.int32, "num_specials_activated": np.int32, "shuffled": np.uint8,
       "mask": np.uint8, "num_moves_left": np.int32, "status": np.uint32, "episode": np.int32}


class OracleVecEnv:
    """CPU statement of the vectorised env semantics the product implements on the GPU."""

    def __init__(self, num_envs, num_rows, num_cols, num_colours, num_moves, colourless_specials=(),
                 colour_specials=(), seed=1, autoreset="disabled", refill="philox", env_id_offset=0,
                 max_reset_iters=16384, num_threads=1, constructive_reset=False):
        self.L = lib()
        modes = {"disabled": 0, "next_step": 1, "same_step": 2}
        self.cfg = _VecConfig(int(num_envs), int(num_rows), int(num_cols), int(num_colours), int(num_moves),
                              specials_mask(colourless_specials, colour_specials), modes[autoreset],
                              {"philox": 0, "injected": 1}[refill], int(seed), int(env_id_offset),
                              int(max_reset_iters), int(num_threads), 8 if constructive_reset else 0)
        self.h = self.L.tmgo_vec_create(C.byref(self.cfg))
        self.N, self.R, self.Cc, self.K = int(num_envs), int(num_rows), int(num_cols), int(num_colours)
        self.A = 2 * self.R * self.Cc - self.R - self.Cc
        bufs = _VecBuffers()
        self.L.tmgo_vec_get_buffers(self.h, C.byref(bufs))
        shapes = {"board": (self.N, 2, self.R, self.Cc), "mask": (self.N, self.A)}
        for name, _ in _VecBuffers._fields_:
            shape = shapes.get(name, (self.N,))
            n = int(np.prod(shape))
            dt = np.dtype(_NP[name])
            raw = (C.c_char * (n * dt.itemsize)).from_address(getattr(bufs, name))
            setattr(self, name, np.frombuffer(raw, dtype=dt).reshape(shape))
        self._inj = None

    def __del__(self):
        try:
            self.L.tmgo_vec_destroy(self.h)
        except Exception:
            pass

    def set_injected_draws(self, draws):
        self._inj = np.ascontiguousarray(draws, dtype=np.uint8)
        assert self._inj.shape[0] == self.N
        self.L.tmgo_vec_set_injected_draws(self.h, _ptr(self._inj), self._inj.shape[1])

    def reset(self, reset_mask=None, init_boards=None):
        m = None if reset_mask is None else np.ascontiguousarray(reset_mask, dtype=np.uint8)
        b = None if init_boards is None else np.ascontiguousarray(init_boards, dtype=np.int8)
        self.L.tmgo_vec_reset(self.h, None if m is None else _ptr(m), None if b is None else _ptr(b))

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        assert a.shape == (self.N,)
        self.L.tmgo_vec_step(self.h, _ptr(a))

    def rollout(self, steps, action_seed, step0=0):
        return int(self.L.tmgo_vec_rollout(self.h, int(steps), int(action_seed), int(step0)))

    def onehot(self):
        planes = self.K + bin(self.cfg.specials).count("1")
        out = np.zeros((self.N, planes, self.R, self.Cc), dtype=np.uint8)
        self.L.tmgo_vec_onehot(self.h, _ptr(out))
        return out

    def diag(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int64()
        self.L.tmgo_vec_diag(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"max_lines": a.value, "max_dfs_depth": b.value, "max_reset_iters": c.value}
