"""TEST INFRASTRUCTURE ONLY -- loader for the *Python reference* (tile-match-gym).

Only `tests/`, `oracle/gen_golden.py` and ad-hoc validation scripts may import this.
The product package (`tile_match_gym_b200`) never does.

The reference lives at /root/reference (this container only -- it does NOT exist on the
GPU box).  `import tile_match_gym` needs `gymnasium` and `pygame`, neither of which is in
the image, so minimal stand-ins are registered in `sys.modules` before the import
(reference call sites: tile_match_env.py:1,4,38,51,60-77; wrappers.py:1-5; renderer.py:5;
__init__.py:1-3).  The stand-ins contribute no arithmetic: the whole transition lives in
board.py, which needs numpy + numba only.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

def _find_reference_root() -> str:
    """TMG_REFERENCE_ROOT, then /root/reference (the build container), then a copy a driver may have left under
    <repo>/baseline/_ref (git-ignored; the GPU box has none unless one was put there)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cands = [os.environ.get("TMG_REFERENCE_ROOT"), "/root/reference"]
    base = os.path.join(here, "baseline", "_ref")
    if os.path.isdir(base):
        for dirpath, dirnames, filenames in os.walk(base):
            if os.path.basename(dirpath) == "tile_match_gym" and "board.py" in filenames:
                src = os.path.dirname(dirpath)
                cands.append(os.path.dirname(src) if os.path.basename(src) == "src" else None)
                break
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "src", "tile_match_gym", "board.py")):
            return c
    return "/root/reference"


REF_ROOT = _find_reference_root()
REF_SRC = os.path.join(REF_ROOT, "src")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "tile_match_gym", "board.py"))


# --------------------------------------------------------------------------------------
# gymnasium / pygame stand-ins
# --------------------------------------------------------------------------------------
def _install_stubs() -> None:
    try:  # a real gymnasium wins if it ever appears in the image
        import gymnasium  # noqa: F401
        have_gym = True
    except Exception:
        have_gym = False

    if not have_gym:
        gym = types.ModuleType("gymnasium")
        spaces = types.ModuleType("gymnasium.spaces")
        envs = types.ModuleType("gymnasium.envs")
        registration = types.ModuleType("gymnasium.envs.registration")

        class Env:
            metadata: dict = {}

            @property
            def np_random(self):
                if getattr(self, "_np_random", None) is None:
                    self._np_random = np.random.default_rng()
                return self._np_random

            @np_random.setter
            def np_random(self, value):
                self._np_random = value

            @property
            def unwrapped(self):
                return self

            def close(self):
                pass

        class _Wrapper:
            def __init__(self, env):
                self.env = env

            def __getattr__(self, name):
                if name.startswith("_") and name not in ("_moves_left_observation_space",):
                    raise AttributeError(name)
                return getattr(self.env, name)

            @property
            def unwrapped(self):
                return self.env.unwrapped

        class ObservationWrapper(_Wrapper):
            def reset(self, **kw):
                obs, info = self.env.reset(**kw)
                return self.observation(obs), info

            def step(self, action):
                obs, r, term, trunc, info = self.env.step(action)
                return self.observation(obs), r, term, trunc, info

        class RewardWrapper(_Wrapper):
            def reset(self, **kw):
                return self.env.reset(**kw)

            def step(self, action):
                obs, r, term, trunc, info = self.env.step(action)
                return obs, self.reward(r), term, trunc, info

        class Discrete:
            def __init__(self, n, seed=None, start=0):
                self.n = int(n)
                self.start = int(start)
                self._rng = np.random.default_rng(seed)
                self.shape = ()
                self.dtype = np.int64

            def sample(self):
                return int(self.start + self._rng.integers(self.n))

            def contains(self, x):
                return self.start <= int(x) < self.start + self.n

        class Box:
            def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
                self.dtype = np.dtype(dtype)
                if shape is None:
                    shape = np.asarray(low).shape
                self.shape = tuple(shape)
                self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape)
                self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape)
                self._rng = np.random.default_rng(seed)

        class Dict:
            def __init__(self, spaces_dict=None, seed=None, **kw):
                self.spaces = dict(spaces_dict or {})
                self.spaces.update(kw)

            def __getitem__(self, k):
                return self.spaces[k]

            def keys(self):
                return self.spaces.keys()

        def register(*a, **k):
            return None

        spaces.Discrete, spaces.Box, spaces.Dict = Discrete, Box, Dict
        gym.Env, gym.ObservationWrapper, gym.RewardWrapper = Env, ObservationWrapper, RewardWrapper
        gym.spaces = spaces
        gym.envs = envs
        envs.registration = registration
        registration.register = register
        gym.__tmg_stub__ = True
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces
        sys.modules["gymnasium.envs"] = envs
        sys.modules["gymnasium.envs.registration"] = registration

    try:
        import pygame  # noqa: F401
    except Exception:
        pg = types.ModuleType("pygame")
        pg.__tmg_stub__ = True
        sys.modules["pygame"] = pg


_cache: dict = {}


def load_reference():
    """Returns a namespace with Board, is_move_effective, swap_coords, TileMatchEnv,
    OneHotWrapper, ProportionRewardWrapper from the unmodified reference."""
    if "ns" in _cache:
        return _cache["ns"]
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT} (it only exists in the build container)")
    _install_stubs()
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    board_mod = importlib.import_module("tile_match_gym.board")
    env_mod = importlib.import_module("tile_match_gym.tile_match_env")
    wrap_mod = importlib.import_module("tile_match_gym.wrappers")
    ns = types.SimpleNamespace(
        board_module=board_mod,
        Board=board_mod.Board,
        is_move_effective=board_mod.is_move_effective,
        swap_coords=board_mod.swap_coords,
        TileMatchEnv=env_mod.TileMatchEnv,
        OneHotWrapper=wrap_mod.OneHotWrapper,
        ProportionRewardWrapper=wrap_mod.ProportionRewardWrapper,
    )
    _cache["ns"] = ns
    return ns


# --------------------------------------------------------------------------------------
# Generators that can be handed to the reference as `np_random` (board.py:49,63 accepts
# anything with .integers/.shuffle -- reference call sites board.py:97,116,129,239).
# --------------------------------------------------------------------------------------
class RecordingGenerator:
    """Wraps a real numpy Generator and logs every call's output (record & replay of the
    reference's PCG64-pinned tests)."""

    def __init__(self, gen):
        self._gen = gen
        self.log = []  # list of ("integers", ndarray) / ("shuffle", permutation ndarray)

    def integers(self, low, high=None, size=None, **kw):
        out = self._gen.integers(low, high, size, **kw)
        self.log.append(("integers", np.array(out, dtype=np.int64).reshape(-1).copy()))
        return out

    def shuffle(self, arr):
        self._gen.shuffle(arr)
        self.log.append(("shuffle", np.array(arr, dtype=np.int64).copy()))

    def __getattr__(self, name):
        return getattr(self._gen, name)
