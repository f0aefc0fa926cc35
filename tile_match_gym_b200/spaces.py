"""Observation/action spaces with the reference's bounds (tile_match_env.py:52-77, wrappers.py:25-30).
gymnasium's classes are used when gymnasium is importable; otherwise minimal duck-typed stand-ins with the
same attributes (`n`, `shape`, `dtype`, `low`, `high`, `spaces`) so that agents can size their networks."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gymnasium is not in the build image
    from gymnasium.spaces import Box, Dict, Discrete  # type: ignore
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Discrete:  # type: ignore[no-redef]
        def __init__(self, n, seed=None, start=0):
            self.n, self.start = int(n), int(start)
            self.shape, self.dtype = (), np.dtype(np.int64)
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return int(self.start + self._rng.integers(self.n))

        def contains(self, x):
            return self.start <= int(x) < self.start + self.n

        def __repr__(self):
            return f"Discrete({self.n})"

    class Box:  # type: ignore[no-redef]
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(np.asarray(low).shape if shape is None else shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape)
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box{self.shape}"

    class Dict:  # type: ignore[no-redef]
        def __init__(self, spaces=None, seed=None, **kw):
            self.spaces = dict(spaces or {})
            self.spaces.update(kw)

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def __repr__(self):
            return f"Dict({self.spaces})"


def board_space(num_rows, num_cols, num_colours, n_colourless, n_colour, seed=None):
    """Box(2,R,C) int32 with the reference's (loose) bounds, tile_match_env.py:52-65."""
    low = np.array([np.zeros((num_rows, num_cols), dtype=np.int32),
                    np.full((num_rows, num_cols), -n_colourless, dtype=np.int32)])
    high = np.array([np.full((num_rows, num_cols), num_colours, dtype=np.int32),
                     np.full((num_rows, num_cols), n_colour + 2, dtype=np.int32)])
    return Box(low=low, high=high, shape=(2, num_rows, num_cols), dtype=np.int32, seed=seed)


def onehot_board_space(num_rows, num_cols, planes):
    """wrappers.py:25"""
    return Box(low=0, high=1, dtype=np.int32, shape=(planes, num_rows, num_cols))
