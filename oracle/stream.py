"""TEST INFRASTRUCTURE ONLY -- numpy statement of the project's counter-based draw stream.

The reference consumes randomness only through `np_random.integers(1, K+1, size=n)`
(board.py:97,129,239) and `np_random.shuffle(arr)` (board.py:116), and takes any
duck-typed generator (board.py:49,63).  The B200 build replaces numpy's PCG64 with a
counter-based stream so that every env can compute its k-th draw independently:

    word(seed, env_id, stream, k) = Philox4x32-10(key = (seed_lo, seed_hi),
                                                   ctr = (k>>2 lo, k>>2 hi, env_id, stream))[k & 3]
    integers(low, high, n)  -> low + mulhi32(word, high-low) for the next n words of stream 0
    shuffle(arr)            -> Fisher-Yates, i = n-1..1, j = mulhi32(next word of stream 1, i+1)

While a board is being generated (TileMatchEnv.reset -> Board.generate_board, board.py:95-112) the draws come
from episode-indexed streams instead, each starting at word 0:

    reset word(seed, env_id, episode j, k)   = Philox(key, ctr = (k>>2, j, env_id, 3))[k & 3]   (integers)
    reset shuffle word(...)                  = Philox(key, ctr = (k>>2, j, env_id, 4))[k & 3]   (shuffle)

so the j-th board of an env is a pure function of (seed, env_id, j): the GPU can generate it ahead of time, off
the step path, and the result is the same as generating it at reset time.

`StreamGenerator` hands exactly this stream to the unmodified reference, which is how
"identical refill draws" (BASELINE north_star) is realised for differential tests.
Philox4x32-10 is the published Random123 algorithm (Salmon et al., SC'11); the
known-answer vectors checked in tests/test_stream.py are Random123's.
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_REFILL = 0
STREAM_SHUFFLE = 1
STREAM_ACTIONS = 2  # used by bench/test drivers to draw synthetic actions
STREAM_RESET = 3
STREAM_RESET_SHUFFLE = 4


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32 array-like, key: (..., 2) uint32 array-like -> (..., 4) uint32."""
    ctr = np.asarray(ctr, dtype=np.uint64)
    key = np.asarray(key, dtype=np.uint64)
    c0, c1, c2, c3 = (ctr[..., i].copy() for i in range(4))
    k0, k1 = key[..., 0].copy(), key[..., 1].copy()
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK32, lo1, (hi0 ^ c3 ^ k1) & MASK32, lo0
        k0 = (k0 + np.uint64(PHILOX_W0)) & MASK32
        k1 = (k1 + np.uint64(PHILOX_W1)) & MASK32
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def stream_words(seed: int, env_id: int, stream: int, start: int, n: int, episode=None) -> np.ndarray:
    """Words start..start+n-1 of one env's stream, as uint32.  `episode` selects an episode-indexed reset stream
    (the episode number replaces the high half of the block counter)."""
    if n <= 0:
        return np.zeros(0, dtype=np.uint32)
    k = np.arange(start, start + n, dtype=np.uint64)
    blk = k >> np.uint64(2)
    ublk, inv = np.unique(blk, return_inverse=True)
    ctr = np.zeros((len(ublk), 4), dtype=np.uint64)
    ctr[:, 0] = ublk & MASK32
    ctr[:, 1] = (ublk >> np.uint64(32)) if episode is None else np.uint64(episode & 0xFFFFFFFF)
    ctr[:, 2] = np.uint64(env_id & 0xFFFFFFFF)
    ctr[:, 3] = np.uint64(stream & 0xFFFFFFFF)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64)
    out = philox4x32_10(ctr, np.broadcast_to(key, (len(ublk), 2)))
    return out[inv, (k & np.uint64(3)).astype(np.int64)]


def mulhi32(w: np.ndarray, n: int) -> np.ndarray:
    return ((w.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


class StreamGenerator:
    """Drop-in for `Board.np_random` driven by the project's stream.

    mode "philox": words come from Philox as above.
    mode "injected": `integers` pops pre-drawn colours from `draws` (a 1-D array of values in
    [low, high)); `shuffle` still uses the Philox shuffle stream (stream 1).
    """

    def __init__(self, seed: int, env_id: int, draws=None):
        self.seed = int(seed)
        self.env_id = int(env_id)
        self.draw_cursor = 0
        self.shuffle_cursor = 0
        self.draws = None if draws is None else np.asarray(draws).reshape(-1)
        self.exhausted = False
        self.episode = -1          # number of the board being / last generated
        self.in_reset = False
        self._rdc = self._rsc = 0  # cursors of the current episode's reset streams

    def begin_reset(self):
        """Call right before the reference's generate_board (env.reset); injected mode has one sequential stream."""
        self.episode += 1
        self.in_reset = self.draws is None
        self._rdc = self._rsc = 0

    def end_reset(self):
        self.in_reset = False

    def integers(self, low, high=None, size=None):
        if high is None:
            low, high = 0, low
        n = 1 if size is None else int(np.prod(size))
        if self.in_reset:
            w = stream_words(self.seed, self.env_id, STREAM_RESET, self._rdc, n, episode=self.episode)
            self._rdc += n
            out = int(low) + mulhi32(w, int(high) - int(low))
            return int(out[0]) if size is None else out.reshape(size)
        if self.draws is None:
            w = stream_words(self.seed, self.env_id, STREAM_REFILL, self.draw_cursor, n)
            out = int(low) + mulhi32(w, int(high) - int(low))
        else:
            if self.draw_cursor + n > len(self.draws):
                self.exhausted = True
                raise IndexError("injected draw stream exhausted")
            out = self.draws[self.draw_cursor:self.draw_cursor + n].astype(np.int64)
        self.draw_cursor += n
        if size is None:
            return int(out[0])
        return out.reshape(size)

    def shuffle(self, arr):
        n = len(arr)
        if n <= 1:
            return
        if self.in_reset:
            w = stream_words(self.seed, self.env_id, STREAM_RESET_SHUFFLE, self._rsc, n - 1, episode=self.episode)
            self._rsc += n - 1
        else:
            w = stream_words(self.seed, self.env_id, STREAM_SHUFFLE, self.shuffle_cursor, n - 1)
            self.shuffle_cursor += n - 1
        t = 0
        for i in range(n - 1, 0, -1):
            j = int((int(w[t]) * (i + 1)) >> 32)
            t += 1
            arr[i], arr[j] = arr[j], arr[i]
