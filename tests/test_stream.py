"""The draw stream (Philox4x32-10 words -> colours / Fisher-Yates) must be identical in the numpy
statement (oracle/stream.py, used to drive the Python reference), the C oracle and the golden file."""
import ctypes as C
import json
import os

import numpy as np

from conftest import GOLDEN
from oracle import oracle as orc
from oracle.stream import StreamGenerator, philox4x32_10, stream_words


def _kat():
    with open(os.path.join(GOLDEN, "stream_kat.json")) as f:
        return json.load(f)


def test_philox_known_answers_numpy_and_c():
    L = orc.lib()
    for v in _kat()["philox4x32_10"]:
        assert [int(x) for x in philox4x32_10(v["ctr"], v["key"])] == v["out"]
        ctr = (C.c_uint32 * 4)(*v["ctr"]); key = (C.c_uint32 * 2)(*v["key"]); out = (C.c_uint32 * 4)()
        L.tmgo_philox4x32_10(ctr, key, out)
        assert list(out) == v["out"]
    # Random123 kat_vectors, philox4x32 10 rounds
    assert _kat()["philox4x32_10"][0]["out"] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


def test_stream_words_match_golden_and_c():
    L = orc.lib()
    for v in _kat()["stream_words"]:
        w = stream_words(v["seed"], v["env_id"], v["stream"], v["start"], len(v["words"]))
        assert [int(x) for x in w] == v["words"]
        for i, x in enumerate(v["words"]):
            assert L.tmgo_stream_word(v["seed"], v["env_id"], v["stream"], v["start"] + i) == x


def test_generator_contract():
    g = _kat()["generator"]
    gen = StreamGenerator(g["seed"], g["env_id"])
    assert [int(x) for x in gen.integers(1, 5, size=12)] == g["integers_1_5_12"]
    arr = np.arange(10)
    gen.shuffle(arr)
    assert [int(x) for x in arr] == g["shuffle_arange10"]
    # contiguity across call sizes (what makes one flat pre-drawn vector equivalent to many calls)
    a = StreamGenerator(7, 3); b = StreamGenerator(7, 3)
    big = a.integers(1, 7, size=50)
    small = np.concatenate([b.integers(1, 7, size=n) for n in (1, 7, 12, 30)])
    assert np.array_equal(big, small)
    # C oracle refill consumes the same stream: all-empty board refill == integers(1,K+1,P)
    ob = orc.OracleBoard(4, 5, 6, seed=7, env_id=3)
    ob.board[:] = 0
    ob.refill()
    assert np.array_equal(ob.board[0].reshape(-1), big[:20])
    assert ob.cursors == (20, 0)
    # and the same Fisher-Yates
    ob2 = orc.OracleBoard(2, 5, 6, seed=g["seed"], env_id=g["env_id"])
    ob2.board[0] = np.arange(10).reshape(2, 5); ob2.board[1] = 1
    gen2 = StreamGenerator(g["seed"], g["env_id"]); idx = np.arange(10); gen2.shuffle(idx)
    ob2.shuffle()
    assert np.array_equal(ob2.board[0].reshape(-1), idx)
