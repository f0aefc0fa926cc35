"""TEST INFRASTRUCTURE ONLY -- generates the committed golden fixtures under tests/golden/.

Runs in the BUILD CONTAINER only (needs /root/reference).  Nothing here is read at run time
by the product; the GPU box only sees the fixtures this script wrote.

    python -m oracle.gen_golden            # regenerate everything

Fixtures
  tests/golden/ref_test_calls.json.gz   every top-level engine / env / wrapper call made by the
                                        reference's own 16 tests while they pass (inputs, outputs,
                                        PCG64 draws consumed) -- the reference's known-answer vectors
                                        turned into data, SURVEY.md section 4 / 8c.
  tests/golden/philox_traces.npz        step-by-step trajectories of the UNMODIFIED reference
                                        TileMatchEnv driven by the project's Philox stream
                                        (oracle/stream.py) for the BASELINE configs, tiny boards that
                                        hit the shuffle path, special subsets and dense-special boards.
  tests/golden/stream_kat.json          Philox4x32-10 (Random123) known answers + stream words.
"""
from __future__ import annotations

import gzip
import json
import os
import sys

import numpy as np

from oracle import ref_loader
from oracle.stream import StreamGenerator, philox4x32_10, stream_words

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

ALL_CL = ("cookie",)
ALL_CS = ("vertical_laser", "horizontal_laser", "bomb")


# ======================================================================================
# A. record the reference's own tests
# ======================================================================================
def _jsonable(x):
    if isinstance(x, np.ndarray):
        return x.astype(np.int64).tolist()
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.bool_,)):
        return bool(x)
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    if isinstance(x, set):
        return sorted(_jsonable(v) for v in x)
    if isinstance(x, dict):
        return {str(k): _jsonable(v) for k, v in x.items()}
    return x


class _Recorder:
    CAP = {"generate_board": 60, "is_move_effective": 400, "possible_move": 200}

    def __init__(self):
        self.records = []
        self.depth = 0
        self.counts = {}

    def full(self, fn):
        return self.counts.get(fn, 0) >= self.CAP.get(fn, 100000)


def record_reference_tests():
    import pytest

    ref = ref_loader.load_reference()
    bm = ref.board_module
    Board = ref.Board
    rec = _Recorder()

    def wrap_rng(board):
        g = board.np_random
        if not isinstance(g, ref_loader.RecordingGenerator) and hasattr(g, "integers"):
            board.np_random = ref_loader.RecordingGenerator(g)
        return board.np_random

    def board_state(b):
        return {
            "R": int(b.num_rows), "C": int(b.num_cols), "K": int(b.num_colours),
            "specials": sorted(b.specials),
            "board": _jsonable(np.asarray(b.board)) if getattr(b, "board", None) is not None else None,
            "new": int(getattr(b, "num_new_specials", 0)), "act": int(getattr(b, "num_specials_activated", 0)),
        }

    def wrap_method(name):
        orig = getattr(Board, name)

        def wrapper(self, *a, **k):
            if rec.depth > 0 or rec.full(name):
                rec.depth += 1
                try:
                    return orig(self, *a, **k)
                finally:
                    rec.depth -= 1
            g = wrap_rng(self)
            log0 = len(g.log) if isinstance(g, ref_loader.RecordingGenerator) else 0
            pre = board_state(self) if hasattr(self, "board") else None
            rec.depth += 1
            try:
                out = orig(self, *a, **k)
                err = None
            except Exception as e:  # noqa: BLE001 - the reference's tests expect some raises
                out, err = None, type(e).__name__
                raise
            finally:
                rec.depth -= 1
                g2 = self.np_random
                calls = g2.log[log0:] if isinstance(g2, ref_loader.RecordingGenerator) else []
                rec.records.append({
                    "fn": name, "args": _jsonable(a), "kwargs": _jsonable(k), "pre": pre,
                    "post": board_state(self), "ret": _jsonable(out), "err": err,
                    "rng": [[kind, _jsonable(v)] for kind, v in calls],
                })
                rec.counts[name] = rec.counts.get(name, 0) + 1
            return out

        setattr(Board, name, wrapper)

    for name in ["generate_board", "shuffle", "get_colour_lines", "detect_colour_matches", "process_colour_lines",
                 "resolve_colour_matches", "get_special_creation_pos", "activate_special", "combination_match",
                 "gravity", "refill", "possible_move", "move", "is_move_legal"]:
        wrap_method(name)

    orig_ime = bm.is_move_effective

    def ime(board, c1, c2):
        if rec.depth > 0 or rec.full("is_move_effective"):
            return orig_ime(board, c1, c2)
        pre = np.array(board).copy()
        out = orig_ime(board, c1, c2)
        rec.records.append({"fn": "is_move_effective", "board": _jsonable(pre), "c1": _jsonable(c1), "c2": _jsonable(c2),
                            "ret": bool(out), "unchanged": bool(np.array_equal(pre, board))})
        rec.counts["is_move_effective"] = rec.counts.get("is_move_effective", 0) + 1
        return out

    bm.is_move_effective = ime
    env_mod = sys.modules["tile_match_gym.tile_match_env"]
    env_mod.is_move_effective = ime

    Env = ref.TileMatchEnv

    def env_cfg(e):
        return {"R": e.num_rows, "C": e.num_cols, "K": e.num_colours, "num_moves": e.num_moves,
                "cl": list(e.colourless_specials), "cs": list(e.colour_specials)}

    orig_reset, orig_step, orig_gea = Env.reset, Env.step, Env._get_effective_actions

    def reset(self, seed=None, options=None):
        wrap_rng(self.board)
        rec.depth += 1
        try:
            if seed is not None:
                self.set_seed(seed)
                wrap_rng(self.board)
                seed = None
            g = self.board.np_random
            log0 = len(g.log)
            obs, info = orig_reset(self, seed=seed, options=options)
        finally:
            rec.depth -= 1
        rec.records.append({"fn": "env.reset", "cfg": env_cfg(self), "board": _jsonable(obs["board"]),
                            "num_moves_left": int(obs["num_moves_left"]), "mask": _jsonable(info["effective_actions"]),
                            "rng": [[k, _jsonable(v)] for k, v in g.log[log0:]]})
        return obs, info

    def step(self, action):
        g = wrap_rng(self.board)
        log0 = len(g.log)
        pre = np.array(self.board.board).copy()
        timer = self.timer
        rec.depth += 1
        try:
            obs, r, done, trunc, info = orig_step(self, action)
        finally:
            rec.depth -= 1
        rec.records.append({"fn": "env.step", "cfg": env_cfg(self), "pre": _jsonable(pre), "timer": timer,
                            "action": int(action), "board": _jsonable(obs["board"]), "reward": int(r), "done": bool(done),
                            "trunc": bool(trunc), "num_moves_left": int(obs["num_moves_left"]),
                            "info": _jsonable(info), "rng": [[k, _jsonable(v)] for k, v in g.log[log0:]]})
        return obs, r, done, trunc, info

    def gea(self):
        top = rec.depth == 0
        rec.depth += 1
        try:
            out = orig_gea(self)
        finally:
            rec.depth -= 1
        if top:
            rec.records.append({"fn": "env.mask", "cfg": env_cfg(self), "board": _jsonable(self.board.board),
                                "timer": self.timer, "mask": _jsonable(out)})
        return out

    Env.reset, Env.step, Env._get_effective_actions = reset, step, gea

    OH = ref.OneHotWrapper
    orig_obs = OH.observation

    def observation(self, obs):
        out = orig_obs(self, obs)
        u = self.unwrapped
        rec.records.append({"fn": "onehot", "cfg": env_cfg(u), "board": _jsonable(obs["board"]),
                            "out": _jsonable(np.asarray(out["board"]))})
        return out

    OH.observation = observation

    sys.path.insert(0, ref_loader.REF_ROOT)
    rc = pytest.main(["-q", "-p", "no:cacheprovider", "--rootdir=/tmp", ref_loader.REF_ROOT + "/tests"])
    assert rc == 0, "the reference's own tests must pass while being recorded"
    return rec.records


# ======================================================================================
# B. Philox-stream trajectories of the unmodified reference env
# ======================================================================================
def _dense_special_board(rng, R, C, K, p_special=0.35, p_cookie=0.06):
    colour = rng.integers(1, K + 1, size=(R, C))
    typ = np.ones((R, C), dtype=np.int64)
    u = rng.random((R, C))
    typ[u < p_special] = rng.integers(2, 5, size=(R, C))[u < p_special]
    ck = rng.random((R, C)) < p_cookie
    typ[ck] = -1
    colour[ck] = 0
    return np.stack([colour, typ]).astype(np.int32)


def run_trace(ref, R, C, K, cl, cs, num_moves, seed, env_id, steps, policy, init_board=None, dense=False):
    env = ref.TileMatchEnv(R, C, K, num_moves, list(cl), list(cs), seed=seed)
    gen = StreamGenerator(seed, env_id)
    env.board.np_random = gen
    rng = np.random.default_rng(seed * 7919 + env_id)
    A = env.num_actions
    def ref_reset():
        gen.begin_reset()
        env.reset()
        gen.end_reset()

    if init_board is None and not dense:
        ref_reset()
    else:
        env.board.board = (np.asarray(init_board, dtype=np.int32).copy() if init_board is not None
                           else _dense_special_board(rng, R, C, K))
        env.timer = 0
    out = {k: [] for k in ["boards", "actions", "rewards", "dones", "comb", "new", "act", "shuf", "masks", "dc", "sc", "resets"]}
    out["init_board"] = env.board.board.astype(np.int8).copy()
    out["init_dc"], out["init_sc"] = gen.draw_cursor, gen.shuffle_cursor
    mask = np.zeros(A, np.uint8)
    mask[env._get_effective_actions()] = 1
    out["init_mask"] = mask.copy()
    for _ in range(steps):
        eff = np.flatnonzero(mask)
        if policy == "uniform" or len(eff) == 0:
            a = int(rng.integers(A))
        else:
            a = int(rng.choice(eff))
        obs, r, done, _, info = env.step(a)
        mask = np.zeros(A, np.uint8)
        mask[info["effective_actions"]] = 1
        out["boards"].append(env.board.board.astype(np.int8).copy())
        out["actions"].append(a); out["rewards"].append(int(r)); out["dones"].append(bool(done))
        out["comb"].append(bool(info["is_combination_match"])); out["new"].append(int(info["num_new_specials"]))
        out["act"].append(int(info["num_specials_activated"])); out["shuf"].append(bool(info["shuffled"]))
        out["masks"].append(mask.copy()); out["dc"].append(gen.draw_cursor); out["sc"].append(gen.shuffle_cursor)
        if done:  # caller-side reset, as in src/examples/random_agent.py:12-31
            if dense:
                env.board.board = _dense_special_board(rng, R, C, K)
                env.timer = 0
            elif init_board is not None:
                env.board.board = np.asarray(init_board, dtype=np.int32).copy()
                env.timer = 0
            else:
                ref_reset()
            mask = np.zeros(A, np.uint8)
            mask[env._get_effective_actions()] = 1
            out["resets"].append({"board": env.board.board.astype(np.int8).copy(), "mask": mask.copy(),
                                  "dc": gen.draw_cursor, "sc": gen.shuffle_cursor})
    return out


def constructive_no_line_board(rng, R, C, K):
    """Full board without any 3-line (for shapes where generate_board does not terminate)."""
    b = np.zeros((R, C), dtype=np.int64)
    for r in range(R):
        for c in range(C):
            while True:
                k = int(rng.integers(1, K + 1))
                if c >= 2 and b[r, c - 1] == k and b[r, c - 2] == k:
                    continue
                if r >= 2 and b[r - 1, c] == k and b[r - 2, c] == k:
                    continue
                b[r, c] = k
                break
    return np.stack([b, np.ones_like(b)]).astype(np.int32)


TRACE_MATRIX = [
    # name, R, C, K, cl, cs, num_moves, steps, policy, kind
    ("c1_10x10k4_none_uniform", 10, 10, 4, (), (), 30, 90, "uniform", "reset"),
    ("c2_10x10k4_all_uniform", 10, 10, 4, ALL_CL, ALL_CS, 30, 90, "uniform", "reset"),
    ("c2_10x10k4_all_mask", 10, 10, 4, ALL_CL, ALL_CS, 30, 120, "mask", "reset"),
    ("c3_9x9k6_all_mask", 9, 9, 6, ALL_CL, ALL_CS, 30, 120, "mask", "reset"),
    ("c5_32x32k7_all_mask", 32, 32, 7, ALL_CL, ALL_CS, 20, 40, "mask", "inject"),
    ("t_3x5k3_all_mask", 3, 5, 3, ALL_CL, ALL_CS, 10, 200, "mask", "reset"),
    ("t_4x4k3_all_mask", 4, 4, 3, ALL_CL, ALL_CS, 10, 200, "mask", "reset"),
    ("t_5x5k4_all_mask", 5, 5, 4, ALL_CL, ALL_CS, 20, 200, "mask", "reset"),
    ("t_6x7k3_all_mask", 6, 7, 3, ALL_CL, ALL_CS, 20, 120, "mask", "reset"),
    ("d_8x8k4_dense", 8, 8, 4, ALL_CL, ALL_CS, 6, 120, "mask", "dense"),
    ("d_5x6k3_dense", 5, 6, 3, ALL_CL, ALL_CS, 6, 120, "mask", "dense"),
    ("d_12x9k5_dense", 12, 9, 5, ALL_CL, ALL_CS, 6, 60, "mask", "dense"),
]
# all 16 subsets of enabled specials (branches at board.py:287,297,299,304)
_names = ["cookie", "vertical_laser", "horizontal_laser", "bomb"]
for _m in range(16):
    _cl = tuple(n for i, n in enumerate(_names[:1]) if _m >> i & 1)
    _cs = tuple(n for i, n in enumerate(_names[1:], start=1) if _m >> i & 1)
    TRACE_MATRIX.append((f"s{_m:02d}_7x7k3_mask", 7, 7, 3, _cl, _cs, 15, 60, "mask", "reset"))


def philox_traces():
    ref = ref_loader.load_reference()
    arrays, meta = {}, []
    for i, (name, R, C, K, cl, cs, nm, steps, policy, kind) in enumerate(TRACE_MATRIX):
        seed, env_id = 2, 100 + i
        init = None
        if kind == "inject":
            init = constructive_no_line_board(np.random.default_rng(1234 + i), R, C, K)
        tr = run_trace(ref, R, C, K, cl, cs, nm, seed, env_id, steps, policy, init_board=init, dense=(kind == "dense"))
        for k in ["boards", "masks"]:
            arrays[f"{name}/{k}"] = np.stack(tr[k])
        for k in ["actions", "rewards", "new", "act"]:
            arrays[f"{name}/{k}"] = np.array(tr[k], dtype=np.int32)
        for k in ["dones", "comb", "shuf"]:
            arrays[f"{name}/{k}"] = np.array(tr[k], dtype=np.uint8)
        for k in ["dc", "sc"]:
            arrays[f"{name}/{k}"] = np.array(tr[k], dtype=np.int64)
        arrays[f"{name}/init_board"] = tr["init_board"]
        arrays[f"{name}/init_mask"] = tr["init_mask"]
        arrays[f"{name}/init_cursors"] = np.array([tr["init_dc"], tr["init_sc"]], dtype=np.int64)
        if tr["resets"]:
            arrays[f"{name}/reset_boards"] = np.stack([x["board"] for x in tr["resets"]])
            arrays[f"{name}/reset_masks"] = np.stack([x["mask"] for x in tr["resets"]])
            arrays[f"{name}/reset_cursors"] = np.array([[x["dc"], x["sc"]] for x in tr["resets"]], dtype=np.int64)
        meta.append({"name": name, "R": R, "C": C, "K": K, "cl": list(cl), "cs": list(cs), "num_moves": nm,
                     "steps": steps, "policy": policy, "kind": kind, "seed": seed, "env_id": env_id,
                     "n_shuffled": int(np.sum(tr["shuf"])), "n_comb": int(np.sum(tr["comb"])),
                     "n_act": int(np.sum(tr["act"]))})
        print("trace", meta[-1])
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return arrays


def stream_kat():
    kat = []
    for ctr, key in [([0, 0, 0, 0], [0, 0]), ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2),
                     ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])]:
        kat.append({"ctr": ctr, "key": key, "out": [int(x) for x in philox4x32_10(ctr, key)]})
    # Random123 kat_vectors for philox4x32-10
    assert kat[0]["out"] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert kat[1]["out"] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert kat[2]["out"] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    words = []
    for seed, env, stream, start, n in [(2, 0, 0, 0, 16), (2, 5, 0, 3, 9), (0xDEADBEEFCAFE, 4000000000, 1, (1 << 34) + 1, 8),
                                        (1, 65535, 2, 0, 8)]:
        words.append({"seed": seed, "env_id": env, "stream": stream, "start": start,
                      "words": [int(x) for x in stream_words(seed, env, stream, start, n)]})
    g = StreamGenerator(2, 5)
    ints = [int(x) for x in g.integers(1, 5, size=12)]
    arr = np.arange(10)
    g.shuffle(arr)
    return {"philox4x32_10": kat, "stream_words": words,
            "generator": {"seed": 2, "env_id": 5, "integers_1_5_12": ints, "shuffle_arange10": [int(x) for x in arr]}}


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    with open(os.path.join(GOLDEN, "stream_kat.json"), "w") as f:
        json.dump(stream_kat(), f, indent=1)
    np.savez_compressed(os.path.join(GOLDEN, "philox_traces.npz"), **philox_traces())
    recs = record_reference_tests()
    with gzip.open(os.path.join(GOLDEN, "ref_test_calls.json.gz"), "wt", compresslevel=9) as f:
        json.dump(recs, f, separators=(",", ":"))
    counts = {}
    for r in recs:
        counts[r["fn"]] = counts.get(r["fn"], 0) + 1
    print("recorded reference-test calls:", counts)


if __name__ == "__main__":
    main()
