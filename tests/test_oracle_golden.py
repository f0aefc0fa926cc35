"""Pins the C oracle (oracle/tmg_oracle.c) to the reference:

 (a) every top-level engine/env/wrapper call the reference's own 16 tests make (recorded while they
     pass, tests/golden/ref_test_calls.json.gz) is replayed through the oracle -- PCG64-dependent ones
     via record-and-replay of the draws;
 (b) Philox-stream trajectories of the unmodified reference env (tests/golden/philox_traces.npz);
 (c) live differential fuzz against the imported reference when /root/reference exists.
"""
import gzip
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle as orc

CL_NAMES = {"cookie"}


def _split_specials(names):
    cl = [n for n in names if n in CL_NAMES]
    cs = [n for n in names if n not in CL_NAMES]
    return cl, cs


def _load_calls():
    with gzip.open(os.path.join(GOLDEN, "ref_test_calls.json.gz"), "rt") as f:
        return json.load(f)


def _draws(rng_log):
    ints = [np.asarray(v, dtype=np.int64) for kind, v in rng_log if kind == "integers"]
    has_shuffle = any(kind == "shuffle" for kind, _ in rng_log)
    flat = np.concatenate(ints) if ints else np.zeros(0, dtype=np.int64)
    return flat.astype(np.uint8), has_shuffle


def _board_from(pre, draws=None):
    cl, cs = _split_specials(pre["specials"])
    b = orc.OracleBoard(pre["R"], pre["C"], pre["K"], cl, cs)
    if pre["board"] is not None:
        arr = np.asarray(pre["board"], dtype=np.int32)
        b.board = arr.copy()
        # the reference lets `board=` override the ctor shape (board.py:73-74)
        assert arr.shape == (2, pre["R"], pre["C"])
    b.set_counters(pre["new"], pre["act"])
    if draws is not None:
        b.set_injected(np.concatenate([draws, np.zeros(1, np.uint8)]))
    return b


def _coords(x):
    return [tuple(c) for c in x]


def test_reference_test_calls_replayed_through_oracle():
    recs = _load_calls()
    done, skipped = {}, {}

    def ok(fn):
        done[fn] = done.get(fn, 0) + 1

    def skip(fn):
        skipped[fn] = skipped.get(fn, 0) + 1

    for rec in recs:
        fn = rec["fn"]
        if rec.get("err"):
            skip(fn); continue
        if fn == "is_move_effective":
            arr = np.asarray(rec["board"], dtype=np.int32)
            b = orc.OracleBoard(arr.shape[1], arr.shape[2], 9, board=arr)
            assert b.is_move_effective(tuple(rec["c1"]), tuple(rec["c2"])) == rec["ret"]
            assert rec["unchanged"]
            ok(fn); continue
        if fn == "onehot":
            cfg = rec["cfg"]
            b = orc.OracleBoard(cfg["R"], cfg["C"], cfg["K"], cfg["cl"], cfg["cs"], board=np.asarray(rec["board"], dtype=np.int32))
            assert np.array_equal(b.onehot(), np.asarray(rec["out"]).astype(np.uint8))
            ok(fn); continue
        if fn in ("env.reset", "env.step", "env.mask"):
            cfg = rec["cfg"]
            draws, has_shuffle = _draws(rec.get("rng", []))
            if has_shuffle:
                skip(fn); continue
            v = orc.OracleVecEnv(1, cfg["R"], cfg["C"], cfg["K"], cfg["num_moves"], cfg["cl"], cfg["cs"], refill="injected")
            v.set_injected_draws(np.concatenate([draws, np.zeros(1, np.uint8)])[None, :])
            A = v.A
            if fn == "env.reset":
                v.reset()
                assert int(v.draw_cursor[0]) == len(draws)
            else:
                pre = np.asarray(rec["pre"] if fn == "env.step" else rec["board"]).astype(np.int8)
                v.reset(init_boards=pre[None])
                v.timer[0] = rec["timer"]
            if fn == "env.step":
                v.step(np.array([rec["action"]], np.int32))
                assert int(v.draw_cursor[0]) == len(draws)
                assert int(v.reward[0]) == rec["reward"]
                assert bool(v.terminated[0]) == rec["done"] and rec["trunc"] is False
                info = rec["info"]
                assert bool(v.is_combination_match[0]) == info["is_combination_match"]
                assert int(v.num_new_specials[0]) == info["num_new_specials"]
                assert int(v.num_specials_activated[0]) == info["num_specials_activated"]
                assert bool(v.shuffled[0]) == info["shuffled"]
                mask = info["effective_actions"]
            else:
                mask = rec["mask"]
                if fn == "env.mask" and rec["timer"] == cfg["num_moves"]:
                    mask = None  # terminal rule is applied by step, not by reset-with-board
            if fn != "env.mask":
                assert np.array_equal(v.board[0], np.asarray(rec["board"]).astype(np.int8))
                assert int(v.num_moves_left[0]) == rec["num_moves_left"]
            if mask is not None:
                m = np.zeros(A, np.uint8); m[mask] = 1
                assert np.array_equal(v.mask[0], m), (fn, rec.get("action"))
            assert int(v.status[0]) == 0
            ok(fn); continue

        # ---- Board methods ----
        pre, post = rec["pre"], rec["post"]
        draws, has_shuffle = _draws(rec.get("rng", []))
        if has_shuffle:
            skip(fn); continue
        if fn == "generate_board":
            b = _board_from(dict(post, board=None, new=0, act=0), draws)
            b.generate_board()
            assert np.array_equal(b.board, np.asarray(post["board"])), fn
            assert b.cursors[0] == len(draws)
            ok(fn); continue
        if pre is None:  # the Board had no `.board` yet (e.g. is_move_legal right after the ctor)
            pre = dict(post, board=None)
        b = _board_from(pre, draws)
        args = rec["args"]
        if fn == "get_colour_lines":
            assert b.get_colour_lines() == [_coords(l) for l in rec["ret"]]
        elif fn == "detect_colour_matches":
            coords, names, colours = b.detect_colour_matches()
            assert coords == [_coords(l) for l in rec["ret"][0]]
            assert names == rec["ret"][1] and colours == rec["ret"][2]
        elif fn == "process_colour_lines":
            if b.get_colour_lines() != [_coords(l) for l in args[0]]:
                skip(fn); continue
            coords, names, colours = b.detect_colour_matches()
            assert coords == [_coords(l) for l in rec["ret"][0]]
            assert names == rec["ret"][1] and colours == rec["ret"][2]
        elif fn == "resolve_colour_matches":
            coords, names, colours = b.detect_colour_matches()
            if coords != [_coords(l) for l in args[0]] or names != args[1] or colours != args[2]:
                skip(fn); continue
            b.resolve_round()
            assert np.array_equal(b.board, np.asarray(post["board"]))
            assert b.counters == (post["new"], post["act"])
        elif fn == "get_special_creation_pos":
            kw = rec["kwargs"]
            coords = _coords(args[0])
            taken = _coords(kw.get("taken_pos", args[1] if len(args) > 1 else []))
            straight = kw.get("straight_match", args[2] if len(args) > 2 else True)
            assert list(b.get_special_creation_pos(coords, taken, straight)) == list(rec["ret"])
        elif fn == "activate_special":
            kw = rec["kwargs"]
            is_comb = kw.get("is_combination_match", args[3] if len(args) > 3 else False)
            b.activate_special(tuple(args[0]), args[1], args[2] if len(args) > 2 else 0, is_comb)
            assert np.array_equal(b.board, np.asarray(post["board"]))
            assert b.counters == (post["new"], post["act"])
        elif fn == "combination_match":
            b.combination_match(tuple(args[0]), tuple(args[1]))
            assert np.array_equal(b.board, np.asarray(post["board"]))
            assert b.counters == (post["new"], post["act"])
        elif fn == "gravity":
            b.gravity()
            assert np.array_equal(b.board, np.asarray(post["board"]))
        elif fn == "refill":
            b.refill()
            assert np.array_equal(b.board, np.asarray(post["board"]))
            assert b.cursors[0] == len(draws)
        elif fn == "possible_move":
            assert b.possible_move() == rec["ret"]
            assert np.array_equal(b.board, np.asarray(post["board"]))
        elif fn == "is_move_legal":
            assert b.is_move_legal(tuple(args[0]), tuple(args[1])) == rec["ret"]
        elif fn == "move":
            out = b.move(tuple(args[0]), tuple(args[1]))
            assert list(out) == [rec["ret"][0], bool(rec["ret"][1]), rec["ret"][2], rec["ret"][3], bool(rec["ret"][4])]
            assert np.array_equal(b.board, np.asarray(post["board"]))
            assert b.cursors[0] == len(draws)
        else:
            skip(fn); continue
        assert b.status == 0
        ok(fn)

    # every kind of call the reference tests make is covered, and (almost) nothing was skipped
    for fn in ["generate_board", "get_colour_lines", "detect_colour_matches", "process_colour_lines",
               "resolve_colour_matches", "get_special_creation_pos", "activate_special", "combination_match",
               "gravity", "refill", "possible_move", "move", "is_move_legal", "is_move_effective", "env.reset",
               "env.step", "env.mask", "onehot"]:
        assert done.get(fn, 0) > 0, (fn, done, skipped)
    assert sum(skipped.values()) <= 0.02 * len(recs), skipped
    print("replayed", done, "skipped", skipped)


# ----------------------------------------------------------------------------------------------
def _trace_meta():
    z = np.load(os.path.join(GOLDEN, "philox_traces.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def replay_trace(make_env, z, m):
    """Shared by the oracle test here and the GPU parity test: `make_env(meta)` returns an object with
    the OracleVecEnv attribute protocol (reset/step + numpy-convertible buffers)."""
    name = m["name"]
    v = make_env(m)
    get = lambda a: np.asarray(a)  # noqa: E731
    if m["kind"] == "reset":
        v.reset()
    else:
        v.reset(init_boards=z[f"{name}/init_board"][None])
    assert np.array_equal(get(v.board)[0], z[f"{name}/init_board"]), name
    assert np.array_equal(get(v.mask)[0], z[f"{name}/init_mask"]), name
    assert [int(get(v.draw_cursor)[0]), int(get(v.shuffle_cursor)[0])] == list(z[f"{name}/init_cursors"])
    n_reset = 0
    for t in range(m["steps"]):
        v.step(np.array([z[f"{name}/actions"][t]], np.int32))
        ctx = (name, t)
        assert np.array_equal(get(v.board)[0], z[f"{name}/boards"][t]), ctx
        assert int(get(v.reward)[0]) == int(z[f"{name}/rewards"][t]), ctx
        assert bool(get(v.terminated)[0]) == bool(z[f"{name}/dones"][t]), ctx
        assert bool(get(v.is_combination_match)[0]) == bool(z[f"{name}/comb"][t]), ctx
        assert int(get(v.num_new_specials)[0]) == int(z[f"{name}/new"][t]), ctx
        assert int(get(v.num_specials_activated)[0]) == int(z[f"{name}/act"][t]), ctx
        assert bool(get(v.shuffled)[0]) == bool(z[f"{name}/shuf"][t]), ctx
        assert np.array_equal(get(v.mask)[0], z[f"{name}/masks"][t]), ctx
        assert int(get(v.draw_cursor)[0]) == int(z[f"{name}/dc"][t]), ctx
        assert int(get(v.shuffle_cursor)[0]) == int(z[f"{name}/sc"][t]), ctx
        if z[f"{name}/dones"][t]:
            if m["kind"] == "reset":
                v.reset()
            else:
                v.reset(init_boards=z[f"{name}/reset_boards"][n_reset][None])
            assert np.array_equal(get(v.board)[0], z[f"{name}/reset_boards"][n_reset]), ctx
            assert np.array_equal(get(v.mask)[0], z[f"{name}/reset_masks"][n_reset]), ctx
            assert [int(get(v.draw_cursor)[0]), int(get(v.shuffle_cursor)[0])] == list(z[f"{name}/reset_cursors"][n_reset])
            n_reset += 1
    assert int(get(v.status)[0]) == 0, name


def test_philox_traces_through_oracle():
    z, meta = _trace_meta()
    assert len(meta) >= 28
    for m in meta:
        replay_trace(lambda m: orc.OracleVecEnv(1, m["R"], m["C"], m["K"], m["num_moves"], m["cl"], m["cs"],
                                                seed=m["seed"], env_id_offset=m["env_id"]), z, m)
    assert sum(m["n_shuffled"] for m in meta) >= 2   # the shuffle path is exercised
    assert sum(m["n_comb"] for m in meta) >= 50      # combination matches are exercised


# ----------------------------------------------------------------------------------------------
@pytest.mark.reference
@pytest.mark.parametrize("cfg", [
    (10, 10, 4, (), (), 30, "uniform"),
    (10, 10, 4, ("cookie",), ("vertical_laser", "horizontal_laser", "bomb"), 30, "mask"),
    (9, 9, 6, ("cookie",), ("vertical_laser", "horizontal_laser", "bomb"), 30, "mask"),
    (4, 4, 3, ("cookie",), ("vertical_laser", "horizontal_laser", "bomb"), 10, "mask"),
    (5, 4, 2, ("cookie",), ("vertical_laser", "bomb"), 10, "mask"),
])
def test_live_differential_fuzz_against_reference(cfg):
    from oracle import ref_loader
    from oracle.stream import StreamGenerator

    ref = ref_loader.load_reference()
    R, C, K, cl, cs, nm, policy = cfg
    for env_id in range(2):
        env = ref.TileMatchEnv(R, C, K, nm, list(cl), list(cs), seed=11)
        env.board.np_random = StreamGenerator(11, env_id)
        o = orc.OracleVecEnv(1, R, C, K, nm, cl, cs, seed=11, env_id_offset=env_id)
        rng = np.random.default_rng(env_id)

        def ref_reset():   # generate_board draws come from the episode-indexed reset streams (oracle/stream.py)
            env.board.np_random.begin_reset()
            out = env.reset()
            env.board.np_random.end_reset()
            return out

        _, info = ref_reset(); o.reset()
        for t in range(150):
            assert np.array_equal(env.board.board.astype(np.int8), o.board[0])
            m = np.zeros(o.A, np.uint8); m[info["effective_actions"]] = 1
            assert np.array_equal(m, o.mask[0])
            eff = np.flatnonzero(m)
            a = int(rng.integers(o.A)) if (policy == "uniform" or len(eff) == 0) else int(rng.choice(eff))
            _, r, done, _, info = env.step(a)
            o.step(np.array([a], np.int32))
            assert (int(r), bool(done)) == (int(o.reward[0]), bool(o.terminated[0]))
            assert int(info["num_specials_activated"]) == int(o.num_specials_activated[0])
            assert env.board.np_random.draw_cursor == int(o.draw_cursor[0])
            if done:
                _, info = ref_reset(); o.reset()
