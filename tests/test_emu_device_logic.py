"""Runs the product's DEVICE code (tile_match_gym_b200/csrc/tmg_device.cuh) on the CPU through the test-only
lane emulator (tests/emu) against the golden traces and the oracle.  This is how kernel logic is debugged in a
container without a GPU; the GPU parity tests (test_gpu_parity.py) are the real gate and do not use it."""
import os
import numpy as np
import pytest

from oracle import oracle as orc
from test_oracle_golden import _trace_meta, replay_trace
from conftest import expected_policy_actions
from emu.emu import EmuVecEnv

ALL_CL = ("cookie",)
ALL_CS = ("vertical_laser", "horizontal_laser", "bomb")


def test_golden_traces_through_emulated_device_code():
    z, meta = _trace_meta()
    for m in meta:
        replay_trace(lambda m: EmuVecEnv(1, m["R"], m["C"], m["K"], m["num_moves"], m["cl"], m["cs"], seed=m["seed"],
                                         env_id_offset=m["env_id"]), z, m)


@pytest.mark.parametrize("cfg", [
    (10, 10, 4, ALL_CL, ALL_CS, 12, "same_step", 24, 40),
    (9, 9, 6, ALL_CL, ALL_CS, 8, "next_step", 24, 30),
    (4, 5, 3, ALL_CL, ALL_CS, 5, "same_step", 48, 40),
    (6, 20, 5, ALL_CL, ALL_CS, 6, "same_step", 8, 20),
    (10, 10, 4, ALL_CL, ALL_CS, 5, "same_step", 16, 12, 2),   # TMG_FLAG_NO_PREGEN: boards generated inside the step
    (3, 5, 3, ALL_CL, ALL_CS, 3, "same_step", 64, 30),        # tiny boards: shuffles inside generate_board
])
def test_emulated_batch_vs_oracle(cfg):
    R, C, K, cl, cs, moves, autoreset, N, steps = cfg[:9]
    flags = cfg[9] if len(cfg) > 9 else 0
    e = EmuVecEnv(N, R, C, K, moves, cl, cs, seed=3, autoreset=autoreset, env_id_offset=50, flags=flags)
    o = orc.OracleVecEnv(N, R, C, K, moves, cl, cs, seed=3, autoreset=autoreset, env_id_offset=50, num_threads=4)
    e.reset(); o.reset()
    rng = np.random.default_rng(7)
    fields = ["board", "timer", "draw_cursor", "shuffle_cursor", "reward", "terminated", "is_combination_match",
              "num_new_specials", "num_specials_activated", "shuffled", "mask", "num_moves_left", "status", "episode"]
    for t in range(steps):
        m = o.mask.astype(np.float64) + 1e-9
        u = rng.random((N, 1)) * m.sum(axis=1, keepdims=True)
        a = (np.cumsum(m, axis=1) < u).sum(axis=1).clip(0, o.A - 1).astype(np.int32)
        e.step(a); o.step(a)
        for f in fields:
            assert np.array_equal(getattr(e, f), getattr(o, f)), (t, f)


@pytest.mark.parametrize("R,C,K,moves,autoreset,policy", [(10, 10, 4, 5, "same_step", "mask"), (9, 9, 6, 4, "next_step", "mask"),
                                                         (10, 10, 4, 7, "same_step", "uniform"), (5, 6, 3, 4, "disabled", "mask")])
def test_emulated_policy_rollout(R, C, K, moves, autoreset, policy):
    """tmg_rollout_policy: the agent inside the kernel takes exactly the actions its stream contract says, and the
    trajectory is the oracle's under those actions."""
    N, T, seed, off = 16, 11, 8, 21
    e = EmuVecEnv(N, R, C, K, moves, ALL_CL, ALL_CS, seed=seed, autoreset=autoreset, env_id_offset=off)
    o = orc.OracleVecEnv(N, R, C, K, moves, ALL_CL, ALL_CS, seed=seed, autoreset=autoreset, env_id_offset=off, num_threads=4)
    e.reset(); o.reset()
    for window in range(2):
        act, rew, term = e.rollout(T, policy)
        for t in range(T):
            want = expected_policy_actions(o, seed, off, moves, policy)
            assert np.array_equal(act[t], want), (window, t)
            o.step(want)
            assert np.array_equal(rew[t], o.reward) and np.array_equal(term[t], o.terminated), (window, t)
        for f in ("board", "timer", "mask", "episode", "draw_cursor", "status", "reward", "num_moves_left"):
            assert np.array_equal(getattr(e, f), getattr(o, f)), (window, f)


@pytest.mark.parametrize("R,C,K,cap", [(10, 10, 4, 6), (9, 9, 6, 2), (6, 7, 3, 4)])
def test_emulated_reset_cap_is_flagged_identically(R, C, K, cap):
    """generate_board stopped by max_reset_iters (TMG_ST_RESET_CAP): same boards, cursors and status bits as the
    oracle, on the packed-row generator (fixed shapes) and the byte-plane one."""
    N, moves = 32, 3
    e = EmuVecEnv(N, R, C, K, moves, ALL_CL, ALL_CS, seed=6, autoreset="same_step", max_reset_iters=cap)
    o = orc.OracleVecEnv(N, R, C, K, moves, ALL_CL, ALL_CS, seed=6, autoreset="same_step", max_reset_iters=cap, num_threads=4)
    e.reset(); o.reset()
    assert (o.status & 8).any()                      # the cap is actually hit
    rng = np.random.default_rng(2)
    for t in range(2 * moves + 1):
        for f in ("board", "mask", "status", "draw_cursor", "shuffle_cursor", "episode"):
            assert np.array_equal(getattr(e, f), getattr(o, f)), (t, f)
        a = rng.integers(0, o.A, N).astype(np.int32)
        e.step(a); o.step(a)


@pytest.mark.parametrize("R,C,K,moves,autoreset,T", [
    (10, 10, 4, 5, "same_step", 13), (9, 9, 6, 4, "next_step", 11), (4, 5, 3, 3, "same_step", 10), (6, 7, 4, 6, "disabled", 9),
    (10, 10, 4, 30, "same_step", 6)])
def test_emulated_step_many_equals_single_steps(R, C, K, moves, autoreset, T):
    """tmg_step_many (k_rollout): T steps in one launch leave the state of T tmg_step calls and return every step's
    reward / termination; two windows in a row so that pool hand-over between launches is covered."""
    N = 24
    e = EmuVecEnv(N, R, C, K, moves, ALL_CL, ALL_CS, seed=4, autoreset=autoreset, env_id_offset=11)
    o = orc.OracleVecEnv(N, R, C, K, moves, ALL_CL, ALL_CS, seed=4, autoreset=autoreset, env_id_offset=11, num_threads=4)
    e.reset(); o.reset()
    e.host_bind()
    rng = np.random.default_rng(12)
    fields = ["board", "timer", "draw_cursor", "shuffle_cursor", "reward", "terminated", "is_combination_match",
              "num_new_specials", "num_specials_activated", "shuffled", "mask", "num_moves_left", "status", "episode"]
    for window in range(2):
        acts = rng.integers(0, o.A, (T, N)).astype(np.int32)
        if window == 1:
            acts[2, 5] = o.A + 3          # bad action: flagged, step does nothing (tile_match_env.py:97)
        rew, term = e.step_many(acts)
        for t in range(T):
            o.step(acts[t])
            assert np.array_equal(rew[t], o.reward) and np.array_equal(term[t], o.terminated), (window, t)
        for f in fields:
            assert np.array_equal(getattr(e, f), getattr(o, f)), (window, f)
        assert np.array_equal(e.h_board, e.board) and np.array_equal(e.h_mask, e.mask)
        assert np.array_equal(e.h_board_packed, ((e.board[:, 0] & 15) | ((e.board[:, 1] & 7) << 4)).astype(np.uint8))
        assert np.array_equal(e.h_reward, e.reward) and np.array_equal(e.h_terminated, e.terminated)
        assert np.array_equal(e.h_moves_left, e.num_moves_left)
        e.step(acts[0]); o.step(acts[0])      # single steps and rollouts interleave
        for f in fields:
            assert np.array_equal(getattr(e, f), getattr(o, f)), (window, "after single step", f)


@pytest.mark.parametrize("R,C,K,moves,autoreset", [(10, 10, 4, 6, "same_step"), (5, 7, 3, 4, "disabled"), (9, 9, 6, 5, "next_step")])
def test_emulated_host_mirror_tracks_device_state(R, C, K, moves, autoreset):
    """tmg_host_bind's write-through: after every step the mirror arrays equal board / mask / packed mask."""
    N = 24
    e = EmuVecEnv(N, R, C, K, moves, ALL_CL, ALL_CS, seed=5, autoreset=autoreset)
    e.reset()
    e.host_bind()
    rng = np.random.default_rng(3)
    for t in range(3 * moves):
        if autoreset == "disabled" and t and t % moves == 0:
            e.reset(); e.host_bind()
        e.step(rng.integers(0, e.A, N).astype(np.int32))
        assert np.array_equal(e.h_board, e.board), t
        assert np.array_equal(e.h_board_packed, ((e.board[:, 0] & 15) | ((e.board[:, 1] & 7) << 4)).astype(np.uint8)), t
        assert np.array_equal(e.h_mask, e.mask), t
        assert np.array_equal(e.h_mask_bits, np.packbits(e.mask, axis=1, bitorder="little")), t
        assert np.array_equal(e.h_reward, e.reward) and np.array_equal(e.h_terminated, e.terminated), t
        assert np.array_equal(e.h_moves_left, e.num_moves_left), t


# ----------------------------------------------------------------------------------------------
# primitive-level differential fuzz: one engine primitive on random boards (dense specials, few colours so that
# crossings / long lines / chains are common), device code (emulated) vs oracle, for several special subsets
OPS = dict(GRAVITY=1, REFILL=2, RESOLVE=3, ACTIVATE=4, COMBINE=5, MOVE=6, EFFECTIVE=7, GENERATE=8, SHUFFLE=9, COUNT=10)
SUBSETS = [(("cookie",), ("vertical_laser", "horizontal_laser", "bomb")), ((), ("bomb",)), (("cookie",), ("horizontal_laser",)),
           ((), ("vertical_laser",)), ((), ())]


def _rand_board(rng, R, C, K, ps, pc, pe=0.0):
    col = rng.integers(1, K + 1, (R, C)); typ = np.ones((R, C), int)
    u = rng.random((R, C)); typ[u < ps] = rng.integers(2, 5, (R, C))[u < ps]
    ck = rng.random((R, C)) < pc; typ[ck] = -1; col[ck] = 0
    em = rng.random((R, C)) < pe; typ[em] = 0; col[em] = 0
    return np.stack([col, typ]).astype(np.int32)


@pytest.mark.parametrize("which,R,C,K,iters", [
    ("RESOLVE", 7, 7, 3, 60), ("RESOLVE", 6, 12, 2, 60), ("RESOLVE", 10, 10, 4, 40), ("MOVE", 10, 10, 4, 40),
    ("MOVE", 5, 4, 2, 60), ("COMBINE", 7, 7, 3, 60), ("COMBINE", 10, 10, 4, 40), ("ACTIVATE", 9, 9, 6, 40),
    ("GRAVITY", 7, 9, 3, 30), ("REFILL", 7, 9, 3, 30), ("MASK", 10, 10, 4, 30), ("COUNT", 12, 6, 2, 40),
])
def test_emulated_primitives_vs_oracle(which, R, C, K, iters):
    rng = np.random.default_rng(hash((which, R, C)) % 2**32)
    for si, (cl, cs) in enumerate(SUBSETS):
        e = EmuVecEnv(1, R, C, K, 1000, cl, cs, seed=11, env_id_offset=7)
        for it in range(iters // len(SUBSETS) + 1):
            b = _rand_board(rng, R, C, K, 0.15, 0.05, pe=(0.2 if which in ("GRAVITY", "REFILL") else 0.0))
            o = orc.OracleBoard(R, C, K, cl, cs, seed=11, env_id=7, board=b)
            e.reset(init_boards=b.astype(np.int8)[None])
            e.status[:] = 0
            dc = int(rng.integers(0, 1000)); e.draw_cursor[0] = dc; o.set_stream(11, 7, dc, 0)
            e.num_new_specials[0] = 0; e.num_specials_activated[0] = 0; o.set_counters(0, 0)
            args = np.zeros((1, 4), np.int32)
            counters = lambda: (int(e.num_new_specials[0]), int(e.num_specials_activated[0]))  # noqa: E731
            if which == "COUNT":
                e.debug_op(OPS["COUNT"], args)
                assert len(o.get_colour_lines()) == int(e.reward[0])
                continue
            if which == "MASK":
                e.legal_mask()
                assert np.array_equal(o.effective_mask(), e.mask[0])
                continue
            if which == "RESOLVE":
                o.resolve_round(); e.debug_op(OPS["RESOLVE"], args)
            elif which == "ACTIVATE":
                sp = np.argwhere((b[1] != 0) & (b[1] != 1))
                if len(sp) == 0:
                    continue
                r, c = sp[rng.integers(len(sp))]; comb = int(rng.integers(2))
                o.activate_special((r, c), b[1, r, c], b[0, r, c], bool(comb)); args[0] = [r, c, b[1, r, c], comb]
                e.debug_op(OPS["ACTIVATE"], args)
            elif which == "COMBINE":
                cands = [a for a in range(o.num_actions)
                         if (lambda t1, t2: (t1 not in (0, 1) and t2 not in (0, 1)) or t1 < 0 or t2 < 0)(
                             b[1][o.action_to_coords[a][0]], b[1][o.action_to_coords[a][1]])]
                if not cands:
                    continue
                (r1, c1), (r2, c2) = o.action_to_coords[cands[rng.integers(len(cands))]]
                o.combination_match((r1, c1), (r2, c2)); args[0] = [r1, c1, r2, c2]; e.debug_op(OPS["COMBINE"], args)
            elif which == "GRAVITY":
                o.gravity(); e.debug_op(OPS["GRAVITY"], args)
            elif which == "REFILL":
                o.refill(); e.debug_op(OPS["REFILL"], args)
                assert o.cursors[0] == int(e.draw_cursor[0])
            elif which == "MOVE":
                (r1, c1), (r2, c2) = o.action_to_coords[int(rng.integers(o.num_actions))]
                want = o.move((r1, c1), (r2, c2)); args[0] = [r1, c1, r2, c2]; e.debug_op(OPS["MOVE"], args)
                got = (int(e.reward[0]), bool(e.is_combination_match[0]), *counters(), bool(e.shuffled[0]))
                assert want == got, (want, got)
                assert o.cursors == (int(e.draw_cursor[0]), int(e.shuffle_cursor[0]))
            assert np.array_equal(o.board.astype(np.int8), e.board[0]), (which, si, it)
            if which in ("RESOLVE", "ACTIVATE", "COMBINE"):
                assert o.counters == counters(), (which, si, it)
            assert int(e.status[0]) == 0


def _decode_lines(words):
    n = int(words[0])
    ent = []
    for i in range(n):
        info, cells = int(words[1 + 2 * i]), int(words[2 + 2 * i])
        kind, idx = (info >> 16) & 1, (info >> 17) & 31
        bits = [b for b in range(32) if (cells >> b) & 1]
        ent.append((info & 0xfff, sorted([(b, idx) for b in bits] if kind else [(idx, b) for b in bits])))
    return [c for _, c in sorted(ent, key=lambda e: e[0])]


def test_emulated_line_tables_list_for_list(monkeypatch):
    """get_colour_lines (board.py:149-215) as the engines' line table, in the reference's list order: recorded calls of the
    reference's tests (tests/board/test_match_detection.py:15-224 among them) through both engines of the emulated device
    code (register-resident bit planes and byte planes; one board per warp)."""
    import gzip
    import json
    from conftest import GOLDEN
    from test_oracle_golden import _split_specials
    monkeypatch.setenv("TMG_EMU_LANES32", "1")
    with gzip.open(os.path.join(GOLDEN, "ref_test_calls.json.gz"), "rt") as f:
        recs = [r for r in json.load(f) if r["fn"] == "get_colour_lines" and not r.get("err") and r["pre"] and r["pre"]["board"] is not None
                and r["pre"]["C"] >= 2]
    n_nonempty = 0
    for rec in [r for i, r in enumerate(recs) if r["ret"] or i % 8 == 0]:
        pre = rec["pre"]
        cl, cs = _split_specials(pre["specials"])
        e = EmuVecEnv(1, pre["R"], pre["C"], max(pre["K"], 1), 10, cl, cs)
        e.reset(init_boards=np.asarray(pre["board"], dtype=np.int8)[None])
        want = [sorted(tuple(c) for c in line) for line in rec["ret"]]
        for byte_planes in (False, True):
            assert _decode_lines(e.debug_lines(byte_planes)[0]) == want, (pre["R"], pre["C"], byte_planes)
        n_nonempty += bool(want)
    assert n_nonempty > 30, n_nonempty


@pytest.mark.parametrize("N,R,Cc,K,moves", [(6, 10, 10, 4, 4), (3, 32, 32, 7, 3), (8, 5, 5, 3, 3), (4, 16, 12, 3, 3)])
def test_emulated_constructive_reset_matches_oracle(N, R, Cc, K, moves):
    """TMG_FLAG_CONSTRUCTIVE_RESET (NOT reference behaviour): the constructive line-free sampler of the device
    code and its restatement in the oracle produce the same boards -- at reset, at every autoreset and through the pool --
    and the boards are line-free with a possible move, also for 32x32 / 7 colours where generate_board never returns."""
    cl, cs = ["cookie"], ["vertical_laser", "horizontal_laser", "bomb"]
    e = EmuVecEnv(N, R, Cc, K, moves, cl, cs, seed=9, autoreset="same_step", flags=8)
    o = orc.OracleVecEnv(N, R, Cc, K, moves, cl, cs, seed=9, autoreset="same_step", num_threads=1, constructive_reset=True)
    e.reset(); o.reset()
    assert np.array_equal(e.board, o.board) and np.array_equal(e.mask, o.mask)
    b = o.board[:, 0]
    assert (o.board[:, 1] == 1).all() and b.min() >= 1 and b.max() <= K and o.mask.any(axis=1).all()
    assert not ((b[:, :, :-2] == b[:, :, 1:-1]) & (b[:, :, 1:-1] == b[:, :, 2:])).any()
    assert not ((b[:, :-2] == b[:, 1:-1]) & (b[:, 1:-1] == b[:, 2:])).any()
    rng = np.random.default_rng(0)
    for t in range(3 * moves + 1):
        a = rng.integers(0, e.A, size=N).astype(np.int32)
        e.step(a); o.step(a)
        assert np.array_equal(e.board, o.board) and np.array_equal(e.reward, o.reward) and np.array_equal(e.mask, o.mask), t
        assert np.array_equal(e.draw_cursor, o.draw_cursor)
    assert (e.status == 0).all() and (o.status == 0).all() and int(o.episode.max()) == 3


@pytest.mark.parametrize("R,C,K", [(10, 10, 7), (10, 10, 8), (9, 9, 4), (9, 9, 8), (10, 10, 5)])
def test_emulated_packed_row_generator_shapes(R, C, K):
    """Board::generate_packed (fixed shapes, colour window of the reset stream): both cell widths (K <= 4: 2-bit stream and
    funnel-shifted rows, K <= 8: byte window) at both row lengths, through reset, the pool and the in-step fallback."""
    N, moves = 24, 3
    for flags in (0, 2):                                     # 2 = TMG_FLAG_NO_PREGEN
        e = EmuVecEnv(N, R, C, K, moves, ALL_CL, ALL_CS, seed=11, autoreset="same_step", env_id_offset=7, flags=flags)
        o = orc.OracleVecEnv(N, R, C, K, moves, ALL_CL, ALL_CS, seed=11, autoreset="same_step", env_id_offset=7, num_threads=4)
        e.reset(); o.reset()
        rng = np.random.default_rng(5)
        for t in range(2 * moves + 1):
            for f in ("board", "mask", "status", "episode", "draw_cursor", "shuffle_cursor"):
                assert np.array_equal(getattr(e, f), getattr(o, f)), (flags, t, f)
            a = rng.integers(0, o.A, N).astype(np.int32)
            e.step(a); o.step(a)
