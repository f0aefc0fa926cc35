"""GPU parity tests, part 2 (run with -m gpu on the B200 box): the gates SURVEY.md section 7 names that part 1 does not reach --
the ordered line table (G3 "LINES"), the remaining recorded call kinds of the reference's own tests replayed on device
(G4), all 16 subsets of enabled specials as batches, BASELINE configs 3 and 5 at full size, a 1e8-env-step run against
the oracle (G2), and the equivalence of the two engines of the step kernels (register-resident bit planes / byte planes).
Everything goes through the C ABI and is bit-exact: array equality on integer state."""
import ctypes as C
import gzip
import itertools
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle as orc
from test_gpu_parity import ALL_CL, ALL_CS, FIELDS, GpuAdapter, assert_same, make_gpu
from test_oracle_golden import _draws, _split_specials

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def _recs():
    with gzip.open(os.path.join(GOLDEN, "ref_test_calls.json.gz"), "rt") as f:
        return json.load(f)


def decode_lines(words, C_):
    """tmg_debug_lines output of one env -> the reference's list of lines (each a sorted list of (r, c)), in list order."""
    n = int(words[0])
    ent = []
    for i in range(n):
        info, cells = int(words[1 + 2 * i]), int(words[2 + 2 * i])
        kind, idx = (info >> 16) & 1, (info >> 17) & 31
        bits = [b for b in range(32) if (cells >> b) & 1]
        coords = [(b, idx) for b in bits] if kind else [(idx, b) for b in bits]
        ent.append((info & 0xfff, sorted(coords)))
    return [c for _, c in sorted(ent, key=lambda e: e[0])]


def lines_of(env, byte_planes):
    torch = _torch()
    from tile_match_gym_b200 import _native as nat
    out = torch.zeros((env.num_envs, nat.LINES_WORDS), dtype=torch.int32, device="cuda")
    nat.check(env._lib.tmg_debug_lines(env._h, C.c_void_p(out.data_ptr()), int(byte_planes), None), "tmg_debug_lines")
    torch.cuda.synchronize()
    return out.cpu().numpy().astype(np.int64) & 0xffffffff


def test_get_colour_lines_table_list_for_list():
    """Every get_colour_lines call the reference's tests make (tests/board/test_match_detection.py:15-224 and the calls
    inside the other tests), replayed on device: the same lines, in the reference's list order, from both engines."""
    torch = _torch()
    groups = {}
    for rec in _recs():
        if rec["fn"] != "get_colour_lines" or rec.get("err") or rec["pre"] is None or rec["pre"]["board"] is None:
            continue
        pre = rec["pre"]
        if pre["C"] < 2:
            continue
        key = (pre["R"], pre["C"], max(pre["K"], 1), pre["specials"] if isinstance(pre["specials"], str) else json.dumps(pre["specials"]))
        groups.setdefault(key, []).append(rec)
    total = 0
    for (R, Cc, K, _), recs in groups.items():
        cl, cs = _split_specials(recs[0]["pre"]["specials"])
        boards = np.stack([np.asarray(r["pre"]["board"], dtype=np.int8) for r in recs])
        env = make_gpu(len(recs), R, Cc, K, 10, cl, cs, autoreset="disabled")
        env.reset(options={"init_boards": torch.from_numpy(boards)})
        for byte_planes in (False, True):
            out = lines_of(env, byte_planes)
            for i, rec in enumerate(recs):
                want = [sorted(tuple(c) for c in line) for line in rec["ret"]]
                got = decode_lines(out[i], Cc)
                assert got == want, (R, Cc, K, byte_planes, i, got, want)
        total += len(recs)
        env.close()
    assert total >= 1500, total
    print("get_colour_lines calls compared list for list on device:", total)


def test_recorded_env_and_wrapper_calls_on_device():
    """The recorded env.reset / env.step / env.mask / onehot / generate_board / possible_move calls of the reference's tests
    (tests/test_env.py:5-120, tests/test_wrappers.py:5-41, tests/board/test_generate_board.py, test_possible_move.py),
    replayed through the C ABI with the recorded PCG64 draws injected."""
    torch = _torch()
    done = {}

    def injected_env(cfg_R, cfg_C, cfg_K, moves, cl, cs, draws):
        env = make_gpu(1, cfg_R, cfg_C, cfg_K, moves, cl, cs, autoreset="disabled", refill="injected")
        d = np.concatenate([draws, np.ones(4, np.uint8)])[None]
        env.set_injected_draws(torch.from_numpy(d).cuda())
        return env

    for rec in _recs():
        fn = rec["fn"]
        if rec.get("err"):
            continue
        if fn == "onehot":
            cfg = rec["cfg"]
            env = make_gpu(1, cfg["R"], cfg["C"], cfg["K"], 10, cfg["cl"], cfg["cs"], autoreset="disabled")
            env.reset(options={"init_boards": torch.as_tensor(np.asarray(rec["board"], dtype=np.int8))[None]})
            want = np.asarray(rec["out"])
            assert np.array_equal(env.onehot().cpu().numpy()[0], want.astype(np.uint8))
            assert np.array_equal(env.onehot(torch.float64).cpu().numpy()[0], want.astype(np.float64))   # the reference's dtype
            done[fn] = done.get(fn, 0) + 1
            continue
        if fn in ("env.reset", "env.step", "env.mask"):
            cfg = rec["cfg"]
            draws, has_shuffle = _draws(rec.get("rng", []))
            if has_shuffle:
                continue
            env = injected_env(cfg["R"], cfg["C"], cfg["K"], cfg["num_moves"], cfg["cl"], cfg["cs"], draws)
            A = env.num_actions
            if fn == "env.reset":
                env.reset()
                assert int(env.draw_cursor[0].item()) == len(draws)
            else:
                pre = np.asarray(rec["pre"] if fn == "env.step" else rec["board"]).astype(np.int8)
                env.reset(options={"init_boards": torch.from_numpy(pre)[None]})
                env.timer[0] = rec["timer"]
            if fn == "env.step":
                _, rew, term, trunc, info = env.step(torch.tensor([rec["action"]], dtype=torch.int32))
                assert int(env.draw_cursor[0].item()) == len(draws)
                assert int(rew[0].item()) == rec["reward"] and bool(term[0].item()) == rec["done"] and not bool(trunc[0].item())
                ri = rec["info"]
                assert bool(info["is_combination_match"][0].item()) == ri["is_combination_match"]
                assert int(info["num_new_specials"][0].item()) == ri["num_new_specials"]
                assert int(info["num_specials_activated"][0].item()) == ri["num_specials_activated"]
                assert bool(info["shuffled"][0].item()) == ri["shuffled"]
                mask = ri["effective_actions"]
            else:
                mask = rec["mask"]
                if fn == "env.mask" and rec["timer"] == cfg["num_moves"]:
                    mask = None
            if fn != "env.mask":
                assert np.array_equal(env.board[0].cpu().numpy(), np.asarray(rec["board"]).astype(np.int8))
                assert int(env.num_moves_left[0].item()) == rec["num_moves_left"]
            if mask is not None:
                m = np.zeros(A, np.uint8); m[mask] = 1
                assert np.array_equal(env.mask[0].cpu().numpy().astype(np.uint8), m), (fn, rec.get("action"))
            assert int(env.status[0].item()) == 0
            done[fn] = done.get(fn, 0) + 1
            continue
        if fn not in ("generate_board", "possible_move"):
            continue
        if fn == "generate_board" and done.get(fn, 0) >= 120:
            continue
        pre, post = rec["pre"], rec["post"]
        draws, has_shuffle = _draws(rec.get("rng", []))
        if has_shuffle:
            continue
        src = post if fn == "generate_board" else pre
        if src is None or src["C"] < 2:
            continue
        cl, cs = _split_specials(src["specials"])
        env = injected_env(src["R"], src["C"], max(src["K"], 1), 10, cl, cs, draws)
        if fn == "generate_board":
            env.reset()                                               # generate_board from the injected draws
            assert np.array_equal(env.board[0].cpu().numpy(), np.asarray(post["board"], dtype=np.int8))
            assert int(env.draw_cursor[0].item()) == len(draws)
        else:
            if pre["board"] is None:
                continue
            env.reset(options={"init_boards": torch.as_tensor(np.asarray(pre["board"], dtype=np.int8))[None]})
            assert bool(env.mask[0].any().item()) == bool(rec["ret"])  # possible_move == any effective action
            assert np.array_equal(env.board[0].cpu().numpy(), np.asarray(post["board"], dtype=np.int8))
        done[fn] = done.get(fn, 0) + 1
    for fn in ("env.reset", "env.step", "env.mask", "onehot", "generate_board", "possible_move"):
        assert done.get(fn, 0) > 0, (fn, done)
    print("device replays of recorded calls:", done)


SPECIALS = ["cookie", "vertical_laser", "horizontal_laser", "bomb"]
SUBSETS16 = [tuple(s for s, on in zip(SPECIALS, bits) if on) for bits in itertools.product((0, 1), repeat=4)]


@pytest.mark.parametrize("subset", SUBSETS16, ids=lambda s: "+".join(s) or "none")
def test_all_sixteen_special_subsets_batched(subset):
    """process_colour_lines branches on the enabled specials (board.py:287,297,299,304): every one of the 16 subsets as a
    batch of 10x10 / 4-colour envs, actions sampled from the mask (every step a move), every buffer every step."""
    cl = tuple(s for s in subset if s == "cookie")
    cs = tuple(s for s in subset if s != "cookie")
    N, R, Cc, K, moves, steps = 768, 10, 10, 4, 14, 32
    g = GpuAdapter(make_gpu(N, R, Cc, K, moves, cl, cs, seed=21, autoreset="same_step", env_id_offset=500))
    o = orc.OracleVecEnv(N, R, Cc, K, moves, cl, cs, seed=21, autoreset="same_step", env_id_offset=500, num_threads=os.cpu_count() or 8)
    g.reset(); o.reset()
    assert_same(g, o, "reset")
    rng = np.random.default_rng(len(subset) * 7 + 1)
    for t in range(steps):
        m = o.mask.astype(np.float64) + 1e-9
        u = rng.random((N, 1)) * m.sum(axis=1, keepdims=True)
        a = (np.cumsum(m, axis=1) < u).sum(axis=1).clip(0, o.A - 1).astype(np.int32)
        g.step(a); o.step(a)
        assert_same(g, o, f"{subset} step {t}")


@pytest.mark.parametrize("R,Cc,K,policy", [(10, 10, 4, "mask"), (9, 9, 6, "uniform"), (5, 5, 4, "mask"), (3, 5, 3, "mask"), (10, 24, 7, "mask")])
def test_register_engine_equals_byte_plane_engine(R, Cc, K, policy):
    """The step kernels have two engines for boards of up to 10 rows and 7 colours (tmg_rb.cuh / the byte planes of
    tmg_device.cuh, TMG_FLAG_BYTE_PLANES): same seeds, same actions -> identical buffers, step by step, and through
    tmg_step_many."""
    torch = _torch()
    N, moves = 1536, 9
    a_env = make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=31, autoreset="same_step")
    b_env = make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=31, autoreset="same_step", byte_planes=True)
    ga, gb = GpuAdapter(a_env), GpuAdapter(b_env)
    ga.reset(); gb.reset()
    gen = torch.Generator(device="cuda"); gen.manual_seed(5)
    for t in range(40):
        if policy == "mask":
            a = torch.multinomial(a_env.mask.float() + 1e-6, 1, generator=gen)[:, 0].to(torch.int32)
        else:
            a = torch.randint(0, a_env.num_actions, (N,), device="cuda", dtype=torch.int32, generator=gen)
        a_env.step(a); b_env.step(a)
        assert_same(ga, gb, f"step {t}")
    acts = torch.randint(0, a_env.num_actions, (12, N), device="cuda", dtype=torch.int32, generator=gen)
    ra, ta = a_env.step_many(acts)
    rb, tb = b_env.step_many(acts)
    assert torch.equal(ra, rb) and torch.equal(ta, tb)
    assert_same(ga, gb, "after step_many")
    assert int((a_env.status != 0).sum().item()) == 0


def test_rollout_windows_on_a_tiny_batch_race_free():
    """Regression for the pool-refill race (two refills of one env in flight, requests overwritten in the ring): many short
    tmg_step_many / tmg_rollout_policy windows on a tiny batch, mixed with single steps, against an env that generates
    every board inside the call (TMG_FLAG_NO_PREGEN).  Requests name their board and one in-order side stream serves
    them, so the two must agree bit for bit however the launches interleave."""
    torch = _torch()
    N, R, Cc, K, moves = 96, 10, 10, 4, 3
    a_env = make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=8, autoreset="same_step")
    b_env = make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=8, autoreset="same_step", pregenerate=False)
    ga, gb = GpuAdapter(a_env), GpuAdapter(b_env)
    ga.reset(); gb.reset()
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    for it in range(300):
        T = 1 + it % 3
        acts = torch.randint(0, a_env.num_actions, (T, N), device="cuda", dtype=torch.int32, generator=gen)
        if it % 7 == 3:
            for t in range(T):
                a_env.step(acts[t]); b_env.step(acts[t])
        elif it % 11 == 5:
            xa = a_env.rollout(T, "mask"); xb = b_env.rollout(T, "mask")
            assert all(torch.equal(p, q) for p, q in zip(xa, xb)), it
        else:
            ra, ta = a_env.step_many(acts); rb, tb = b_env.step_many(acts)
            assert torch.equal(ra, rb) and torch.equal(ta, tb), it
        if it % 10 == 9:
            assert_same(ga, gb, f"window {it}")
    assert_same(ga, gb, "end")
    assert int(a_env.episode.max().item()) > 150


def test_config3_full_size_onehot_and_mask():
    """BASELINE configs[2] at the size one GPU holds of it (131 072 envs = 1M / 8): 9x9, 6 colours, all specials,
    num_moves=30; rewards and terminations every step, every buffer plus the one-hot observation at checkpoints."""
    torch = _torch()
    N, R, Cc, K, moves = 131072, 9, 9, 6, 30
    env = make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=2, autoreset="same_step", obs="onehot")
    g = GpuAdapter(env)
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=2, autoreset="same_step", num_threads=os.cpu_count() or 8)
    g.reset(); o.reset()
    assert_same(g, o, "reset")
    rng = np.random.default_rng(3)
    for t in range(36):
        a = rng.integers(0, o.A, size=N).astype(np.int32)
        obs, rew, term, _, _ = env.step(torch.from_numpy(a).cuda()); o.step(a)
        assert np.array_equal(rew.cpu().numpy(), o.reward) and np.array_equal(term.cpu().numpy().astype(np.uint8), o.terminated), t
        if t in (0, 17, 29, 30, 35):
            assert_same(g, o, f"step {t}")
            assert np.array_equal(obs["board"].cpu().numpy(), o.onehot()), t
    assert int(o.status.sum()) == 0


def test_config5_full_size_injected_boards():
    """BASELINE configs[4] at bench size: 8 192 envs of 32x32 / 7 colours from injected line-free boards (generate_board
    does not terminate for this shape in the reference), actions sampled from the mask."""
    import bench
    torch = _torch()
    N, R, Cc, K, moves = 8192, 32, 32, 7, 12
    boards = bench.no_line_boards(N, R, Cc, K, 9).astype(np.int8)
    g = GpuAdapter(make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=12, autoreset="disabled"))
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=12, autoreset="disabled", num_threads=os.cpu_count() or 8)
    g.reset(init_boards=boards); o.reset(init_boards=boards)
    assert_same(g, o, "inject")
    rng = np.random.default_rng(5)
    for t in range(moves):
        if t % 2:
            a = rng.integers(0, o.A, size=N).astype(np.int32)
        else:
            a = np.array([rng.choice(np.flatnonzero(o.mask[e])) if o.mask[e].any() else 0 for e in range(N)], dtype=np.int32)
        g.step(a); o.step(a)
        assert np.array_equal(g.reward, o.reward), t
        if t in (0, 5, moves - 1):
            assert_same(g, o, f"step {t}")
    assert int(o.status.sum()) == 0


def test_hundred_million_env_steps_against_the_oracle():
    """SURVEY gate G2: CUDA == C restatement on >= 1e8 env-steps.  BASELINE configs[1] (65 536 envs, 10x10, 4 colours, all
    specials, num_moves=30, same-step autoreset) for 1 530 steps = 1.0e8 env-steps and 51 episodes per env: rewards,
    terminations and the activation counter every step, every buffer every 90 steps."""
    torch = _torch()
    N, R, Cc, K, moves, steps = 65536, 10, 10, 4, 30, 1530
    env = make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=2, autoreset="same_step")
    g = GpuAdapter(env)
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=2, autoreset="same_step", num_threads=os.cpu_count() or 8)
    g.reset(); o.reset()
    rng = np.random.default_rng(2026)
    total = 0
    for t in range(steps):
        a = rng.integers(0, o.A, size=N).astype(np.int32)
        _, rew, term, _, info = env.step(torch.from_numpy(a).cuda())
        o.step(a)
        assert np.array_equal(rew.cpu().numpy(), o.reward), t
        assert np.array_equal(info["num_specials_activated"].cpu().numpy(), o.num_specials_activated), t
        if t % 90 == 89 or t == steps - 1:
            assert_same(g, o, f"step {t}")
        total += N
    assert total >= 100_000_000 and int(o.episode.max()) == steps // moves
    assert int(o.status.sum()) == 0
    print("env-steps compared:", total, "oracle diag (max lines, max DFS depth, max reset iterations):", o.diag())


def test_state_dict_records_the_modes_and_refuses_a_mismatch():
    env = make_gpu(64, 6, 6, 4, 5, ALL_CL, ALL_CS, seed=3, autoreset="next_step")
    env.reset()
    sd = env.state_dict()
    assert sd["config"]["autoreset"] == "next_step" and sd["config"]["refill"] == "philox"
    other = make_gpu(64, 6, 6, 4, 5, ALL_CL, ALL_CS, seed=3, autoreset="same_step")
    with pytest.raises(ValueError):
        other.load_state_dict(sd)
    same = make_gpu(64, 6, 6, 4, 5, ALL_CL, ALL_CS, seed=3, autoreset="next_step")
    same.load_state_dict(sd)


def test_host_mirror_outlives_a_dropped_stepper():
    """The pinned arrays of a bound host mirror belong to the env until they are unbound: dropping the HostStepper without
    close() must neither free them under the kernels nor leave the env writing into freed memory."""
    import gc
    torch = _torch()
    from tile_match_gym_b200 import HostStepper
    N = 4096
    env = make_gpu(N, 10, 10, 4, 30, ALL_CL, ALL_CS, seed=4, autoreset="same_step")
    env.reset()
    hs = HostStepper(env, outputs=("board", "reward", "terminated", "mask_bits", "num_moves_left"), mirror=True)
    keep = env._mirror_keep
    assert keep is hs.host
    del hs
    gc.collect()
    assert env._mirror_keep is None                      # the finaliser unbound the mirror (and synchronised) ...
    junk = [torch.empty(4096, dtype=torch.uint8).pin_memory() for _ in range(64)]   # ... so recycled pinned memory is safe
    for j in junk:
        j.fill_(7)
    a = torch.randint(0, env.num_actions, (N,), device="cuda", dtype=torch.int32)
    for _ in range(5):
        env.step(a)
    torch.cuda.synchronize()
    assert all(bool((j == 7).all()) for j in junk)
    hs2 = HostStepper(env, outputs=("board", "reward", "terminated", "mask_bits", "num_moves_left"), mirror=True)
    hs2.step(a.cpu().numpy())
    assert np.array_equal(hs2.host["board"].numpy(), env.board.cpu().numpy())
    env.close()                                          # unbinds before destroying the handle
    assert env._mirror_keep is None


@pytest.mark.parametrize("N,R,Cc,K,moves,autoreset", [(4096, 32, 32, 7, 6, "same_step"), (2048, 10, 10, 4, 5, "next_step"),
                                                     (1024, 20, 20, 5, 4, "same_step")])
def test_constructive_reset_on_device(N, R, Cc, K, moves, autoreset):
    """SURVEY 8f.2 / TMG_FLAG_CONSTRUCTIVE_RESET (NOT reference behaviour): config 5's shape with NO host-side board injection
    -- boards from the on-device constructive line-free sampler at reset, at every autoreset and through the pool -- against
    the same contract restated in the oracle; boards are line-free, all normal, and have a possible move."""
    torch = _torch()
    g = GpuAdapter(make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=14, autoreset=autoreset, constructive_reset=True))
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=14, autoreset=autoreset, num_threads=os.cpu_count() or 8,
                         constructive_reset=True)
    g.reset(); o.reset()
    assert_same(g, o, "reset")
    b = o.board[:, 0]
    assert (o.board[:, 1] == 1).all() and b.min() >= 1 and b.max() <= K and o.mask.any(axis=1).all()
    assert not ((b[:, :, :-2] == b[:, :, 1:-1]) & (b[:, :, 1:-1] == b[:, :, 2:])).any()
    assert not ((b[:, :-2] == b[:, 1:-1]) & (b[:, 1:-1] == b[:, 2:])).any()
    rng = np.random.default_rng(6)
    for t in range(3 * moves + 2):
        a = rng.integers(0, o.A, size=N).astype(np.int32)
        g.step(a); o.step(a)
        assert_same(g, o, f"step {t}")
    assert int(o.status.sum()) == 0 and int(o.episode.max()) >= 2


@pytest.mark.parametrize("mirror", [True, False])
def test_packed_board_host_output(mirror):
    """tmg_host_io.board_packed: the board as one byte per cell (colour | (type & 7) << 4), as a host mirror the step
    kernel writes in place and as a plain copy -- decoded, it is the board of every env after every step."""
    torch = _torch()
    from tile_match_gym_b200 import HostStepper
    N, R, Cc, K, moves = 3000, 10, 10, 4, 7
    env = make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=6, autoreset="same_step")
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=6, autoreset="same_step", num_threads=8)
    env.reset(); o.reset()
    hs = HostStepper(env, outputs=("board_packed", "reward", "terminated", "mask_bits", "num_moves_left"), mirror=mirror)
    rng = np.random.default_rng(8)
    for t in range(3 * moves + 1):
        m = o.mask.astype(np.float64) + 1e-9
        u = rng.random((N, 1)) * m.sum(axis=1, keepdims=True)
        a = (np.cumsum(m, axis=1) < u).sum(axis=1).clip(0, o.A - 1).astype(np.int32)
        out = hs.step(a); o.step(a)
        assert np.array_equal(hs.board(), o.board), t
        assert np.array_equal(out["reward"].numpy(), o.reward) and np.array_equal(out["terminated"].numpy(), o.terminated), t
        assert np.array_equal(np.unpackbits(out["mask_bits"].numpy(), axis=1, bitorder="little")[:, :o.A], o.mask), t
    hs.close()
    big = make_gpu(8, 5, 5, 16, 5, ALL_CL, ALL_CS, seed=1, autoreset="disabled")   # 16 colours do not fit the nibble
    big.reset()
    with pytest.raises(RuntimeError):
        HostStepper(big, outputs=("board_packed",), mirror=True)


@pytest.mark.parametrize("R,Cc,K", [(10, 10, 3), (10, 10, 4), (10, 10, 5), (10, 10, 8), (9, 9, 3), (9, 9, 4), (9, 9, 6), (9, 9, 8)])
def test_packed_row_generator_all_widths(R, Cc, K):
    """Board::generate_packed with the colour window of the reset stream: both cell widths (K <= 4: 2-bit stream and
    funnel-shifted rows; K <= 8: byte window) at both fixed row lengths, through tmg_reset, the pool (k_pregen) and the
    in-step fallback (TMG_FLAG_NO_PREGEN) -- boards, masks, cursors and the iteration-cap status against the oracle."""
    N, moves = 512, 3
    for kw in ({}, {"pregenerate": False}):
        g = GpuAdapter(make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=11, autoreset="same_step", **kw))
        o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=11, autoreset="same_step", num_threads=os.cpu_count() or 8)
        g.reset(); o.reset()
        assert_same(g, o, "reset")
        rng = np.random.default_rng(5)
        for t in range(2 * moves + 1):
            a = rng.integers(0, o.A, size=N).astype(np.int32)
            g.step(a); o.step(a)
            assert_same(g, o, f"step {t} {kw}")
        assert int(o.episode.max()) >= 2
