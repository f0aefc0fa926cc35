"""Trace / replay archive (tile_match_gym_b200/trace.py): the file format on the CPU, and on the GPU the
two directions of the exchange -- a trajectory recorded by the CUDA env replays bit for bit on the oracle, a trajectory
recorded by the oracle (tests/golden/oracle_trace.npz, made by tests/golden/gen_trace.py) replays bit for bit on the GPU."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import trace as otrace

ALL_CL = ("cookie",)
ALL_CS = ("vertical_laser", "horizontal_laser", "bomb")


def _golden():
    from tile_match_gym_b200.trace import load_trace
    return load_trace(os.path.join(GOLDEN, "oracle_trace.npz"))


# ------------------------------------------------------------------------------------------ CPU: format + oracle
def test_golden_trace_is_well_formed_and_replays_on_the_oracle():
    tr = _golden()
    assert tr["actions"].shape == (40, 48) and tr["board"].shape == (40, 48, 2, 10, 10)
    assert tr["mask"].shape == (40, 48, (180 + 7) // 8)
    assert int(tr["is_combination_match"].sum()) > 0 and int(tr["num_specials_activated"].sum()) > 0
    assert int(tr["terminated"].sum()) == 48 * (40 // 6)          # same-step autoreset, 6-move episodes
    assert otrace.replay(tr) == 40


def test_product_and_oracle_agree_on_the_key_names():
    from tile_match_gym_b200 import trace as ptrace
    from tile_match_gym_b200.vec_env import TileMatchVecEnv
    assert ptrace.FORMAT == otrace.FORMAT and ptrace.CFG_KEYS == otrace.CFG_KEYS
    assert ptrace.STEP_I32 == otrace.STEP_I32 and ptrace.STEP_U8 == otrace.STEP_U8
    assert TileMatchVecEnv.STATE_FIELDS == otrace.STATE_FIELDS


def test_save_load_round_trip_and_structure_checks(tmp_path):
    from tile_match_gym_b200.trace import check_trace, load_trace, save_trace
    cfg = {"seed": 3, "num_envs": 5, "num_rows": 4, "num_cols": 6, "num_colours": 3, "num_moves": 4, "specials": 15,
           "env_id_offset": 9, "autoreset": 1, "refill": 1}
    rng = np.random.default_rng(0)
    draws = rng.integers(1, 4, size=(5, 4000)).astype(np.uint8)
    o = otrace.oracle_from_config(cfg)
    o.set_injected_draws(draws)
    o.reset()
    acts = rng.integers(0, o.A, size=(11, 5)).astype(np.int32)
    tr = otrace.record(o, cfg, acts, draws=draws)
    p = tmp_path / "t.npz"
    save_trace(p, tr)
    back = load_trace(p)
    assert sorted(back) == sorted(tr)
    for k in tr:
        assert np.array_equal(back[k], tr[k]) and back[k].dtype == tr[k].dtype, k
    assert otrace.replay(back) == 11
    # a tampered trajectory is caught, a malformed archive is refused
    bad = dict(back); bad["reward"] = back["reward"].copy(); bad["reward"][7, 2] += 1
    with pytest.raises(AssertionError, match="step 7: reward"):
        otrace.replay(bad)
    for breaker in (lambda d: d.pop("cfg/seed"), lambda d: d.pop("draws"), lambda d: d.pop("state0/timer"),
                    lambda d: d.__setitem__("mask", d["mask"][:, :, :-1]), lambda d: d.__setitem__("format", np.array("x")),
                    lambda d: d.__setitem__("terminated", d["terminated"][:-1])):
        d = dict(back); breaker(d)
        with pytest.raises(ValueError):
            check_trace(d)


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_oracle_recorded_golden_trace_replays_on_the_gpu():
    from tile_match_gym_b200.trace import TraceMismatch, env_from_trace, replay_trace
    tr = _golden()
    env = env_from_trace(tr)
    assert replay_trace(env, tr) == 40
    assert int((env.status != 0).sum().item()) == 0
    # the same handle replays again from the archive's start state (load_state_dict rewinds the cursors) ...
    assert replay_trace(env, tr) == 40
    # ... and a trajectory that is not the recorded one is reported with its step
    bad = dict(tr); bad["board"] = tr["board"].copy(); bad["board"][13, 4, 0, 9, 9] ^= 1
    with pytest.raises(TraceMismatch, match="step 13: board"):
        replay_trace(env, bad)


@pytest.mark.gpu
@pytest.mark.parametrize("refill,autoreset,R,Cc,K,moves", [("philox", "same_step", 10, 10, 4, 7),
                                                           ("philox", "next_step", 9, 9, 6, 5),
                                                           ("injected", "same_step", 6, 7, 4, 6)])
def test_gpu_recorded_trace_replays_on_the_oracle_and_on_a_fresh_gpu_env(tmp_path, refill, autoreset, R, Cc, K, moves):
    import torch
    from tile_match_gym_b200 import TileMatchVecEnv
    from tile_match_gym_b200.trace import env_from_trace, load_trace, record_trace, replay_trace, save_trace
    N, T = (512 if refill == "injected" else 1500), 24
    env = TileMatchVecEnv(N, R, Cc, K, moves, list(ALL_CL), list(ALL_CS), seed=21, device="cuda:0", autoreset=autoreset,
                          refill=refill, env_id_offset=300)
    if refill == "injected":
        rng = np.random.default_rng(8)
        env.set_injected_draws(torch.from_numpy(rng.integers(1, K + 1, size=(N, 12000)).astype(np.uint8)).cuda())
    env.reset()
    gen = torch.Generator(device="cuda"); gen.manual_seed(5)
    warm = torch.randint(0, env.num_actions, (3, N), device="cuda", dtype=torch.int32, generator=gen)
    for a in warm:                       # the trace starts mid-episode, with non-zero cursors
        env.step(a)
    # two steps out of three take an effective action, so that the trace holds cascades and specials
    acts = torch.randint(0, env.num_actions, (T, N), device="cuda", dtype=torch.int32, generator=gen)
    parts = []
    for t in range(T):
        if t % 3:
            m = env.mask.float()
            has = m.sum(1) > 0
            pick = torch.multinomial(torch.where(has[:, None], m, torch.ones_like(m)), 1, generator=gen)[:, 0].int()
            acts[t] = torch.where(has, pick, acts[t])
        parts.append(record_trace(env, acts[t:t + 1]))
    tr = dict(parts[0])
    tr["actions"] = acts.cpu().numpy()
    for k in ("reward", "num_new_specials", "num_specials_activated", "num_moves_left", "terminated",
              "is_combination_match", "shuffled", "board", "mask"):
        tr[k] = np.concatenate([p[k] for p in parts])
    assert int((env.status != 0).sum().item()) == 0
    assert int(tr["num_specials_activated"].sum()) > 0 and int(tr["terminated"].sum()) > 0
    p = tmp_path / "gpu.npz"
    save_trace(p, tr)
    back = load_trace(p)
    assert otrace.replay(back, num_threads=4) == T            # the oracle reproduces what the GPU recorded
    fresh = env_from_trace(back)
    assert replay_trace(fresh, back) == T                     # and so does a fresh handle (checkpoint with its proof)
