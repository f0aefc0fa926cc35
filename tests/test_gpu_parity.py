"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against the
oracle on the same seeded inputs and against the committed golden fixtures.  Bit-exact: every comparison is
array equality on integer state."""
import ctypes as C
import gzip
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, expected_policy_actions
from oracle import oracle as orc
from test_oracle_golden import _trace_meta, replay_trace, _draws, _split_specials

pytestmark = pytest.mark.gpu

ALL_CL = ("cookie",)
ALL_CS = ("vertical_laser", "horizontal_laser", "bomb")
FIELDS = ["board", "timer", "draw_cursor", "shuffle_cursor", "reward", "terminated", "is_combination_match",
          "num_new_specials", "num_specials_activated", "shuffled", "mask", "num_moves_left", "status", "episode"]


def _torch():
    import torch
    return torch


class GpuAdapter:
    """Gives TileMatchVecEnv the numpy attribute protocol of OracleVecEnv."""

    def __init__(self, env):
        self.env = env
        self.A = env.num_actions

    def reset(self, reset_mask=None, init_boards=None):
        torch = _torch()
        opts = {}
        if reset_mask is not None:
            opts["reset_mask"] = torch.as_tensor(np.asarray(reset_mask, dtype=np.uint8))
        if init_boards is not None:
            opts["init_boards"] = torch.as_tensor(np.asarray(init_boards, dtype=np.int8))
        self.env.reset(options=opts)

    def step(self, actions):
        torch = _torch()
        self.env.step(torch.as_tensor(np.asarray(actions, dtype=np.int32)))

    def __getattr__(self, name):
        if name in FIELDS:
            t = getattr(self.env, name)
            _torch().cuda.synchronize()
            a = t.cpu().numpy()
            if a.dtype == np.bool_:
                a = a.astype(np.uint8)
            if name in ("draw_cursor", "shuffle_cursor"):
                a = a.astype(np.uint64)
            if name == "status":
                a = a.astype(np.uint32)
            return a
        raise AttributeError(name)


def make_gpu(N, R, Cc, K, moves, cl, cs, **kw):
    from tile_match_gym_b200 import TileMatchVecEnv
    return TileMatchVecEnv(N, R, Cc, K, moves, list(cl), list(cs), device="cuda:0", **kw)


def assert_same(g, o, ctx, fields=FIELDS):
    for f in fields:
        a, b = getattr(g, f), getattr(o, f)
        if not np.array_equal(a, b):
            bad = np.flatnonzero((a.reshape(a.shape[0], -1) != b.reshape(b.shape[0], -1)).any(axis=1))
            raise AssertionError(f"{ctx}: field {f} differs for {len(bad)} envs, first {bad[:5]}\n"
                                 f"gpu {a[bad[0]]}\noracle {b[bad[0]]}")


def test_library_loaded_and_is_the_cuda_one():
    from tile_match_gym_b200 import _native
    L = _native.lib()
    assert L.tmg_abi_version() == 2
    assert os.path.basename(_native.LIB_PATH) == "libtmg_b200.so"


def test_golden_traces_on_gpu():
    z, meta = _trace_meta()
    for m in meta:
        replay_trace(lambda m: GpuAdapter(make_gpu(1, m["R"], m["C"], m["K"], m["num_moves"], m["cl"], m["cs"],
                                                   seed=m["seed"], env_id_offset=m["env_id"], autoreset="disabled")), z, m)


@pytest.mark.parametrize("cfg", [
    # R, C, K, cl, cs, moves, autoreset, policy, N, steps
    (10, 10, 4, (), (), 30, "same_step", "uniform", 2048, 70),
    (10, 10, 4, ALL_CL, ALL_CS, 30, "same_step", "uniform", 2048, 70),
    (10, 10, 4, ALL_CL, ALL_CS, 12, "next_step", "mask", 2048, 60),
    (9, 9, 6, ALL_CL, ALL_CS, 10, "same_step", "mask", 2048, 60),
    (5, 5, 4, ALL_CL, ALL_CS, 9, "next_step", "mask", 4096, 60),
    (3, 5, 3, ALL_CL, ALL_CS, 7, "same_step", "mask", 4096, 60),
    (4, 4, 3, ALL_CL, ALL_CS, 5, "same_step", "mask", 4096, 60),
    (8, 16, 5, ALL_CL, ALL_CS, 8, "same_step", "mask", 1024, 40),
    (16, 7, 4, ("cookie",), ("bomb",), 8, "next_step", "mask", 1024, 40),
    (12, 20, 6, ALL_CL, ALL_CS, 6, "same_step", "mask", 512, 30),
    (7, 7, 3, (), ("horizontal_laser",), 9, "disabled", "mask", 1024, 9),
])
def test_batched_parity_vs_oracle(cfg):
    torch = _torch()
    R, Cc, K, cl, cs, moves, autoreset, policy, N, steps = cfg
    g = GpuAdapter(make_gpu(N, R, Cc, K, moves, cl, cs, seed=5, autoreset=autoreset, env_id_offset=1000))
    o = orc.OracleVecEnv(N, R, Cc, K, moves, cl, cs, seed=5, autoreset=autoreset, env_id_offset=1000, num_threads=8)
    g.reset(); o.reset()
    assert_same(g, o, "reset")
    rng = np.random.default_rng(42)
    for t in range(steps):
        if policy == "uniform":
            a = rng.integers(0, o.A, size=N)
        else:  # sample from the legal-move mask when it is non-empty (src/examples/random_agent.py:12-31)
            m = o.mask.astype(np.float64) + 1e-9
            u = rng.random((N, 1)) * m.sum(axis=1, keepdims=True)
            a = (np.cumsum(m, axis=1) < u).sum(axis=1).clip(0, o.A - 1)
        a = a.astype(np.int32)
        g.step(a); o.step(a)
        assert_same(g, o, f"step {t}")
    d = o.diag()
    print(cfg[:3], "oracle diag", d)


def test_injected_draws_parity():
    torch = _torch()
    N, R, Cc, K, moves = 512, 6, 7, 4, 6
    rng = np.random.default_rng(3)
    draws = rng.integers(1, K + 1, size=(N, 6000)).astype(np.uint8)
    g_env = make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=9, refill="injected", autoreset="same_step")
    g_env.set_injected_draws(torch.from_numpy(draws).cuda())
    g = GpuAdapter(g_env)
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=9, refill="injected", autoreset="same_step", num_threads=8)
    o.set_injected_draws(draws)
    g.reset(); o.reset()
    assert_same(g, o, "reset")
    for t in range(40):
        a = rng.integers(0, o.A, size=N).astype(np.int32)
        g.step(a); o.step(a)
        assert_same(g, o, f"step {t}")
    # short streams must exhaust identically (flag + colour 1)
    g2_env = make_gpu(64, R, Cc, K, moves, ALL_CL, ALL_CS, seed=9, refill="injected", autoreset="same_step", max_reset_iters=50)
    g2_env.set_injected_draws(torch.from_numpy(draws[:64, :200].copy()).cuda())
    g2 = GpuAdapter(g2_env)
    o2 = orc.OracleVecEnv(64, R, Cc, K, moves, ALL_CL, ALL_CS, seed=9, refill="injected", autoreset="same_step",
                          max_reset_iters=50, num_threads=4)
    o2.set_injected_draws(draws[:64, :200].copy())
    g2.reset(); o2.reset()
    assert_same(g2, o2, "short reset")
    assert (o2.status & orc.ST_DRAWS_EXHAUSTED).any()


def test_no_mask_flag_and_status_bits():
    N, R, Cc, K, moves = 1024, 8, 8, 4, 5
    g = GpuAdapter(make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=4, autoreset="disabled", compute_mask=False))
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=4, autoreset="disabled", num_threads=8)
    no_mask = [f for f in FIELDS if f != "mask"]
    rng = np.random.default_rng(0)
    a = rng.integers(0, o.A, size=N).astype(np.int32)
    g.step(a); o.step(a)                       # step before reset -> NEEDS_RESET everywhere
    assert_same(g, o, "before reset", no_mask)
    assert (g.status & 2).all()
    g.reset(); o.reset()
    for t in range(moves + 2):                 # two steps past the end -> NEEDS_RESET again
        a = rng.integers(-3, o.A + 3, size=N).astype(np.int32)   # includes out-of-range ids -> BAD_ACTION
        g.step(a); o.step(a)
        assert_same(g, o, f"step {t}", no_mask)
    assert (g.status & 1).any() and (g.status & 2).all()
    from tile_match_gym_b200 import _native
    with pytest.raises(Exception):
        g.env.check_status()
    assert int(g.status.sum()) == 0            # cleared


def test_sharding_is_invisible():
    """Env e's trajectory depends on (seed, global id) only: two shards == one big batch."""
    N, R, Cc, K, moves = 1000, 10, 10, 4, 7
    from tile_match_gym_b200 import TileMatchVecEnv
    whole = GpuAdapter(make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=8, autoreset="same_step"))
    parts = [GpuAdapter(TileMatchVecEnv.sharded(N, r, 3, R, Cc, K, moves, list(ALL_CL), list(ALL_CS), seed=8,
                                                device="cuda:0", autoreset="same_step")) for r in range(3)]
    whole.reset(); [p.reset() for p in parts]
    rng = np.random.default_rng(1)
    for t in range(20):
        a = rng.integers(0, whole.A, size=N).astype(np.int32)
        whole.step(a)
        lo = 0
        for p in parts:
            n = p.env.num_envs
            p.step(a[lo:lo + n]); lo += n
        for f in ["board", "reward", "terminated", "mask", "draw_cursor"]:
            assert np.array_equal(getattr(whole, f), np.concatenate([getattr(p, f) for p in parts])), (t, f)


def test_onehot_and_int32_obs():
    torch = _torch()
    N, R, Cc, K, moves = 777, 9, 9, 6, 30
    for cl, cs in [(ALL_CL, ALL_CS), ((), ("bomb",)), (("cookie",), ("vertical_laser",)), ((), ())]:
        env = make_gpu(N, R, Cc, K, moves, cl, cs, seed=3, autoreset="same_step", obs="onehot")
        o = orc.OracleVecEnv(N, R, Cc, K, moves, cl, cs, seed=3, autoreset="same_step", num_threads=8)
        obs, info = env.reset(); o.reset()
        rng = np.random.default_rng(2)
        for t in range(12):
            m = o.mask.astype(np.float64) + 1e-9
            u = rng.random((N, 1)) * m.sum(axis=1, keepdims=True)
            a = (np.cumsum(m, axis=1) < u).sum(axis=1).clip(0, o.A - 1).astype(np.int32)
            obs, *_ = env.step(torch.from_numpy(a).cuda()); o.step(a)
        want = o.onehot()
        assert obs["board"].shape == want.shape and obs["board"].dtype == torch.uint8
        assert np.array_equal(obs["board"].cpu().numpy(), want)
        assert np.array_equal(env.onehot(torch.float32).cpu().numpy(), want.astype(np.float32))
        assert env.single_observation_space["board"].shape == want.shape[1:]
    env = make_gpu(4, 5, 4, 3, 5, ALL_CL, ALL_CS, obs="int32")
    obs, _ = env.reset()
    assert obs["board"].dtype == torch.int32 and obs["board"].shape == (4, 2, 5, 4)
    assert int(obs["num_moves_left"][0]) == 5


def test_host_buffer_path_matches_device_path():
    from tile_match_gym_b200 import HostStepper
    N, R, Cc, K, moves = 3000, 10, 10, 4, 9
    env = make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=6, autoreset="same_step")
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=6, autoreset="same_step", num_threads=8)
    env.reset(); o.reset()
    hs = HostStepper(env, outputs=("board", "reward", "terminated", "mask", "num_moves_left", "status"))
    rng = np.random.default_rng(5)
    for t in range(25):
        a = rng.integers(0, o.A, size=N).astype(np.int32)
        out = hs.step(a); o.step(a)
        assert np.array_equal(out["board"].numpy(), o.board)
        assert np.array_equal(out["reward"].numpy(), o.reward)
        assert np.array_equal(out["terminated"].numpy(), o.terminated)
        assert np.array_equal(out["mask"].numpy(), o.mask)
        assert np.array_equal(out["num_moves_left"].numpy(), o.num_moves_left)
    assert hs.h2d_bytes == 4 * N and hs.d2h_bytes == N * (200 + 4 + 1 + 180 + 4 + 4)
    # bit-packed mask form of the same call
    hb = HostStepper(env, outputs=("reward", "mask_bits"))
    for t in range(5):
        a = rng.integers(0, o.A, size=N).astype(np.int32)
        out = hb.step(a); o.step(a)
        bits = np.unpackbits(out["mask_bits"].numpy(), axis=1, bitorder="little")[:, :o.A]
        assert np.array_equal(bits, o.mask) and np.array_equal(out["reward"].numpy(), o.reward)
    assert hb.effective_actions(3) == np.flatnonzero(o.mask[3]).tolist()


@pytest.mark.parametrize("autoreset,R,Cc,K", [("same_step", 10, 10, 4), ("next_step", 9, 9, 6), ("disabled", 7, 12, 5)])
def test_host_mirror_write_through_matches_oracle(autoreset, R, Cc, K):
    """tmg_host_bind: the step kernel updates pinned host arrays in place for the envs it changes; after every call
    the arrays must hold the complete current board / mask / packed mask."""
    from tile_match_gym_b200 import HostStepper
    N, moves = 2500, 7
    env = make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=8, autoreset=autoreset)
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=8, autoreset=autoreset, num_threads=8)
    env.reset(); o.reset()
    hs = HostStepper(env, outputs=("board", "reward", "terminated", "mask", "mask_bits", "num_moves_left"), mirror=True)
    _torch().cuda.synchronize()
    assert np.array_equal(hs.host["board"].numpy(), o.board) and np.array_equal(hs.host["mask"].numpy(), o.mask)
    rng = np.random.default_rng(15)
    for t in range(3 * moves + 2):
        if autoreset == "disabled" and t and t % moves == 0:
            env.reset(); o.reset()          # tmg_reset re-copies the bound arrays in full
        a = rng.integers(0, o.A, size=N).astype(np.int32)
        out = hs.step(a); o.step(a)
        assert np.array_equal(out["board"].numpy(), o.board), t
        assert np.array_equal(out["mask"].numpy(), o.mask), t
        assert np.array_equal(np.unpackbits(out["mask_bits"].numpy(), axis=1, bitorder="little")[:, :o.A], o.mask), t
        assert np.array_equal(out["reward"].numpy(), o.reward) and np.array_equal(out["terminated"].numpy(), o.terminated)
        assert np.array_equal(out["num_moves_left"].numpy(), o.num_moves_left)
    hs.close()
    # pageable memory is refused, not silently staged
    import ctypes as C
    bad = np.zeros((N, 2, R, Cc), np.int8)
    from tile_match_gym_b200 import _native as nat
    io = nat.HostIO(); io.board = bad.ctypes.data
    assert env._lib.tmg_host_bind(env._h, C.byref(io), env._stream()) != 0


@pytest.mark.parametrize("R,Cc,K,moves,autoreset,T", [(10, 10, 4, 30, "same_step", 30), (10, 10, 4, 7, "same_step", 20),
                                                      (9, 9, 6, 5, "next_step", 12), (32, 32, 7, 6, "disabled", 6),
                                                      (6, 14, 5, 8, "same_step", 16)])
def test_step_many_equals_single_steps(R, Cc, K, moves, autoreset, T):
    """tmg_step_many (fused rollout): the state after T steps in one launch and every step's reward / termination
    equal T TileMatchEnv.step calls of the oracle; rollouts and single steps interleave."""
    torch = _torch()
    N = 1500 if R < 32 else 96
    inject = R == 32      # generate_board does not terminate for 32x32 / 7 colours: injected boards
    g = GpuAdapter(make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=21, autoreset=autoreset))
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=21, autoreset=autoreset, num_threads=8)
    if inject:
        z, meta = _trace_meta()
        m32 = [m for m in meta if m["R"] == 32][0]
        b0 = z[m32["name"] + "/init_board"].astype(np.int8)
        boards = np.repeat(b0[None], N, axis=0)
        g.reset(init_boards=boards); o.reset(init_boards=boards)
    else:
        g.env.reset(); o.reset()
    rng = np.random.default_rng(2)
    for window in range(3):
        acts = rng.integers(0, o.A, (T, N)).astype(np.int32)
        rew, term = g.env.step_many(torch.from_numpy(acts).cuda())
        rew, term = rew.cpu().numpy(), term.cpu().numpy()
        for t in range(T):
            o.step(acts[t])
            assert np.array_equal(rew[t], o.reward), (window, t)
            assert np.array_equal(term[t].astype(np.uint8), o.terminated), (window, t)
        assert_same(g, o, f"after rollout window {window}")
        if autoreset == "disabled":
            break
        a = rng.integers(0, o.A, N).astype(np.int32)
        g.step(a); o.step(a)
        assert_same(g, o, f"single step after window {window}")


@pytest.mark.parametrize("R,Cc,K,moves,autoreset,policy", [(10, 10, 4, 30, "same_step", "mask"), (9, 9, 6, 5, "next_step", "mask"),
                                                          (10, 10, 4, 8, "same_step", "uniform"), (7, 9, 5, 6, "disabled", "mask")])
def test_policy_rollout_takes_the_contract_actions(R, Cc, K, moves, autoreset, policy):
    """tmg_rollout_policy: the agent inside the kernel takes exactly the actions its stream contract says (word
    board*num_moves + timer of stream 2; uniform, or the n-th effective action), and the trajectory is the oracle's."""
    N, T, seed, off = 600, 2 * moves + 3 if moves < 30 else 33, 31, 1000
    env = make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=seed, autoreset=autoreset, env_id_offset=off)
    g = GpuAdapter(env)
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=seed, autoreset=autoreset, env_id_offset=off, num_threads=8)
    env.reset(); o.reset()
    act, rew, term = env.rollout(T, policy)
    act, rew, term = act.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy().astype(np.uint8)
    n_eff = 0
    for t in range(T):
        want = expected_policy_actions(o, seed, off, moves, policy)
        assert np.array_equal(act[t], want), t
        o.step(want)
        assert np.array_equal(rew[t], o.reward) and np.array_equal(term[t], o.terminated), t
        n_eff += int((o.reward > 0).sum())
    assert_same(g, o, "after the policy rollout")
    if policy == "mask" and autoreset == "same_step":
        assert n_eff == N * T          # every sampled action is effective


def test_full_occupancy_batch_matches_oracle_across_episode_boundaries():
    """BASELINE configs[1] at its full size: 65 536 envs, every buffer compared with the oracle while the persistent
    kernels are fully occupied, the work list is long, and every env crosses an episode boundary (pool hand-over);
    then one 30-step rollout on top."""
    torch = _torch()
    N, R, Cc, K, moves = 65536, 10, 10, 4, 30
    g = GpuAdapter(make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=2, autoreset="same_step"))
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=2, autoreset="same_step", num_threads=os.cpu_count() or 8)
    g.env.reset(); o.reset()
    assert_same(g, o, "reset")
    rng = np.random.default_rng(77)
    for t in range(34):
        a = rng.integers(0, o.A, size=N).astype(np.int32)
        g.step(a); o.step(a)
        if t in (0, 13, 28, 29, 30, 33):
            assert_same(g, o, f"step {t}")
        else:
            assert np.array_equal(g.reward, o.reward) and np.array_equal(g.terminated, o.terminated), t
    acts = rng.integers(0, o.A, (30, N)).astype(np.int32)
    rew, term = g.env.step_many(torch.from_numpy(acts).cuda())
    rew, term = rew.cpu().numpy(), term.cpu().numpy()
    for t in range(30):
        o.step(acts[t])
        assert np.array_equal(rew[t], o.reward) and np.array_equal(term[t].astype(np.uint8), o.terminated), t
    assert_same(g, o, "after the rollout")


def test_many_episodes_stay_bit_exact():
    """1 500 steps of 3-move episodes = 500 boards per env: the pool of boards generated ahead of time, its request ring
    and its event slots wrap many times; every buffer still equals the oracle's (checked every 50 steps)."""
    N, R, Cc, K, moves = 2048, 10, 10, 4, 3
    g = GpuAdapter(make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=5, autoreset="same_step", env_id_offset=1000))
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=5, autoreset="same_step", env_id_offset=1000, num_threads=8)
    g.reset(); o.reset()
    rng = np.random.default_rng(1)
    for t in range(1500):
        a = rng.integers(0, o.A, size=N).astype(np.int32)
        g.step(a); o.step(a)
        if t % 50 == 49 or t < 5:
            assert_same(g, o, f"step {t}")
    assert_same(g, o, "end")
    assert int(o.episode.max()) == 500


def test_checkpoint_resume_reproduces_the_trajectory():
    """state_dict / load_state_dict: an env resumed from a checkpoint -- the same handle or a fresh one --
    continues bit for bit like the original, including the boards of later episodes and a bound host mirror."""
    torch = _torch()
    N, R, Cc, K, moves = 2000, 10, 10, 4, 6
    mk = lambda: make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=12, autoreset="same_step", env_id_offset=100)  # noqa: E731
    env = mk()
    env.reset()
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    acts = [torch.randint(0, env.num_actions, (N,), device="cuda", dtype=torch.int32, generator=gen) for _ in range(30)]
    for a in acts[:9]:
        env.step(a)
    sd = env.state_dict()
    want = []
    for a in acts[9:]:
        _, rew, term, _, _ = env.step(a)
        want.append((env.board.cpu().clone(), rew.cpu().clone(), term.cpu().clone(), env.mask.cpu().clone()))
    fresh = mk()
    for e in (env, fresh):
        e.load_state_dict(sd)
        for a, (wb, wr, wt, wm) in zip(acts[9:], want):
            _, rew, term, _, _ = e.step(a)
            assert torch.equal(e.board.cpu(), wb) and torch.equal(rew.cpu(), wr) and torch.equal(term.cpu(), wt)
            assert torch.equal(e.mask.cpu(), wm)
        assert int((e.status != 0).sum().item()) == 0


def test_proportion_reward_wrapper():
    """ProportionRewardWrapper (wrappers.py:71-77): reward / (rows * cols) as the reference's float."""
    from tile_match_gym_b200 import ProportionRewardWrapper
    N, R, Cc, K, moves = 512, 9, 9, 6, 12
    env = ProportionRewardWrapper(make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=3, autoreset="same_step"))
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=3, autoreset="same_step", num_threads=4)
    env.reset(); o.reset()
    rng = np.random.default_rng(4)
    for t in range(15):
        m = o.mask.astype(np.float64) + 1e-9
        u = rng.random((N, 1)) * m.sum(axis=1, keepdims=True)
        a = (np.cumsum(m, axis=1) < u).sum(axis=1).clip(0, o.A - 1).astype(np.int32)
        _, rew, _, _, _ = env.step(_torch().from_numpy(a).cuda()); o.step(a)
        want = np.array([float(int(r) / (R * Cc)) for r in o.reward])      # the reference: reward / self.flat_size
        assert rew.dtype == _torch().float64 and np.array_equal(rew.cpu().numpy(), want)


def test_reset_with_seed_and_partial_reset():
    torch = _torch()
    N, R, Cc, K, moves = 600, 6, 6, 4, 50
    g = GpuAdapter(make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=1, autoreset="disabled"))
    g.env.reset(seed=77)
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=77, autoreset="disabled", num_threads=4)
    o.reset()
    assert_same(g, o, "reset(seed)")
    rng = np.random.default_rng(9)
    for t in range(5):
        a = rng.integers(0, o.A, size=N).astype(np.int32)
        g.step(a); o.step(a)
    sel = (rng.random(N) < 0.3).astype(np.uint8)
    g.reset(reset_mask=sel); o.reset(reset_mask=sel)
    assert_same(g, o, "partial reset")
    assert (g.timer[sel == 1] == 0).all() and (g.timer[sel == 0] == 5).all()


def test_large_shape_properties():
    """BASELINE-size batch: properties that do not need the oracle (sortedness of nothing, but invariants)."""
    torch = _torch()
    N = 65536
    env = make_gpu(N, 10, 10, 4, 30, ALL_CL, ALL_CS, seed=2, autoreset="same_step")
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    total_reward = 0
    for t in range(64):
        a = torch.randint(0, env.num_actions, (N,), device="cuda", dtype=torch.int32, generator=g)
        before = env.board.clone()
        gate = env.mask.gather(1, a.long()[:, None])[:, 0].clone()
        term_next = env.timer == 29
        _, r, term, _, info = env.step(a)
        # ineffective moves leave the board untouched and score 0 (board.py:352-353)
        same = (env.board == before).flatten(1).all(1)
        assert bool((same | gate | term_next).all())
        assert bool((r[~gate] == 0).all()) and bool((r[gate] >= 3).all())
        # boards stay full and valid: colours 1..K for typed tiles, (0,-1) cookies only
        col, typ = env.board[:, 0], env.board[:, 1]
        assert bool(((typ == -1) == (col == 0)).all()) and bool(((typ >= -1) & (typ <= 4) & (typ != 0)).all())
        assert bool((col <= 4).all())
        # no 3-line survives a step (cascade ran to a fixed point)
        h = (col[:, :, :-2] == col[:, :, 1:-1]) & (col[:, :, 1:-1] == col[:, :, 2:]) & (col[:, :, :-2] > 0)
        v = (col[:, :-2] == col[:, 1:-1]) & (col[:, 1:-1] == col[:, 2:]) & (col[:, :-2] > 0)
        assert not bool(h.any()) and not bool(v.any())
        assert bool((term == (t % 30 == 29)).all())
        total_reward += int(r.sum())
    assert int(env.status.abs().sum()) == 0 and total_reward > 0
    # determinism: same seed, same actions -> same state
    env2 = make_gpu(N, 10, 10, 4, 30, ALL_CL, ALL_CS, seed=2, autoreset="same_step")
    env2.reset()
    g.manual_seed(0)
    for t in range(64):
        a = torch.randint(0, env.num_actions, (N,), device="cuda", dtype=torch.int32, generator=g)
        env2.step(a)
    assert bool((env2.board == env.board).all()) and bool((env2.draw_cursor == env.draw_cursor).all())


def test_reference_known_answers_on_device():
    """G3: the function-level calls of the reference's own tests, replayed through tmg_debug_op."""
    torch = _torch()
    with gzip.open(os.path.join(GOLDEN, "ref_test_calls.json.gz"), "rt") as f:
        recs = json.load(f)
    done = {}
    for rec in recs:
        fn = rec["fn"]
        if fn not in ("activate_special", "combination_match", "gravity", "resolve_colour_matches", "move", "refill",
                      "get_colour_lines", "is_move_effective") or rec.get("err"):
            continue
        if fn == "is_move_effective":
            arr = np.asarray(rec["board"], dtype=np.int32)
            R, Cc = arr.shape[1:]
            env = make_gpu(1, R, Cc, 9, 10, ALL_CL, ALL_CS, autoreset="disabled")
            env.reset(options={"init_boards": torch.as_tensor(arr.astype(np.int8))[None]})
            args = np.array([[*rec["c1"], *rec["c2"]]], dtype=np.int32)
            env.debug_op("effective", args)
            assert bool(env.reward[0].item()) == rec["ret"]
            done[fn] = done.get(fn, 0) + 1
            continue
        pre, post = rec["pre"], rec["post"]
        if pre is None or pre["board"] is None or (fn == "get_colour_lines" and done.get(fn, 0) >= 150):
            continue
        draws, has_shuffle = _draws(rec.get("rng", []))
        if has_shuffle:
            continue
        cl, cs = _split_specials(pre["specials"])
        R, Cc, K = pre["R"], pre["C"], pre["K"]
        if Cc < 2:
            continue
        b = np.asarray(pre["board"], dtype=np.int8)
        ob = orc.OracleBoard(R, Cc, K, cl, cs, board=np.asarray(pre["board"], dtype=np.int32))
        if fn == "resolve_colour_matches":   # only when called on detect_colour_matches' own output (it always is)
            coords, names, colours = ob.detect_colour_matches()
            if coords != [[tuple(c) for c in l] for l in rec["args"][0]]:
                continue
        env = make_gpu(1, R, Cc, max(K, 1), 10, cl, cs, autoreset="disabled", refill="injected")
        d = np.concatenate([draws, np.ones(4, np.uint8)])[None]
        env.set_injected_draws(torch.from_numpy(d).cuda())
        env.reset(options={"init_boards": torch.from_numpy(b)[None]})
        env._lib.tmg_clear_status(env._h, None)   # hand-made test boards may hold empties: "invalid" for reset, fine here
        env.num_new_specials[0] = pre["new"]; env.num_specials_activated[0] = pre["act"]
        a = rec["args"]
        if fn == "activate_special":
            is_comb = rec["kwargs"].get("is_combination_match", a[3] if len(a) > 3 else False)
            env.debug_op("activate", np.array([[a[0][0], a[0][1], a[1], int(bool(is_comb))]], np.int32))
        elif fn == "combination_match":
            env.debug_op("combine", np.array([[a[0][0], a[0][1], a[1][0], a[1][1]]], np.int32))
        elif fn == "gravity":
            env.debug_op("gravity")
        elif fn == "refill":
            env.debug_op("refill")
        elif fn == "resolve_colour_matches":
            env.debug_op("resolve_round")
        elif fn == "get_colour_lines":
            env.debug_op("count_lines")
            assert int(env.reward[0].item()) == len(rec["ret"])
            done[fn] = done.get(fn, 0) + 1
            continue
        elif fn == "move":
            env.debug_op("move", np.array([[a[0][0], a[0][1], a[1][0], a[1][1]]], np.int32))
            ret = rec["ret"]
            assert int(env.reward[0].item()) == ret[0] and bool(env.is_combination_match[0].item()) == bool(ret[1])
            assert bool(env.shuffled[0].item()) == bool(ret[4])
            assert int(env.draw_cursor[0].item()) == len(draws)
        assert np.array_equal(env.board[0].cpu().numpy(), np.asarray(post["board"], dtype=np.int8)), (fn, a)
        assert (int(env.num_new_specials[0].item()), int(env.num_specials_activated[0].item())) == (post["new"], post["act"]), fn
        assert int(env.status[0].item()) == 0
        done[fn] = done.get(fn, 0) + 1
    for fn in ("activate_special", "combination_match", "gravity", "resolve_colour_matches", "move", "refill",
               "get_colour_lines", "is_move_effective"):
        assert done.get(fn, 0) > 0, (fn, done)
    print("device KAT replays:", done)


def test_32x32_seven_colours_injected_boards():
    """BASELINE config 5: generate_board never terminates for 32x32 / 7 colours in the reference, so both
    sides start from injected line-free boards; long cascades, divergence-heavy."""
    torch = _torch()
    N, R, Cc, K, moves = 192, 32, 32, 7, 12
    rng = np.random.default_rng(77)

    def no_line_board():
        b = np.zeros((R, Cc), dtype=np.int8)
        for r in range(R):
            for c in range(Cc):
                while True:
                    k = int(rng.integers(1, K + 1))
                    if c >= 2 and b[r, c - 1] == k and b[r, c - 2] == k:
                        continue
                    if r >= 2 and b[r - 1, c] == k and b[r - 2, c] == k:
                        continue
                    b[r, c] = k
                    break
        return np.stack([b, np.ones_like(b)])

    boards = np.stack([no_line_board() for _ in range(N)]).astype(np.int8)
    g = GpuAdapter(make_gpu(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=12, autoreset="disabled"))
    o = orc.OracleVecEnv(N, R, Cc, K, moves, ALL_CL, ALL_CS, seed=12, autoreset="disabled", num_threads=8)
    g.reset(init_boards=boards); o.reset(init_boards=boards)
    assert_same(g, o, "inject")
    for t in range(moves):
        m = o.mask.astype(np.float64) + 1e-9
        u = rng.random((N, 1)) * m.sum(axis=1, keepdims=True)
        a = (np.cumsum(m, axis=1) < u).sum(axis=1).clip(0, o.A - 1).astype(np.int32)
        g.step(a); o.step(a)
        assert_same(g, o, f"step {t}")
    assert int(o.status.sum()) == 0
