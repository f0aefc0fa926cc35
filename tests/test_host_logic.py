"""Host-side logic that needs no GPU: sharding arithmetic, spaces, and the episode-statistics all-reduce over a
world_size-2 gloo group (the only collective in the design; NCCL on the GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tile_match_gym_b200 import spaces as sp
from tile_match_gym_b200.vec_env import EpisodeStatistics, shard_range


def test_shard_range_partitions_exactly():
    for n in (1, 7, 65536, 1048576, 1000003):
        for g in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, g) for r in range(g)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(g - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_spaces_match_reference_bounds():
    # tile_match_env.py:52-77 for TileMatchEnv(3, 5, 3, 4, ["cookie"], ["bomb","vertical_laser","horizontal_laser"])
    b = sp.board_space(3, 5, 3, 1, 3)
    assert b.shape == (2, 3, 5) and b.dtype == np.int32
    assert b.low[0].max() == 0 and b.low[1].min() == -1 and b.high[0].min() == 3 and b.high[1].min() == 5
    assert sp.Discrete(2 * 3 * 5 - 3 - 5).n == 22
    oh = sp.onehot_board_space(4, 3, 5)       # tests/test_wrappers.py:8
    assert oh.shape == (5, 4, 3)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(10, rank, world)
    n = hi - lo
    st = EpisodeStatistics(n, "cpu")
    rng = np.random.default_rng(100)
    all_r = rng.integers(0, 9, size=(6, 10)); all_t = np.zeros((6, 10), bool); all_t[2] = True; all_t[5] = True
    for t in range(6):
        r = torch.from_numpy(all_r[t, lo:hi]).int(); term = torch.from_numpy(all_t[t, lo:hi])
        info = {"num_new_specials": torch.ones(n, dtype=torch.int32), "num_specials_activated": torch.zeros(n, dtype=torch.int32),
                "shuffled": torch.zeros(n, dtype=torch.bool), "is_combination_match": term.clone()}
        st.update(r, term, info)
    q.put((rank, st.allreduce()))
    dist.destroy_process_group()


def test_episode_statistics_allreduce_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = dict(q.get(timeout=120) for _ in range(world))
    [p.join(60) for p in ps]
    assert res[0] == res[1]
    rng = np.random.default_rng(100)
    all_r = rng.integers(0, 9, size=(6, 10))
    assert res[0]["episodes"] == 20 and res[0]["steps"] == 60
    assert res[0]["return_sum"] == int(all_r.sum()) and res[0]["length_sum"] == 60
    assert res[0]["specials_created"] == 60 and res[0]["combination_matches"] == 20


def test_numa_helpers_parse_and_degrade():
    from tile_match_gym_b200 import numa
    assert numa._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert numa._parse_cpulist("") == set()
    assert numa.gpu_numa_node(0) is None or isinstance(numa.gpu_numa_node(0), int)   # no GPU here: None, never raises


def test_episode_statistics_next_step_skips_the_reset_step():
    """gymnasium's vector RecordEpisodeStatistics convention: under next_step autoreset the call after a termination only
    resets the env -- it is not a step of the new episode."""
    import torch
    from tile_match_gym_b200.vec_env import EpisodeStatistics
    st = EpisodeStatistics(2, "cpu", autoreset="next_step")
    zeros = {k: torch.zeros(2, dtype=torch.int32) for k in ("num_new_specials", "num_specials_activated", "shuffled", "is_combination_match")}
    seq = [([3, 1], [0, 0]), ([2, 0], [1, 0]), ([0, 4], [0, 1]), ([5, 0], [0, 0]), ([1, 2], [1, 0])]
    # env 0: episode of 2 steps (return 5), reset step, then an episode of 2 steps (5 + 1 = 6)
    # env 1: episode of 3 steps (return 5), reset step, one live step so far
    for r, t in seq:
        st.update(torch.tensor(r), torch.tensor(t, dtype=torch.bool), zeros)
    out = st.allreduce()
    assert out["episodes"] == 3 and out["return_sum"] == 5 + 5 + 6 and out["length_sum"] == 2 + 3 + 2
    assert out["steps"] == 10 - 2          # two of the ten env-steps were reset steps
