"""BASELINE config 3 as written: 9x9, 6 colours, all specials, legal-move mask + one-hot observation, 1 048 576 envs
sharded by env index over the GPUs of one box (TileMatchVecEnv.sharded: rank g owns [g*N/G, (g+1)*N/G), the draw
stream of an env depends on its global id only), no collective on the step path, episode statistics all-reduced over
NCCL at the end (the design's only collective).  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      scripts/config3_sharded.py [--envs 1048576] [--steps 60] [--warmup 30]
(without torchrun it runs the whole batch on one GPU).  Rank 0 prints one JSON line; timing = CUDA events around
every step (tmg_step + one-hot encode), L2 flushed between steps, barrier on both sides, max over ranks."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from tile_match_gym_b200 import EpisodeStatistics, TileMatchVecEnv

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--warmup", type=int, default=30)
a = ap.parse_args()

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
R, C, K, MOVES = 9, 9, 6, 30
env = TileMatchVecEnv.sharded(a.envs, rank, world, R, C, K, MOVES, ["cookie"], ["vertical_laser", "horizontal_laser", "bomb"],
                              seed=2, device=dev, autoreset="same_step", obs="onehot")
N = env.num_envs
stats = EpisodeStatistics(N, dev)
env.reset()
g = torch.Generator(device=dev); g.manual_seed(1000 + rank)
acts = [torch.randint(0, env.num_actions, (N,), device=dev, dtype=torch.int32, generator=g) for _ in range(8)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for i in range(a.warmup):
    env.step(acts[i % 8])
env.join()
barrier()
evs = []
for i in range(a.steps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    obs, rew, term, trunc, info = env.step(acts[i % 8])       # tmg_step + tmg_encode_onehot
    e1.record(st)
    evs.append((e0, e1))
    stats.update(rew, term, info)                              # off the timed path (elementwise torch ops)
d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
d0.record(st); env.join(); d1.record(st)
barrier()
ms = sum(x.elapsed_time(y) for x, y in evs) + d0.elapsed_time(d1)
t = torch.tensor([ms], device=dev, dtype=torch.float64)
bad = torch.tensor([int((env.status != 0).sum().item())], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(bad)
tot = stats.allreduce()                                        # NCCL all-reduce of eight int64 counters
assert obs["board"].shape == (N, env.onehot_planes, R, C) and info["effective_actions"].shape == (N, env.num_actions)
if rank == 0:
    ms = float(t.item())
    A = env.num_actions
    bytes_step = 4 * R * C + 48 + A + env.onehot_planes * R * C          # 1 326 B
    v = a.envs * a.steps / (ms * 1e-3)
    print(json.dumps({"name": "config3 9x9/6 all specials, one-hot + mask, envs sharded by index", "global_envs": a.envs,
                      "n_gpus": world, "envs_per_gpu": N, "steps": a.steps, "warmup": a.warmup, "env_steps_per_s": v,
                      "ms_per_step": ms / a.steps, "bytes_per_env_step": bytes_step,
                      "achieved_GBps_per_gpu": v * bytes_step / 1e9 / world, "status_flags": int(bad.item()),
                      "episode_statistics_allreduced": tot, "step_path_collectives": 0,
                      "l2": "flushed between steps (256 MiB memset)"}), flush=True)
env.close()
if world > 1:
    dist.destroy_process_group()
