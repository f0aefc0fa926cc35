"""Throughput of tmg_step for the other BASELINE configs and a batch-size sweep (diagnostics, not the bench line).
Prints one JSON object per measurement.  Usage: python scripts/sweep.py [--quick]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tile_match_gym_b200 import TileMatchVecEnv

ALL = (["cookie"], ["vertical_laser", "horizontal_laser", "bomb"])


def no_line_boards(n, R, C, K, seed):
    """Constructive line-free boards (for shapes where generate_board does not terminate), a few base boards tiled."""
    rng = np.random.default_rng(seed)
    base = []
    for _ in range(64):
        b = np.zeros((R, C), dtype=np.int8)
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
        base.append(np.stack([b, np.ones_like(b)]))
    base = np.stack(base)
    return torch.from_numpy(base[np.arange(n) % len(base)].copy())


def measure(name, N, R, C, K, moves, cl, cs, steps=60, warmup=30, obs="int8", inject=False, policy="uniform"):
    env = TileMatchVecEnv(N, R, C, K, moves, cl, cs, seed=2, autoreset="disabled" if inject else "same_step", obs=obs)
    if inject:
        boards = no_line_boards(N, R, C, K, 5).cuda()
        env.reset(options={"init_boards": boards})
    else:
        env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    A = env.num_actions
    acts = [torch.randint(0, A, (N,), device="cuda", dtype=torch.int32, generator=g) for _ in range(8)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream()

    def one(i):
        if policy == "mask":
            m = env.mask.float() + 1e-6
            a = torch.multinomial(m, 1, generator=g)[:, 0].to(torch.int32)
        else:
            a = acts[i % 8]
        return env.step(a)

    for i in range(warmup):
        if inject and i % moves == 0 and i:
            env.reset(options={"init_boards": boards})
        one(i)
    env.join(); torch.cuda.synchronize()
    evs = []
    n_timed = 0
    for i in range(steps):   # same timing as bench.py: events around every launch, no host sync inside, side streams drained at the end
        if inject and (warmup + i) % moves == 0:
            env.reset(options={"init_boards": boards})
        if N * (2 * R * C + A) < (200 << 20):
            flush.zero_()
        if policy == "mask":
            m = env.mask.float() + 1e-6
            a = torch.multinomial(m, 1, generator=g)[:, 0].to(torch.int32)
        else:
            a = acts[i % 8]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); env.step(a); e1.record(st)
        evs.append((e0, e1)); n_timed += 1
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record(st); env.join(); d1.record(st); torch.cuda.synchronize()
    tot = sum(a.elapsed_time(b) for a, b in evs) + d0.elapsed_time(d1)
    bytes_step = 4 * R * C + 48 + A + (env.onehot_planes * R * C if obs == "onehot" else 0)
    v = N * n_timed / (tot * 1e-3)
    out = {"name": name, "envs": N, "shape": f"{R}x{C}", "colours": K, "num_moves": moves, "obs": obs, "policy": policy,
           "env_steps_per_s": v, "ms_per_step": tot / n_timed, "bytes_per_env_step": bytes_step,
           "achieved_GBps": v * bytes_step / 1e9, "status_flags": int((env.status != 0).sum().item())}
    print(json.dumps(out), flush=True)
    env.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser(); ap.add_argument("--quick", action="store_true"); a = ap.parse_args()
    measure("config2 10x10/4 all specials", 65536, 10, 10, 4, 30, *ALL)
    measure("config2, actions sampled from the mask", 65536, 10, 10, 4, 30, *ALL, policy="mask")
    measure("config1 shape, no specials", 65536, 10, 10, 4, 30, [], [])
    measure("config3 9x9/6 one-hot + mask, 131072 envs (1/8 of 1M)", 131072, 9, 9, 6, 30, *ALL, obs="onehot")
    measure("config5 32x32/7 injected boards", 8192, 32, 32, 7, 30, *ALL, inject=True)
    sizes = [1024, 4096, 16384, 65536, 262144, 1048576] + ([] if a.quick else [4194304, 16777216])
    for n in sizes:
        measure(f"sweep {n}", n, 10, 10, 4, 30, *ALL, steps=30 if n > 1000000 else 60, warmup=30)
