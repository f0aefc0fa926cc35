"""Diagnostics: per-env SM cycles of one tmg_step launch (tmg_set_profile_buffer), to see what the slowest
warps of a launch are doing.  Usage: python scripts/profile_env_cycles.py [--num-moves M] [--envs N]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tile_match_gym_b200 import TileMatchVecEnv

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--num-moves", type=int, default=30)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--stagger", default="env")
a = ap.parse_args()
env = TileMatchVecEnv(a.envs, 10, 10, 4, a.num_moves, ["cookie"], ["vertical_laser", "horizontal_laser", "bomb"], seed=2,
                      autoreset="same_step")
env.reset()
if a.stagger == "env":
    env.timer.copy_(torch.arange(a.envs, device="cuda") % a.num_moves)
elif a.stagger == "pair":
    env.timer.copy_((torch.arange(a.envs, device="cuda") // 2) % a.num_moves)
prof = torch.zeros((a.envs, 8), dtype=torch.int32, device="cuda")
env._lib.tmg_set_profile_buffer(env._h, C.c_void_p(prof.data_ptr()))
g = torch.Generator(device="cuda"); g.manual_seed(0)
for t in range(a.steps):
    act = torch.randint(0, env.num_actions, (a.envs,), device="cuda", dtype=torch.int32, generator=g)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof.zero_()
    e0.record(); env.step(act); e1.record(); torch.cuda.synchronize()
    if t >= a.steps - 3:
        pr = prof.cpu().numpy().astype("int64")
        cyc, ser, rounds, iters = pr[:, 0], pr[:, 1], pr[:, 2], pr[:, 3]
        import numpy as np
        order = np.argsort(-cyc)[:8]
        print(f"step {t}: kernel {e0.elapsed_time(e1)*1e3:.0f} us; env cycles: mean {cyc.mean():.0f} p50 {np.percentile(cyc,50):.0f} "
              f"p99 {np.percentile(cyc,99):.0f} p99.9 {np.percentile(cyc,99.9):.0f} max {cyc.max()} "
              f"(= {cyc.max()/1.965e3:.0f} us at 1.965 GHz); sum {cyc.sum()/1e6:.0f} Mcyc")
        act = env.num_specials_activated.cpu().numpy(); new = env.num_new_specials.cpu().numpy(); rew = env.reward.cpu().numpy()
        for i in order:
            print(f"   env {i}: cycles {cyc[i]} general rounds (byte planes: slow-path cycles) {ser[i]} rounds {rounds[i]} redraws {iters[i]} "
                  f"activated {act[i]} new_specials {new[i]} reward {rew[i]} | cycles: scan+fast round {pr[i,4]} general path {pr[i,5]} "
                  f"fall+refill {pr[i,6]} (byte planes: scan / table / classify / resolve {pr[i,4]} {pr[i,5]} {pr[i,6]} {pr[i,7]})")
        eff = rounds > 0
        if eff.any():
            print(f"   effective moves {eff.sum()}: cycles/round {cyc[eff & (iters == 0)].sum() / max(1, rounds[eff & (iters == 0)].sum()):.0f}, "
                  f"serial share {ser[eff].sum() / cyc[eff].sum():.2f}")
        rs = iters > 0
        if rs.any():
            print(f"   resets {rs.sum()}: redraws mean {iters[rs].mean():.0f} max {iters[rs].max()}; cycles/redraw {cyc[rs].sum() / iters[rs].sum():.0f}")
