"""Diagnostics: where the end-to-end step time goes (host-mirror path).  Prints per-step times of
(a) tmg_step_host with the mirror, synchronised per step (= bench.py's e2e), (b) tmg_step with the mirror bound and
device actions, one sync at the end (GPU time incl. the PCIe writes), (c) the same without a mirror."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tile_match_gym_b200 import HostStepper, TileMatchVecEnv

N, STEPS = 65536, 90
env = TileMatchVecEnv(N, 10, 10, 4, 30, ["cookie"], ["vertical_laser", "horizontal_laser", "bomb"], seed=2, autoreset="same_step")
env.reset()
if "--sync-episodes" not in sys.argv:      # staggered episode phases, as in bench.py
    env.timer.copy_(torch.arange(N, device="cuda") % 30)
    env.num_moves_left.copy_(30 - env.timer)
g = torch.Generator(device="cuda"); g.manual_seed(0)
acts = [torch.randint(0, env.num_actions, (N,), device="cuda", dtype=torch.int32, generator=g) for _ in range(16)]
hacts = [a.cpu().pin_memory() for a in acts]


def device_loop(label):
    for i in range(10):
        env.step(acts[i % 16])
    env.join(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(STEPS):
        env.step(acts[i % 16])
    env.join(); torch.cuda.synchronize()
    print(f"{label}: {(time.perf_counter() - t0) / STEPS * 1e3:.3f} ms/step")


device_loop("(c) tmg_step, no mirror, no per-step sync")
hs = HostStepper(env, outputs=("board_packed", "reward", "terminated", "mask_bits", "num_moves_left"), mirror=True)
device_loop("(b) tmg_step, mirror bound, no per-step sync")
for i in range(10):
    hs.io.actions = hacts[i % 16].data_ptr(); hs.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
t_call = 0.0
for i in range(STEPS):
    hs.io.actions = hacts[i % 16].data_ptr()
    hs.step()
dt = (time.perf_counter() - t0) / STEPS
print(f"(a) tmg_step_host, mirror, sync per step: {dt * 1e3:.3f} ms/step")
# (d) the same call with CUDA events around it: GPU time from the call reaching the stream to the end of the step kernels
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(STEPS)]
st = torch.cuda.current_stream()
t_in = 0.0
for i in range(STEPS):
    hs.io.actions = hacts[i % 16].data_ptr()
    ev[i][0].record(st)
    t1 = time.perf_counter()
    hs.step()
    t_in += time.perf_counter() - t1
    ev[i][1].record(st)
torch.cuda.synchronize()
gpu = sum(a.elapsed_time(b) for a, b in ev) / STEPS
print(f"(d) inside tmg_step_host (wall) {t_in / STEPS * 1e3:.3f} ms/step, GPU span of the call {gpu:.3f} ms/step")
hs.close()
# (e) actions staged by an async copy instead of read in place, outputs copied instead of mirrored
hs2 = HostStepper(env, outputs=("reward", "terminated", "num_moves_left"), mirror=False)
for i in range(10):
    hs2.io.actions = hacts[i % 16].data_ptr(); hs2.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(STEPS):
    hs2.io.actions = hacts[i % 16].data_ptr(); hs2.step()
print(f"(e) tmg_step_host, no mirror, scalars only by copy: {(time.perf_counter() - t0) / STEPS * 1e3:.3f} ms/step")
