#!/usr/bin/env python
"""bench.py -- env-steps/sec of the tile-match board-transition hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one TileMatchEnv.step (one tmg_step call = k_gate + k_work) over the whole batch of synthetic envs: swap, effectiveness gate,
combination match, full cascade loop (detect / classify / activate / gravity / refill until stable),
playability repair, timer/termination, legal-move mask, and the autoreset (generate_board) of every env whose
episode ended.  Workload = BASELINE.json configs[1]: 10x10, 4 colours, cookie + v/h laser + bomb,
num_moves=30, 65536 envs per GPU (weak scaling), uniform random actions.  Episodes are synchronised, as they are
for the reference (fixed-length episodes that all start at reset()): every 30th step ends all episodes and every
env gets its next board in that step.  The default --steps 120 covers four whole episodes; --stagger env|pair
spreads the episode phases instead (1/30 of the envs reset in every step).

Prints ONE JSON line (rank 0).  `value` times tmg_step with inputs resident in HBM (CUDA events around each
launch, L2 flushed between steps); `e2e` times the host-buffer call tmg_step_host (actions from pinned host
memory; board, reward, terminated, bit-packed mask and num_moves_left complete in pinned host memory after every step,
the board and mask through the host mirror the kernel writes directly, stream synchronised per step); `rollout`
reports tmg_step_many beside them.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS, COLS, COLOURS, NUM_MOVES = 10, 10, 4, 30
CL, CS = ["cookie"], ["vertical_laser", "horizontal_laser", "bomb"]
ENVS_PER_GPU = 65536
SEED = 2
P = ROWS * COLS
A = 2 * P - ROWS - COLS
# algorithmic bytes per env-step = 4P + 48 + A = 628 B for 10x10 (int8 planes in+out, scalars, mask)
BYTES_PER_STEP = 4 * P + 48 + A
# dram__bytes_read.sum + dram__bytes_write.sum of one k_work launch (ncu --set full, profiles/r01e_k_work_ncu.txt):
# 5.84 MB read + 0.22 MB written (no-op steps never load their board; writes stay in the 126 MB L2 within a launch)
NCU_TRAFFIC_BYTES_PER_LAUNCH = 6.07e6
METRIC = "env-steps/sec (full cascade, bit-exact)"
UNIT = "env-steps/s"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.stop_flag = index, period, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def summary(self):
        self.stop_flag.set()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def cpu_port_throughput(threads, target_s=10.0):
    """The CPU statement of the same path (oracle/tmg_oracle.c, a literal C port of the reference's algorithm),
    looped on the host cores over a bounded sample of the same workload (same shape, specials, num_moves, same-step
    autoreset, uniform actions; ~target_s seconds of work).  Checker code used as a *baseline*, never as product."""
    from oracle.oracle import OracleVecEnv
    num_envs = max(threads * 256, 2048)
    o = OracleVecEnv(num_envs, ROWS, COLS, COLOURS, NUM_MOVES, CL, CS, seed=SEED, autoreset="same_step", num_threads=threads)
    o.reset()
    t0 = time.perf_counter()
    o.rollout(NUM_MOVES, 99, 0)                      # one whole episode: warm-up + calibration
    cal = time.perf_counter() - t0
    episodes = max(1, min(200, int(target_s / max(cal, 1e-3))))
    steps = episodes * NUM_MOVES
    t0 = time.perf_counter()
    o.rollout(steps, 99, NUM_MOVES)
    dt = time.perf_counter() - t0
    return num_envs * steps / dt, dt, num_envs, steps


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The reference is pure
    Python and its tree does not exist on the GPU box, so this arm times the C port of it (kind "port") with every
    host thread -- a much stronger baseline than the Python original (~1e3 steps/s/core)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # --steps / --warmup size the GPU arm; this arm times whole episodes of a bounded sample instead
    vals, samples = [], []
    for _ in range(2):
        v, dt, n, steps = cpu_port_throughput(threads, target_s=8.0)
        vals.append(v); samples.append((n, steps, dt))
    v = max(vals)
    n, per_call, _ = samples[vals.index(v)]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * ENVS_PER_GPU / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": workload_config(1),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n} envs x {per_call} steps (whole episodes) per timed call, best of {len(vals)} calls; "
                                       "oracle/tmg_oracle.c (C port of the Python reference) on all host threads"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "TileMatchEnv 10x10, 4 colours, specials=[vertical_laser,horizontal_laser,bomb,cookie], "
                        "65536 envs per GPU, num_moves=30, uniform random actions, same-step autoreset, synchronised episodes",
            "envs_per_gpu": ENVS_PER_GPU, "global_envs": ENVS_PER_GPU * n_gpus, "num_moves": NUM_MOVES,
            "parallelism": f"env-index sharding x{n_gpus}, no step-path collective",
            "l2": "flushed between timed steps (256 MiB memset); per-GPU state 29 MB is L2-resident otherwise",
            "refill": "philox4x32-10 counter stream"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=120)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--stagger", default="none", choices=["none", "env", "pair"],
                    help="episode phases: none = synchronised (reference behaviour), env = timer0 = env %% num_moves")
    ap.add_argument("--no-stagger", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--num-moves", type=int, default=NUM_MOVES, help="diagnostic: episode length (huge = no resets)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--step-stream-priority", type=int, default=0,
                    help="diagnostic: run the steps on a CUDA stream of this priority (-1 = above the library's side streams)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    from tile_match_gym_b200 import HostStepper, TileMatchVecEnv

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if args.step_stream_priority:
        torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=args.step_stream_priority))
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner (and anything else) to stdout by default
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    n_local = args.envs_per_gpu
    num_moves = args.num_moves
    env = TileMatchVecEnv(n_local, ROWS, COLS, COLOURS, num_moves, CL, CS, seed=SEED, device=dev, autoreset="same_step",
                          env_id_offset=rank * n_local)
    env.reset()
    if args.stagger != "none" and not args.no_stagger:
        ids = torch.arange(n_local, device=dev) + rank * n_local
        env.timer.copy_((ids // 2 if args.stagger == "pair" else ids) % num_moves)
        env.num_moves_left.copy_(num_moves - env.timer)
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    n_act = 16
    actions = [torch.randint(0, A, (n_local,), device=dev, dtype=torch.int32, generator=gen) for _ in range(n_act)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing: CUDA events around every tmg_step launch, L2 flushed in between -------------
    for i in range(args.warmup):
        env.step(actions[i % n_act])
    # rank 0 samples its own GPU's clocks (every rank spawning nvidia-smi in a loop would load the host cores that
    # the end-to-end path runs on)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    env.join()   # start from an empty side stream so that the timed region owns all of its board generations
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    drain = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        if not args.no_flush:
            flush.zero_()
        ev[i][0].record(stream)
        env.step(actions[i % n_act])
        ev[i][1].record(stream)
    drain[0].record(stream)
    env.join()   # board generations still running beside the steps belong to the timed work
    drain[1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    drain_ms = drain[0].elapsed_time(drain[1])
    total_ms = sum(step_ms) + drain_ms
    status_bad = int((env.status != 0).sum().item())

    # ---- fused rollout (tmg_step_many): the same env-steps, num_moves of them per launch ---------------------------
    T = min(num_moves, 30)
    n_win = max(2, args.steps // T)
    ro_actions = [torch.randint(0, A, (T, n_local), device=dev, dtype=torch.int32, generator=gen) for _ in range(2)]
    env.step_many(ro_actions[0]); env.join()
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(stream)
    for i in range(n_win):
        if not args.no_flush:
            flush.zero_()
        env.step_many(ro_actions[i % 2])
    env.join()
    r1.record(stream)
    barrier()
    rollout_ms = r0.elapsed_time(r1)
    # the agent inside the kernel: samples from the effective actions (every step is a move: ~4x the cascade work)
    env.rollout(T, "mask"); env.join()
    barrier()
    r0.record(stream)
    for i in range(n_win):
        env.rollout(T, "mask")
    env.join()
    r1.record(stream)
    barrier()
    rollout_mask_ms = r0.elapsed_time(r1)
    status_bad += int((env.status != 0).sum().item())

    # ---- end-to-end timing through the host-buffer call (pinned host memory in and out) -----------------------
    def time_host_path(outputs, mirror=False):
        hs = HostStepper(env, outputs=outputs, mirror=mirror)
        for i in range(3):
            hs.io.actions = host_actions[i % n_act].data_ptr()
            hs.step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(e2e_steps):
            hs.io.actions = host_actions[i % n_act].data_ptr()
            hs.step()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if mirror:   # bytes the kernel wrote over PCIe: counted on one further, untimed episode
            changed = 0
            for i in range(num_moves if num_moves < 1000 else 30):
                hs.io.actions = host_actions[i % n_act].data_ptr()
                out = hs.step()
                changed += int(((out["reward"] > 0) | (out["terminated"] != 0)).sum())
            hs.avg_changed = changed / (num_moves if num_moves < 1000 else 30)
            hs.d2h_bytes = int(hs.d2h_bytes_fixed + hs.avg_changed * hs.d2h_bytes_per_changed_env)
            hs.close()
        return hs, ms

    host_actions = [a.cpu().pin_memory() for a in actions]
    e2e_steps = 3 if args.skip_e2e else max(10, min(args.steps, 60))
    # the full result of TileMatchEnv.step for every env: board, reward, terminated, legal-move mask, num_moves_left.
    # Headline form: mask as bits (same information as the reference's effective_actions list, 23 B instead of 180 B
    # per env over PCIe); the byte-mask form is reported beside it.
    # Headline form: board and bit-packed mask bound as the host mirror (tmg_host_bind) -- the step kernel writes the
    # entries of the envs it changed straight into the pinned arrays, the scalars come back by copy.  The full-copy
    # forms (every array copied after every step) are reported beside it.
    hs, e2e_ms = time_host_path(("board", "reward", "terminated", "mask_bits", "num_moves_left"), mirror=True)
    hs_full, e2e_full_ms = time_host_path(("board", "reward", "terminated", "mask_bits", "num_moves_left"))
    hs_bytes, e2e_bytes_ms = time_host_path(("board", "reward", "terminated", "mask", "num_moves_left"))
    clocks = sampler.summary() if rank == 0 else None

    # max over ranks
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms, e2e_bytes_ms, e2e_full_ms, rollout_ms, rollout_mask_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, e2e_bytes_ms, e2e_full_ms, rollout_ms, rollout_mask_ms = t.tolist()
        bad = torch.tensor([status_bad], device=dev); dist.all_reduce(bad); status_bad = int(bad.item())
    n_global = n_local * world
    value = n_global * args.steps / (total_ms * 1e-3)
    e2e_value = n_global * e2e_steps / (e2e_ms * 1e-3)
    peak, peak_src = measured_peak_gbs()
    kernel_ms = total_ms / args.steps
    achieved = BYTES_PER_STEP * n_local / (kernel_ms * 1e-3) / 1e9
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
            "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": hs.h2d_bytes * world,
                    "d2h_bytes_per_step": hs.d2h_bytes * world, "steps": e2e_steps,
                    "returns": "board,reward,terminated,mask(bit-packed),num_moves_left in pinned host memory, complete and "
                               "current after every step, stream synchronised per step; board and mask are a host mirror "
                               "(tmg_host_bind) that the step kernel updates in place for the envs it changed "
                               f"({hs.avg_changed / n_local:.3f} of the envs per step), scalars are copied in full",
                    "full_copy_every_step": {"value": n_global * e2e_steps / (e2e_full_ms * 1e-3),
                                             "d2h_bytes_per_step": hs_full.d2h_bytes * world},
                    "full_copy_byte_mask": {"value": n_global * e2e_steps / (e2e_bytes_ms * 1e-3),
                                            "d2h_bytes_per_step": hs_bytes.d2h_bytes * world}},
            # k_gate + k_work per step, plus one k_pregen per step on a side stream
            "gpu_launches": 3 * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH, "peak_source": peak_src, "kernel": "tmg::k_gate + tmg::k_work<32,10,10> (one tmg_step)",
                         "bytes_per_env_step": BYTES_PER_STEP, "envs_per_launch": n_local,
                         "note": "integer/divergence-bound kernel: the HBM fraction is low by construction, see DESIGN.md"},
            "rollout": {"value": n_global * T * n_win / (rollout_ms * 1e-3), "unit": UNIT, "steps_per_launch": T, "launches": n_win,
                        "what": "tmg_step_many: the same env-steps with the actions of a whole window given up front "
                                "(random-agent loop), boards kept on chip between steps; not the headline",
                        "mask_policy_in_kernel": {"value": n_global * T * n_win / (rollout_mask_ms * 1e-3),
                                                  "what": "tmg_rollout_policy(TMG_POLICY_MASK): actions sampled from the legal-move "
                                                          "mask inside the kernel, every step an effective move"}},
            "clocks": clocks,
            "step_ms": {"min": min(step_ms), "median": statistics.median(step_ms), "max": max(step_ms)},
            "drain_ms": drain_ms, "wall_s": t_wall, "status_flags_set": status_bad,
        }
        if not args.no_cpu_baseline and world == 1:   # reported at N=1 only
            threads = os.cpu_count() or 1
            v, dt, n, steps = cpu_port_throughput(threads, target_s=10.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{n} envs x {steps} steps of the same workload ({dt:.1f} s), oracle/tmg_oracle.c "
                                              "(C port of the Python reference; the Python original runs ~1e3 steps/s/core) "
                                              "on all host threads"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
