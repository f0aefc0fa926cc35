#!/usr/bin/env python
"""bench.py -- env-steps/sec of the tile-match board-transition hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|5]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one TileMatchEnv.step (one tmg_step call = k_gate + k_work) over the whole batch of synthetic envs: swap,
effectiveness gate, combination match, full cascade loop (detect / classify / activate / gravity / refill until
stable), playability repair, timer/termination, legal-move mask, and the autoreset (generate_board) of every env
whose episode ended.  Workloads (BASELINE.json `configs`):
  --config 2 (default, the config the metric is quoted on): 10x10, 4 colours, cookie + v/h laser + bomb, num_moves=30,
             65536 envs per GPU (weak scaling), uniform random actions.
  --config 3: 9x9, 6 colours, all specials, legal-move mask + one-hot observation every step, 131072 envs per GPU
             (1M envs over 8 GPUs).
  --config 5: 32x32, 7 colours, all specials, 8192 envs per GPU; the reference's generate_board does not terminate for
             this shape, so every board comes from the on-device constructive line-free sampler
             (TMG_FLAG_CONSTRUCTIVE_RESET; the CPU baseline runs the same contract) -- no host-side board injection.
Episode phases are STAGGERED (env e starts at move e mod num_moves, as in any long-running RL loop), so every timed
step carries exactly 1/num_moves of the batch's episode ends and board generations whatever --steps is; the line
reports the episode ends counted inside the timed region and, beside `value`, the reset-free figure.

Prints ONE JSON line (rank 0).  `value` times tmg_step with inputs resident in HBM (CUDA events around each launch on
the launching stream, L2 flushed between steps, board generations still in flight at the end drained inside the timed
total); `e2e` times the host-buffer call tmg_step_host (actions from pinned host memory; board, reward, terminated,
bit-packed mask and num_moves_left complete in pinned host memory after every step, stream synchronised per step);
`rollout` reports tmg_step_many beside them.
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

ALL_CL, ALL_CS = ["cookie"], ["vertical_laser", "horizontal_laser", "bomb"]
SEED = 2
METRIC = "env-steps/sec (full cascade, bit-exact)"
UNIT = "env-steps/s"

# BASELINE.json configs that fit one GPU; bytes = SURVEY.md 8(d): 4P + 48 + A (+ (K+S)P with the one-hot observation)
CONFIGS = {
    2: dict(rows=10, cols=10, colours=4, envs=65536, onehot=False, inject=False, constructive=False, num_moves=30,
            name="TileMatchEnv 10x10, 4 colours, specials=[vertical_laser,horizontal_laser,bomb,cookie]"),
    3: dict(rows=9, cols=9, colours=6, envs=131072, onehot=True, inject=False, constructive=False, num_moves=30,
            name="9x9, 6 colours, all specials, legal-move mask + one-hot obs (1M envs over 8 GPUs = 131072 per GPU)"),
    5: dict(rows=32, cols=32, colours=7, envs=8192, onehot=False, inject=False, constructive=True, num_moves=30,
            name="large-board stress 32x32, 7 colours, all specials; boards from the on-device constructive line-free sampler "
                 "(TMG_FLAG_CONSTRUCTIVE_RESET, not the reference's generate_board, which never returns for this shape)"),
}
# dram__bytes_read.sum + dram__bytes_write.sum of one k_work launch of config 2 (ncu --set full, profiles/): no-op steps
# never load their board and writes stay in the 126 MB L2 within a launch
NCU_TRAFFIC_BYTES_PER_LAUNCH = {2: 8.0e6}   # profiles/r02_kernels_ncu.txt: 7.87 MB read + 0.13 MB written


def workload(cfg_id):
    c = dict(CONFIGS[cfg_id])
    P = c["rows"] * c["cols"]
    c["P"], c["A"] = P, 2 * P - c["rows"] - c["cols"]
    c["planes"] = c["colours"] + 4
    c["bytes_per_step"] = 4 * P + 48 + c["A"] + (c["planes"] * P if c["onehot"] else 0)
    return c


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


def no_line_boards(n, R, C, K, seed):
    """Constructive line-free boards for shapes where generate_board does not terminate (a few base boards, tiled)."""
    import numpy as np
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
    return base[np.arange(n) % len(base)].copy()


def cpu_port_throughput(wl, threads, target_s=10.0):
    """The CPU statement of the same path (oracle/tmg_oracle.c, a literal C port of the reference's algorithm),
    looped on the host cores over a bounded sample of the same workload (same shape, specials, num_moves, same-step
    autoreset, uniform actions; ~target_s seconds of work).  Checker code used as a *baseline*, never as product."""
    from oracle.oracle import OracleVecEnv
    moves = wl["num_moves"]
    if wl["inject"]:
        import numpy as np
        num_envs = max(threads * 8, 64)
        boards = no_line_boards(num_envs, wl["rows"], wl["cols"], wl["colours"], 5)
        o = OracleVecEnv(num_envs, wl["rows"], wl["cols"], wl["colours"], moves, ALL_CL, ALL_CS, seed=SEED,
                         autoreset="disabled", num_threads=threads)
        t0 = time.perf_counter()
        steps = 0
        while time.perf_counter() - t0 < target_s or steps == 0:
            o.reset(init_boards=boards)
            o.rollout(moves, 99, steps)
            steps += moves
        dt = time.perf_counter() - t0
        return num_envs * steps / dt, dt, num_envs, steps
    num_envs = max(threads * 256, 2048)
    if wl["constructive"]:
        num_envs = max(threads * 16, 128)
    o = OracleVecEnv(num_envs, wl["rows"], wl["cols"], wl["colours"], moves, ALL_CL, ALL_CS, seed=SEED, autoreset="same_step",
                     num_threads=threads, constructive_reset=wl["constructive"])
    o.reset()
    t0 = time.perf_counter()
    o.rollout(moves, 99, 0)                          # one whole episode: warm-up + calibration
    cal = time.perf_counter() - t0
    episodes = max(1, min(200, int(target_s / max(cal, 1e-3))))
    steps = episodes * moves
    t0 = time.perf_counter()
    o.rollout(steps, 99, moves)
    dt = time.perf_counter() - t0
    return num_envs * steps / dt, dt, num_envs, steps


def _python_reference_worker(args):
    """One process looping the UNMODIFIED Python reference env (tile_match_env.py:84-112) for ~target_s seconds."""
    wl, seed, target_s = args
    import numpy as np
    from oracle.ref_loader import load_reference
    ref = load_reference()
    env = ref.TileMatchEnv(wl["rows"], wl["cols"], wl["colours"], wl["num_moves"], ALL_CL, ALL_CS, seed=seed)
    env.reset()
    rng = np.random.default_rng(seed)
    for _ in range(40):                              # numba JIT warm-up
        _, _, done, _, _ = env.step(int(rng.integers(0, env.num_actions)))
        if done:
            env.reset()
    env.reset()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < target_s:
        for _ in range(50):
            _, _, done, _, _ = env.step(int(rng.integers(0, env.num_actions)))
            n += 1
            if done:
                env.reset()
    return n, time.perf_counter() - t0


def python_reference_throughput(wl, target_s=8.0):
    """BASELINE.md section 3 legs (i) and (ii): the Python reference env looped in one process and in os.cpu_count()
    processes (what gymnasium's AsyncVectorEnv does, minus the pipes).  Only possible where the reference tree exists
    (the build container; it is absent on the GPU box) -- returns a one-line reason otherwise."""
    try:
        from oracle.ref_loader import reference_available
        if wl["inject"] or wl["constructive"]:
            return {"unavailable": "generate_board does not terminate for this shape in the reference (SURVEY 0.7)"}
        if not reference_available():
            return {"unavailable": "reference tree not present on this machine (it exists only in the build container); "
                                   "profiles/ holds the run made there"}
        import multiprocessing as mp
        n1, dt1 = _python_reference_worker((wl, SEED, target_s))
        procs = os.cpu_count() or 1
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_python_reference_worker, [(wl, SEED + 1 + i, target_s) for i in range(procs)])
        agg = sum(n / dt for n, dt in res)
        import importlib.util
        gm = sys.modules.get("gymnasium")
        real_gym = (gm is not None and getattr(gm, "__file__", None)) or (gm is None and importlib.util.find_spec("gymnasium") is not None)
        gym_note = ("gymnasium importable: AsyncVectorEnv leg not run by this script" if real_gym else
                    "gymnasium absent from the image: the AsyncVectorEnv leg is substituted by the multiprocessing leg (what "
                    "AsyncVectorEnv does, minus the pipes)")
        return {"single_process": n1 / dt1, "multiprocess": agg, "processes": procs, "unit": UNIT,
                "sample": f"~{target_s:.0f} s per process, uniform actions, reset on done", "async_vector_env": gym_note}
    except Exception as ex:  # noqa: BLE001
        return {"unavailable": f"{type(ex).__name__}: {ex}"}


def workload_config(wl, cfg_id, n_gpus, envs_per_gpu, stagger):
    phases = {"env": "staggered: env e starts at move e mod num_moves, 1/num_moves of the envs end an episode in every step",
              "none": "synchronised: every env starts at reset(), all episodes end in the same step",
              "pair": "staggered in pairs"}[stagger]
    if wl["inject"]:
        phases = "synchronised (every episode starts from injected boards: tmg_reset inside the timed region every num_moves steps)"
    return {"workload": f"{wl['name']}, {envs_per_gpu} envs per GPU, num_moves={wl['num_moves']}, uniform random actions, "
                        + ("episodes restart from injected boards" if wl["inject"] else "same-step autoreset"),
            "baseline_config": cfg_id, "envs_per_gpu": envs_per_gpu, "global_envs": envs_per_gpu * n_gpus,
            "num_moves": wl["num_moves"], "episode_phases": phases,
            "parallelism": f"env-index sharding x{n_gpus}, no step-path collective",
            "l2": "flushed between timed steps (256 MiB memset)",
            "refill": "philox4x32-10 counter stream"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The reference is pure
    Python and its tree does not exist on the GPU box, so this arm times the C port of it (kind "port") with every
    host thread -- a much stronger baseline than the Python original (~1e3 steps/s/core); where the
    reference tree is present the Python original is timed beside it (`python_reference`)."""
    if rank != 0:
        return
    wl = workload(args.config)
    threads = os.cpu_count() or 1
    # --steps / --warmup size the GPU arm; this arm times whole episodes of a bounded sample instead
    vals, samples = [], []
    for _ in range(2):
        v, dt, n, steps = cpu_port_throughput(wl, threads, target_s=8.0)
        vals.append(v); samples.append((n, steps, dt))
    v = max(vals)
    n, per_call, _ = samples[vals.index(v)]
    envs_per_gpu = args.envs_per_gpu or wl["envs"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * envs_per_gpu * world / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": workload_config(wl, args.config, world, envs_per_gpu, args.stagger),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "per_core": v / threads,
                             "sample": f"{n} envs x {per_call} steps (whole episodes) per timed call, best of {len(vals)} calls; "
                                       "oracle/tmg_oracle.c (C port of the Python reference) on all host threads; the host "
                                       "cores are one shared pool, so the figure does not grow with n_gpus",
                             "python_reference": python_reference_throughput(wl)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=120)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--envs-per-gpu", type=int, default=0, help="override the config's batch (batch-size sweeps)")
    ap.add_argument("--stagger", default="env", choices=["none", "env", "pair"],
                    help="episode phases: env = timer0 = env %% num_moves (default), none = synchronised")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--num-moves", type=int, default=0, help="diagnostic: episode length (huge = no resets)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-rollout", action="store_true")
    ap.add_argument("--skip-no-reset", action="store_true")
    ap.add_argument("--step-stream-priority", type=int, default=0,
                    help="diagnostic: run the steps on a CUDA stream of this priority (-1 = above the library's side stream)")
    ap.add_argument("--byte-planes", action="store_true", help="diagnostic: TMG_FLAG_BYTE_PLANES (shared-memory byte-plane engine)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    wl = workload(args.config)
    if args.num_moves:
        wl["num_moves"] = args.num_moves

    import numpy as np
    import torch
    import torch.distributed as dist

    from tile_match_gym_b200 import HostStepper, TileMatchVecEnv

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from tile_match_gym_b200.numa import bind_to_gpu_node
    numa_node = bind_to_gpu_node(local_rank)        # pinned host arrays and the driving thread on the GPU's NUMA node
    if args.step_stream_priority:
        torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=args.step_stream_priority))
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner (and anything else) to stdout by default
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    n_local = args.envs_per_gpu or wl["envs"]
    num_moves = wl["num_moves"]
    R, Cc, K, A = wl["rows"], wl["cols"], wl["colours"], wl["A"]
    inject = wl["inject"]
    stagger = "none" if inject else args.stagger

    def make_env(moves):
        env = TileMatchVecEnv(n_local, R, Cc, K, moves, ALL_CL, ALL_CS, seed=SEED, device=dev,
                              autoreset="disabled" if inject else "same_step", env_id_offset=rank * n_local,
                              obs="onehot" if wl["onehot"] else "int8", byte_planes=args.byte_planes,
                              constructive_reset=wl["constructive"])
        if inject:
            env.reset(options={"init_boards": init_boards})
        else:
            env.reset()
        if stagger != "none":
            ids = torch.arange(n_local, device=dev) + rank * n_local
            env.timer.copy_((ids // 2 if stagger == "pair" else ids) % moves)
            env.num_moves_left.copy_(moves - env.timer)
        return env

    init_boards = torch.from_numpy(no_line_boards(n_local, R, Cc, K, 5 + rank)).to(dev) if inject else None
    env = make_env(num_moves)
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

    def timed_steps(env, n_steps, n_warm, moves):
        """CUDA events around every tmg_step call, L2 flushed in between; returns per-step ms, drain ms, episode ends."""
        step_no = 0
        for i in range(n_warm):
            if inject and step_no and step_no % moves == 0:
                env.reset(options={"init_boards": init_boards})
            env.step(actions[i % n_act]); step_no += 1
        env.join()   # start from an empty side stream so that the timed region owns all of its board generations
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        drain = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ends = torch.zeros((), dtype=torch.int64, device=dev)
        for i in range(n_steps):
            if not args.no_flush:
                flush.zero_()
            ev[i][0].record(stream)
            if inject and step_no % moves == 0:
                env.reset(options={"init_boards": init_boards})   # the episode restart is part of the timed work
            env.step(actions[i % n_act]); step_no += 1
            ev[i][1].record(stream)
            ends += env.terminated.sum()                          # (outside the timed brackets)
        drain[0].record(stream)
        env.join()   # board generations still running beside the steps belong to the timed work
        drain[1].record(stream)
        barrier()
        return [a.elapsed_time(b) for a, b in ev], drain[0].elapsed_time(drain[1]), int(ends.item())

    # ---- device-resident timing ------------------------------------------------------------------------------------------
    # rank 0 samples its own GPU's clocks (every rank spawning nvidia-smi in a loop would load the host cores that
    # the end-to-end path runs on)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    step_ms, drain_ms, episode_ends = timed_steps(env, args.steps, args.warmup, num_moves)
    t_wall = time.perf_counter() - t_wall0
    total_ms = sum(step_ms) + drain_ms
    status_bad = int((env.status != 0).sum().item())

    # the same steps with no episode end inside the run ("with and without autoreset cost")
    no_reset_ms = None
    if not inject and not args.num_moves and not args.skip_no_reset:
        env_nr = make_env(1 << 20)
        s_nr = min(args.steps, 40)
        ms_nr, _, _ = timed_steps(env_nr, s_nr, max(5, min(args.warmup, 20)), 1 << 20)
        no_reset_ms = sum(ms_nr) / s_nr
        env_nr.close()
        del env_nr

    # ---- fused rollout (tmg_step_many): the same env-steps, num_moves of them per launch ---------------------------
    rollout_ms = rollout_mask_ms = None
    T = min(num_moves, 30)
    n_win = max(2, args.steps // T)
    if not args.skip_rollout and not inject:
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
        step_no = [0]

        def one(i):
            if inject and step_no[0] % num_moves == 0:
                env.reset(options={"init_boards": init_boards})
            hs.io.actions = host_actions[i % n_act].data_ptr()
            step_no[0] += 1
            return hs.step()

        if inject:
            step_no[0] = 0
        for i in range(3):
            one(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(e2e_steps):
            one(i)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if mirror:   # bytes the kernel wrote over PCIe: counted on one further, untimed episode
            changed = 0
            n_count = num_moves if num_moves < 1000 else 30
            for i in range(n_count):
                out = one(i)
                changed += int(((out["reward"] > 0) | (out["terminated"] != 0)).sum())
            hs.avg_changed = changed / n_count
            hs.d2h_bytes = int(hs.d2h_bytes_fixed + hs.avg_changed * hs.d2h_bytes_per_changed_env)
        hs.close()
        return hs, ms

    host_actions = [a.cpu().pin_memory() for a in actions]
    e2e_steps = 3 if args.skip_e2e else max(10, min(args.steps, 60))
    # the full result of TileMatchEnv.step for every env: board, reward, terminated, legal-move mask, num_moves_left.
    # Headline form: board and bit-packed mask (the same information as the reference's effective_actions list, 23 B
    # instead of 180 B per env over PCIe) bound as the host mirror (tmg_host_bind) -- the step kernel writes the entries
    # of the envs it changed straight into the pinned arrays.  The full-copy forms are reported beside it.
    if inject:
        env.reset(options={"init_boards": init_boards})
    board_out = "board_packed" if K <= 15 else "board"     # one byte per cell (colour | type << 4): half the PCIe bytes
    hs, e2e_ms = time_host_path((board_out, "reward", "terminated", "mask_bits", "num_moves_left"), mirror=True)
    if args.skip_e2e:
        hs_planes, e2e_planes_ms, hs_full, e2e_full_ms, hs_bytes, e2e_bytes_ms = hs, e2e_ms, hs, e2e_ms, hs, e2e_ms
    else:
        hs_planes, e2e_planes_ms = time_host_path(("board", "reward", "terminated", "mask_bits", "num_moves_left"), mirror=True)
        hs_full, e2e_full_ms = time_host_path(("board", "reward", "terminated", "mask_bits", "num_moves_left"))
        hs_bytes, e2e_bytes_ms = time_host_path(("board", "reward", "terminated", "mask", "num_moves_left"))
    clocks = sampler.summary() if rank == 0 else None

    # max over ranks
    vals = [total_ms, e2e_ms, e2e_bytes_ms, e2e_full_ms, rollout_ms or 0.0, rollout_mask_ms or 0.0, no_reset_ms or 0.0, e2e_planes_ms]
    if world > 1:
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = t.tolist()
        cnt = torch.tensor([status_bad, episode_ends], device=dev); dist.all_reduce(cnt)
        status_bad, episode_ends = int(cnt[0].item()), int(cnt[1].item())
    total_ms, e2e_ms, e2e_bytes_ms, e2e_full_ms, rollout_ms_, rollout_mask_ms_, no_reset_ms_, e2e_planes_ms = vals
    n_global = n_local * world
    value = n_global * args.steps / (total_ms * 1e-3)
    e2e_value = n_global * e2e_steps / (e2e_ms * 1e-3)
    peak, peak_src = measured_peak_gbs()
    kernel_ms = total_ms / args.steps
    achieved = wl["bytes_per_step"] * n_local / (kernel_ms * 1e-3) / 1e9
    if rank == 0:
        lib_every = max(1, min(4, num_moves // 4)) if num_moves >= 8 else 1
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
            "data": "synthetic", "config": workload_config(wl, args.config, world, n_local, stagger),
            "episode_ends_in_timed_region": episode_ends,
            "episode_ends_expected": n_global * args.steps / num_moves,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": hs.h2d_bytes * world,
                    "d2h_bytes_per_step": hs.d2h_bytes * world, "steps": e2e_steps,
                    "returns": "board (one byte per cell: colour | type << 4),reward,terminated,mask(bit-packed),num_moves_left in "
                               "pinned host memory, complete and current after every step, the call returns when the stream has "
                               "finished; board and mask are a host mirror (tmg_host_bind) that the step kernel updates in place "
                               f"for the envs it changed ({hs.avg_changed / n_local:.3f} of the envs per step), the scalars of "
                               "every env are written by the gate kernel",
                    "board_as_byte_planes": {"value": n_global * e2e_steps / (e2e_planes_ms * 1e-3),
                                             "d2h_bytes_per_step": hs_planes.d2h_bytes * world,
                                             "what": "the same with the board mirrored as the reference's two int8 planes"},
                    "full_copy_every_step": {"value": n_global * e2e_steps / (e2e_full_ms * 1e-3),
                                             "d2h_bytes_per_step": hs_full.d2h_bytes * world},
                    "full_copy_byte_mask": {"value": n_global * e2e_steps / (e2e_bytes_ms * 1e-3),
                                            "d2h_bytes_per_step": hs_bytes.d2h_bytes * world}},
            # k_gate + k_work per step, one k_pregen per `lib_every` steps on the side stream (+ k_onehot with the one-hot obs)
            "gpu_launches": (2 + (1 if wl["onehot"] else 0)) * args.steps + args.steps // lib_every,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH.get(args.config), "peak_source": peak_src,
                         "kernel": f"tmg::k_gate + tmg::k_work<32,{R},{Cc}> (one tmg_step)",
                         "bytes_per_env_step": wl["bytes_per_step"], "envs_per_launch": n_local,
                         "note": "integer/divergence-bound kernel: the HBM fraction is low by construction, see DESIGN.md"},
            "clocks": clocks,
            "step_ms": {"min": min(step_ms), "median": statistics.median(step_ms), "max": max(step_ms)},
            "drain_ms": drain_ms, "wall_s": t_wall, "status_flags_set": status_bad,
            "numa_node": numa_node,
            "engine": "byte planes (TMG_FLAG_BYTE_PLANES)" if args.byte_planes else "register-resident bit planes where the shape allows",
        }
        if no_reset_ms is not None:
            line["value_no_reset"] = {"value": n_global / (no_reset_ms_ * 1e-3), "ms_per_step": no_reset_ms_,
                                      "what": "the same steps with num_moves = 2^20: no episode end, no board generation"}
        if rollout_ms is not None:
            line["rollout"] = {"value": n_global * T * n_win / (rollout_ms_ * 1e-3), "unit": UNIT, "steps_per_launch": T, "launches": n_win,
                               "what": "tmg_step_many: the same env-steps with the actions of a whole window given up front "
                                       "(random-agent loop), boards kept on chip between steps; not the headline",
                               "mask_policy_in_kernel": {"value": n_global * T * n_win / (rollout_mask_ms_ * 1e-3),
                                                         "what": "tmg_rollout_policy(TMG_POLICY_MASK): actions sampled from the legal-move "
                                                                 "mask inside the kernel, every step an effective move"}}
        if not args.no_cpu_baseline and world == 1:   # reported at N=1 only
            threads = os.cpu_count() or 1
            v, dt, n, steps = cpu_port_throughput(wl, threads, target_s=10.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "per_core": v / threads,
                                    "sample": f"{n} envs x {steps} steps of the same workload ({dt:.1f} s), oracle/tmg_oracle.c "
                                              "(C port of the Python reference; the Python original runs ~1e3 steps/s/core) "
                                              "on all host threads",
                                    "python_reference": python_reference_throughput(wl, target_s=5.0)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
