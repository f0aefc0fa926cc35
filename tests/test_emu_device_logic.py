"""Runs the product's DEVICE code (tile_match_gym_b200/csrc/tmg_device.cuh) on the CPU through the test-only
lane emulator (tests/emu) against the golden traces and the oracle.  This is how kernel logic is debugged in a
container without a GPU; the GPU parity tests (test_gpu_parity.py) are the real gate and do not use it."""
import numpy as np
import pytest

from oracle import oracle as orc
from test_oracle_golden import _trace_meta, replay_trace
from emu.emu import EmuVecEnv

ALL_CL = ("cookie",)
ALL_CS = ("vertical_laser", "horizontal_laser", "bomb")


def test_golden_traces_through_emulated_device_code():
    z, meta = _trace_meta()
    for m in meta:
        replay_trace(lambda m: EmuVecEnv(1, m["R"], m["C"], m["K"], m["num_moves"], m["cl"], m["cs"], seed=m["seed"],
                                         env_id_offset=m["env_id"]), z, m)


@pytest.mark.parametrize("cfg", [
    (10, 10, 4, ALL_CL, ALL_CS, 12, "same_step", 24, 40),
    (9, 9, 6, ALL_CL, ALL_CS, 8, "next_step", 24, 30),
    (4, 5, 3, ALL_CL, ALL_CS, 5, "same_step", 48, 40),
    (6, 20, 5, ALL_CL, ALL_CS, 6, "same_step", 8, 20),
    (10, 10, 4, ALL_CL, ALL_CS, 5, "same_step", 16, 12, 2),   # TMG_FLAG_NO_PREGEN: boards generated inside the step
    (3, 5, 3, ALL_CL, ALL_CS, 3, "same_step", 64, 30),        # tiny boards: shuffles inside generate_board
])
def test_emulated_batch_vs_oracle(cfg):
    R, C, K, cl, cs, moves, autoreset, N, steps = cfg[:9]
    flags = cfg[9] if len(cfg) > 9 else 0
    e = EmuVecEnv(N, R, C, K, moves, cl, cs, seed=3, autoreset=autoreset, env_id_offset=50, flags=flags)
    o = orc.OracleVecEnv(N, R, C, K, moves, cl, cs, seed=3, autoreset=autoreset, env_id_offset=50, num_threads=4)
    e.reset(); o.reset()
    rng = np.random.default_rng(7)
    fields = ["board", "timer", "draw_cursor", "shuffle_cursor", "reward", "terminated", "is_combination_match",
              "num_new_specials", "num_specials_activated", "shuffled", "mask", "num_moves_left", "status", "episode"]
    for t in range(steps):
        m = o.mask.astype(np.float64) + 1e-9
        u = rng.random((N, 1)) * m.sum(axis=1, keepdims=True)
        a = (np.cumsum(m, axis=1) < u).sum(axis=1).clip(0, o.A - 1).astype(np.int32)
        e.step(a); o.step(a)
        for f in fields:
            assert np.array_equal(getattr(e, f), getattr(o, f)), (t, f)
