"""Writes tests/golden/oracle_trace.npz: a "tmg-trace-1" archive (tile_match_gym_b200/trace.py) recorded by the CPU
oracle -- 48 envs of the headline shape (10x10, 4 colours, all specials), 6-move episodes with same-step autoreset,
40 steps whose actions are sampled from the effective-action mask two times out of three (so cascades, specials and
combination matches occur), Philox refill.  The GPU test replays it with `replay_trace`.
Run from the repo root:  python tests/golden/gen_trace.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import trace as otrace  # noqa: E402

CFG = {"seed": 77, "num_envs": 48, "num_rows": 10, "num_cols": 10, "num_colours": 4, "num_moves": 6, "specials": 15,
       "env_id_offset": 5000, "autoreset": 2, "refill": 0}


def main():
    o = otrace.oracle_from_config(CFG)
    o.reset()
    rng = np.random.default_rng(2024)
    T = 40
    # the actions depend on the state (mask sampling), so record step by step
    acts = np.zeros((T, o.N), np.int32)
    parts = []
    for t in range(T):
        for e in range(o.N):
            idx = np.flatnonzero(o.mask[e])
            acts[t, e] = rng.choice(idx) if (len(idx) and rng.random() < 0.67) else rng.integers(0, o.A)
        parts.append(otrace.record(o, CFG, acts[t:t + 1]))
    tr = dict(parts[0])
    tr["actions"] = acts
    for k in otrace.STEP_I32 + otrace.STEP_U8 + ("board", "mask"):
        tr[k] = np.concatenate([p[k] for p in parts])
    assert otrace.replay(tr) == T
    out = os.path.join(ROOT, "tests", "golden", "oracle_trace.npz")
    np.savez_compressed(out, **tr)
    print(out, os.path.getsize(out), "bytes; reward sum", int(tr["reward"].sum()), "combos", int(tr["is_combination_match"].sum()),
          "specials activated", int(tr["num_specials_activated"].sum()))


if __name__ == "__main__":
    main()
