"""TEST INFRASTRUCTURE (like everything under oracle/): the CPU oracle's reader and writer of the trace archive that
tile_match_gym_b200/trace.py defines ("tmg-trace-1").  A trace recorded on the GPU is replayed here
through oracle/tmg_oracle.c (the restatement of board.py:330-395 / tile_match_env.py:93-112), and a trace recorded
here is replayed on the GPU by `replay_trace` -- the archive is the only thing the two sides share.  The key names
are restated, not imported, so that a change of the product's format shows up as a test failure."""
import numpy as np

from .oracle import OracleVecEnv

FORMAT = "tmg-trace-1"
CFG_KEYS = ("seed", "num_envs", "num_rows", "num_cols", "num_colours", "num_moves", "specials", "env_id_offset",
            "autoreset", "refill")
STATE_FIELDS = ("board", "timer", "draw_cursor", "shuffle_cursor", "episode", "num_moves_left", "status", "reward",
                "terminated", "is_combination_match", "num_new_specials", "num_specials_activated", "shuffled")
STEP_I32 = ("reward", "num_new_specials", "num_specials_activated", "num_moves_left")
STEP_U8 = ("terminated", "is_combination_match", "shuffled")
_BITS = (("cookie", 1), ("vertical_laser", 2), ("horizontal_laser", 4), ("bomb", 8))


def oracle_from_config(c: dict, num_threads: int = 1) -> OracleVecEnv:
    on = [n for n, b in _BITS if c["specials"] & b]
    return OracleVecEnv(c["num_envs"], c["num_rows"], c["num_cols"], c["num_colours"], c["num_moves"],
                        [s for s in on if s == "cookie"], [s for s in on if s != "cookie"], seed=c["seed"],
                        autoreset=("disabled", "next_step", "same_step")[c["autoreset"]],
                        refill=("philox", "injected")[c["refill"]], env_id_offset=c["env_id_offset"],
                        num_threads=num_threads)


def record(o: OracleVecEnv, cfg: dict, actions, boards=True, masks=True, draws=None) -> dict:
    """Steps the oracle env `o` (already reset) through actions (T, N) and returns the archive dict."""
    a = np.ascontiguousarray(actions, dtype=np.int32)
    tr = {"format": np.array(FORMAT)}
    for k in CFG_KEYS:
        tr["cfg/" + k] = np.array(cfg[k], dtype=np.int64)
    for name in STATE_FIELDS:
        tr["state0/" + name] = np.array(getattr(o, name), copy=True)
    if draws is not None:
        tr["draws"] = np.ascontiguousarray(draws, dtype=np.uint8)
    tr["actions"] = a
    names = STEP_I32 + STEP_U8 + (("board",) if boards else ()) + (("mask",) if masks else ())
    steps = {k: [] for k in names}
    for t in range(a.shape[0]):
        o.step(a[t])
        for k in names:
            v = np.array(getattr(o, k), copy=True)
            steps[k].append(np.packbits(v, axis=1) if k == "mask" else v)
    for k in names:
        tr[k] = np.stack(steps[k])
    return tr


def replay(tr: dict, num_threads: int = 1) -> int:
    """Replays the archive on the oracle from its start state; AssertionError at the first differing step."""
    assert str(tr["format"]) == FORMAT
    c = {k: int(tr["cfg/" + k]) for k in CFG_KEYS}
    o = oracle_from_config(c, num_threads)
    if "draws" in tr:
        o.set_injected_draws(tr["draws"])
    for name in STATE_FIELDS:
        getattr(o, name)[...] = tr["state0/" + name]
    for t in range(tr["actions"].shape[0]):
        o.step(tr["actions"][t])
        for k in STEP_I32 + STEP_U8 + ("board", "mask"):
            if k not in tr:
                continue
            got = np.packbits(o.mask, axis=1) if k == "mask" else getattr(o, k)
            if not np.array_equal(got, tr[k][t]):
                bad = np.flatnonzero((got.reshape(got.shape[0], -1) != tr[k][t].reshape(got.shape[0], -1)).any(axis=1))
                raise AssertionError(f"oracle replay, step {t}: {k} differs for {len(bad)} envs, first {bad[:5].tolist()}")
    return int(tr["actions"].shape[0])
