"""Trace / replay files: one `.npz` that holds everything needed to reproduce a stretch of play bit for
bit -- the env configuration, the state the stretch starts from (`TileMatchVecEnv.state_dict()`), the actions, the
injected draws if the env uses them, and what every step returned.  The same file is read by the GPU env
(`replay_trace`, below) and by the CPU checker of the test suite (which restates the key names instead of importing
this module), so a trajectory recorded on one side can be checked on the other; it also doubles as a checkpoint that
carries its own proof.

What a step returns is what `TileMatchEnv.step` returns (tile_match_env.py:93-112): reward, done, the info counters,
`num_moves_left`, plus -- optionally, they dominate the file size -- the board and the effective-action mask after the
step.  With the Philox refill the draws are not stored: draw j of env e is a pure function of (seed, global env id,
cursor), and the cursors are part of the state.

Keys of the archive (all numpy arrays; T steps, N envs, A actions):
  format                      "tmg-trace-1"
  cfg/<name>                  int64 scalars: seed, num_envs, num_rows, num_cols, num_colours, num_moves, specials,
                              env_id_offset, autoreset (0 disabled, 1 next_step, 2 same_step), refill (0 philox, 1 injected)
  state0/<field>              TileMatchVecEnv.STATE_FIELDS before the first step
  draws                       (N, L) uint8, injected refill only
  actions                     (T, N) int32
  reward, num_new_specials, num_specials_activated, num_moves_left     (T, N) int32
  terminated, is_combination_match, shuffled                           (T, N) uint8
  board                       (T, N, 2, R, C) int8        [boards=True]
  mask                        (T, N, ceil(A/8)) uint8, np.packbits along the action axis   [masks=True]
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nat
from .vec_env import TileMatchVecEnv

FORMAT = "tmg-trace-1"
CFG_KEYS = ("seed", "num_envs", "num_rows", "num_cols", "num_colours", "num_moves", "specials", "env_id_offset",
            "autoreset", "refill")
STEP_I32 = ("reward", "num_new_specials", "num_specials_activated", "num_moves_left")
STEP_U8 = ("terminated", "is_combination_match", "shuffled")


class TraceMismatch(AssertionError):
    """replay_trace: the env did not reproduce the recorded trajectory."""


def _np(t: torch.Tensor) -> np.ndarray:
    a = t.detach().cpu().numpy()
    return a.astype(np.uint8) if a.dtype == np.bool_ else a


def _config(env: TileMatchVecEnv) -> dict:
    return {"seed": env.seed, "num_envs": env.num_envs, "num_rows": env.num_rows, "num_cols": env.num_cols,
            "num_colours": env.num_colours, "num_moves": env.num_moves, "specials": env.specials,
            "env_id_offset": env.env_id_offset, "autoreset": nat.AUTORESET[env.autoreset_mode],
            "refill": nat.REFILL[env.refill]}


def record_trace(env: TileMatchVecEnv, actions, boards: bool = True, masks: bool = True) -> dict:
    """Steps `env` through actions (T, N) from its current state and returns the trace as a dict of numpy arrays.
    The env is left in the state after the last step."""
    a = torch.as_tensor(actions).to(device=env.device, dtype=torch.int32).contiguous()
    if a.dim() != 2 or a.shape[1] != env.num_envs:
        raise ValueError("actions must have shape (T, num_envs)")
    tr = {"format": np.array(FORMAT)}
    for k, v in _config(env).items():
        tr["cfg/" + k] = np.array(v, dtype=np.int64)
    sd = env.state_dict()
    for name in env.STATE_FIELDS:
        tr["state0/" + name] = _np(sd[name])
    if env.refill == "injected":
        if env._injected is None:
            raise RuntimeError("refill='injected' but no draws were set")
        tr["draws"] = _np(env._injected)
    tr["actions"] = _np(a)
    T = a.shape[0]
    steps = {k: [] for k in STEP_I32 + STEP_U8 + (("board",) if boards else ()) + (("mask",) if masks else ())}
    for t in range(T):
        env.step(a[t])
        for k in STEP_I32 + STEP_U8:
            steps[k].append(_np(getattr(env, k)))
        if boards:
            steps["board"].append(_np(env.board))
        if masks:
            steps["mask"].append(np.packbits(_np(env.mask), axis=1))
    for k, v in steps.items():
        dt = np.int32 if k in STEP_I32 else (np.int8 if k == "board" else np.uint8)
        tr[k] = (np.stack(v) if T else np.zeros((0, env.num_envs), dt)).astype(dt, copy=False)
    return tr


def save_trace(path, trace: dict) -> None:
    check_trace(trace)
    np.savez_compressed(path, **trace)


def load_trace(path) -> dict:
    with np.load(path, allow_pickle=False) as z:
        tr = {k: z[k] for k in z.files}
    check_trace(tr)
    return tr


def check_trace(tr: dict) -> None:
    """Structure of the archive: format tag, every key present, shapes consistent with cfg."""
    if "format" not in tr or str(tr["format"]) != FORMAT:
        raise ValueError(f"not a {FORMAT} archive")
    for k in CFG_KEYS:
        if "cfg/" + k not in tr:
            raise ValueError(f"trace lacks cfg/{k}")
    c = trace_config(tr)
    N, R, Cc = c["num_envs"], c["num_rows"], c["num_cols"]
    A = 2 * R * Cc - R - Cc
    for name in TileMatchVecEnv.STATE_FIELDS:
        if "state0/" + name not in tr:
            raise ValueError(f"trace lacks state0/{name}")
    if tr["state0/board"].shape != (N, 2, R, Cc):
        raise ValueError("state0/board has the wrong shape")
    if "actions" not in tr or tr["actions"].ndim != 2 or tr["actions"].shape[1] != N:
        raise ValueError("actions must have shape (T, num_envs)")
    T = tr["actions"].shape[0]
    for k in STEP_I32 + STEP_U8:
        if k not in tr or tr[k].shape != (T, N):
            raise ValueError(f"{k} must have shape (T, num_envs)")
    if "board" in tr and tr["board"].shape != (T, N, 2, R, Cc):
        raise ValueError("board must have shape (T, num_envs, 2, num_rows, num_cols)")
    if "mask" in tr and tr["mask"].shape != (T, N, (A + 7) // 8):
        raise ValueError("mask must have shape (T, num_envs, ceil(A/8))")
    if c["refill"] == nat.REFILL["injected"] and ("draws" not in tr or tr["draws"].shape[0] != N):
        raise ValueError("a trace of an injected-refill env carries its draws")


def trace_config(tr: dict) -> dict:
    return {k: int(tr["cfg/" + k]) for k in CFG_KEYS}


def env_from_trace(tr: dict, device="cuda:0", **kw) -> TileMatchVecEnv:
    """A fresh env with the trace's configuration (specials given as the bit mask the trace stores)."""
    c = trace_config(tr)
    on = [name for name, bit in nat.SPECIAL_BITS.items() if c["specials"] & bit]
    cl = [s for s in on if s == "cookie"]          # the one colourless special (board.py:18-25)
    cs = [s for s in on if s != "cookie"]
    inv = lambda d, v: next(k for k, x in d.items() if x == v)  # noqa: E731
    return TileMatchVecEnv(c["num_envs"], c["num_rows"], c["num_cols"], c["num_colours"], c["num_moves"], cl, cs,
                           seed=c["seed"], device=device, autoreset=inv(nat.AUTORESET, c["autoreset"]),
                           refill=inv(nat.REFILL, c["refill"]), env_id_offset=c["env_id_offset"], **kw)


def replay_trace(env: TileMatchVecEnv, tr: dict) -> int:
    """Puts `env` into the trace's start state, takes the recorded actions and compares every recorded output of
    every step bit for bit.  Raises TraceMismatch at the first step that differs; returns the number of steps."""
    check_trace(tr)
    c = trace_config(tr)
    mine = _config(env)
    for k in CFG_KEYS:
        if k != "seed" and c[k] != mine[k]:
            raise ValueError(f"trace was recorded with {k}={c[k]}, this env has {mine[k]}")
    if "draws" in tr:
        env.set_injected_draws(torch.from_numpy(np.ascontiguousarray(tr["draws"])).to(env.device))
    sd = {name: torch.from_numpy(np.ascontiguousarray(tr["state0/" + name])) for name in env.STATE_FIELDS}
    sd["config"] = {k: c[k] for k in ("seed", "num_envs", "num_rows", "num_cols", "num_colours", "num_moves", "specials",
                                      "env_id_offset")}
    env.load_state_dict(sd)
    acts = torch.from_numpy(np.ascontiguousarray(tr["actions"])).to(env.device)
    for t in range(acts.shape[0]):
        env.step(acts[t])
        for k in STEP_I32 + STEP_U8 + ("board", "mask"):
            if k not in tr:
                continue
            got = _np(env.mask if k == "mask" else getattr(env, k))
            if k == "mask":
                got = np.packbits(got, axis=1)
            want = tr[k][t]
            if not np.array_equal(got, want):
                bad = np.flatnonzero((got.reshape(got.shape[0], -1) != want.reshape(want.shape[0], -1)).any(axis=1))
                raise TraceMismatch(f"step {t}: {k} differs for {len(bad)} envs, first {bad[:5].tolist()}")
    return int(acts.shape[0])
