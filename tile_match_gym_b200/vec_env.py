"""TileMatchVecEnv -- drop-in *vectorised* TileMatchEnv (reference tile_match_env.py:14-150) whose
reset/step run as hand-written sm_100a kernels behind the C ABI of include/tmg_b200.h.

Same constructor arguments, spaces, reward, termination and info keys as the reference env; every
per-env scalar becomes a length-N tensor and `info["effective_actions"]` becomes an (N, A) bool mask.
PyTorch is only the plumbing here (device tensors that alias the engine's buffers, streams); all
transition arithmetic happens in libtmg_b200.so.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch

from . import _native as nat
from . import spaces as sp

ENV_ID = "TileMatch-v0"  # reference __init__.py:3


class _DevArray:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can alias it (no copy)."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner


_TYPESTR = {"board": "|i1", "timer": "<i4", "draw_cursor": "<i8", "shuffle_cursor": "<i8", "reward": "<i4",
            "terminated": "|u1", "is_combination_match": "|u1", "num_new_specials": "<i4",
            "num_specials_activated": "<i4", "shuffled": "|u1", "mask": "|u1", "num_moves_left": "<i4",
            "status": "<i4", "episode": "<i4"}


class TileMatchVecEnv:
    """N independent TileMatchEnv instances on one B200.

    Args mirror `TileMatchEnv(num_rows, num_cols, num_colours, num_moves, colourless_specials,
    colour_specials, seed=1)` (tile_match_env.py:17-27) plus:
      num_envs        envs held by this shard
      device          "cuda:i"
      autoreset       "next_step" (gymnasium default), "same_step" or "disabled" (reference behaviour:
                      stepping a finished env is an error, reported through `status`)
      refill          "philox" (counter-based stream, see include/tmg_b200.h) or "injected"
      env_id_offset   global id of local env 0; the draw stream of an env depends only on (seed, global id),
                      so results are independent of how the batch is sharded over GPUs
      compute_mask    maintain info["effective_actions"] (the reference always does)
      pregenerate     generate each env's next board ahead of time on a side stream (same bytes as generating it
                      inside the step; only the timing differs)
      obs             "int8" (aliases engine state, zero-copy), "int32" (reference dtype, one cast per call)
                      or "onehot" (OneHotWrapper planes, uint8)
      copy_outputs    step() returns fresh copies of reward / terminated / the info tensors instead of views of the
                      engine's buffers (see step(): the views are overwritten in place by the next call)
      constructive_reset  NOT reference behaviour: boards come from the constructive line-free sampler of
                      TMG_FLAG_CONSTRUCTIVE_RESET (include/tmg_b200.h) instead of generate_board -- for shapes such as
                      32x32 / 7 colours, where the reference's generate_board never returns
      byte_planes     diagnostics: run the moves on the shared-memory byte-plane engine (TMG_FLAG_BYTE_PLANES) instead of
                      the register-resident bit-plane engine; results are identical
    """

    metadata = {"render_modes": ["string"], "render_fps": 2}

    def __init__(self, num_envs: int, num_rows: int, num_cols: int, num_colours: int, num_moves: int,
                 colourless_specials, colour_specials, seed: Optional[int] = 1, device="cuda:0",
                 autoreset: str = "next_step", refill: str = "philox", env_id_offset: int = 0,
                 compute_mask: bool = True, obs: str = "int8", max_reset_iters: int = 0,
                 render_mode: str = "string", pregenerate: bool = True, copy_outputs: bool = False,
                 byte_planes: bool = False, constructive_reset: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError("tile_match_gym_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._lib = nat.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("device must be a CUDA device")
        self.num_envs = int(num_envs)
        self.num_rows, self.num_cols, self.num_colours = int(num_rows), int(num_cols), int(num_colours)
        self.num_moves = int(num_moves)
        self.colourless_specials = list(colourless_specials)
        self.colour_specials = list(colour_specials)
        self.num_colour_specials = len(self.colour_specials)
        self.num_colourless_specials = len(self.colourless_specials)
        self.seed = 1 if seed is None else int(seed)
        self.render_mode = render_mode
        self.autoreset_mode = autoreset
        self.refill = refill
        self.env_id_offset = int(env_id_offset)
        self.obs_mode = obs
        if obs not in ("int8", "int32", "onehot"):
            raise ValueError("obs must be 'int8', 'int32' or 'onehot'")
        self.num_actions = 2 * self.num_rows * self.num_cols - self.num_rows - self.num_cols  # board.py:77
        self.specials = nat.specials_mask(self.colourless_specials, self.colour_specials)
        self.onehot_planes = self._lib.tmg_onehot_planes(self.num_colours, self.specials)
        self.compute_mask = bool(compute_mask)
        self.copy_outputs = bool(copy_outputs)
        self._mirror_keep = None   # pinned host arrays bound as the host mirror (tmg_host_bind): kept alive while bound

        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        torch.cuda.init()
        with torch.cuda.device(dev_index):
            torch.zeros(1, device=self.device)  # make sure the primary context exists before the library uses it
        cfg = nat.Config(C.sizeof(nat.Config), dev_index, self.num_envs, self.num_rows, self.num_cols,
                         self.num_colours, self.num_moves, self.specials, nat.AUTORESET[autoreset],
                         nat.REFILL[refill],
                         (0 if compute_mask else nat.FLAG_NO_MASK) | (0 if pregenerate else nat.FLAG_NO_PREGEN)
                         | (nat.FLAG_BYTE_PLANES if byte_planes else 0)
                         | (nat.FLAG_CONSTRUCTIVE_RESET if constructive_reset else 0),
                         int(max_reset_iters),
                         self.seed & 0xFFFFFFFFFFFFFFFF, self.env_id_offset)
        h = C.c_void_p()
        nat.check(self._lib.tmg_create(C.byref(cfg), C.byref(h)), "tmg_create")
        self._h = h
        bufs = nat.Buffers()
        nat.check(self._lib.tmg_get_buffers(self._h, C.byref(bufs)), "tmg_get_buffers")
        N, R, Cc, A = self.num_envs, self.num_rows, self.num_cols, self.num_actions
        shapes = {"board": (N, 2, R, Cc), "mask": (N, A)}
        self._t = {}
        for name in nat.BUFFER_FIELDS:
            arr = _DevArray(getattr(bufs, name), shapes.get(name, (N,)), _TYPESTR[name], self)
            self._t[name] = torch.as_tensor(arr, device=self.device)
        # public aliases of engine state (zero-copy)
        self.board = self._t["board"]
        self.timer = self._t["timer"]
        self.draw_cursor = self._t["draw_cursor"]
        self.shuffle_cursor = self._t["shuffle_cursor"]
        self.reward = self._t["reward"]
        self.terminated = self._t["terminated"].view(torch.bool)
        self.is_combination_match = self._t["is_combination_match"].view(torch.bool)
        self.num_new_specials = self._t["num_new_specials"]
        self.num_specials_activated = self._t["num_specials_activated"]
        self.shuffled = self._t["shuffled"].view(torch.bool)
        self.mask = self._t["mask"].view(torch.bool)
        self.num_moves_left = self._t["num_moves_left"]
        self.status = self._t["status"]
        self.episode = self._t["episode"]
        self.truncated = torch.zeros(N, dtype=torch.bool, device=self.device)  # tile_match_env.py:112: always False
        self._onehot = (torch.empty((N, self.onehot_planes, R, Cc), dtype=torch.uint8, device=self.device)
                        if obs == "onehot" else None)
        self._injected = None

        # spaces (tile_match_env.py:52-77, wrappers.py:25-30)
        self._moves_left_observation_space = sp.Discrete(self.num_moves + 1, seed=self.seed)
        if obs == "onehot":
            board_space = sp.onehot_board_space(R, Cc, self.onehot_planes)
        else:
            board_space = sp.board_space(R, Cc, self.num_colours, self.num_colourless_specials,
                                         self.num_colour_specials, seed=self.seed)
        self._board_observation_space = board_space
        self.single_observation_space = sp.Dict({"board": board_space,
                                                 "num_moves_left": self._moves_left_observation_space})
        self.single_action_space = sp.Discrete(self.num_actions, seed=self.seed)
        self.observation_space = self.single_observation_space
        self.action_space = self.single_action_space
        # (r1,c1),(r2,c2) table, board.py:78-93
        self._action_to_coords = tuple(self.action_to_coords(a) for a in range(self.num_actions))

    # ------------------------------------------------------------------------------------------------
    @classmethod
    def sharded(cls, global_num_envs: int, rank: int, world_size: int, *args, **kwargs) -> "TileMatchVecEnv":
        """Shard [0, global_num_envs) by env index: rank g owns [g*N/G, (g+1)*N/G).  No cross-GPU traffic."""
        lo, hi = shard_range(global_num_envs, rank, world_size)
        kwargs = dict(kwargs)
        kwargs["env_id_offset"] = kwargs.get("env_id_offset", 0) + lo
        return cls(hi - lo, *args, **kwargs)

    def action_to_coords(self, action: int):
        out = (C.c_int32 * 4)()
        nat.check(self._lib.tmg_action_to_coords(self.num_rows, self.num_cols, int(action), C.byref(out)),
                  "tmg_action_to_coords")
        return ((out[0], out[1]), (out[2], out[3]))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def unbind_host_mirror(self) -> None:
        """Detach the host mirror (tmg_host_bind(NULL)) and only then let go of the pinned arrays: the step kernels write
        into them over PCIe, so they must outlive every launch that was issued while they were bound."""
        if getattr(self, "_h", None) is not None and self._h and self._mirror_keep is not None:
            nat.check(self._lib.tmg_host_bind(self._h, None, self._stream()), "tmg_host_bind")
            torch.cuda.synchronize(self.device)
        self._mirror_keep = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            try:
                self.unbind_host_mirror()
            finally:
                torch.cuda.synchronize(self.device)
                self._lib.tmg_destroy(self._h)
                self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # ------------------------------------------------------------------------------------------------
    def set_seed(self, seed: int) -> None:
        """tile_match_env.py:79-82: replaces the board generator (new key, cursors back to zero)."""
        self.seed = int(seed)
        nat.check(self._lib.tmg_set_seed(self._h, self.seed & 0xFFFFFFFFFFFFFFFF, self._stream()), "tmg_set_seed")

    # ---- checkpoint / resume: the state of an env is its board, timer, stream cursors and episode number
    STATE_FIELDS = ("board", "timer", "draw_cursor", "shuffle_cursor", "episode", "num_moves_left", "status", "reward",
                    "terminated", "is_combination_match", "num_new_specials", "num_specials_activated", "shuffled")

    def state_dict(self) -> dict:
        """Host copy of everything the next steps depend on (boards generated ahead of time are not state: board j of
        env e is a pure function of (seed, e, j))."""
        self.join()
        torch.cuda.synchronize(self.device)
        sd = {name: self._t[name].detach().cpu().clone() for name in self.STATE_FIELDS}
        sd["config"] = {"format": "tmg-state-2", "seed": self.seed, "num_envs": self.num_envs, "num_rows": self.num_rows,
                        "num_cols": self.num_cols, "num_colours": self.num_colours, "num_moves": self.num_moves,
                        "specials": self.specials, "env_id_offset": self.env_id_offset, "autoreset": self.autoreset_mode,
                        "refill": self.refill}
        return sd

    def load_state_dict(self, sd: dict) -> None:
        """Resume from state_dict(): the following steps reproduce, bit for bit, what the saved env would have done."""
        cfg = sd["config"]
        for k in ("num_envs", "num_rows", "num_cols", "num_colours", "num_moves", "specials", "env_id_offset"):
            if cfg[k] != getattr(self, k):
                raise ValueError(f"state was saved with {k}={cfg[k]}, this env has {getattr(self, k)}")
        # the autoreset mode decides what a saved terminal timer means and the refill mode where the next draws come
        # from: a state only continues bit for bit under the modes it was saved with (older states carry neither)
        for k, mine in (("autoreset", self.autoreset_mode), ("refill", self.refill)):
            if k in cfg and cfg[k] != mine:
                raise ValueError(f"state was saved with {k}={cfg[k]!r}, this env has {mine!r}")
        if cfg["seed"] != self.seed:
            self.set_seed(cfg["seed"])       # also forgets the boards generated ahead of time under the old key
        self.join()
        for name in self.STATE_FIELDS:
            self._t[name].copy_(sd[name].to(self._t[name].dtype))
        self.legal_mask()                    # masks (and a bound host mirror) follow from the boards

    def set_injected_draws(self, draws: torch.Tensor) -> None:
        """draws: (N, Ldraws) uint8 on this device, values 1..K; consumed in the reference's draw order."""
        if self.refill != "injected":
            raise RuntimeError("env was not created with refill='injected'")
        d = draws.to(device=self.device, dtype=torch.uint8).contiguous()
        if d.dim() != 2 or d.shape[0] != self.num_envs:
            raise ValueError("draws must have shape (num_envs, per_env_len)")
        self._injected = d
        nat.check(self._lib.tmg_set_injected_draws(self._h, C.c_void_p(d.data_ptr()), d.shape[1]),
                  "tmg_set_injected_draws")

    def _obs(self):
        if self.obs_mode == "onehot":
            nat.check(self._lib.tmg_encode_onehot(self._h, C.c_void_p(self._onehot.data_ptr()), self._stream()),
                      "tmg_encode_onehot")
            board = self._onehot
        elif self.obs_mode == "int32":
            board = self.board.to(torch.int32)
        else:
            board = self.board  # aliases live state, like the reference (tile_match_env.py:115)
        return OrderedDict([("board", board), ("num_moves_left", self.num_moves_left)])

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        """tile_match_env.py:84-91.  options: {"reset_mask": bool (N,), "init_boards": int8 (N,2,R,C)}."""
        if seed is not None:
            self.set_seed(seed)
        options = options or {}
        mask = options.get("reset_mask")
        boards = options.get("init_boards")
        mptr = bptr = None
        if mask is not None:
            mask = torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()
            if mask.shape != (self.num_envs,):
                raise ValueError("reset_mask must have shape (num_envs,)")
            mptr = C.c_void_p(mask.data_ptr())
        if boards is not None:
            boards = torch.as_tensor(boards).to(device=self.device, dtype=torch.int8).contiguous()
            if boards.shape != (self.num_envs, 2, self.num_rows, self.num_cols):
                raise ValueError("init_boards must have shape (num_envs, 2, num_rows, num_cols)")
            bptr = C.c_void_p(boards.data_ptr())
        nat.check(self._lib.tmg_reset(self._h, mptr, bptr, self._stream()), "tmg_reset")
        self._keep = (mask, boards)  # keep the inputs alive until the stream has consumed them
        return self._obs(), {"effective_actions": self.mask}

    def step(self, actions):
        """tile_match_env.py:93-112 for every env.  actions: (N,) integer tensor.

        Aliasing contract: obs["board"] (obs="int8"), obs["num_moves_left"], reward, terminated and every info tensor
        are zero-copy VIEWS of the engine's device buffers -- like the reference's obs, which aliases live state
        (tile_match_env.py:115) -- and are overwritten in place by the next step() / step_many() / rollout() / reset().
        Code that keeps them across calls (`rewards.append(rew)`) must clone them, or construct the env with
        copy_outputs=True, which returns fresh copies of the small per-env tensors (reward, terminated, info)."""
        a = torch.as_tensor(actions)
        if a.device != self.device or a.dtype != torch.int32 or not a.is_contiguous():
            a = a.to(device=self.device, dtype=torch.int32).contiguous()
        if a.shape != (self.num_envs,):
            raise ValueError("actions must have shape (num_envs,)")
        nat.check(self._lib.tmg_step(self._h, C.c_void_p(a.data_ptr()), self._stream()), "tmg_step")
        self._last_actions = a
        info = {
            "is_combination_match": self.is_combination_match,
            "num_new_specials": self.num_new_specials,
            "num_specials_activated": self.num_specials_activated,
            "shuffled": self.shuffled,
            "effective_actions": self.mask,
        }
        if self.copy_outputs:
            info = {k: (v if k == "effective_actions" else v.clone()) for k, v in info.items()}
            return self._obs(), self.reward.clone(), self.terminated.clone(), self.truncated, info
        return self._obs(), self.reward, self.terminated, self.truncated, info

    def step_many(self, actions):
        """T successive steps in one launch (tmg_step_many).  actions: (T, N) integer tensor, env e takes
        actions[t, e] at step t.  Returns (rewards (T, N) int32, terminations (T, N) bool); the env's buffers and
        observations afterwards are those after the last step."""
        a = torch.as_tensor(actions)
        if a.device != self.device or a.dtype != torch.int32 or not a.is_contiguous():
            a = a.to(device=self.device, dtype=torch.int32).contiguous()
        if a.dim() != 2 or a.shape[1] != self.num_envs:
            raise ValueError("actions must have shape (T, num_envs)")
        T = a.shape[0]
        rew = torch.empty((T, self.num_envs), dtype=torch.int32, device=self.device)
        term = torch.empty((T, self.num_envs), dtype=torch.uint8, device=self.device)
        nat.check(self._lib.tmg_step_many(self._h, C.c_void_p(a.data_ptr()), T, C.c_void_p(rew.data_ptr()),
                                          C.c_void_p(term.data_ptr()), self._stream()), "tmg_step_many")
        self._last_actions = a
        return rew, term.view(torch.bool)

    def rollout(self, num_steps: int, policy: str = "mask"):
        """num_steps steps with the agent inside the kernel (tmg_rollout_policy): policy "uniform" draws any action,
        "mask" samples from the effective actions.  Returns (actions (T, N) int32, rewards (T, N) int32,
        terminations (T, N) bool); the actions are a pure function of (seed, env id, board number, timer)."""
        T = int(num_steps)
        act = torch.empty((T, self.num_envs), dtype=torch.int32, device=self.device)
        rew = torch.empty((T, self.num_envs), dtype=torch.int32, device=self.device)
        term = torch.empty((T, self.num_envs), dtype=torch.uint8, device=self.device)
        nat.check(self._lib.tmg_rollout_policy(self._h, nat.POLICY[policy], T, C.c_void_p(act.data_ptr()),
                                               C.c_void_p(rew.data_ptr()), C.c_void_p(term.data_ptr()), self._stream()),
                  "tmg_rollout_policy")
        return act, rew, term.view(torch.bool)

    def join(self) -> None:
        """Make the current stream wait for the board generations queued on the library's side stream."""
        nat.check(self._lib.tmg_join(self._h, self._stream()), "tmg_join")

    def legal_mask(self) -> torch.Tensor:
        """Recompute the mask from the current boards (tile_match_env.py:118-124)."""
        nat.check(self._lib.tmg_legal_mask(self._h, self._stream()), "tmg_legal_mask")
        return self.mask

    def onehot(self, dtype=torch.uint8) -> torch.Tensor:
        """OneHotWrapper._one_hot_encode_board (wrappers.py:54-69) of the current boards."""
        N, R, Cc = self.num_envs, self.num_rows, self.num_cols
        out = torch.empty((N, self.onehot_planes, R, Cc), dtype=dtype, device=self.device)
        if dtype == torch.uint8:
            nat.check(self._lib.tmg_encode_onehot(self._h, C.c_void_p(out.data_ptr()), self._stream()), "onehot")
        elif dtype == torch.float32:
            nat.check(self._lib.tmg_encode_onehot_f32(self._h, C.c_void_p(out.data_ptr()), self._stream()), "onehot")
        elif dtype == torch.float64:       # the reference's own dtype (wrappers.py:57,64)
            nat.check(self._lib.tmg_encode_onehot_f64(self._h, C.c_void_p(out.data_ptr()), self._stream()), "onehot")
        else:
            raise ValueError("dtype must be uint8, float32 or float64")
        return out

    def check_status(self, clear: bool = True) -> None:
        """Raises like the reference would (Exception / IndexError / ValueError) if any env flagged an error."""
        st = self.status
        if not bool((st != 0).any().item()):
            return
        agg = 0
        for v in torch.unique(st).tolist():
            agg |= int(v) & 0xFFFFFFFF
        if clear:
            nat.check(self._lib.tmg_clear_status(self._h, self._stream()), "tmg_clear_status")
        msg = self._lib.tmg_status_string(agg).decode()
        if agg & nat.ST_NEEDS_RESET:
            raise Exception("You must call reset before calling step")  # tile_match_env.py:95
        if agg & nat.ST_BAD_ACTION:
            raise IndexError(f"action out of range: {msg}")
        raise RuntimeError(f"env status: {msg}")

    def debug_op(self, op: str, args=None) -> None:
        """Runs one engine primitive on every device board (known-answer tests)."""
        a = None
        if args is not None:
            a = torch.as_tensor(args).to(device=self.device, dtype=torch.int32).contiguous()
            if a.shape != (self.num_envs, 4):
                raise ValueError("args must have shape (num_envs, 4)")
        nat.check(self._lib.tmg_debug_op(self._h, nat.OPS[op], None if a is None else C.c_void_p(a.data_ptr()),
                                         self._stream()), "tmg_debug_op")
        self._dbg_keep = a

    def render(self):
        """String rendering of env 0 (tile_match_env.py:127-143 without the ANSI colour map)."""
        b = self.board[0].cpu().numpy()
        lines = [" " + "-" * (self.num_cols * 2 + 1)]
        for r in range(self.num_rows):
            lines.append("| " + " ".join(f"{b[0, r, c]}{'*' if b[1, r, c] not in (0, 1) else ''}" for c in range(self.num_cols)) + " |")
        lines.append(" " + "-" * (self.num_cols * 2 + 1))
        print("\n".join(lines))


class ProportionRewardWrapper:
    """wrappers.py:71-77 for the vector env: reward / (num_rows * num_cols), float64 like the reference's Python
    float (int / int).  Everything else is forwarded to the wrapped TileMatchVecEnv."""

    def __init__(self, env: TileMatchVecEnv):
        self.env = env
        self.flat_size = env.num_rows * env.num_cols

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reward(self, reward: torch.Tensor) -> torch.Tensor:
        # tensor / tensor is an IEEE division; tensor / python-scalar is a multiplication by the reciprocal on CUDA,
        # which differs from the reference's int / int in the last bit
        return reward.to(torch.float64) / torch.full((), float(self.flat_size), dtype=torch.float64, device=reward.device)

    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, actions):
        obs, reward, terminated, truncated, info = self.env.step(actions)
        return obs, self.reward(reward), terminated, truncated, info

    def step_many(self, actions):
        rewards, terminated = self.env.step_many(actions)
        return self.reward(rewards), terminated


class HostStepper:
    """The host-buffer path (tmg_step_host): actions come from pinned host memory, observations / rewards /
    terminations / masks are copied back to pinned host memory, stream synchronised per call.  This is what a
    CPU-side consumer of the reference API pays, and what bench.py reports as `e2e`."""

    def __init__(self, env: TileMatchVecEnv, outputs=("board", "reward", "terminated", "mask", "num_moves_left"),
                 mirror: bool = False):
        """mirror=True binds the pinned arrays as the host mirror (tmg_host_bind): the step kernels read the actions in
        place and write board / mask entries of the envs they change, and reward / terminated / num_moves_left of
        every env, straight into these arrays over PCIe instead of copying every array in full after every step; the
        arrays hold the complete current state after each call either way."""
        self.env = env
        self.mirror = mirror
        N, R, Cc, A = env.num_envs, env.num_rows, env.num_cols, env.num_actions
        shapes = {"actions": ((N,), torch.int32), "board": ((N, 2, R, Cc), torch.int8), "reward": ((N,), torch.int32),
                  "terminated": ((N,), torch.uint8), "mask": ((N, A), torch.uint8),
                  "mask_bits": ((N, (A + 7) // 8), torch.uint8), "num_moves_left": ((N,), torch.int32),
                  "board_packed": ((N, R, Cc), torch.uint8),
                  "is_combination_match": ((N,), torch.uint8), "num_new_specials": ((N,), torch.int32),
                  "num_specials_activated": ((N,), torch.int32), "shuffled": ((N,), torch.uint8),
                  "status": ((N,), torch.int32)}
        self.host = {}
        self.io = nat.HostIO()
        for name in ("actions",) + tuple(outputs):
            shape, dt = shapes[name]
            t = torch.empty(shape, dtype=dt).pin_memory()
            self.host[name] = t
            setattr(self.io, name, t.data_ptr())
        self.h2d_bytes = self.host["actions"].numel() * 4
        self.d2h_bytes = sum(t.numel() * t.element_size() for n, t in self.host.items() if n != "actions")
        if mirror:
            if env._mirror_keep is not None:
                raise RuntimeError("this env already has a host mirror bound; close() the other HostStepper first")
            nat.check(env._lib.tmg_host_bind(env._h, C.byref(self.io), env._stream()), "tmg_host_bind")
            # the env owns the bound arrays from here on: the kernels of ANY later step write into them, whether or not
            # this object is still alive (they are released by close() / env.unbind_host_mirror() / env.close())
            env._mirror_keep = self.host
            self._mirrored = [n for n in ("board", "board_packed", "mask", "mask_bits") if n in self.host]
            scalars = [n for n in ("reward", "terminated", "num_moves_left") if n in self.host]
            # per step over PCIe: the mirrored scalars of every env + one entry of every mirrored array per env that
            # changed + whatever else is copied in full
            self.d2h_bytes_fixed = sum(t.numel() * t.element_size() for n, t in self.host.items()
                                       if n != "actions" and n not in self._mirrored)
            self.d2h_bytes_per_changed_env = sum(self.host[n][0].numel() * self.host[n].element_size() for n in self._mirrored)
            self.d2h_bytes_per_changed_env += 4 if "reward" in scalars else 0

    def close(self):
        """Unbinds the mirror (if any) and waits for the kernels that may still be writing into the pinned arrays."""
        if self.mirror:
            self.mirror = False
            if self.env._mirror_keep is self.host:
                self.env.unbind_host_mirror()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def board(self) -> np.ndarray:
        """The boards as the reference's (N, 2, R, C) int8 planes, from whichever board form was copied back: `board`, or
        `board_packed` (one byte per cell: colour | (type & 7) << 4, cookie type -1 stored as 7)."""
        if "board" in self.host:
            return self.host["board"].numpy()
        pk = self.host["board_packed"].numpy()
        typ = (pk >> 4).astype(np.int8)
        typ[typ == 7] = -1
        return np.stack([(pk & 15).astype(np.int8), typ], axis=1)

    def effective_actions(self, i: int):
        """The reference's info["effective_actions"] list of env i from whichever mask form was copied back."""
        if "mask" in self.host:
            return np.flatnonzero(self.host["mask"][i].numpy()).tolist()
        bits = np.unpackbits(self.host["mask_bits"][i].numpy(), bitorder="little")[:self.env.num_actions]
        return np.flatnonzero(bits).tolist()

    def step(self, actions=None):
        """actions: optional array-like copied into the pinned staging buffer first (int32, (N,))."""
        if actions is not None:
            self.host["actions"].copy_(torch.as_tensor(actions, dtype=torch.int32))
        nat.check(self.env._lib.tmg_step_host(self.env._h, C.byref(self.io), self.env._stream()), "tmg_step_host")
        return self.host


def shard_range(global_num_envs: int, rank: int, world_size: int):
    """Env-index shard owned by `rank`: [rank*N/G, (rank+1)*N/G)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    lo = global_num_envs * rank // world_size
    hi = global_num_envs * (rank + 1) // world_size
    return lo, hi


class EpisodeStatistics:
    """Optional episode statistics accumulated with a handful of elementwise torch ops per step (off the
    transition path), with an on-demand all-reduce over the process group -- the only collective the design has
    (NCCL over NVLink on GPUs, gloo in the CPU tests)."""

    FIELDS = ("episodes", "return_sum", "length_sum", "steps", "specials_created", "specials_activated", "shuffles",
              "combination_matches")

    def __init__(self, num_envs: int, device, autoreset: str = "same_step"):
        """autoreset: the env's mode.  Under "next_step" the call after a termination is the reset step (action ignored,
        reward 0): like gymnasium's vector RecordEpisodeStatistics it is not counted as a step of the new episode."""
        self.device = torch.device(device)
        self.autoreset = autoreset
        self.prev_terminated = torch.zeros(num_envs, dtype=torch.bool, device=self.device)
        self.running_return = torch.zeros(num_envs, dtype=torch.int64, device=self.device)
        self.running_length = torch.zeros(num_envs, dtype=torch.int64, device=self.device)
        self.totals = torch.zeros(len(self.FIELDS), dtype=torch.int64, device=self.device)

    def update(self, reward, terminated, info) -> None:
        r = reward.to(torch.int64)
        term = terminated.to(torch.bool)
        if self.autoreset == "next_step":
            live = (~self.prev_terminated).to(torch.int64)     # the step after a termination only resets the env
            self.prev_terminated = term.clone()
        else:
            live = torch.ones_like(r)
        self.running_return += r * live
        self.running_length += live
        t = self.totals
        t[0] += term.sum()
        t[1] += (self.running_return * term).sum()
        t[2] += (self.running_length * term).sum()
        t[3] += live.sum()
        t[4] += info["num_new_specials"].to(torch.int64).sum()
        t[5] += info["num_specials_activated"].to(torch.int64).sum()
        t[6] += info["shuffled"].to(torch.int64).sum()
        t[7] += info["is_combination_match"].to(torch.int64).sum()
        keep = (~term).to(torch.int64)
        self.running_return *= keep
        self.running_length *= keep

    def allreduce(self, group=None) -> dict:
        """Sum of the totals over all ranks (no-op without an initialised process group)."""
        tot = self.totals.clone()
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
        return dict(zip(self.FIELDS, tot.tolist()))
