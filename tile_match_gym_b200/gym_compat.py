"""gymnasium glue (reference __init__.py:3 registers "TileMatch-v0").  Nothing here is imported by the
package itself: gymnasium is optional (it is absent from the build image), and `TileMatchVecEnv` already follows the
`gymnasium.vector.VectorEnv` calling convention.  With gymnasium installed,

    from tile_match_gym_b200.gym_compat import register
    register()                                         # id "TileMatch-v0", vector entry point
    envs = gymnasium.make_vec("TileMatch-v0", num_envs=65536, num_rows=10, num_cols=10, num_colours=4, num_moves=30,
                              colourless_specials=["cookie"], colour_specials=["vertical_laser", "horizontal_laser", "bomb"])

gives an `isinstance(envs, gymnasium.vector.VectorEnv)` whose arrays stay on the GPU (torch tensors, as documented for
`TileMatchVecEnv`).  The aliasing contract of `TileMatchVecEnv.step` applies to the subclass too: the returned reward /
terminated / info tensors are views of the engine's buffers and are overwritten by the next call -- pass
`copy_outputs=True` for gymnasium-style fresh arrays.

gymnasium is not in the build image, so this module is exercised against a stand-in module only
(tests/test_gym_compat.py); it has not been run against the real package."""
from __future__ import annotations

from .vec_env import ENV_ID, TileMatchVecEnv

_cls = None


def gymnasium_vector_env_class():
    """`TileMatchVecEnv` as a subclass of `gymnasium.vector.VectorEnv` (built on first use; ImportError without gymnasium)."""
    global _cls
    if _cls is not None:
        return _cls
    import gymnasium as gym
    from gymnasium.vector.utils import batch_space

    modes = getattr(gym.vector, "AutoresetMode", None)

    class GymnasiumTileMatchVecEnv(TileMatchVecEnv, gym.vector.VectorEnv):
        def __init__(self, *args, **kwargs):
            TileMatchVecEnv.__init__(self, *args, **kwargs)
            # VectorEnv: `observation_space` / `action_space` are the batched spaces, `single_*` the per-env ones
            self.observation_space = batch_space(self.single_observation_space, self.num_envs)
            self.action_space = batch_space(self.single_action_space, self.num_envs)
            self.metadata = dict(self.metadata)
            if modes is not None:
                self.metadata["autoreset_mode"] = {"next_step": modes.NEXT_STEP, "same_step": modes.SAME_STEP,
                                                   "disabled": modes.DISABLED}[self.autoreset_mode]
            self.closed = False

        def close_extras(self, **kwargs):
            TileMatchVecEnv.close(self)

        def close(self, **kwargs):
            if not getattr(self, "closed", False):
                self.close_extras(**kwargs)
                self.closed = True

    _cls = GymnasiumTileMatchVecEnv
    return _cls


def register(env_id: str = ENV_ID) -> None:
    """`gymnasium.register` with a vector entry point, so that `gymnasium.make_vec(env_id, num_envs=..., ...)` builds
    the B200 env.  The reference registers the same id for its single CPU env (`entry_point`); the two can coexist
    under different ids."""
    import gymnasium as gym
    cls = gymnasium_vector_env_class()
    gym.register(id=env_id, vector_entry_point=lambda **kw: cls(**kw))
