"""tile_match_gym_b200 -- B200-native (sm_100a) batched implementation of tile-match-gym's board-transition
hot path behind a vectorised TileMatchEnv API.  The compute path is libtmg_b200.so (hand-written CUDA behind
the C ABI of include/tmg_b200.h); importing this package never falls back to a CPU implementation."""
from ._native import build as build_native  # noqa: F401
from .trace import env_from_trace, load_trace, record_trace, replay_trace, save_trace  # noqa: F401
from .vec_env import ENV_ID, EpisodeStatistics, HostStepper, ProportionRewardWrapper, TileMatchVecEnv, shard_range  # noqa: F401

__all__ = ["TileMatchVecEnv", "ProportionRewardWrapper", "HostStepper", "EpisodeStatistics", "shard_range", "build_native", "ENV_ID",
           "record_trace", "replay_trace", "save_trace", "load_trace", "env_from_trace"]
__version__ = "0.1.0"
