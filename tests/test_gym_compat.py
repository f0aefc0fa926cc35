"""gym_compat against a stand-in `gymnasium` module -- the real package is not in the image, so this checks
the glue's own logic (subclassing, batched spaces, registration with a vector entry point, close), not gymnasium."""
import sys
import types

import pytest


@pytest.fixture
def fake_gymnasium(monkeypatch):
    gym = types.ModuleType("gymnasium")
    vector = types.ModuleType("gymnasium.vector")
    utils = types.ModuleType("gymnasium.vector.utils")

    class VectorEnv:
        closed = False

    class AutoresetMode:
        NEXT_STEP, SAME_STEP, DISABLED = "NextStep", "SameStep", "Disabled"

    registry = {}
    utils.batch_space = lambda space, n: ("batched", space, n)
    vector.VectorEnv, vector.AutoresetMode, vector.utils = VectorEnv, AutoresetMode, utils
    gym.vector = vector
    gym.register = lambda id, **kw: registry.__setitem__(id, kw)
    gym.registry = registry
    for name, mod in (("gymnasium", gym), ("gymnasium.vector", vector), ("gymnasium.vector.utils", utils)):
        monkeypatch.setitem(sys.modules, name, mod)
    import tile_match_gym_b200.gym_compat as gc
    monkeypatch.setattr(gc, "_cls", None)
    return gym


def test_register_installs_a_vector_entry_point(fake_gymnasium):
    from tile_match_gym_b200 import ENV_ID, TileMatchVecEnv
    from tile_match_gym_b200.gym_compat import gymnasium_vector_env_class, register
    register()
    assert ENV_ID == "TileMatch-v0" and callable(fake_gymnasium.registry[ENV_ID]["vector_entry_point"])
    cls = gymnasium_vector_env_class()
    assert issubclass(cls, TileMatchVecEnv) and issubclass(cls, fake_gymnasium.vector.VectorEnv)
    assert cls is gymnasium_vector_env_class()


def test_without_gymnasium_the_glue_says_so(monkeypatch):
    import tile_match_gym_b200.gym_compat as gc
    monkeypatch.setattr(gc, "_cls", None)
    monkeypatch.setitem(sys.modules, "gymnasium", None)
    with pytest.raises(ImportError):
        gc.gymnasium_vector_env_class()


@pytest.mark.gpu
def test_vector_env_through_the_entry_point(fake_gymnasium):
    import torch
    from tile_match_gym_b200.gym_compat import register
    register("TileMatchB200-v0")
    make = fake_gymnasium.registry["TileMatchB200-v0"]["vector_entry_point"]
    envs = make(num_envs=300, num_rows=10, num_cols=10, num_colours=4, num_moves=5, colourless_specials=["cookie"],
                colour_specials=["vertical_laser", "horizontal_laser", "bomb"], seed=4, autoreset="same_step")
    assert isinstance(envs, fake_gymnasium.vector.VectorEnv)
    assert envs.observation_space == ("batched", envs.single_observation_space, 300)
    assert envs.action_space == ("batched", envs.single_action_space, 300)
    assert envs.metadata["autoreset_mode"] == "SameStep" and "autoreset_mode" not in type(envs).metadata
    obs, info = envs.reset(seed=4)
    assert obs["board"].shape == (300, 2, 10, 10) and info["effective_actions"].shape == (300, 180)
    for _ in range(7):
        a = torch.randint(0, 180, (300,), dtype=torch.int32, device="cuda")
        obs, rew, term, trunc, info = envs.step(a)
    assert rew.shape == (300,) and not bool(trunc.any()) and int((envs.status != 0).sum().item()) == 0
    envs.close(); envs.close()
    assert envs.closed
