import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the Python reference under /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isfile("/root/reference/src/tile_match_gym/board.py")
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="reference tree not present on this machine"))
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))


import numpy as np  # noqa: E402


def expected_policy_actions(o, seed, env_id_offset, num_moves, policy):
    """The action tmg_rollout_policy must take for every env of the oracle's current state (include/tmg_b200.h)."""
    from oracle.stream import mulhi32, stream_words
    out = np.zeros(o.N, np.int32)
    for e in range(o.N):
        t = int(o.timer[e])
        if t < 0 or t >= num_moves:
            continue                                   # no action is taken (recorded as 0)
        k = int(o.episode[e]) * num_moves + t
        w = stream_words(seed, env_id_offset + e, 2, k, 1)
        idx = np.flatnonzero(o.mask[e]) if policy == "mask" else np.arange(0)
        out[e] = idx[int(mulhi32(w, len(idx))[0])] if len(idx) else int(mulhi32(w, o.A)[0])
    return out
