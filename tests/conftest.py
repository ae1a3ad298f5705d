import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE_SRC = "/root/reference/src"
HAVE_REFERENCE = os.path.isdir(REFERENCE_SRC)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: imports the read-only reference tree (build container only)")


def pytest_collection_modifyitems(config, items):
    skip_ref = pytest.mark.skip(reason="/root/reference not present on this machine")
    for item in items:
        if "reference" in item.keywords and not HAVE_REFERENCE:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def build_lib():
    """Build libddqc.so if it is missing (nvcc cross-compiles without a GPU)."""
    from data_driven_quad_control_b200.csrc.build import build

    return build()


def default_cfgs(action_type_name="CTBR_FIXED_YAW"):
    """get_cfgs() + the training reward config of the action type (reference: learning/train.py)."""
    from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, get_cfgs
    from data_driven_quad_control_b200.learning.config.reward_config import get_reward_cfg_by_action_type

    env_cfg, obs_cfg, _, command_cfg = get_cfgs()
    at = EnvActionType[action_type_name]
    return env_cfg, obs_cfg, get_reward_cfg_by_action_type(at), command_cfg, at
