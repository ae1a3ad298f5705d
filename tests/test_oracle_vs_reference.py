"""Build-container-only tests (need /root/reference): the restated Genesis physics satisfies the
reference's own behavioural pins, and the env oracle tracks the reference's hover_env.py live."""

import copy
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "genesis_shim")


def test_reference_own_tests_pass_on_the_shim():
    """tests/env_tests, tests/controller_tests, test_math_utils, test_drone_utilities of the reference
    (incl. the Genesis-marked integration tests: hover |w| < 0.1 after 10 steps, rate error <= 1e-3
    after 20 steps, two drones reach their targets, save/restore bit equality) run unmodified."""
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([SHIM, os.path.join(ROOT, "oracle"), "/root/reference/src", "/root/reference"]))
    cmd = [sys.executable, "-m", "pytest", "-p", "no:cacheprovider", "-q", "/root/reference/tests/env_tests",
           "/root/reference/tests/controller_tests", "/root/reference/tests/utility_tests/test_math_utils.py",
           "/root/reference/tests/utility_tests/test_drone_utilities.py"]  # fmt: skip
    out = subprocess.run(cmd, cwd="/tmp", env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "21 passed" in out.stdout


@pytest.mark.parametrize("action_type", [2, 0])
def test_env_oracle_tracks_reference_live(action_type):
    sys.path[:0] = [SHIM, "/root/reference/src"]
    import torch
    from data_driven_quad_control.envs.hover_env import HoverEnv
    from data_driven_quad_control.envs.hover_env_config import EnvActionType, get_cfgs

    import env_oracle as eo

    N, T = 48, 300
    cfgs = get_cfgs()
    ref = HoverEnv(N, *copy.deepcopy(cfgs), device="cpu", action_type=EnvActionType(action_type))
    kw = {}
    if ref.uses_ctbr_actions:
        kw = dict(mixer=ref.ctbr_controller.mixer.numpy(), inv_sqrt_kf=float(ref.ctbr_controller.inv_sqrt_KF))
    orc = eo.EnvOracle(N, *copy.deepcopy(cfgs), action_type=action_type, rng=eo.TorchRng(), **kw)
    acts = np.random.RandomState(9).uniform(-1, 1, (T, N, ref.num_actions)).astype(np.float32)
    torch.manual_seed(3)
    ref.reset()
    rec = []
    for t in range(T):
        o, r, d, ex = ref.step(torch.from_numpy(acts[t]))
        rec.append((o.numpy().copy(), r.numpy().copy(), d.numpy().copy()))
    torch.manual_seed(3)
    orc.reset()
    for t in range(T):
        orc.step(acts[t])
        assert np.array_equal(orc.reset_buf, rec[t][2]), t
        np.testing.assert_allclose(orc.obs_buf, rec[t][0], rtol=0, atol=0 if action_type == 0 else 2e-3)
        np.testing.assert_allclose(orc.rew_buf, rec[t][1], rtol=0, atol=1e-6 if action_type == 0 else 1e-3)
