"""Fused tcgen05 policy MLP (`ddqc_policy_mlp_forward`) against the numpy actor oracle and the golden actions
torch.nn computes from the reference's own checkpoint (tests/golden/policy_demo.npz)."""

import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "policy_demo.npz")
TOL_SPLIT = 1e-4  # stated tolerance of the bf16 hi/lo split products (measured 3.2e-5 on the demo checkpoint, 4e-6 on
# random-init weights; torch fp32 itself is 1.7e-6 from the fp64 forward)
TOL_BF16 = 5e-2  # single bf16 product: plain bf16 accuracy


def _weights(g):
    return {
        "actor.0.weight": torch.from_numpy(g["W0"]), "actor.0.bias": torch.from_numpy(g["W0_b"]),
        "actor.2.weight": torch.from_numpy(g["W2"]), "actor.2.bias": torch.from_numpy(g["W2_b"]),
        "actor.4.weight": torch.from_numpy(g["W4"]), "actor.4.bias": torch.from_numpy(g["W4_b"]),
    }  # fmt: skip


def test_demo_checkpoint_actions_match_reference_torch_forward(build_lib):
    from data_driven_quad_control_b200.learning.policy import FusedActorMLP

    g = np.load(GOLD)
    pol = FusedActorMLP(_weights(g), precision="bf16x2")
    act = pol.act_inference(torch.from_numpy(g["obs"]).cuda()).cpu().numpy()
    np.testing.assert_allclose(act, g["actions_fp32"], rtol=0, atol=TOL_SPLIT)
    np.testing.assert_allclose(act, g["actions_fp64"], rtol=0, atol=TOL_SPLIT)
    fast = FusedActorMLP(_weights(g), precision="bf16").act_inference(torch.from_numpy(g["obs"]).cuda()).cpu().numpy()
    np.testing.assert_allclose(fast, g["actions_fp32"], rtol=0, atol=TOL_BF16)
    assert np.abs(fast - g["actions_fp32"]).max() > 10 * np.abs(act - g["actions_fp32"]).max()  # the split really matters


@pytest.mark.parametrize("n,num_actions", [(1, 3), (127, 3), (128, 4), (129, 3), (5000, 4), (300001, 3)])
def test_random_init_policy_matches_oracle_on_ragged_batches(build_lib, n, num_actions):
    import policy_oracle as po
    from data_driven_quad_control_b200.learning.policy import FusedActorMLP

    pol = FusedActorMLP.random_init(num_actions=num_actions, seed=n)
    rng = np.random.default_rng(n)
    obs = rng.uniform(-1, 1, (n, 16)).astype(np.float32)
    act = pol.act_inference(torch.from_numpy(obs).cuda()).cpu().numpy()
    ref = po.actor_forward(obs, pol.W1.cpu().numpy(), pol.b1.cpu().numpy(), pol.W2.cpu().numpy(), pol.b2.cpu().numpy(),
                           pol.W3.cpu().numpy(), pol.b3.cpu().numpy(), dtype=np.float64)  # fmt: skip
    assert act.shape == (n, num_actions)
    np.testing.assert_allclose(act, ref, rtol=0, atol=TOL_SPLIT)


def test_policy_drives_the_env(build_lib):
    """obs -> policy -> env.step on the device (BASELINE configs[2] loop): finite, bounded, and identical to
    stepping the env with the oracle's actions for the same observations."""
    import policy_oracle as po
    from data_driven_quad_control_b200.envs.hover_env import HoverEnv
    from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, get_cfgs
    from data_driven_quad_control_b200.learning.config.reward_config import get_reward_cfg_by_action_type
    from data_driven_quad_control_b200.learning.policy import FusedActorMLP

    env_cfg, obs_cfg, _, cmd_cfg = get_cfgs()
    at = EnvActionType.CTBR_FIXED_YAW
    env = HoverEnv(1024, env_cfg, obs_cfg, get_reward_cfg_by_action_type(at), cmd_cfg, device="cuda", action_type=at, seed=3)
    pol = FusedActorMLP.random_init(num_actions=env.num_actions, seed=1)
    env.reset()
    obs, _ = env.get_observations()
    for _ in range(20):
        act = pol.act_inference(obs)
        ref = po.actor_forward(obs.cpu().numpy(), *(t.cpu().numpy() for t in (pol.W1, pol.b1, pol.W2, pol.b2, pol.W3, pol.b3)), dtype=np.float64)
        np.testing.assert_allclose(act.cpu().numpy(), ref, rtol=0, atol=TOL_SPLIT)
        obs, rew, reset, _ = env.step(act)
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()


def test_unsupported_shapes_fail_loudly(build_lib):
    from data_driven_quad_control_b200.learning.policy import FusedActorMLP

    w = {"actor.0.weight": torch.zeros(64, 16), "actor.0.bias": torch.zeros(64), "actor.2.weight": torch.zeros(64, 64),
         "actor.2.bias": torch.zeros(64), "actor.4.weight": torch.zeros(3, 64), "actor.4.bias": torch.zeros(3)}  # fmt: skip
    with pytest.raises(NotImplementedError):
        FusedActorMLP(w)
    with pytest.raises(ValueError):
        FusedActorMLP.random_init().act_inference(torch.zeros(4, 15, device="cuda"))
