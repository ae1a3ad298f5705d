"""TEST INFRASTRUCTURE ONLY (oracle) -- numpy restatement of the actor forward the reference's rollouts
run: rsl_rl `ActorCritic.act_inference` = actor(obs) with actor = nn.Sequential(Linear, Tanh, Linear, Tanh,
Linear) (learning/config/hover_ppo_config.py:30-40; state-dict layout of
models/demo_ctbr_fixed_yaw/model_1500.pt).  Pinned by tests/golden/policy_demo.npz, produced by
oracle/make_golden_policy.py from the reference's own checkpoint run through torch.nn on the CPU."""

from __future__ import annotations

import numpy as np


def actor_forward(obs, W1, b1, W2, b2, W3, b3, dtype=np.float32):
    x = np.asarray(obs, dtype)
    h1 = np.tanh(x @ np.asarray(W1, dtype).T + np.asarray(b1, dtype))
    h2 = np.tanh(h1 @ np.asarray(W2, dtype).T + np.asarray(b2, dtype))
    return h2 @ np.asarray(W3, dtype).T + np.asarray(b3, dtype)
