"""Actor MLP of the reference's PPO policies on the fused tcgen05 kernel (`ddqc_policy_mlp_forward`).

The reference trains / plays `rsl_rl.modules.ActorCritic` policies (learning/config/hover_ppo_config.py:
30-40: actor_hidden_dims [128, 128], activation "tanh"); rollouts call `policy.act_inference(obs)`.
`FusedActorMLP` is that inference path for the shipped shape (num_obs 16 -> 128 -> 128 -> num_actions):
the whole forward is one persistent kernel per call, hidden activations never leave the SM.
"""

from __future__ import annotations

import torch

from data_driven_quad_control_b200 import _lib

PRECISIONS = {"bf16x2": 0, "bf16": 1}


class FusedActorMLP:
    """`act_inference(obs) -> actions` for actor = Linear(16,128)-Tanh-Linear(128,128)-Tanh-Linear(128,A)."""

    def __init__(self, weights: dict[str, torch.Tensor], device="cuda", precision: str = "bf16x2"):
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self._lib = _lib.load()
        self.device = _lib.require_cuda(device)
        self.precision = precision
        f32 = dict(dtype=torch.float32, device=self.device)
        get = lambda k: weights[k].detach().to(**f32).contiguous()  # noqa: E731
        self.W1, self.b1 = get("actor.0.weight"), get("actor.0.bias")
        self.W2, self.b2 = get("actor.2.weight"), get("actor.2.bias")
        self.W3, self.b3 = get("actor.4.weight"), get("actor.4.bias")
        self.num_obs, self.hidden, self.num_actions = self.W1.shape[1], self.W1.shape[0], self.W3.shape[0]
        if self.W2.shape != (self.hidden, self.hidden) or self.W3.shape[1] != self.hidden:
            raise ValueError("inconsistent actor weight shapes")
        if (self.num_obs, self.hidden) != (16, 128) or not 1 <= self.num_actions <= 4:
            raise NotImplementedError("the fused kernel supports the shipped actor shape 16 -> 128 -> 128 -> A (A <= 4) only")

    @classmethod
    def from_checkpoint(cls, path: str, device="cuda", precision: str = "bf16x2") -> "FusedActorMLP":
        """rsl_rl checkpoint (`model_state_dict` with actor.{0,2,4}.{weight,bias}), e.g. models/demo_ctbr_fixed_yaw/model_1500.pt."""
        sd = torch.load(path, map_location="cpu", weights_only=True)  # tensors only: no pickle code from a checkpoint
        return cls(sd.get("model_state_dict", sd), device, precision)

    @classmethod
    def random_init(cls, num_actions: int = 3, seed: int = 0, device="cuda", precision: str = "bf16x2") -> "FusedActorMLP":
        """Random-init actor with torch's default nn.Linear initialisation (BASELINE configs[2])."""
        torch.manual_seed(seed)
        net = torch.nn.Sequential(torch.nn.Linear(16, 128), torch.nn.Tanh(), torch.nn.Linear(128, 128), torch.nn.Tanh(),
                                  torch.nn.Linear(128, num_actions))  # fmt: skip
        return cls({f"actor.{k}": v for k, v in net.state_dict().items()}, device, precision)

    def act_inference(self, observations: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """Actions for `observations` (N, 16).  Like the reference's `policy.act_inference`, every call returns a fresh
        tensor; pass `out=` (float32, (N, A), contiguous, on this device) to reuse a buffer, e.g. inside a CUDA graph."""
        obs = observations
        if obs.dtype != torch.float32 or obs.device != self.device or not obs.is_contiguous():
            obs = obs.to(device=self.device, dtype=torch.float32).contiguous()
        if obs.ndim != 2 or obs.shape[1] != self.num_obs:
            raise ValueError(f"observations must have shape (N, {self.num_obs})")
        n = obs.shape[0]
        if out is None:
            out = torch.empty((n, self.num_actions), dtype=torch.float32, device=self.device)
        elif out.shape != (n, self.num_actions) or out.dtype != torch.float32 or out.device != self.device or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 tensor of shape {(n, self.num_actions)} on {self.device}")
        with _lib.device_guard(self.device):
            rc = self._lib.ddqc_policy_mlp_forward(
                obs.data_ptr(), n, self.num_obs, self.hidden, self.num_actions, self.W1.data_ptr(), self.b1.data_ptr(),
                self.W2.data_ptr(), self.b2.data_ptr(), self.W3.data_ptr(), self.b3.data_ptr(), PRECISIONS[self.precision],
                out.data_ptr(), _lib.stream_ptr(self.device),
            )  # fmt: skip
        _lib.check(rc, "ddqc_policy_mlp_forward")
        return out

    __call__ = act_inference
