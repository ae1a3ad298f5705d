"""TEST INFRASTRUCTURE -- writes tests/golden/policy_demo.npz from the REFERENCE checkpoint:
actor weights of /root/reference/models/demo_ctbr_fixed_yaw/model_1500.pt, 512 seeded observations in
[-1, 1] (the env clips every observation entry to that range, envs/hover_env.py:597-627) and the actions
torch.nn computes for them on the CPU in fp32 (the arithmetic of rsl_rl's actor) and in fp64."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CKPT = "/root/reference/models/demo_ctbr_fixed_yaw/model_1500.pt"


def main():
    sd = torch.load(CKPT, map_location="cpu", weights_only=False)["model_state_dict"]
    net = torch.nn.Sequential(torch.nn.Linear(16, 128), torch.nn.Tanh(), torch.nn.Linear(128, 128), torch.nn.Tanh(),
                              torch.nn.Linear(128, 3))  # fmt: skip
    net.load_state_dict({k[len("actor."):]: v for k, v in sd.items() if k.startswith("actor.")})
    g = torch.Generator().manual_seed(0)
    obs = torch.rand((512, 16), generator=g) * 2 - 1
    obs[:8] = torch.tensor([[0.0] * 16, [1.0] * 16, [-1.0] * 16] + [[0.5 * (-1) ** (i + j) for j in range(16)] for i in range(5)])
    with torch.no_grad():
        act32 = net(obs).numpy()
        act64 = net.double()(obs.double()).numpy()
    out = {k.replace("actor.", "W").replace(".weight", "").replace(".bias", "_b"): v.numpy() for k, v in sd.items() if k.startswith("actor.")}
    np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "policy_demo.npz"), obs=obs.numpy(), actions_fp32=act32,
                        actions_fp64=act64, **out)
    print({k: v.shape for k, v in out.items()}, np.abs(act32 - act64).max())


if __name__ == "__main__":
    main()
