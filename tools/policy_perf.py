"""Developer probe: fused policy MLP throughput (1M envs) and policy-in-the-loop rollout."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench
from data_driven_quad_control_b200.learning.policy import FusedActorMLP
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device("cuda")
obs = torch.rand((N, 16), device=dev) * 2 - 1
out = {}
for prec in ("bf16x2", "bf16"):
    pol = FusedActorMLP.random_init(3, 0, dev, prec)
    for _ in range(5): pol.act_inference(obs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): pol.act_inference(obs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    prods = 3 if prec == "bf16x2" else 1
    flops = N * (prods * 2 * (16 * 128 + 128 * 128) + 2 * 128 * 3)
    out[prec] = {"ms": ms, "env_per_s": N / ms * 1e3, "tflops": flops / ms / 1e9}
# torch reference (cuBLAS fp32) for comparison
net = torch.nn.Sequential(torch.nn.Linear(16, 128), torch.nn.Tanh(), torch.nn.Linear(128, 128), torch.nn.Tanh(), torch.nn.Linear(128, 3)).to(dev)
with torch.no_grad():
    for _ in range(3): net(obs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): net(obs)
    e1.record(); torch.cuda.synchronize()
out["torch_fp32_cublas_ms"] = e0.elapsed_time(e1) / 20
env = bench.make_env(N, dev, 1)
env.reset(); o, _ = env.get_observations()
pol = FusedActorMLP.random_init(3, 0, dev, "bf16x2")
for _ in range(10): o, *_ = env.step(pol.act_inference(o))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): o, *_ = env.step(pol.act_inference(o))
e1.record(); torch.cuda.synchronize()
out["rollout_ms_per_step"] = e0.elapsed_time(e1) / 100
out["rollout_env_steps_per_s"] = N / out["rollout_ms_per_step"] * 1e3
print(json.dumps(out))
