"""Developer probe: policy-MLP kernel variant (DDQC_LIB) - forward time at 2^20 envs and max |action - float64 forward|
with the demo checkpoint weights (tests/golden/policy_demo.npz) scaled x1 and x4 (the latter drives tanh into saturation)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from data_driven_quad_control_b200.learning.policy import FusedActorMLP
dev = torch.device("cuda")
g = np.load(os.path.join(ROOT, "tests", "golden", "policy_demo.npz"))
out = {"lib": os.environ.get("DDQC_LIB", "default")}
for scale in (1.0, 4.0):
    w = {f"actor.{k}.{n}": torch.from_numpy(g[f"W{k}" + ("" if n == "weight" else "_b")] * (scale if k == 0 else 1.0)) for k in (0, 2, 4) for n in ("weight", "bias")}
    pol = FusedActorMLP(w, dev)
    rng = np.random.default_rng(1)
    obs = rng.uniform(-1, 1, (8192, 16)).astype(np.float32)
    W = {k: v.double().numpy() for k, v in w.items()}
    h = np.tanh(obs.astype(np.float64) @ W["actor.0.weight"].T + W["actor.0.bias"])
    h = np.tanh(h @ W["actor.2.weight"].T + W["actor.2.bias"])
    ref = h @ W["actor.4.weight"].T + W["actor.4.bias"]
    act = pol.act_inference(torch.from_numpy(obs).to(dev)).cpu().numpy()
    out[f"max_abs_err_x{scale:g}"] = float(np.abs(act - ref).max())
N = 1 << 20
obs = torch.rand((N, 16), device=dev) * 2 - 1
pol = FusedActorMLP.random_init(3, 0, dev, "bf16x2")
best = 1e9
for rep in range(3):
    for _ in range(5): pol.act_inference(obs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): pol.act_inference(obs)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 50)
out["ms"] = best
print(json.dumps(out))
