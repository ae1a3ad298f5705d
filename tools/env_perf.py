"""Developer probe: time env_step for the kernel variant selected with DDQC_LIB."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_env
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
env = make_env(N, torch.device("cuda"), 1)
acts = torch.rand((16, N, 3), device="cuda") * 2 - 1
env.reset()
for t in range(20): env.step(acts[t % 16])
torch.cuda.synchronize()
K = 200
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(K): env.step(acts[t % 16])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({"lib": os.environ.get("DDQC_LIB", "default"), "N": N, "ms": ms, "env_steps_per_s": N / ms * 1e3, "GBps": 357 * N / ms / 1e6}))
