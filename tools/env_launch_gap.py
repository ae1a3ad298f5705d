"""Developer probe: env-step throughput with per-step events, without, and as a CUDA graph replay."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, bench
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20; dev = torch.device("cuda")
env = bench.make_env(N, dev, 1); env.reset()
acts = torch.rand((16, N, 3), device=dev) * 2 - 1
for t in range(50): env.step(acts[t % 16])
K = 1000; out = {}
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for t in range(K): env.step(acts[t % 16])
e1.record(); t_host = time.perf_counter() - t0; torch.cuda.synchronize()
out["no_events_us_per_step"] = e0.elapsed_time(e1) / K * 1e3; out["host_issue_us_per_step"] = t_host / K * 1e6
ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
ev[0].record()
for t in range(K):
    env.step(acts[t % 16]); ev[t + 1].record()
torch.cuda.synchronize()
out["with_events_us_per_step"] = ev[0].elapsed_time(ev[K]) / K * 1e3
try:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for t in range(3): env.step(acts[t % 16])
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for t in range(16): env.step(acts[t])
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(60): g.replay()
    e1.record(); torch.cuda.synchronize()
    out["graph_us_per_step"] = e0.elapsed_time(e1) / (60 * 16) * 1e3
except Exception as ex:
    out["graph_error"] = repr(ex)[:300]
out["N"] = N; out["lib"] = os.environ.get("DDQC_LIB", "default"); out["fused_max_n"] = os.environ.get("DDQC_ENV_FUSED_MAX_N", "default")
print(json.dumps(out))
