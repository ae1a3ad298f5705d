#!/bin/bash
# Read-only probe of the GPU box for the reference's third-party stack.
# It installs and downloads NOTHING: import checks, offline wheelhouse / site-packages scan, and a
# DNS + TCP reachability check.  (Installing genesis-world / direct-data-driven-mpc from their git
# repositories would build and run external code and needs a human's authorisation; not attempted.)
out=gpurun_out/r2_reference_stack_probe.log
mkdir -p gpurun_out
{
echo "== date: $(date -u +%FT%TZ)"
echo "== nvidia-smi -L"; nvidia-smi -L
echo "== lscpu"; lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket|NUMA node\(s\)'
echo "== python / pip"; python --version; python -m pip --version
echo "== importable already?"
python - <<'PY'
import importlib
for m in ["genesis", "taichi", "gstaichi", "cvxpy", "osqp", "clarabel", "scs", "direct_data_driven_mpc", "rsl_rl", "mujoco", "qpsolvers"]:
    try:
        importlib.import_module(m); print(f"  import {m}: OK")
    except Exception as e:
        print(f"  import {m}: {type(e).__name__}: {e}")
PY
echo "== network reachability (DNS + TCP connect only, 5 s limits; nothing is fetched)"
for h in pypi.org files.pythonhosted.org github.com; do
  timeout 8 python - "$h" <<'PY'
import socket, sys
h = sys.argv[1]
try:
    a = socket.getaddrinfo(h, 443)[0][4]
    s = socket.create_connection(a[:2], timeout=5); s.close(); print(f"  {h}: tcp/443 reachable ({a[0]})")
except Exception as e:
    print(f"  {h}: {type(e).__name__}: {e}")
PY
done
echo "== offline wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i -E 'cvxpy|osqp|genesis|taichi|clarabel|scs|mujoco|rsl' || echo "  none of cvxpy/osqp/genesis/taichi/clarabel/scs/mujoco/rsl in /opt/wheelhouse"
echo "== site-packages scan"; python - <<'PY'
import site, os
hits = []
for sp in site.getsitepackages():
    for n in os.listdir(sp):
        if any(k in n.lower() for k in ("cvxpy", "osqp", "genesis", "taichi", "clarabel", "scs", "mujoco", "rsl_rl", "direct_data")):
            hits.append(os.path.join(sp, n))
print("  hits:", hits or "none")
PY
echo "== install step: NOT ATTEMPTED (needs human authorisation to build external git code)"
} > "$out" 2>&1
cat "$out"
