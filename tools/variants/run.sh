#!/bin/bash
# developer probe: compare libddqc.so kernel variants on the DD-MPC phase probe (same box, same run)
for f in "$@"; do
  if [ "$f" = base ]; then out=$(python tools/mpc_perf.py 2>&1 | tail -1); else out=$(DDQC_LIB=tools/variants/lib_$f.so python tools/mpc_perf.py 2>&1 | tail -1); fi
  python - "$f" "$out" <<'P'
import sys, json
d = json.loads(sys.argv[2]); k = d["phase_kcycles_per_solve"]
print(sys.argv[1], round(d["solves_per_s"]), "bad", d["status_bad"], {a: k[a] for a in ("gram", "A_factor", "A_update", "B_factor", "B_backsub", "C_factor_schur", "sweep", "qp_assemble", "pivoting")})
P
done
