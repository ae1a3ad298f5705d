"""Developer tool: per-kernel counts of the SASS mnemonics that prove which hardware paths a kernel uses
(B200_PROFILING.md: UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, DMMA = FP64 tensor, UTMALDG/UTMASTG = TMA tensor
copies, UBLKCP = cp.async.bulk, LDGSTS = cp.async, ACQBULK/ARRIVES/SYNCS = mbarrier, FFMA2 = packed fp32x2).
  python tools/sass_counts.py > profiles/r2_sass_counts.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "data_driven_quad_control_b200", "libddqc.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "DMMA", "HMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "FFMA2",
       "FADD2", "FMUL2", "MUFU", "DFMA", "LDG", "STG", "LDS", "STS", "ATOMG", "RED", "MEMBAR", "BAR", "SHFL", "LDL", "STL", "ACQBULK"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    out = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            cur = re.sub(r"\(.*", "", cur)
            out[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            out[cur]["_total"] += 1
            for p in PAT:
                if op == p or op.startswith(p + "."):
                    out[cur][p] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a; static instruction counts per kernel)")
    for k, c in out.items():
        parts = [f"{p}={c[p]}" for p in PAT if c[p]]
        print(f"{k}\n    total={c['_total']}  " + "  ".join(parts))


if __name__ == "__main__":
    sys.exit(main())
