"""Builds libddqc.so (sm_100a) in-tree with nvcc.  Used by `__graft_entry__.build()`.

  python -m data_driven_quad_control_b200.csrc.build [--force]

The env / controller translation units are compiled with -fmad=false (bit-exact parity
contract with oracle/env_oracle.py, see ddqc_env.cu); the DD-MPC solver with the default
FMA contraction.
"""

from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libddqc.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I", os.path.join(PKG, "..", "include")]

UNITS = [
    ("ddqc_env.cu", ["-fmad=false"]),
    ("ddqc_env_aux.cu", ["-fmad=false"]),
    ("ddqc_ctrl.cu", ["-fmad=false"]),
    ("ddqc_ddmpc.cu", []),
    ("ddqc_ddmpc_recover.cu", []),
    ("ddqc_policy.cu", []),
]


def _newer(src: str, dst: str) -> bool:
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    hdrs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(PKG, "..", "include", "ddqc.h"))
    relink = force
    for name, extra in UNITS:
        src = os.path.join(HERE, name)
        if not os.path.exists(src):
            continue
        obj = os.path.join(HERE, name.replace(".cu", ".o"))
        if force or _newer(src, obj) or any(_newer(h, obj) for h in hdrs):
            cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            subprocess.check_call(cmd)
            relink = True
        objs.append(obj)
    if relink or not os.path.exists(LIB):
        subprocess.check_call([nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
