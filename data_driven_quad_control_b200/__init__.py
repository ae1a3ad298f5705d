"""B200-native (sm_100a) hot path of pavelacamposp/data-driven-quad-control.

Python host mirrors of the reference's classes for the env-step / DD-MPC path; all compute
is done by hand-written CUDA kernels in `libddqc.so` (C ABI: `include/ddqc.h`).  There is no
CPU fallback: constructing an env / controller without the library or without a CUDA device
raises.
"""

__version__ = "0.1.0"
