"""GPU parity of the alternative launch shapes of ddqc_env_step against the numpy oracle.  Besides the default kernels
(one thread per env: fused finalisation up to 32 768 envs, split finalisation + programmatic dependent launch above)
the library holds three experimental ones, all bit-exact, all measured SLOWER on B200 and therefore off unless asked for
(DESIGN.md section 3 has the numbers and the ncu evidence):
  pair       two envs per thread on packed fp32x2 arithmetic (FFMA2 / FADD2), 64-bit state accesses
  tile       persistent CTAs, the next tile's state staged through shared memory by cp.async.bulk + mbarrier
  pair_tile  both
`ddqc_env_step` reads the thresholds DDQC_ENV_{PAIR,TILE,PAIR_TILE}_MIN_N once per process, so every case runs in a
subprocess that sets them (tests/tile_kernel_check.py)."""

import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


KINDS = {  # environment that forces one launch shape of ddqc_env_step at small batch sizes
    "tile": dict(DDQC_ENV_TILE_MIN_N="256", DDQC_ENV_FUSED_MAX_N="0", DDQC_ENV_PAIR_TILE_MIN_N=str(1 << 40)),
    "pair_tile": dict(DDQC_ENV_PAIR_TILE_MIN_N="256", DDQC_ENV_FUSED_MAX_N="0"),
    "pair": dict(DDQC_ENV_PAIR_MIN_N="2", DDQC_ENV_FUSED_MAX_N="0", DDQC_ENV_TILE_MIN_N=str(1 << 40), DDQC_ENV_PAIR_TILE_MIN_N=str(1 << 40)),
}


def _run(at_name, N, T, auto=True, independent=False, park=False, kind="tile"):
    env = dict(os.environ, **KINDS[kind])
    cmd = [sys.executable, os.path.join(HERE, "tile_kernel_check.py"), at_name, str(N), str(T), *("1" if f else "0" for f in (auto, independent, park))]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("kind", ["pair_tile", "pair", "tile"])
@pytest.mark.parametrize("at_name", ["CTBR_FIXED_YAW", "CTBR", "ROTOR_RPMS"])
def test_bit_exact_with_ragged_tail(build_lib, at_name, kind):
    """1300 envs (tile kernel: 5 full 256-env tiles + a tail of 20; pair kernel: 650 pairs, last CTA partly idle), 250
    random-action steps, ~2.5k resets, an external reset_idx in the middle: every output identical to the oracle,
    statistics and extras["episode"] included."""
    res = _run(at_name, 1300, 250, kind=kind)
    assert res["resets"] > 1500


def test_pair_kernel_hover_fixups_and_independent_envs(build_lib):
    """The per-lane paths of the pair kernel: parked drones (hover counter, Philox target resample of one env of a pair,
    A15/A16 fix-up records), then the independent-env mode without auto targets."""
    res = _run("CTBR_FIXED_YAW", 4098, 120, park=True, kind="pair")
    assert res["resets"] > 1000 and res["hover_events"] > 300
    res = _run("CTBR", 2050, 200, auto=False, independent=True, kind="pair")
    assert res["resets"] > 1000


def test_pair_kernel_odd_batch_falls_back(build_lib):
    res = _run("CTBR_FIXED_YAW", 1301, 60, kind="pair")
    assert res["resets"] > 300


@pytest.mark.parametrize("kind", ["pair_tile", "tile"])
def test_tile_kernel_many_tiles_per_stream_and_hover_fixups(build_lib, kind):
    """148 x 4 tile streams: 166 404 envs = 650 tiles + a tail of 4 -> some streams walk two tiles (buffer refill under
    the arithmetic, mbarrier phase flips); parked drones produce hover / resample events and A15/A16 fix-ups on steps
    with resets elsewhere."""
    res = _run("CTBR_FIXED_YAW", 166404, 40, park=True, kind=kind)
    assert res["resets"] > 1000 and res["hover_events"] > 1000


@pytest.mark.parametrize("kind", ["pair_tile", "tile"])
def test_tile_kernel_independent_envs_and_no_auto_targets(build_lib, kind):
    res = _run("CTBR_FIXED_YAW", 2048 + 4, 200, auto=False, independent=True, kind=kind)
    assert res["resets"] > 1000


def test_unaligned_batch_falls_back_to_the_plain_kernel(build_lib):
    """n % 4 != 0: the rows of the struct-of-arrays state are not 16-byte aligned, the bulk copies cannot be used and
    the step runs on the one-thread-per-env kernel - same results."""
    res = _run("CTBR_FIXED_YAW", 1301, 60, kind="pair_tile")
    assert res["resets"] > 300


def test_default_launch_shape_ragged_full_size(build_lib):
    """2**20 - 76 envs on the default launch shape (one thread per env, split finalisation), 2 steps."""
    from test_env_parity_gpu import _compare_step, _make_pair

    N, T = (1 << 20) - 76, 2
    env, orc = _make_pair(N, "CTBR_FIXED_YAW", True, materialize=False)
    acts = np.random.RandomState(12).uniform(-1.0, 1.0, (T, N, 3)).astype(np.float32)
    env.reset()
    orc.reset()
    for t in range(T):
        env.step(torch.from_numpy(acts[t]).cuda())
        orc.step(acts[t])
        _compare_step(env, orc, t)
