"""Developer probe: the sequence of tests/test_env_parity_gpu.py::test_state_save_restore_matches_oracle with a
field-by-field comparison after every step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from test_env_parity_gpu import _make_pair, _actions
N = 6
env, orc = _make_pair(N, "CTBR", False)
acts = _actions(40, N, 4, 7, gentle_frac=1.0) * 10
env.reset(); orc.reset()
def cmp(t):
    bad = []
    for name, a, b in (("pos", env.base_pos, orc.base_pos), ("quat", env.base_quat, orc.base_quat), ("lin", env.base_lin_vel, orc.base_lin_vel),
                       ("ang", env.base_ang_vel, orc.base_ang_vel), ("reset", env.reset_buf, orc.reset_buf), ("obs", env.obs_buf, orc.obs_buf),
                       ("rew", env.rew_buf, orc.rew_buf), ("eplen", env.episode_length_buf, orc.episode_length_buf)):
        a = a.cpu().numpy()
        if not np.array_equal(a, b): bad.append((name, np.argwhere(a != b)[:3].tolist()))
    print(t, "ok" if not bad else bad, flush=True)
    return bad
for t in range(8):
    env.step(torch.from_numpy(acts[t]).cuda()); orc.step(acts[t]); cmp(t)
idx = [2]
s_env = env.get_current_state(torch.tensor(idx, device="cuda")); s_orc = orc.get_current_state(idx)
for t in range(8, 16):
    env.step(torch.from_numpy(acts[t]).cuda()); orc.step(acts[t]); cmp(t)
env.restore_from_state(torch.tensor(idx, device="cuda"), s_env); orc.restore_from_state(idx, s_orc)
print("after restore"); cmp(-1)
for t in range(16, 20):
    print("pre-step gpu: pos", env._pos[:, 2].cpu().numpy(), "quat", env._quat[:, 2].cpu().numpy(), "vel", env._vel[:, 2].cpu().numpy(), "ang", env._ang[:, 2].cpu().numpy(),
          "pidI", env._pid_integral[:, 2].cpu().numpy(), "pidE", env._pid_prev_error[:, 2].cpu().numpy(), "lact", env._last_actions[:, 2].cpu().numpy(), "eplen", env._episode_length[2].item() if hasattr(env, "_episode_length") else None)
    print("pre-step orc: pos", orc.sim_pos[2], "quat", orc.sim_quat[2], "vel", orc.sim_vel[2], "ang", orc.sim_ang[2], "pidI", orc.pid_integral[2], "pidE", orc.pid_prev_error[2], "lact", orc.last_actions[2], "bav", orc.base_ang_vel[2])
    print("cmd gpu", env._commands[:, 2].cpu().numpy(), "cmd orc", orc.commands[2], "act", acts[t][2])
    print("gpu bav", env._bav[:, 2].cpu().numpy() if env._bav is not None else None)
    env.step(torch.from_numpy(acts[t]).cuda()); orc.step(acts[t])
    print("crash gpu", env._crash.cpu().numpy(), "orc", orc.crash_condition, "rpm gpu", env._rpm[:, 2].cpu().numpy(), "rpm orc", orc.last_rpm[2], "euler gpu", env._base_euler[:, 2].cpu().numpy() if env._base_euler is not None else None, "orc euler", orc.base_euler[2])
    if cmp(t):
        for name, a, b in (("pos", env.base_pos, orc.base_pos), ("quat", env.base_quat, orc.base_quat), ("lin", env.base_lin_vel, orc.base_lin_vel), ("ang", env.base_ang_vel, orc.base_ang_vel)):
            print(name, a.cpu().numpy()[2], b[2])
        print("sim vel", env._vel[:, 2].cpu().numpy(), orc.sim_vel[2], "sim ang", env._ang[:, 2].cpu().numpy(), orc.sim_ang[2])
        break
