"""Developer probe (CPU): candidate orders of the rigid-body substep update against the digitised reference figure
(tests/golden/ref_comparison_plot_curves.npz) - the experiment that selected the pose-first order of oracle/drone_physics.py.
The scenario of tests/test_reference_plot_pin.py is flown on the oracle env with `env_oracle.substep` replaced by a float64
variant; per controller (tracking / PPO) and axis: median / 90th percentile / max pixel distance of the published curve to
the flight, the PPO command curves, and the PPO drone's hover jitter.

  python tools/physics_order_fit.py f64 explicit posfirst lag_f64 explicit_pos explicit_quat rotfirst midpoint nogyro
    f64      velocity first, pose with the new velocities (the order used until round 2)
    explicit pose with the old velocities, forces at the old pose
    posfirst pose with the old velocities, then velocities with the forces at the NEW pose   (adopted)
    lag_f64  velocity first with the rotor command delayed by one substep
    base     whatever oracle/drone_physics.py::substep currently is (exact fp32 arithmetic)
"""
import sys, copy, numpy as np
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import env_oracle as eo, ddmpc_oracle as mo, drone_physics as dp, policy_oracle as po
from controllers_oracle import TrackingOracle
F32=np.float32
g=np.load('tests/golden/ref_comparison_plot_curves.npz')
gp=np.load('tests/golden/policy_demo.npz'); W=(gp['W0'],gp['W0_b'],gp['W2'],gp['W2_b'],gp['W4'],gp['W4_b'])

def rot(v,q):  # rotate v by q (arrays)
    return dp.rotate_by_quat(v[0],v[1],v[2],q[0],q[1],q[2],q[3])
def rot_inv(v,q):
    return dp.rotate_by_quat(v[0],v[1],v[2],q[0],-q[1],-q[2],-q[3])
def qnorm(q):
    n=np.sqrt(q[0]**2+q[1]**2+q[2]**2+q[3]**2); return tuple(c/n for c in q)
def dq(h):  # exp of half-rotation vector h
    s=h[0]**2+h[1]**2+h[2]**2; a=np.sqrt(s); c=np.cos(a); k=np.where(a>1e-12,np.sin(a)/np.maximum(a,1e-12),1.0)
    return (c,k*h[0],k*h[1],k*h[2])

def make_substep(kind0):
    def sub(P,wrench,p,q,v,w):
        a_x,a_y,a_z,c_xy,c_z=wrench
        p=[np.asarray(x,np.float64) for x in p]; q=[np.asarray(x,np.float64) for x in q]
        v=[np.asarray(x,np.float64) for x in v]; w=[np.asarray(x,np.float64) for x in w]
        dt=float(P.dt); b=[float(x) for x in P.b_w]
        gy=0.0 if 'nogyro' in kind0 else 1.0
        kind='explicit' if kind0=='explicit_nogyro' else kind0
        wn=[w[0]+a_x-gy*b[0]*w[1]*w[2], w[1]+a_y-gy*b[1]*w[2]*w[0], w[2]+a_z-gy*b[2]*w[0]*w[1]]
        def thrust_dir(q):
            return (q[1]*q[3]+q[0]*q[2], q[2]*q[3]-q[0]*q[1], 1-2*(q[1]**2+q[2]**2))
        def qstep(q,wuse):
            h=[wuse[i]*dt/2 for i in range(3)]
            return qnorm(dp.quat_mul(q[0],q[1],q[2],q[3],*dq(h)))
        if kind=='posfirst':  # pose first with the OLD velocities, then velocities with the force at the NEW pose
            qn=qstep(q,w); pn=[p[i]+dt*v[i] for i in range(3)]
            td=thrust_dir(qn)
            vn=[v[0]+c_xy*td[0], v[1]+c_xy*td[1], v[2]+c_z*td[2]-float(P.dt_g)]
            f=lambda t: tuple(np.asarray(x,np.float32) for x in t)
            return f(pn),f(qn),f(vn),f(wn)
        if 'rotfirst' in kind:   # orientation updated first, thrust along the NEW axis
            qn=qstep(q,wn); td=thrust_dir(qn)
        else:
            td=thrust_dir(q); qn=None
        vn=[v[0]+c_xy*td[0], v[1]+c_xy*td[1], v[2]+c_z*td[2]-float(P.dt_g)]
        if kind=='explicit':
            pn=[p[i]+dt*v[i] for i in range(3)]; qn=qstep(q,w)
        elif kind=='explicit_pos':
            pn=[p[i]+dt*v[i] for i in range(3)]; qn=qstep(q,wn)
        elif kind=='explicit_quat':
            pn=[p[i]+dt*vn[i] for i in range(3)]; qn=qstep(q,w)
        else:
            pn=[p[i]+dt*vn[i] for i in range(3)]
            if qn is None: qn=qstep(q,wn)
        if 'midpoint' in kind:  # thrust direction at the half-step orientation
            qm=qstep(q,[0.5*x for x in wn]); td=thrust_dir(qm)
            vn=[v[0]+c_xy*td[0], v[1]+c_xy*td[1], v[2]+c_z*td[2]-float(P.dt_g)]
            pn=[p[i]+dt*vn[i] for i in range(3)]
        f=lambda t: tuple(np.asarray(x,np.float32) for x in t)
        return f(pn),f(qn),f(vn),f(wn)
    return sub

def make_alllag(inner=None):
    inner = inner or dp.substep
    st={'n':0,'prev':None}
    def sub(P,wrench,p,q,v,w):
        k=st['n']%4; st['n']+=1
        use=wrench
        if k==0:
            use=st['prev'] if st['prev'] is not None else wrench
            st['prev']=wrench
        return inner(P,use,p,q,v,w)
    return sub
def scenario(kind, which):
    eo.substep = dp.substep if kind=='base' else (make_alllag() if kind=='alllag' else (make_alllag(make_substep(kind[4:])) if kind.startswith('lag_') else make_substep(kind)))
    env_cfg = dict(dt=0.01, decimation=4, simulate_action_latency=True, clip_actions=1.0, actuator_noise_std=0.0,
        termination_if_roll_greater_than=180, termination_if_pitch_greater_than=180, termination_if_close_to_ground=0.0, termination_if_x_greater_than=100.0,
        termination_if_y_greater_than=100.0, termination_if_z_greater_than=10.0, termination_if_ang_vel_greater_than=20, termination_if_lin_vel_greater_than=20,
        base_init_pos=[0.0, 0.0, 1.0], base_init_quat=[1.0, 0.0, 0.0, 0.0], episode_length_s=100, at_target_threshold=0.1, min_hover_time_s=0.5, resampling_time_s=3.0)
    obs_cfg = dict(obs_scales=dict(rel_pos=1 / 3.0, lin_vel=1 / 3.0, ang_vel=1 / 3.14159), obs_noise_std=0.0)
    rew = dict(yaw_lambda=-10.0, reward_scales=dict(target=10.0, closeness=1.5, hover_time=0.01, smooth=-0.01, yaw=0.0, angular=-2e-4, crash=-10.0))
    cmd = dict(num_commands=3, pos_x_range=[-1.0, 1.0], pos_y_range=[-1.0, 1.0], pos_z_range=[1.0, 3.0])
    mixer, isk = mo.ctbr_mixer()
    env = eo.EnvOracle(1, env_cfg, obs_cfg, copy.deepcopy(rew), cmd, action_type=eo.CTBR_FIXED_YAW, auto_target_updates=False, mixer=mixer, inv_sqrt_kf=isk, rng=eo.PhiloxRng(1))
    env.reset()
    ctrl = TrackingOracle(0.027, [[2.0, 0.0, 2.2], [2.0, 0.0, 2.2], [18.0, 0.0, 6.0]], 5.0, 1, env.step_dt)
    q_sp = np.array([[1.0, 0.0, 0.0, 0.0]], np.float32)
    lo, span = env.act_lo.astype(np.float64), env.act_span.astype(np.float64)
    def track(target, steps, log):
        tgt=np.array([target],np.float32); env.update_target_pos([0],tgt)
        for _ in range(steps):
            c = ctrl.compute(env.base_pos.astype(np.float32), env.base_quat, tgt, q_sp)[:, :3].astype(np.float64)
            a = -1.0 + (c - lo) * 2.0 / span
            env.step(np.clip(a, -1, 1).astype(np.float32))
            if log is not None: log.append(np.concatenate([env.base_pos[0].copy(), (lo+(np.clip(a,-1,1)+1)*span/2)[0]]))
    track([0,0,1.5],500,None)
    log=[]
    if which=='tracking':
        track([1,-1,2.5],175,log); track([0,0,1.5],175,log)
    else:
        obs=env.obs_buf.copy()
        for target in ([1,-1,2.5],[0,0,1.5]):
            env.update_target_pos([0],np.array([target],np.float32))
            for _ in range(175):
                a=po.actor_forward(obs,*W)
                ac=np.clip(a,-1,1)
                env.step(ac.astype(np.float32)); obs=env.obs_buf.copy()
                log.append(np.concatenate([env.base_pos[0].copy(), (lo+(ac.astype(np.float64)+1)*span/2)[0]]))
    return np.array(log)

def pxdist(pts, tt, vv, pxs, pxm):
    tf=np.linspace(tt[0],tt[-1],len(tt)*20); vf=np.interp(tf,tt,vv)
    P=np.stack([tf*pxs,vf*pxm],1); Q=np.stack([pts[:,0]*pxs,pts[:,1]*pxm],1)
    return np.sqrt(((Q[:,None,:]-P[None,:,:])**2).sum(2)).min(1)
tt=np.arange(350)*0.04
if __name__=='__main__':
    for kind in sys.argv[1:]:
        for which,key in (('tracking','tracking'),('rl','rl')):
            pos=scenario(kind,which)
            out=[]
            for i,ax in enumerate('xyz'):
                d=pxdist(g[f'{ax}_{key}'], tt, pos[:,i], float(g[f'{ax}_px_per_s']), float(g[f'{ax}_px_per_m']))
                out.append('%s med %.2f p90 %.2f max %.1f'%(ax,np.median(d),np.percentile(d,90),d.max()))
            for i,ax in enumerate(('thrust','roll','pitch')):
                d=pxdist(g[f'{ax}_{key}'], tt, pos[:,3+i], float(g[f'{ax}_px_per_s']), float(g[f'{ax}_px_per_m']))
                out.append('%s med %.2f p90 %.2f'%(ax,np.median(d),np.percentile(d,90)))
            jit=pos[100:170,:3].std(0)
            print(kind,which,' | '.join(out),' jitter(4-6.8s) std mm',np.round(jit*1000,1),flush=True)
