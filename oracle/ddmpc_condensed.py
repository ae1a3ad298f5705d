"""TEST INFRASTRUCTURE ONLY -- numpy float64 mirror of the CONDENSED algorithm the CUDA solver
(`data_driven_quad_control_b200/csrc/ddqc_ddmpc.cu`) implements, kept to (a) cross-check the
algebra against the literal-QP oracle (`oracle/ddmpc_oracle.py`) on the CPU and (b) localise a
GPU discrepancy to a phase.  Not used by the product.

Condensation (derivation in DESIGN.md "DD-MPC solve"):  with H = [H_u; H_y; 1^T] (r x C rows),
G = H H^T and v = H alpha, everything except lamb_alpha|alpha - alpha_sr|^2 depends on alpha only
through v, and for fixed v the minimal regulariser is lamb_alpha (v - v_sr)^T G^-1 (v - v_sr).
Eliminating sigma / ybar row-wise turns every row of H into either a HARD row (v_i prescribed:
past inputs, terminal inputs = u_s, the ones row, and the box-constrained predicted inputs y_B
which become decision variables) or a SOFT row with weight w_i (outputs).  Minimising over alpha
for fixed targets is a ridge regression in dual form,

    J(y_B, u_s, y_s) = lamb_alpha rho^T (G + D0)^-1 rho + sum_B R_i (y_Bi - u_s,i)^2 + g(u_s, y_s),
    rho = E_B y_B + T [u_s; y_s] + tau - v_sr,   D0 = diag(lamb_alpha / w_i) (0 on hard rows),

i.e. a strictly convex BOX-constrained QP in (L-n) m + m + p variables (93 at the default
sizes) whose only constraints are simple bounds.  It is solved exactly by block principal
pivoting on the active bounds.  alpha itself (309 numbers) is never formed.
"""

from __future__ import annotations

import numpy as np

from ddmpc_oracle import APPROXIMATED, MPCParams


def hankel_gram(u_win, y_win, ell):
    """G = H H^T for H = [H_ell(u); H_ell(y); 1^T], built from the Hankel structure:
    G[(j,a),(k,b)] = sum_c w_a[j+c] w_b[k+c]."""
    N, m = u_win.shape
    p = y_win.shape[1]
    C = N - ell + 1
    H = np.empty(((m + p) * ell + 1, C))
    for j in range(ell):
        H[j * m : (j + 1) * m] = u_win[j : j + C].T
        H[m * ell + j * p : m * ell + (j + 1) * p] = y_win[j : j + C].T
    H[-1] = 1.0
    return H @ H.T, H


def row_classes(prm: MPCParams):
    n, m, p, L, ell = prm.n, prm.m, prm.p, prm.L, prm.ell
    r = (m + p) * ell + 1
    w = np.full(r, np.inf)  # soft weight, inf = hard
    B = []  # box rows (predicted inputs k = 0..L-n-1)
    T = np.zeros((r, m + p))
    wq = prm.Q * prm.lamb_sigma / (prm.Q + prm.lamb_sigma)
    for j in range(ell):
        for i in range(m):
            row = j * m + i
            if n <= j < L:
                B.append(row)
            elif j >= L:
                T[row, i] = 1.0
        for i in range(p):
            row = m * ell + j * p + i
            if j < n:
                w[row] = prm.lamb_sigma
            elif j < L:
                w[row] = wq[i]
                T[row, m + i] = 1.0
            else:
                w[row] = prm.lamb_sigma
                T[row, m + i] = 1.0
    return w, np.array(B), T


def box_qp_active_set(Hq, f, lo, hi, x0=None, max_iter=200):
    """min 0.5 x'Hq x + f'x, lo <= x <= hi (Hq SPD): block principal pivoting in the space of the
    active bounds, exactly as `box_qp_bpp` of ddqc_ddmpc.cu (W = Hq^-1 explicit; per pass
    mu = W_AA^-1 (x0_A - b_A), x = x0 - W[:,A] mu; all infeasible indices exchanged at once, single
    largest-index exchanges after three passes without progress)."""
    nx = len(f)
    W = np.linalg.inv(Hq)
    W = 0.5 * (W + W.T)
    xu = -W @ f
    state = np.zeros(nx, int)
    tol_x, tol_g = 1e-10, 1e-11 * max(1.0, np.abs(f).max())
    best, patience = nx + 1, 3
    for it in range(max_iter):
        A = np.nonzero(state)[0]
        bnd = np.where(state < 0, lo, hi)
        x = xu.copy()
        mu = np.zeros(0)
        if len(A):
            mu = np.linalg.solve(W[np.ix_(A, A)], xu[A] - bnd[A])
            x = xu - W[:, A] @ mu
            x[A] = bnd[A]
        verdict = np.zeros(nx, int)
        g_act = -mu
        for a, i in enumerate(A):
            if (state[i] < 0 and g_act[a] < -tol_g) or (state[i] > 0 and g_act[a] > tol_g):
                verdict[i] = 2
        free = state == 0
        verdict[free & (x < lo - tol_x * (1 + np.abs(lo)))] = -1
        verdict[free & (x > hi + tol_x * (1 + np.abs(hi)))] = 1
        bad = np.nonzero(verdict)[0]
        if len(bad) == 0:
            return x, it + 1
        if len(bad) < best:
            best, patience, full = len(bad), 3, True
        elif patience > 0:
            patience, full = patience - 1, True
        else:
            full = False
        for i in bad if full else bad[-1:]:
            state[i] = 0 if verdict[i] == 2 else verdict[i]
    raise RuntimeError("block principal pivoting did not converge")


def v_sr_optimal_reachable(prm: MPCParams, G, y_r):
    """v_sr = H alpha_sr for alpha_sr = least-norm alpha of the optimal reachable equilibrium for y_r
    (ddmpc_oracle.solve_alpha_sr), in row space: with M_s = G + D_s, rho = A_u u_s + A_y y_s + e_one,
      (u_s, y_s) = argmin lamb_alpha_s rho' M_s^-1 rho + |y_s - y_r|_S^2,  u_s in U_s   (a 6-variable box QP)
      v_sr = rho - D_s M_s^-1 rho.
    Rows in the canonical order of `hankel_gram`: inputs, outputs, ones."""
    m, p, ell = prm.m, prm.p, prm.ell
    r = G.shape[0]
    Ds = np.zeros(r)
    Ds[m * ell : (m + p) * ell] = prm.lamb_alpha_s / prm.lamb_sigma_s
    Ks = np.zeros((r, m + p))
    for j in range(ell):
        Ks[j * m : (j + 1) * m, :m] = np.eye(m)
        Ks[m * ell + j * p : m * ell + (j + 1) * p, m:] = np.eye(p)
    e1 = np.zeros(r)
    e1[-1] = 1.0
    Lc = np.linalg.cholesky(G + np.diag(Ds))
    X = np.linalg.solve(Lc, np.concatenate([Ks, e1[:, None]], axis=1))
    Z = X.T @ X
    Hs = 2 * prm.lamb_alpha_s * Z[: m + p, : m + p]
    fs = 2 * prm.lamb_alpha_s * Z[: m + p, m + p]
    for i in range(p):
        Hs[m + i, m + i] += 2 * prm.S[i]
        fs[m + i] -= 2 * prm.S[i] * y_r[i]
    lo = np.concatenate([prm.Us[:, 0], np.full(p, -np.inf)])
    hi = np.concatenate([prm.Us[:, 1], np.full(p, np.inf)])
    t, _ = box_qp_active_set(Hs, fs, lo, hi)
    rho = Ks @ t + e1
    z = np.linalg.solve(Lc.T, np.linalg.solve(Lc, rho))
    return rho - Ds * z, t[:m], t[m:]


def solve_condensed(prm: MPCParams, u_win, y_win, y_r, us_prev=None, ys_prev=None, x0=None, v_sr_ext=None,
                    u_ic=None, y_ic=None):
    """`v_sr_ext` (alpha_reg_type PREVIOUS): H alpha_sr for an arbitrary alpha_sr, canonical row order.
    `u_ic` / `y_ic` (update_cost_threshold): windows whose last n samples are the initial condition when the
    Hankel data (u_win, y_win) is a frozen snapshot."""
    n, m, p, L, ell = prm.n, prm.m, prm.p, prm.L, prm.ell
    G, _ = hankel_gram(u_win, y_win, ell)
    r = G.shape[0]
    w, B, T = row_classes(prm)
    u_past = (u_win if u_ic is None else u_ic)[-n:]
    y_past = (y_win if y_ic is None else y_ic)[-n:]
    tau = np.zeros(r)
    tau[: n * m] = u_past.ravel()
    tau[m * ell : m * ell + n * p] = y_past.ravel()
    tau[-1] = 1.0

    # alpha_sr only enters through v_sr = H alpha_sr (us_prev / ys_prev are unused: kept for call compatibility)
    v_sr = np.zeros(r)
    if prm.alpha_reg_type == APPROXIMATED:
        v_sr, _, _ = v_sr_optimal_reachable(prm, G, np.asarray(y_r, float).ravel())
    elif v_sr_ext is not None:
        v_sr = np.asarray(v_sr_ext, float)

    D0 = np.where(np.isfinite(w), prm.lamb_alpha / w, 0.0)
    Lc = np.linalg.cholesky(G + np.diag(D0))
    nB = len(B)
    K = np.zeros((r, nB + m + p))
    K[B, np.arange(nB)] = 1.0
    K[:, nB:] = T
    c = tau - v_sr
    X = np.linalg.solve(Lc, np.concatenate([K, c[:, None]], axis=1))  # forward substitution only
    Sm = X.T @ X
    nx = nB + m + p
    Hq = 2 * prm.lamb_alpha * Sm[:nx, :nx]
    f = 2 * prm.lamb_alpha * Sm[:nx, nx]
    for b_idx, row in enumerate(B):
        i = row % m
        Hq[b_idx, b_idx] += 2 * prm.R[i]
        Hq[b_idx, nB + i] -= 2 * prm.R[i]
        Hq[nB + i, b_idx] -= 2 * prm.R[i]
        Hq[nB + i, nB + i] += 2 * prm.R[i]
    for i in range(m):
        Hq[nB + i, nB + i] += 2 * n * prm.R[i]
        f[nB + i] += -2 * prm.R[i] * u_past[:, i].sum()
    for i in range(p):
        Hq[nB + m + i, nB + m + i] += 2 * (n * prm.Q[i] + prm.S[i])
        f[nB + m + i] += -2 * (prm.Q[i] * y_past[:, i].sum() + prm.S[i] * y_r[i])
    lo = np.concatenate([np.tile(prm.U[:, 0], L - n), np.maximum(prm.U[:, 0], prm.Us[:, 0]), np.full(p, -np.inf)])
    hi = np.concatenate([np.tile(prm.U[:, 1], L - n), np.minimum(prm.U[:, 1], prm.Us[:, 1]), np.full(p, np.inf)])
    x, iters = box_qp_active_set(Hq, f, lo, hi, x0)
    u_s, y_s = x[nB : nB + m], x[nB + m :]
    u_pred = np.vstack([x[:nB].reshape(L - n, m), np.tile(u_s, (n + 1, 1))])  # ubar_0..ubar_L
    return dict(u_opt=u_pred, us=u_s, ys=y_s, x=x, iters=iters, Hq=Hq, f=f, lo=lo, hi=hi, v_sr=v_sr)


def recover_dual(prm: MPCParams, u_win, y_win, y_r, u_opt, us, ys, v_sr=None, u_ic=None, y_ic=None):
    """Mirror of `ddmpc_recover_kernel` (csrc/ddqc_ddmpc_recover.cu): with the box-QP optimum x* = (ubar, u_s, y_s)
    known, every row of H has a prescribed value (hard rows, D = 0) or a target with weight w (soft rows,
    D = lamb_alpha / w); the minimising alpha is alpha_sr + H'z with

        (G + D) z = t - v_sr,     t = prescribed values / targets,

    from which  alpha* - alpha_sr = H'z  (alpha_reg_type PREVIOUS needs alpha*),  ybar_k - y_s = -(lamb_alpha / Q) z
    on the predicted output rows, and the TRACKING COST sum |ubar - u_s|_R^2 + |ybar - y_s|_Q^2 + |y_s - y_r|_S^2
    (update_cost_threshold).  Canonical row order of `hankel_gram`.  Returns (z, H'z, tracking cost)."""
    n, m, p, L, ell = prm.n, prm.m, prm.p, prm.L, prm.ell
    G, H = hankel_gram(u_win, y_win, ell)
    r = G.shape[0]
    w, _, _ = row_classes(prm)
    u_past = (u_win if u_ic is None else u_ic)[-n:]
    y_past = (y_win if y_ic is None else y_ic)[-n:]
    ubar = np.vstack([u_past, np.asarray(u_opt, float)])  # (ell, m): k = -n..L
    t = np.empty(r)
    t[: m * ell] = ubar.ravel()
    t[m * ell : (m + p) * ell] = np.concatenate([y_past.ravel(), np.tile(ys, ell - n)])
    t[-1] = 1.0
    D = np.where(np.isfinite(w), prm.lamb_alpha / w, 0.0)
    rhs = t - (0.0 if v_sr is None else np.asarray(v_sr, float))
    Lc = np.linalg.cholesky(G + np.diag(D))
    z = np.linalg.solve(Lc.T, np.linalg.solve(Lc, rhs))
    cost = float((prm.R * (ubar - us) ** 2).sum() + (prm.Q * (y_past - ys) ** 2).sum() + (prm.S * (ys - np.asarray(y_r, float).ravel()) ** 2).sum())
    zy = z[m * ell + n * p : m * ell + L * p].reshape(L - n, p)  # predicted output rows k = 0..L-n-1
    cost += float((prm.lamb_alpha**2 * zy**2 / prm.Q).sum())
    return z, H.T @ z, cost


# =============================================================================================
# v2: the elimination order the CUDA solver uses since round 1b (ddqc_ddmpc.cu, phases A-D)
# =============================================================================================
# Rows are ordered [R1 | R2 | R3 | aug]: R1 = past inputs, ones row, terminal inputs (hard, shared by
# both regularised systems), R2 = output rows (soft), R3 = predicted-input rows (the box-constrained
# decision variables), aug = right-hand sides carried as extra rows (T_u, T_y, tau, b_s).  Eliminating
# R1 then R2 leaves the Schur complement on [R3 | aug]; a Gauss-Jordan sweep on R3 yields
#   Z = K^T (G + D0)^-1 K   for K = [E_B | T | c],
# i.e. exactly the Hessian / gradient blocks `solve_condensed` obtains from 94 forward substitutions.
# The box QP is solved by block principal pivoting in the space of the FREE variables (warm-startable)
# with a Mehrotra interior-point fallback followed by a pivoting polish.
def perm_rows(prm):
    n,m,p,L,ell=prm.n,prm.m,prm.p,prm.L,prm.ell
    # canonical rows (as in hankel_gram): u rows j*m+i | y rows m*ell+j*p+i | ones (last)
    r=(m+p)*ell+1
    R1=[j*m+i for j in range(n) for i in range(m)]+[r-1]+[j*m+i for j in range(L,ell) for i in range(m)]
    R2=[m*ell+j*p+i for j in range(ell) for i in range(p)]
    R3=[j*m+i for j in range(n,L) for i in range(m)]
    return np.array(R1),np.array(R2),np.array(R3)

def chol_partial(A, k):
    """eliminate first k columns of symmetric A (right-looking); return Schur complement of trailing"""
    A=A.copy()
    L11=np.linalg.cholesky(A[:k,:k])
    L21=np.linalg.solve(L11, A[:k,k:]).T
    return A[k:,k:]-L21@L21.T, L11, L21

def sweep(S,k):
    S=S.copy()
    for j in range(k):
        d=S[j,j]; col=S[:,j].copy()
        S-=np.outer(col,col)/d
        S[:,j]=col/d; S[j,:]=col/d; S[j,j]=-1.0/d
    return S

def box_qp_free(Hq,f,lo,hi,state0=None,max_pass=15):
    nx=len(f); state=np.zeros(nx,int) if state0 is None else state0.copy()
    # sanitize: infinite bounds cannot be active
    state[(state<0)&~np.isfinite(lo)]=0; state[(state>0)&~np.isfinite(hi)]=0
    tol_x=1e-10; tol_g=1e-11*max(1.0,np.abs(f).max())
    best,patience=nx+1,3
    for it in range(max_pass):
        F=np.nonzero(state==0)[0]; A=np.nonzero(state)[0]
        x=np.where(state<0,lo,np.where(state>0,hi,0.0))
        x=np.where(state==0,0.0,x)
        rhs=-(f[F]+Hq[np.ix_(F,A)]@x[A])
        if len(F): x[F]=np.linalg.solve(Hq[np.ix_(F,F)],rhs)
        g=Hq@x+f
        verdict=np.zeros(nx,int)
        verdict[(state<0)&(g<-tol_g)]=2; verdict[(state>0)&(g>tol_g)]=2
        fr=state==0
        verdict[fr&(x<lo-tol_x*(1+np.abs(lo)))]=-1
        verdict[fr&(x>hi+tol_x*(1+np.abs(hi)))]=1
        bad=np.nonzero(verdict)[0]
        if len(bad)==0: return x,state,it+1,True
        if len(bad)<best: best,patience,full=len(bad),3,True
        elif patience>0: patience-=1; full=True
        else: full=False
        for i in (bad if full else bad[-1:]): state[i]=0 if verdict[i]==2 else verdict[i]
    return x,state,max_pass,False

def box_qp_ipm(Hq,f,lo,hi,max_iter=60,tol=1e-10):
    nx=len(f)
    il=np.isfinite(lo); iu=np.isfinite(hi)
    # start: midpoint (or 0 for free vars)
    x=np.where(il&iu,0.5*(np.where(il,lo,0)+np.where(iu,hi,0)),0.0)
    sl=np.where(il,x-lo,1.0); su=np.where(iu,hi-x,1.0)
    zl=np.where(il,1.0,0.0); zu=np.where(iu,1.0,0.0)
    # scale duals to problem
    g=Hq@x+f; sc=max(1.0,np.abs(g).max()); zl*=sc; zu*=sc
    nc=il.sum()+iu.sum()
    for it in range(max_iter):
        g=Hq@x+f
        rd=g-zl+zu                    # dual residual
        mu=(sl[il]@zl[il]+su[iu]@zu[iu])/nc
        if np.abs(rd).max()<tol*sc and mu<tol*sc*1e-2: break
        dl=np.where(il,zl/sl,0.0); du=np.where(iu,zu/su,0.0)
        Lc=np.linalg.cholesky(Hq+np.diag(dl+du))
        def solve(rcl,rcu):
            # complementarity residual targets: sl*zl + sl*dzl + zl*dsl = rcl ; dsl=dx, dsu=-dx
            rhs=-rd+np.where(il,(rcl-sl*zl)/sl,0.0)-np.where(iu,(rcu-su*zu)/su,0.0)
            dx=np.linalg.solve(Lc.T,np.linalg.solve(Lc,rhs))
            dzl=np.where(il,(rcl-sl*zl-zl*dx)/sl,0.0); dzu=np.where(iu,(rcu-su*zu+zu*dx)/su,0.0)
            return dx,dzl,dzu
        def maxstep(dx,dzl,dzu):
            a=1.0
            for v,dv,msk in ((sl,dx,il),(su,-dx,iu),(zl,dzl,il),(zu,dzu,iu)):
                neg=msk&(dv<0)
                if neg.any(): a=min(a,(-v[neg]/dv[neg]).min())
            return a
        dxa,dzla,dzua=solve(np.zeros(nx),np.zeros(nx))
        aa=maxstep(dxa,dzla,dzua)
        mu_aff=(((sl+aa*dxa)*(zl+aa*dzla))[il].sum()+((su-aa*dxa)*(zu+aa*dzua))[iu].sum())/nc
        sig=(mu_aff/mu)**3
        rcl=sig*mu-dxa*dzla; rcu=sig*mu+dxa*dzua
        dx,dzl,dzu=solve(rcl,rcu)
        a=min(1.0,0.995*maxstep(dx,dzl,dzu)) if maxstep(dx,dzl,dzu)<1 else 1.0
        x=x+a*dx; sl=np.where(il,sl+a*dx,1.0); su=np.where(iu,su-a*dx,1.0); zl=zl+a*dzl; zu=zu+a*dzu
    state=np.zeros(nx,int); state[il&(zl>sl)]=-1; state[iu&(zu>su)]=1
    return x,state,it

def solve_condensed_v2(prm,u_win,y_win,y_r,us_prev=None,ys_prev=None,state0=None):
    n,m,p,L,ell=prm.n,prm.m,prm.p,prm.L,prm.ell
    G,_=hankel_gram(u_win,y_win,ell)
    r=G.shape[0]
    R1,R2,R3=perm_rows(prm)
    n1,n2,n3=len(R1),len(R2),len(R3)
    w,B,T=row_classes(prm)
    assert (B==R3).all()
    tau=np.zeros(r); tau[:n*m]=u_win[-n:].ravel(); tau[m*ell:m*ell+n*p]=y_win[-n:].ravel(); tau[-1]=1.0
    use_sr=prm.alpha_reg_type==APPROXIMATED
    y_r=np.asarray(y_r,float).ravel()
    # augmented right-hand sides carried as extra rows, in the kernel's order:
    #   group 0: A_u (all input rows, per component), A_y (all output rows), e_one, tau   (alpha_sr system + c-row)
    #   group 1: T_u (terminal input rows), T_y (free + terminal output rows)             (main system)
    Au=np.zeros((r,m)); Ay=np.zeros((r,p))
    for j in range(ell):
        Au[j*m:(j+1)*m,:]=np.eye(m); Ay[m*ell+j*p:m*ell+(j+1)*p,:]=np.eye(p)
    e1=np.zeros(r); e1[-1]=1.0
    order=np.concatenate([R1,R2,R3])
    K=np.concatenate([Au,Ay,e1[:,None],tau[:,None],T],axis=1)  # r x (2(m+p)+2)
    na0,na=m+p+2,K.shape[1]
    A=np.zeros((r+na,r+na)); A[:r,:r]=G[np.ix_(order,order)]; A[r:,:r]=K[order].T; A[:r,r:]=K[order]
    # phase A: eliminate R1 (shared by both regularised systems)
    S1,_,_=chol_partial(A,n1)      # order [R2 | R3 | aug group 0 | aug group 1]
    nm=n2+n3
    e_row=np.zeros(n2); coef=np.zeros(na0); coef[m+p+1]=0.0
    if use_sr:
        ds=prm.lamb_alpha_s/prm.lamb_sigma_s
        # phase B: S1 + D_s with aug group 0: factor everything; the aug x aug Schur block is -K_s' M_s^-1 K_s
        Sb=S1[:nm+na0,:nm+na0].copy()
        Sb[np.arange(n2),np.arange(n2)]+=ds
        Sa,Lb,Lab=chol_partial(Sb,nm)           # Lab: forward-substituted aug rows (na0 x nm)
        Hs=-2*prm.lamb_alpha_s*Sa[:m+p,:m+p]; fs=-2*prm.lamb_alpha_s*Sa[:m+p,m+p]
        for i in range(p):
            Hs[m+i,m+i]+=2*prm.S[i]; fs[m+i]-=2*prm.S[i]*y_r[i]
        lo_s=np.concatenate([prm.Us[:,0],np.full(p,-np.inf)]); hi_s=np.concatenate([prm.Us[:,1],np.full(p,np.inf)])
        t_sr,_=box_qp_active_set(Hs,fs,lo_s,hi_s)
        coef=np.concatenate([t_sr,[1.0,0.0]])   # rho = A_u u_sr + A_y y_sr + e_one
        yv=Lab.T@coef                           # L^-1 rho restricted to [R2 | R3]
        z=np.linalg.solve(Lb.T,yv)              # back-substitution
        e_row=ds*z[:n2]                         # v_sr = rho - [e on R2]
    # phase C: S1 + D_0 with aug rows (T_u, T_y, c),  c = tau - rho + e
    W=np.zeros((nm+na,nm+m+p+1)); W[:nm,:nm]=np.eye(nm)
    for k in range(m+p): W[nm+na0+k,nm+k]=1.0  # T rows
    W[nm+m+p+1,nm+m+p]=1.0                     # tau
    W[nm:nm+na0,nm+m+p]-=coef                  # - rho (coef has 0 on the tau slot)
    Sc=W.T@S1@W
    Sc[nm+m+p,:n2]+=e_row; Sc[:n2,nm+m+p]+=e_row
    D0=prm.lamb_alpha/w[R2]
    Sc[np.arange(n2),np.arange(n2)]+=D0
    S7,_,_=chol_partial(Sc,n2)       # (n3 + 7): Schur complement on [R3 | T_u, T_y, c]
    nB=n3
    Z=-sweep(S7,nB); Z[:nB,nB:]*=-1; Z[nB:,:nB]*=-1
    nx=nB+m+p
    Hq=2*prm.lamb_alpha*Z[:nx,:nx]; f=2*prm.lamb_alpha*Z[:nx,nx]
    u_past,y_past=u_win[-n:],y_win[-n:]
    for b_idx,row in enumerate(B):
        i=row%m
        Hq[b_idx,b_idx]+=2*prm.R[i]; Hq[b_idx,nB+i]-=2*prm.R[i]; Hq[nB+i,b_idx]-=2*prm.R[i]; Hq[nB+i,nB+i]+=2*prm.R[i]
    for i in range(m):
        Hq[nB+i,nB+i]+=2*n*prm.R[i]; f[nB+i]+=-2*prm.R[i]*u_past[:,i].sum()
    for i in range(p):
        Hq[nB+m+i,nB+m+i]+=2*(n*prm.Q[i]+prm.S[i]); f[nB+m+i]+=-2*(prm.Q[i]*y_past[:,i].sum()+prm.S[i]*y_r[i])
    lo=np.concatenate([np.tile(prm.U[:,0],L-n),np.maximum(prm.U[:,0],prm.Us[:,0]),np.full(p,-np.inf)])
    hi=np.concatenate([np.tile(prm.U[:,1],L-n),np.minimum(prm.U[:,1],prm.Us[:,1]),np.full(p,np.inf)])
    x,state,it,ok=box_qp_free(Hq,f,lo,hi,state0)
    info=dict(passes=it,ipm=0)
    if not ok:
        xi,st,iti=box_qp_ipm(Hq,f,lo,hi)
        x2,state2,it2,ok2=box_qp_free(Hq,f,lo,hi,st,max_pass=5)
        info.update(ipm=iti,polish=it2,polish_ok=ok2,ipm_err=np.abs(x2-xi).max())
        x,state=(x2,state2) if ok2 else (xi,st)
    u_s, y_s = x[nB:nB+m], x[nB+m:]
    u_pred = np.vstack([x[:nB].reshape(L-n, m), np.tile(u_s, (n+1, 1))])  # ubar_0..ubar_L
    return dict(x=x,state=state,Hq=Hq,f=f,lo=lo,hi=hi,info=info,u_opt=u_pred,us=u_s,ys=y_s)

