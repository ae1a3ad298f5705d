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


def solve_condensed(prm: MPCParams, u_win, y_win, y_r, us_prev=None, ys_prev=None, x0=None):
    n, m, p, L, ell = prm.n, prm.m, prm.p, prm.L, prm.ell
    G, _ = hankel_gram(u_win, y_win, ell)
    r = G.shape[0]
    w, B, T = row_classes(prm)
    u_past, y_past = u_win[-n:], y_win[-n:]
    tau = np.zeros(r)
    tau[: n * m] = u_past.ravel()
    tau[m * ell : m * ell + n * p] = y_past.ravel()
    tau[-1] = 1.0

    # alpha_sr only enters through v_sr = H alpha_sr
    v_sr = np.zeros(r)
    if prm.alpha_reg_type == APPROXIMATED and us_prev is not None:
        Ds = np.zeros(r)
        Ds[m * ell : (m + p) * ell] = prm.lamb_alpha_s / prm.lamb_sigma_s
        b_s = np.concatenate([np.tile(us_prev, ell), np.tile(ys_prev, ell), [1.0]])
        c_s = np.linalg.solve(G + np.diag(Ds), b_s)
        v_sr = b_s - Ds * c_s

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
