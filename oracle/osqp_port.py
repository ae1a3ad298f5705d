"""TEST / BASELINE INFRASTRUCTURE ONLY (oracle) -- an OSQP-style ADMM QP solver in numpy / scipy.sparse.

The reference solves its data-driven MPC problems with `cvxpy` (unpinned; default QP solver OSQP, which cvxpy calls with
eps_abs = eps_rel = 1e-5 and max_iter = 10000) inside the third-party `direct-data-driven-mpc@94150cc` - none of which is
installable offline (profiles/r2_reference_stack_probe.log).  This module restates the PUBLISHED algorithm of that solver
(B. Stellato, G. Banjac, P. Goulart, A. Bemporad, S. Boyd, "OSQP: an operator splitting solver for quadratic programs",
Math. Prog. Comp. 12, 2020: Algorithm 1, Section 5.1 termination criteria, Section 5.2 adaptive rho, Section 5.1 Ruiz
equilibration) so that

  * bench.py has a CPU baseline of the SAME CLASS as the reference's path (first-order solver at 1e-5, sparse
    quasi-definite KKT factorisation reused across iterations) instead of the 40x slower dense interior-point oracle, and
  * the tests can report how far a 1e-5 first-order solution - what the reference itself computes - lies from the
    high-accuracy oracle, which is the yardstick the north star's "within 1e-4 of the reference's cvxpy solution" needs.

It is NOT the reference and NOT OSQP's code: no polishing, scipy's SuperLU instead of QDLDL, rho adapted every
`adapt_every` iterations instead of OSQP's timing heuristic.  PARITY UNPINNED like the rest of the DD-MPC oracle.

    minimise 0.5 x'Px + q'x   subject to   l <= A x <= u
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

RHO_EQ_FACTOR = 1e3  # OSQP: rho of equality rows = 1e3 * rho
RHO_MIN, RHO_MAX = 1e-6, 1e6


def ruiz_equilibrate(P, A, q, iters: int = 10):
    """Ruiz equilibration of the KKT matrix [[P, A'], [A, 0]] (OSQP Algorithm 2): returns the scaled (P, A, q), the
    diagonal scalings D (variables), E (constraints) and the cost scaling c."""
    n, m = P.shape[0], A.shape[0]
    D, E, c = np.ones(n), np.ones(m), 1.0
    Ps, As, qs = P.copy(), A.copy(), q.copy()
    for _ in range(iters):
        col_P = np.abs(Ps).max(axis=0).toarray().ravel() if Ps.nnz else np.zeros(n)
        col_A = np.abs(As).max(axis=0).toarray().ravel() if As.nnz else np.zeros(n)
        dn = np.maximum(col_P, col_A)
        en = np.abs(As).max(axis=1).toarray().ravel() if As.nnz else np.zeros(m)
        dn[dn < 1e-4], en[en < 1e-4] = 1.0, 1.0
        d, e = 1.0 / np.sqrt(dn), 1.0 / np.sqrt(en)
        Dm, Em = sp.diags(d), sp.diags(e)
        Ps, As, qs = (Dm @ Ps @ Dm).tocsc(), (Em @ As @ Dm).tocsc(), d * qs
        D, E = D * d, E * e
        col_norm = np.abs(Ps).max(axis=0).toarray().ravel().mean() if Ps.nnz else 0.0
        g = 1.0 / max(col_norm, np.abs(qs).max(), 1e-4)
        Ps, qs, c = Ps * g, qs * g, c * g
    return Ps.tocsc(), As.tocsc(), qs, D, E, c


def osqp_solve(P, q, A, l, u, eps_abs=1e-5, eps_rel=1e-5, max_iter=10000, rho=0.1, sigma=1e-6, alpha=1.6,
               scaling=10, check_every=25, adapt_every=100, x0=None, y0=None):  # fmt: skip
    P, A = sp.csc_matrix(P), sp.csc_matrix(A)
    q, l, u = np.asarray(q, float), np.asarray(l, float), np.asarray(u, float)
    n, m = P.shape[0], A.shape[0]
    Ps, As, qs, D, E, c = ruiz_equilibrate(P, A, q, scaling) if scaling else (P, A, q, np.ones(n), np.ones(m), 1.0)
    ls, us = E * l, E * u
    is_eq = np.abs(us - ls) < 1e-4 * np.maximum(1.0, np.abs(us))  # OSQP's equality detection
    Pu = sp.triu(Ps).tocsc()
    AsT = As.T.tocsc()
    n_fact = 0

    def factor(rho_):
        nonlocal n_fact
        rv = np.where(is_eq, RHO_EQ_FACTOR * rho_, rho_)
        K = sp.bmat([[Ps + sigma * sp.identity(n), AsT], [As, -sp.diags(1.0 / rv)]], format="csc")
        n_fact += 1
        return spla.splu(K), rv

    lu, rho_v = factor(rho)
    x = np.zeros(n) if x0 is None else np.asarray(x0, float) / D
    y = np.zeros(m) if y0 is None else np.asarray(y0, float) / E * c
    z = np.clip(As @ x, ls, us)
    status, it = "max_iter", 0
    r_prim = r_dual = np.inf
    for it in range(1, max_iter + 1):
        sol = lu.solve(np.concatenate([sigma * x - qs, z - y / rho_v]))
        xt, nu = sol[:n], sol[n:]
        zt = z + (nu - y) / rho_v
        x = alpha * xt + (1 - alpha) * x
        zr = alpha * zt + (1 - alpha) * z
        z_new = np.clip(zr + y / rho_v, ls, us)
        y = y + rho_v * (zr - z_new)
        z = z_new
        if it % check_every == 0 or it == max_iter:
            # residuals and tolerances in the ORIGINAL scaling (OSQP section 5.1)
            Ax, Px, Aty = As @ x, Ps @ x, AsT @ y
            r_prim = np.abs((Ax - z) / E).max()
            r_dual = np.abs((Px + qs + Aty) / D).max() / c
            e_prim = eps_abs + eps_rel * max(np.abs(Ax / E).max(), np.abs(z / E).max())
            e_dual = eps_abs + eps_rel * max(np.abs(Px / D).max(), np.abs(Aty / D).max(), np.abs(qs / D).max()) / c
            if r_prim <= e_prim and r_dual <= e_dual:
                status = "solved"
                break
            if adapt_every and it % adapt_every == 0:  # section 5.2: balance the scaled residuals
                num = r_prim / max(np.abs(Ax / E).max(), np.abs(z / E).max(), 1e-12)
                den = r_dual / max(np.abs(Px / D).max() / c, np.abs(Aty / D).max() / c, np.abs(qs / D).max() / c, 1e-12)
                new_rho = float(np.clip(rho * np.sqrt(num / max(den, 1e-300)), RHO_MIN, RHO_MAX))
                if new_rho > 5 * rho or new_rho < rho / 5:
                    rho = new_rho
                    lu, rho_v = factor(rho)
    return D * x, E * y / c, dict(status=status, iters=it, r_prim=float(r_prim), r_dual=float(r_dual), rho=rho, factorisations=n_fact)


# ------------------------------------------------------------------------------------------------------
# the two QPs of one nonlinear DD-MPC step in the sparse form a modelling layer hands to OSQP
# ------------------------------------------------------------------------------------------------------
def mpc_qp_osqp_form(prm, u_win, y_win, y_r, alpha_sr):
    """The literal problem of oracle/ddmpc_oracle.build_literal_qp as (P, q, A, l, u): equalities as l = u rows, the
    box constraints as two-sided rows."""
    import ddmpc_oracle as mo

    P, q, Aeq, beq, G, h, offs = mo.build_literal_qp(prm, u_win, y_win, y_r, alpha_sr)
    # G holds (+row <= hi, -row <= -lo) pairs in this order: fold each pair into one two-sided row
    rows, lo, hi = [], [], []
    for k in range(0, G.shape[0], 2):
        rows.append(G[k])
        hi.append(h[k])
        lo.append(-h[k + 1])
    A = sp.vstack([sp.csc_matrix(Aeq), sp.csc_matrix(np.array(rows))]).tocsc()
    return sp.csc_matrix(P), q, A, np.concatenate([beq, lo]), np.concatenate([beq, hi]), offs


def alpha_sr_qp_osqp_form(prm, u_win, y_win, y_r):
    """The optimal-reachable-equilibrium problem of oracle/ddmpc_oracle.solve_alpha_sr with sigma_s kept as a variable
    (sparse, as a modelling layer would state it): z = [a (C), s (p ell), u_s (m), y_s (p)]."""
    import ddmpc_oracle as mo

    ell, C, m, p = prm.ell, prm.C, prm.m, prm.p
    Hu, Hy = mo.hankel(u_win, ell), mo.hankel(y_win, ell)
    o_s, o_us, o_ys = C, C + p * ell, C + p * ell + m
    nz = o_ys + p
    Pd = np.zeros(nz)
    Pd[:C] = 2 * prm.lamb_alpha_s
    Pd[o_s:o_us] = 2 * prm.lamb_sigma_s
    Pd[o_ys:] = 2 * prm.S
    q = np.zeros(nz)
    q[o_ys:] = -2 * prm.S * np.asarray(y_r, float).ravel()
    Eu, Ey = np.kron(np.ones((ell, 1)), np.eye(m)), np.kron(np.ones((ell, 1)), np.eye(p))
    A1 = sp.hstack([sp.csc_matrix(Hu), sp.csc_matrix((m * ell, p * ell)), sp.csc_matrix(-Eu), sp.csc_matrix((m * ell, p))])
    A2 = sp.hstack([sp.csc_matrix(Hy), -sp.identity(p * ell), sp.csc_matrix((p * ell, m)), sp.csc_matrix(-Ey)])
    A3 = sp.hstack([sp.csc_matrix(np.ones((1, C))), sp.csc_matrix((1, nz - C))])
    A4 = sp.hstack([sp.csc_matrix((m, o_us)), sp.identity(m), sp.csc_matrix((m, p))])
    A = sp.vstack([A1, A2, A3, A4]).tocsc()
    ne = m * ell + p * ell
    l = np.concatenate([np.zeros(ne), [1.0], prm.Us[:, 0]])
    u = np.concatenate([np.zeros(ne), [1.0], prm.Us[:, 1]])
    return sp.diags(Pd).tocsc(), q, A, l, u


class OSQPStyleDDMPC:
    """The closed-loop step API of the DD-MPC controller with both QPs of a step (alpha_sr, then the MPC problem) solved by
    the OSQP-style port at cvxpy's OSQP settings, warm-started from the previous step like a parametrised cvxpy problem."""

    def __init__(self, prm, u, y, y_r, **osqp_kw):
        self.prm, self.u, self.y = prm, np.array(u, float), np.array(y, float)
        self.y_r = np.asarray(y_r, float).ravel()
        self.kw = osqp_kw
        self.optimal_u = None
        self.info = {}
        self._warm = [None, None]

    def update_and_solve_data_driven_mpc(self):
        import ddmpc_oracle as mo

        prm = self.prm
        a_sr = np.zeros(prm.C)
        if prm.alpha_reg_type == mo.APPROXIMATED:
            P, q, A, l, u = alpha_sr_qp_osqp_form(prm, self.u, self.y, self.y_r)
            w = self._warm[0]
            z, yv, info_s = osqp_solve(P, q, A, l, u, x0=w[0] if w else None, y0=w[1] if w else None, **self.kw)
            self._warm[0] = (z, yv)
            a_sr = z[: prm.C]
            self.info["alpha_sr"] = info_s
        P, q, A, l, u, offs = mpc_qp_osqp_form(prm, self.u, self.y, self.y_r, a_sr)
        w = self._warm[1]
        z, yv, info = osqp_solve(P, q, A, l, u, x0=w[0] if w else None, y0=w[1] if w else None, **self.kw)
        self._warm[1] = (z, yv)
        self.info["mpc"] = info
        ell = prm.ell
        self.optimal_u = z[offs["ubar"] : offs["ubar"] + prm.m * ell].reshape(ell, prm.m)[prm.n :].copy()
        self.us = z[offs["us"] : offs["us"] + prm.m].copy()
        self.ys = z[offs["ys"] : offs["ys"] + prm.p].copy()

    def get_optimal_control_input_at_step(self, n_step=0):
        return self.optimal_u[n_step].copy()

    def store_input_output_measurement(self, u_current, y_current, du_current=None):
        self.u = np.vstack([self.u[1:], np.asarray(u_current, float).reshape(1, -1)])
        self.y = np.vstack([self.y[1:], np.asarray(y_current, float).reshape(1, -1)])
