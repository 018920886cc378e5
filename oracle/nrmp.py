"""CPU restatement of the NRMP half of the PAN hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference builds a DPP cvxpy problem once (neupan/blocks/nrmp.py:263-383 with the cost /
constraint fragments of neupan/robot/robot.py:73-236) and solves it every PAN iteration through
cvxpylayers -> diffcp -> ECOS (nrmp.py:144).  None of those packages is installed here (and the
reference pins no version of them), so this file restates the *mathematical program* and solves it
in float64 with HiGHS' active-set QP solver (scipy.optimize._highspy) on the slack-lifted form.
``oracle/ipm.py`` holds an independently written second solver; the two must agree (tests).

    min   sum (q_s*S - gamma_a)^2  [omni: rows 0:2]          robot.py:142-170
        + sum (p_u*U[0,:] - gamma_b)^2                        robot.py:151,168
        + 0.5*bk*sum (S - nom_s)^2                            nrmp.py:350, robot.py:172-180
        - eta*sum D                                           nrmp.py:382-383
        + 0.5*ro_obs*sum_{t,m} neg(fa_t[m].S[0:2,t+1] - fb_t[m] - D_t)^2   robot.py:183-198
    s.t.  S[:,t+1] = A_t S[:,t] + B_t U[:,t] + C_t            robot.py:200-221
          S[:,0] = nom_s[:,0]                                 robot.py:234
          |U[:,t+1]-U[:,t]| <= max_acce*dt ; |U| <= max_speed robot.py:232-233, 69
          d_min <= D <= d_max ; D >= 0                        nrmp.py:371-380, 264-266

Parameter values are float32 (the reference hands float32 tensors to cvxpylayers, which solves in
float64); the solution is cast back to float32 (nrmp.py:145-148).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from math import cos, sin, tan

import numpy as np


# --------------------------------------------------------------------------------------
# kinematics linearisation, robot.py:239-316.  Python float math on float32 inputs, result
# stored as float32 (torch.Tensor(...) in the reference).
# --------------------------------------------------------------------------------------
def linearise(kinematics: str, nom_s: np.ndarray, nom_u: np.ndarray, dt: float, L: float | None):
    """A_t, B_t, C_t exactly as robot.py:272-316 evaluates them: ``phi``, ``v``, ``psi`` are 0-d float32
    tensors there, so products *with them* are float32 operations (a Python float operand is first
    rounded to float32), whereas math.sin/cos/tan and float-only sub-expressions are float64; every
    entry is finally stored as float32 by torch.Tensor(...)."""
    f = np.float32
    T = nom_u.shape[1]
    A = np.zeros((T, 3, 3), np.float32)
    B = np.zeros((T, 3, 2), np.float32)
    C = np.zeros((T, 3), np.float32)
    fdt = f(dt)
    for t in range(T):
        if kinematics in ("acker", "diff"):
            phi, v = f(nom_s[2, t]), f(nom_u[0, t])
        else:  # omni: phi is the heading *command* nom_u[1] (robot.py:306)
            phi, v = f(nom_u[1, t]), f(nom_u[0, t])
        sn, cs = sin(float(phi)), cos(float(phi))
        A[t] = np.eye(3, dtype=np.float32)
        if kinematics != "omni":
            A[t, 0, 2] = ((-v) * fdt) * f(sn)
            A[t, 1, 2] = (v * fdt) * f(cs)
        C[t, 0] = ((phi * v) * f(sn)) * fdt
        C[t, 1] = (((-phi) * v) * f(cs)) * fdt
        B[t, 0, 0] = f(cs * dt)
        B[t, 1, 0] = f(sn * dt)
        if kinematics == "diff":  # robot.py:289-302
            B[t, 2, 1] = fdt
        elif kinematics == "acker":  # robot.py:272-286
            psi = f(nom_u[1, t])
            den = f(L * (cos(float(psi))) ** 2)
            B[t, 2, 0] = f(tan(float(psi)) * dt / L)
            B[t, 2, 1] = (v * fdt) / den
            C[t, 2] = (((-psi) * v) * fdt) / den
        elif kinematics == "omni":  # robot.py:304-316
            B[t, 0, 1] = ((-v) * f(sn)) * fdt
            B[t, 1, 1] = (v * f(cs)) * fdt
        else:
            raise ValueError("kinematics currently only supports acker or diff")  # robot.py:256
    return A, B, C


@dataclass
class RobotSpec:
    """The numeric facts of neupan/robot/robot.py:32-71 the program needs."""
    kinematics: str
    G: np.ndarray  # (E,2) float64 as produced by gen_inequal_from_vertex
    h: np.ndarray  # (E,1)
    max_speed: np.ndarray  # (2,)
    max_acce: np.ndarray  # (2,)
    dt: float = 0.1
    L: float | None = None

    def __post_init__(self):
        self.max_speed = np.asarray(self.max_speed, np.float64).reshape(2).copy()
        self.max_acce = np.asarray(self.max_acce, np.float64).reshape(2).copy()
        if self.kinematics == "acker" and self.max_speed[1] >= 1.57:  # robot.py:63-66
            self.max_speed[1] = 1.57

    @property
    def speed_bound(self):
        return self.max_speed

    @property
    def acce_bound(self):  # robot.py:69
        return self.max_acce * self.dt


@dataclass
class Adjust:
    """nrmp.py:35-112 defaults; pan.py:70-82 passes eta=10 etc. when the yaml omits them."""
    q_s: object = 1.0  # scalar or 3-vector
    p_u: float = 1.0
    eta: float = 10.0
    d_max: float = 1.0
    d_min: float = 0.1
    ro_obs: float = 400.0
    bk: float = 0.1

    def q_vec32(self) -> np.ndarray:
        q = np.asarray(self.q_s, np.float32).reshape(-1)
        return np.repeat(q, 3) if q.size == 1 else q


@dataclass
class NrmpProblem:
    """All data of one solve, float64 copies of the float32 parameter values."""
    T: int
    M: int  # 0 => no_obs
    omni: bool
    nom_s: np.ndarray  # (3,T+1)
    gamma_a: np.ndarray  # (3,T+1)
    gamma_b: np.ndarray  # (T,)
    A: np.ndarray
    B: np.ndarray
    C: np.ndarray
    fa: np.ndarray  # (T,M,2)
    fb: np.ndarray  # (T,M)
    q: np.ndarray  # (3,)
    p_u: float
    eta: float
    d_max: float
    d_min: float
    ro_obs: float
    bk: float
    speed_bound: np.ndarray
    acce_bound: np.ndarray
    meta: dict = field(default_factory=dict)


def build_problem(robot: RobotSpec, adj: Adjust, nom_s, nom_u, ref_s, ref_us, fa=None, fb=None, M: int = 10) -> NrmpProblem:
    """nrmp.py:152-166 + robot.py:239-268: parameter values.  Inputs float32 arrays."""
    nom_s = np.asarray(nom_s, np.float32)
    nom_u = np.asarray(nom_u, np.float32)
    ref_s = np.asarray(ref_s, np.float32)
    ref_us = np.asarray(ref_us, np.float32).reshape(-1)
    T = nom_u.shape[1]
    q32 = adj.q_vec32()
    gamma_a = (q32.reshape(3, 1) * ref_s).astype(np.float32)  # nrmp.py:158 (float32 product)
    gamma_b = (np.float32(adj.p_u) * ref_us).astype(np.float32)
    A, B, C = linearise(robot.kinematics, nom_s, nom_u, robot.dt, robot.L)
    if M > 0:
        if fa is None:  # nrmp.py:237-241 (no points): zeros
            fa = np.zeros((T, M, 2), np.float32)
            fb = np.zeros((T, M), np.float32)
        fa = np.asarray(fa, np.float32).reshape(T, M, 2)
        fb = np.asarray(fb, np.float32).reshape(T, M)
    else:
        fa = np.zeros((T, 0, 2), np.float32)
        fb = np.zeros((T, 0), np.float32)
    f64 = lambda a: np.asarray(a, np.float64)
    return NrmpProblem(
        T=T, M=M, omni=(robot.kinematics == "omni"), nom_s=f64(nom_s), gamma_a=f64(gamma_a), gamma_b=f64(gamma_b),
        A=f64(A), B=f64(B), C=f64(C), fa=f64(fa), fb=f64(fb), q=f64(q32),
        p_u=float(np.float32(adj.p_u)), eta=float(np.float32(adj.eta)), d_max=float(np.float32(adj.d_max)),
        d_min=float(np.float32(adj.d_min)), ro_obs=float(adj.ro_obs), bk=float(adj.bk),
        speed_bound=f64(robot.speed_bound), acce_bound=f64(robot.acce_bound))


def objective(p: NrmpProblem, S, U, D) -> float:
    """The cost of the program at a point (float64); used by the tests as a solver-free check."""
    rows = 2 if p.omni else 3
    c = np.sum((p.q[:rows, None] * S[:rows] - p.gamma_a[:rows]) ** 2)
    c += np.sum((p.p_u * U[0] - p.gamma_b) ** 2)
    c += 0.5 * p.bk * np.sum((S - p.nom_s) ** 2)
    if p.M > 0:
        c -= p.eta * np.sum(D)
        for t in range(p.T):
            I = p.fa[t] @ S[0:2, t + 1] - p.fb[t] - D[t]
            c += 0.5 * p.ro_obs * np.sum(np.minimum(I, 0.0) ** 2)
    return float(c)


# --------------------------------------------------------------------------------------
# Solver 1: HiGHS QP on the lifted form.  x = [S (3(T+1)) | U (2T) | D (T) | W (T*M)],
# W_tm >= 0, W_tm >= D_t + fb_tm - fa_tm.S[0:2,t+1], cost 0.5*ro*W^2 (equivalent to neg()^2).
# --------------------------------------------------------------------------------------
def solve_highs(p: NrmpProblem):
    from scipy.optimize._highspy import _core as hp
    import scipy.sparse as sp

    T, M = p.T, p.M
    nS, nU, nD, nW = 3 * (T + 1), 2 * T, (T if M > 0 else 0), T * M
    n = nS + nU + nD + nW
    iS = lambda i, t: i * (T + 1) + t
    iU = lambda i, t: nS + i * T + t
    iD = lambda t: nS + nU + t
    iW = lambda t, m: nS + nU + nD + t * M + m
    inf = hp.kHighsInf

    hdiag = np.zeros(n)
    c = np.zeros(n)
    rows_s = 2 if p.omni else 3
    for i in range(3):
        for t in range(T + 1):
            qq = p.q[i] if i < rows_s else 0.0
            hdiag[iS(i, t)] = 2 * qq * qq + p.bk
            c[iS(i, t)] = -2 * qq * p.gamma_a[i, t] - p.bk * p.nom_s[i, t]
    for t in range(T):
        hdiag[iU(0, t)] = 2 * p.p_u ** 2
        c[iU(0, t)] = -2 * p.p_u * p.gamma_b[t]
    lb = np.full(n, -inf)
    ub = np.full(n, inf)
    for i in range(2):
        sb = p.speed_bound[i]
        for t in range(T):
            if np.isfinite(sb):
                lb[iU(i, t)], ub[iU(i, t)] = -sb, sb
    if M > 0:
        for t in range(T):
            c[iD(t)] = -p.eta
            lb[iD(t)] = max(p.d_min, 0.0)
            ub[iD(t)] = p.d_max
            for m in range(M):
                hdiag[iW(t, m)] = p.ro_obs
                lb[iW(t, m)] = 0.0

    ri, ci, vals, rl, ru = [], [], [], [], []
    r = 0

    def add(entries, lo, hi):
        nonlocal r
        for cc, v in entries:
            ri.append(r); ci.append(cc); vals.append(v)
        rl.append(lo); ru.append(hi)
        r += 1

    for i in range(3):  # initial state
        add([(iS(i, 0), 1.0)], p.nom_s[i, 0], p.nom_s[i, 0])
    for t in range(T):  # dynamics
        for i in range(3):
            e = [(iS(i, t + 1), 1.0)]
            e += [(iS(j, t), -p.A[t, i, j]) for j in range(3) if p.A[t, i, j] != 0.0]
            e += [(iU(j, t), -p.B[t, i, j]) for j in range(2) if p.B[t, i, j] != 0.0]
            add(e, p.C[t, i], p.C[t, i])
    for i in range(2):  # rate
        ab = p.acce_bound[i]
        if np.isfinite(ab):
            for t in range(T - 1):
                add([(iU(i, t + 1), 1.0), (iU(i, t), -1.0)], -ab, ab)
    for t in range(T):  # hinge slack: W - D + fa.S >= fb
        for m in range(M):
            e = [(iW(t, m), 1.0), (iD(t), -1.0)]
            e += [(iS(j, t + 1), p.fa[t, m, j]) for j in range(2) if p.fa[t, m, j] != 0.0]
            add(e, p.fb[t, m], inf)

    Amat = sp.csc_matrix((vals, (ri, ci)), shape=(r, n))
    lp = hp.HighsLp()
    lp.num_col_, lp.num_row_ = n, r
    lp.col_cost_, lp.col_lower_, lp.col_upper_ = c, lb, ub
    lp.row_lower_, lp.row_upper_ = np.array(rl), np.array(ru)
    lp.a_matrix_.format_ = hp.MatrixFormat.kColwise
    lp.a_matrix_.start_ = Amat.indptr.astype(np.int32)
    lp.a_matrix_.index_ = Amat.indices.astype(np.int32)
    lp.a_matrix_.value_ = Amat.data.astype(np.float64)
    model = hp.HighsModel()
    model.lp_ = lp
    nz = np.nonzero(hdiag)[0]
    start = np.zeros(n + 1, np.int32)
    start[1:] = np.cumsum(hdiag != 0)
    model.hessian_.dim_ = n
    model.hessian_.format_ = hp.HessianFormat.kTriangular
    model.hessian_.start_ = start
    model.hessian_.index_ = nz.astype(np.int32)
    model.hessian_.value_ = hdiag[nz]
    h = hp._Highs()
    h.setOptionValue("output_flag", False)
    h.setOptionValue("primal_feasibility_tolerance", 1e-8)
    h.setOptionValue("dual_feasibility_tolerance", 1e-8)
    h.passModel(model)
    h.run()
    status = h.getModelStatus()
    x = np.array(h.getSolution().col_value)
    S = x[:nS].reshape(3, T + 1)
    U = x[nS:nS + nU].reshape(2, T)
    D = x[nS + nU:nS + nU + nD].reshape(1, T) if M > 0 else None
    ok = status == hp.HighsModelStatus.kOptimal
    return S, U, D, ok


# --------------------------------------------------------------------------------------
# Solver-free optimality certificate (used to decide which solver is right when they differ).
# --------------------------------------------------------------------------------------
def kkt_certificate(p: NrmpProblem, S, U, D, act_tol: float = 1e-6):
    """Returns (stationarity_residual, primal_violation) of (S,U,D) for the program above.

    The program is convex, so a feasible point whose cost gradient (w.r.t. the free variables
    x = (U, D); S follows from the dynamics) is a non-negative combination of the normals of
    its active inequality constraints is optimal.  Multipliers are found by NNLS.
    """
    from scipy.optimize import nnls

    T, M = p.T, p.M
    nU = 2 * T
    nD = T if M > 0 else 0
    n = nU + nD
    S = np.asarray(S, float); U = np.asarray(U, float)
    D = np.zeros(0) if M == 0 else np.asarray(D, float).reshape(-1)
    # primal: dynamics + initial state
    viol = np.abs(S[:, 0] - p.nom_s[:, 0]).max()
    for t in range(T):
        viol = max(viol, np.abs(S[:, t + 1] - (p.A[t] @ S[:, t] + p.B[t] @ U[:, t] + p.C[t])).max())
    # sensitivities dS_{t+1}/dU
    F = np.zeros((T, 3, nU)); Fp = np.zeros((3, nU))
    for t in range(T):
        F[t] = p.A[t] @ Fp
        F[t][:, 2 * t:2 * t + 2] += p.B[t]
        Fp = F[t]
    rows = 2 if p.omni else 3
    gS = np.zeros((3, T + 1))
    gS[:rows] += 2 * p.q[:rows, None] * (p.q[:rows, None] * S[:rows] - p.gamma_a[:rows])
    gS += p.bk * (S - p.nom_s)
    g = np.zeros(n)
    g[0:nU:2] += 2 * p.p_u * (p.p_u * U[0] - p.gamma_b)
    if M > 0:
        g[nU:] = -p.eta
        for t in range(T):
            neg = np.minimum(p.fa[t] @ S[0:2, t + 1] - p.fb[t] - D[t], 0.0)
            gS[0:2, t + 1] += p.ro_obs * (p.fa[t].T @ neg)
            g[nU + t] += -p.ro_obs * neg.sum()
    for t in range(T):
        g[:nU] += F[t].T @ gS[:, t + 1]
    normals = []

    def consider(entries, slack):
        nonlocal viol
        viol = max(viol, -slack)
        if slack <= act_tol:
            a = np.zeros(n)
            for j, v in entries:
                a[j] = v
            normals.append(a)

    for i in range(2):
        sb, ab = p.speed_bound[i], p.acce_bound[i]
        for t in range(T):
            if np.isfinite(sb):
                consider([(2 * t + i, 1.0)], sb - U[i, t])
                consider([(2 * t + i, -1.0)], sb + U[i, t])
            if np.isfinite(ab) and t < T - 1:
                d = U[i, t + 1] - U[i, t]
                consider([(2 * t + 2 + i, 1.0), (2 * t + i, -1.0)], ab - d)
                consider([(2 * t + 2 + i, -1.0), (2 * t + i, 1.0)], ab + d)
    for t in range(nD):
        consider([(nU + t, 1.0)], p.d_max - D[t])
        consider([(nU + t, -1.0)], D[t] - max(p.d_min, 0.0))
    if normals:
        Aact = np.array(normals).T  # n x k
        z, rnorm = nnls(Aact, -g, maxiter=50 * Aact.shape[1] + 200)
        res = np.abs(g + Aact @ z).max()
    else:
        res = np.abs(g).max()
    return float(res), float(viol)
