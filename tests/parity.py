"""Parity definition shared by the GPU tests.

(i)  primal: max_i |x_gpu - x_ora| / max(1, |x_ora|_inf) <= 1e-6
(ii) active set: equal on every row that is strongly determined, i.e. excluding rows where the oracle's
     |multiplier| < 1e-9 and |slack| < 1e-9 do not both clear the threshold (weakly active / degenerate rows);
     those rows are counted and reported.
(iii) instances where the oracle's solver returned non-zero are excluded and counted.
"""
import numpy as np

PRIMAL_TOL = 1e-6   # BASELINE.json north_star: "within 1e-6 relative on the primal"


def primal_rel_err(x_gpu, x_ora):
    x_gpu = np.asarray(x_gpu, dtype=np.float64); x_ora = np.asarray(x_ora, dtype=np.float64)
    scale = np.maximum(1.0, np.abs(x_ora).max(axis=-1, keepdims=True))
    return (np.abs(x_gpu - x_ora) / scale).max(axis=-1)


def active_set_mismatch(as_gpu, as_ora, duals_ora, weak_tol=1e-9):
    """Number of strongly-determined rows that differ per instance, and number of weak rows.
    A row is weak when the oracle's multiplier is ~0 (it may sit in either working set)."""
    as_gpu = np.asarray(as_gpu).astype(np.int32); as_ora = np.asarray(as_ora).astype(np.int32)
    weak = np.abs(duals_ora) < weak_tol
    diff = (as_gpu != as_ora) & ~weak
    # a weak row may legitimately be reported active by one side only; but it must not be active on opposite sides
    opp = (as_gpu * as_ora == -1)
    return (diff | opp).sum(axis=-1), (weak & (as_gpu != as_ora)).sum(axis=-1)


def kkt_certificate(Hd, g, A, lb, ub, x, active, n_eq=2):
    """Solver-independent optimality check of x for  min 1/2 x'diag(Hd)x + g'x,  lb <= Ax <= ub  (first n_eq rows are
    equalities), given the claimed working set `active` (-1 lower / 0 / +1 upper per inequality row).
    Returns (feasibility violation, stationarity residual per variable relative to max(1, |H_jj x_j| + |g_j|), worst
    wrong-signed multiplier relative to the largest multiplier): all three ~0 certify the minimiser of a strictly convex
    QP."""
    r = A @ x
    scale = np.maximum(1.0, np.maximum(np.abs(lb), np.abs(ub)))
    feas = max(((lb - r) / scale).max(), ((r - ub) / scale).max(), 0.0)
    W = np.concatenate([np.arange(n_eq), n_eq + np.nonzero(active)[0]])
    grad = Hd * x + g
    y, *_ = np.linalg.lstsq(A[W].T, grad, rcond=None)                 # grad = A_W' y: y >= 0 at lower, <= 0 at upper bounds
    stat = (np.abs(A[W].T @ y - grad) / np.maximum(1.0, np.abs(Hd * x) + np.abs(g))).max()
    yi = y[n_eq:] * np.where(active[np.nonzero(active)[0]] < 0, 1.0, -1.0)   # must be >= 0
    wrong = max(0.0, -yi.min()) / max(1e-300, np.abs(y).max()) if len(yi) else 0.0
    return feas, stat, wrong


def forma_certified_outliers(O, model, inst, ft, plan, g, o, ok, max_outliers, loose_tol=1e-5):
    """Formulation A against qpOASES with the reference's options (setToMPC: termination tolerance 2.2e-7 on the
    homotopy, utils.cpp:122): on heavily constrained QPs qpOASES stops up to ~2e-7 outside its bounds, which moves ITS
    primal by a few 1e-6.  Instances over PRIMAL_TOL are therefore not waved through but CERTIFIED:
    the GPU's point must be the minimiser by a solver-independent KKT certificate built from the oracle's dense builder
    only (feasible 1e-9, stationary 1e-7 on its working set, multipliers of the right sign), agree with the oracle's
    second, exact solver (portable dual active set) to PRIMAL_TOL, stay within loose_tol of qpOASES, and there may be at
    most max_outliers of them.  Returns the outlier indices."""
    err = primal_rel_err(g["primal"], o["primal"])
    out = np.nonzero(ok & (err > PRIMAL_TOL))[0]
    assert len(out) <= max_outliers, "%d instances over %.0e against qpOASES (max err %.3e)" % (len(out), PRIMAL_TOL, err[ok].max())
    if len(out) == 0:
        return out
    C, F = int(model["C"][0]), int(model["F"][0])
    port = O.forma_batch(model, inst[out], ft, plan, solver=O.SOLVER_PORT, nthreads=4)
    assert (port["ret"] == 0).all()
    assert primal_rel_err(g["primal"][out], port["primal"]).max() <= PRIMAL_TOL, "GPU differs from the exact CPU solver too"
    for i in out:
        it = inst[i]
        p = O.FormAParams(float(model["dt"][0]), float(np.sqrt(model["g_eta"][0] / it["height"])), float(it["wx"]), float(it["wy"]),
                          float(model["disp_forw"][0]), float(model["disp_forw_dummy"][0]), float(model["disp_L"][0]),
                          float(model["q_zdot"][0]), float(model["q_foot"][0]), C, int(model["P"][0]), F)
        tf, nt, a, nf = int(it["timing_first"]), int(it["n_timing"]), int(it["plan_first_row"]), int(it["n_fs"])
        Hd, gq, A, lb, ub = O.forma_build(p, it["st"], it["cur_fs"], it["fs_store"], int(it["j"]), int(it["fs_counter"]),
                                          ft[tf:tf + nt], int(it["ds"]), plan[a:a + nf], int(it["cl_first_ramp"]))
        feas, stat, wrong = kkt_certificate(Hd, gq, A, lb, ub, g["primal"][i], g["active"][i])
        assert feas <= 1e-9 and stat <= 1e-7 and wrong <= 1e-7, (i, feas, stat, wrong)
        assert err[i] <= loose_tol, "instance %d: %.3e from qpOASES" % (i, err[i])
    return out
