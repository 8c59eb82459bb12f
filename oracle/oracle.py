"""ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg and --impl reference) may import
this module.  The product package never does; it fails loudly if its CUDA library is missing.

Two shared objects (see oracle/Makefile):
  oracle/_ref/libismpc_oracle_ref.so   restated builders + the reference's qpOASES 3.2 ("reference")
  oracle/_build/libismpc_oracle.so     restated builders + portable dual active-set         ("port")
"""
import ctypes as C
import os
import subprocess

import numpy as np

from quadruped_gait_generation_ismpc_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libismpc_oracle_ref.so")
PORT_SO = os.path.join(_HERE, "_build", "libismpc_oracle.so")
SOLVER_PORT, SOLVER_QPOASES = 0, 1

_libs = {}


def build(ref=True):
    """Compile the oracle (port always; ref only where /root/reference exists)."""
    targets = ["port"]
    if ref and os.path.exists("/root/reference/AMR_code_DART/qpOASES/QProblem.cpp"):
        targets.append("ref")
    subprocess.check_call(["make", "-s", "-C", _HERE, "-j8"] + targets)


def have_ref():
    return os.path.exists(REF_SO)


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def lib(kind="auto"):
    """kind: 'ref' (qpOASES-backed), 'port', or 'auto' (ref if present else port)."""
    if kind == "auto":
        kind = "ref" if have_ref() else "port"
    if kind in _libs:
        return _libs[kind]
    path = REF_SO if kind == "ref" else PORT_SO
    if not os.path.exists(path):
        build(ref=(kind == "ref"))
    L = C.CDLL(path)
    L.oracle_have_qpoases.restype = C.c_int
    L.oracle_hw_threads.restype = C.c_int
    _libs[kind] = L
    return L


def default_solver(kind="auto"):
    return SOLVER_QPOASES if lib(kind).oracle_have_qpoases() else SOLVER_PORT


def hw_threads():
    return int(lib().oracle_hw_threads())


def qp_batch(H, g, A, lbA, ubA, solver=None, nwsr_cap=300, nthreads=1, kind="auto"):
    """n dense QPs of one shape; returns dict(x, y, ws, ret, nwsr)."""
    L = lib(kind)
    if solver is None:
        solver = default_solver(kind)
    H = np.ascontiguousarray(H, dtype=np.float64)
    n, nV = H.shape[0], H.shape[1]
    A = np.ascontiguousarray(A, dtype=np.float64)
    nC = A.shape[1]
    g = np.ascontiguousarray(g, dtype=np.float64)
    lbA = np.ascontiguousarray(lbA, dtype=np.float64)
    ubA = np.ascontiguousarray(ubA, dtype=np.float64)
    x = np.zeros((n, nV)); y = np.zeros((n, nC))
    ws = np.zeros((n, nC), dtype=np.int32); ret = np.zeros(n, dtype=np.int32); it = np.zeros(n, dtype=np.int32)
    rc = L.oracle_qp_batch(C.c_int(solver), C.c_int(nwsr_cap), C.c_int(n), C.c_int(nV), C.c_int(nC),
                           _p(H), _p(g), _p(A), _p(lbA), _p(ubA), _p(x), _p(y), _p(ws, C.c_int),
                           _p(ret, C.c_int), _p(it, C.c_int), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError("oracle_qp_batch: solver kind %d unavailable" % solver)
    return dict(x=x, y=y, ws=ws, ret=ret, nwsr=it)


def qp_solve(H, g, A, lbA, ubA, **kw):
    r = qp_batch(H[None], g[None], A[None], lbA[None], ubA[None], **kw)
    return {k: v[0] for k, v in r.items()}


def formc_batch(model, state, walk, inst, plan, solver=None, nwsr_cap=300, nthreads=1, kind="auto",
                want_full=True):
    """Mirror of ismpc_formc_solve_batch on the CPU. Returns dict(out, primal, active, duals, ret, nwsr)."""
    L = lib(kind)
    if solver is None:
        solver = default_solver(kind)
    n = len(state)
    N = int(model["N"][0])
    plan = np.ascontiguousarray(plan, dtype=np.float64)
    out = np.zeros(n, dtype=abi.FORMC_OUT)
    primal = np.zeros((n, 3 * N)) if want_full else None
    active = np.zeros((n, 3 * N), dtype=np.int32) if want_full else None
    duals = np.zeros((n, 3 * N)) if want_full else None
    ret = np.zeros((n, 3), dtype=np.int32); it = np.zeros((n, 3), dtype=np.int32)
    rc = L.oracle_formc_batch(C.c_int(solver), C.c_int(nwsr_cap), _vp(model), C.c_int(n), _vp(state), _vp(walk),
                              _vp(inst), _p(plan), C.c_int(plan.shape[0]), _vp(out), _p(primal),
                              _p(active, C.c_int), _p(duals), _p(ret, C.c_int), _p(it, C.c_int),
                              C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError("oracle_formc_batch: solver kind %d unavailable" % solver)
    return dict(out=out, primal=primal, active=active, duals=duals, ret=ret, nwsr=it)


def forma_batch(model, inst, fs_timing, fs_plan, solver=None, nwsr_cap=300, nthreads=1, kind="auto",
                want_full=True):
    """Mirror of ismpc_forma_solve_batch on the CPU."""
    L = lib(kind)
    if solver is None:
        solver = default_solver(kind)
    n = len(inst)
    nV = 2 * (int(model["C"][0]) + int(model["F"][0]))
    fs_timing = np.ascontiguousarray(fs_timing, dtype=np.int32)
    fs_plan = np.ascontiguousarray(fs_plan, dtype=np.float64)
    out = np.zeros(n, dtype=abi.FORMA_OUT)
    primal = np.zeros((n, nV))  # always needed for pred_fs
    active = np.zeros((n, nV), dtype=np.int32) if want_full else None
    duals = np.zeros((n, nV)) if want_full else None
    ret = np.zeros(n, dtype=np.int32); it = np.zeros(n, dtype=np.int32)
    rc = L.oracle_forma_batch(C.c_int(solver), C.c_int(nwsr_cap), _vp(model), C.c_int(n), _vp(inst),
                              _p(fs_timing, C.c_int32), C.c_int(len(fs_timing)), _p(fs_plan),
                              C.c_int(fs_plan.shape[0]), _vp(out), _p(primal), _p(active, C.c_int),
                              _p(duals), _p(ret, C.c_int), _p(it, C.c_int), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError("oracle_forma_batch: solver kind %d unavailable" % solver)
    return dict(out=out, primal=primal, active=active, duals=duals, ret=ret, nwsr=it)


class FormCParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("dtc", C.c_double), ("h", C.c_double), ("mass", C.c_double),
                ("g", C.c_double), ("box_w", C.c_double), ("box_w_init", C.c_double),
                ("q_p", C.c_double), ("q_v", C.c_double), ("q_u", C.c_double), ("fz_max", C.c_double),
                ("N", C.c_int), ("S", C.c_int), ("F", C.c_int)]


class FormAParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("eta", C.c_double), ("wx", C.c_double), ("wy", C.c_double),
                ("disp_forw", C.c_double), ("disp_forw_dummy", C.c_double), ("disp_L", C.c_double),
                ("Qzdot", C.c_double), ("Qfoot", C.c_double), ("C", C.c_int), ("P", C.c_int), ("F", C.c_int)]


def formc_default_params(kind="auto"):
    p = FormCParams()
    lib(kind).oracle_formc_default_params(C.byref(p))
    return p


def formc_horizontal_qp(p, lam, cs, mid_q, footstep_counter, kind="auto"):
    """Literal stage-3 build for one axis: returns a, b, lo, hi, g, phi_state."""
    N = p.N
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    mid_q = np.ascontiguousarray(mid_q, dtype=np.float64)
    cs = np.ascontiguousarray(cs, dtype=np.float64)
    a = np.zeros(N); lo = np.zeros(N); hi = np.zeros(N); g = np.zeros(N); phi = np.zeros(4)
    b = C.c_double(0.0)
    lib(kind).oracle_formc_horizontal_qp(C.byref(p), _p(lam), _p(cs), _p(mid_q), C.c_int(footstep_counter),
                                         _p(a), C.byref(b), _p(lo), _p(hi), _p(g), _p(phi))
    return a, b.value, lo, hi, g, phi.reshape(2, 2)


def formc_midpoint(plan, S, F, kind="auto"):
    plan = np.ascontiguousarray(plan, dtype=np.float64)
    n = plan.shape[0]
    mid = np.zeros((n * (S + F), 3))
    lib(kind).oracle_formc_midpoint(_p(plan), C.c_int(n), C.c_int(S), C.c_int(F), _p(mid))
    return mid


def formc_vertical_qp(p, z0, mid_z, mpc_iter, footstep_counter, kind="auto"):
    N = p.N
    z0 = np.ascontiguousarray(z0, dtype=np.float64)
    mid_z = np.ascontiguousarray(mid_z, dtype=np.float64)
    H = np.zeros((N, N)); g = np.zeros(N)
    A = np.zeros((p.F + p.S + N + 2, N)); lb = np.zeros(p.F + p.S + N + 2); ub = np.zeros(p.F + p.S + N + 2)
    ne = C.c_int(0)
    L = lib(kind)
    L.oracle_formc_vertical_qp.restype = C.c_int
    nC = L.oracle_formc_vertical_qp(C.byref(p), _p(z0), _p(mid_z), C.c_int(mpc_iter), C.c_int(footstep_counter),
                                    _p(H), _p(g), _p(A), _p(lb), _p(ub), C.byref(ne))
    return H, g, A[:nC].copy(), lb[:nC].copy(), ub[:nC].copy(), ne.value


def forma_build(p, st, cur_fs, fs_store, j, fs_counter, fs_timing, ds, fs_plan, cl_first_ramp, kind="auto"):
    """Dense stacked QP of one form-A tick: returns Hdiag, g, A, lbA, ubA."""
    nV = 2 * (p.C + p.F); nC = nV + 2
    st = np.ascontiguousarray(st, dtype=np.float64)
    cur_fs = np.ascontiguousarray(cur_fs, dtype=np.float64)
    fs_store = np.ascontiguousarray(fs_store, dtype=np.float64)
    fs_timing = np.ascontiguousarray(fs_timing, dtype=np.int32)
    fs_plan = np.ascontiguousarray(fs_plan, dtype=np.float64)
    Hd = np.zeros(nV); g = np.zeros(nV); A = np.zeros((nC, nV)); lb = np.zeros(nC); ub = np.zeros(nC)
    lib(kind).oracle_forma_build(C.byref(p), _p(st), _p(cur_fs), _p(fs_store), C.c_int(j), C.c_int(fs_counter),
                                 _p(fs_timing, C.c_int), C.c_int(len(fs_timing)), C.c_int(ds), _p(fs_plan),
                                 C.c_int(fs_plan.shape[0]), C.c_int(cl_first_ramp),
                                 _p(Hd), _p(g), _p(A), _p(lb), _p(ub))
    return Hd, g, A, lb, ub


def kf_filter(model, state, samples, kind="auto"):
    """CPU restatement of StateFiltering::FilterWithKalman for n filters: model (1,) KF_MODEL, state (n,) KF_STATE,
    samples (n, n_steps) KF_SAMPLE.  Returns (state advanced, zmp (n, n_steps, 2))."""
    L = lib(kind)
    n, n_steps = samples.shape
    state = state.copy(); samples = np.ascontiguousarray(samples)
    zmp = np.zeros((n, n_steps, 2), dtype=np.float32)
    for i in range(n):
        L.oracle_kf_filter(_vp(model), C.c_void_p(state.ctypes.data + i * state.itemsize),
                           C.c_void_p(samples.ctypes.data + i * n_steps * samples.itemsize), C.c_int(n_steps),
                           C.c_void_p(zmp.ctypes.data + i * n_steps * 8))
    return state, zmp


class FeetParams(C.Structure):
    _fields_ = [("disp_forw", C.c_double), ("disp_i", C.c_double), ("disp_o", C.c_double),
                ("disp_forw_dummy", C.c_double), ("disp_i_dummy", C.c_double), ("disp_o_dummy", C.c_double)]


def feet_params(disp_forw=0.5, disp_i=0.4, disp_o=0.4):
    """init_quadruped.m:31-36 / init_quadruped2.m:31-36."""
    return FeetParams(disp_forw, disp_i, disp_o, disp_forw / 2, disp_i / 2, disp_o / 2)


def feet_trot_tick(fp, fs_counter, pred, phi, foot_plan, kind="auto"):
    """In place on foot_plan (rows x 8, C-contiguous float64). Returns changed."""
    L = lib(kind)
    pred = np.ascontiguousarray(pred, dtype=np.float64)
    L.oracle_feet_trot_tick.restype = C.c_int
    return L.oracle_feet_trot_tick(C.byref(fp), C.c_int(fs_counter), _p(pred), C.c_double(phi), _p(foot_plan),
                                   C.c_int(foot_plan.shape[0]))


def feet_walk_tick(fp, counter, fs_counter, pred, foot_plan, kind="auto"):
    L = lib(kind)
    pred = np.ascontiguousarray(pred, dtype=np.float64)
    L.oracle_feet_walk_tick.restype = C.c_int
    return L.oracle_feet_walk_tick(C.byref(fp), C.c_int(counter), C.c_int(fs_counter), _p(pred), _p(foot_plan),
                                   C.c_int(foot_plan.shape[0]))


def feet_export(foot_plan, n_steps, gait, fixed=30, swing=50, step_duration=50, kind="auto"):
    """Returns dict(fl, fr, rl, rr) of (samples x 3) arrays as the scripts write them to foot_*.txt."""
    L = lib(kind)
    foot_plan = np.ascontiguousarray(foot_plan, dtype=np.float64)
    n_steps = min(n_steps, foot_plan.shape[0] - 1)
    per = fixed + swing if gait == "trot" else step_duration
    out = {k: np.zeros((n_steps * per, 3)) for k in ("fl", "fr", "rl", "rr")}
    if gait == "trot":
        L.oracle_feet_export_trot(_p(foot_plan), C.c_int(foot_plan.shape[0]), C.c_int(n_steps), C.c_int(fixed),
                                  C.c_int(swing), _p(out["fl"]), _p(out["fr"]), _p(out["rl"]), _p(out["rr"]))
    else:
        L.oracle_feet_export_walk(_p(foot_plan), C.c_int(foot_plan.shape[0]), C.c_int(n_steps), C.c_int(step_duration),
                                  _p(out["fl"]), _p(out["fr"]), _p(out["rl"]), _p(out["rr"]))
    return out


def forma_closed_loop_pred(p, st, fs_plan, fs_timing, ds, n_ticks, push=(0, 0, 0, 0.0, 0.0), solver=None,
                           kind="auto", nwsr_cap=300):
    """forma_closed_loop that also returns the per-tick predicted footstep (n_ticks x 2) and fsCounter (n_ticks)."""
    L = lib(kind)
    if solver is None:
        solver = default_solver(kind)
    fn = L.oracle_qpoases_solve if solver == SOLVER_QPOASES else L.oracle_qp_dual_active_set
    if solver == SOLVER_QPOASES:
        L.oracle_qpoases_set_nwsr(C.c_int(nwsr_cap))
    st = np.array(st, dtype=np.float64)
    fs_plan = np.array(fs_plan, dtype=np.float64)
    fs_timing = np.ascontiguousarray(fs_timing, dtype=np.int32)
    traj = np.zeros((n_ticks, 6)); pred = np.zeros((n_ticks, 2)); fsc = np.zeros(n_ticks, dtype=np.int32)
    wsr = C.c_int(0)
    L.oracle_forma_closed_loop2.restype = C.c_int
    fails = L.oracle_forma_closed_loop2(C.byref(p), fn, _p(st), _p(fs_plan), C.c_int(fs_plan.shape[0]),
                                        _p(fs_timing, C.c_int), C.c_int(len(fs_timing)), C.c_int(ds),
                                        C.c_int(n_ticks), C.c_int(push[0]), C.c_int(push[1]), C.c_int(push[2]),
                                        C.c_double(push[3]), C.c_double(push[4]), _p(traj), C.byref(wsr),
                                        _p(pred), _p(fsc, C.c_int))
    return traj, fails, pred, fsc, fs_plan


def forma_closed_loop(p, st, fs_plan, fs_timing, ds, n_ticks, push=(0, 0, 0, 0.0, 0.0), solver=None,
                      kind="auto", nwsr_cap=300):
    """MATLAB closed loop (QP-1 only). Returns traj (n_ticks x 6: x,y,xd,yd,xz,yz), fails, nwsr_total, final plan."""
    L = lib(kind)
    if solver is None:
        solver = default_solver(kind)
    fn = L.oracle_qpoases_solve if solver == SOLVER_QPOASES else L.oracle_qp_dual_active_set
    if solver == SOLVER_QPOASES:
        L.oracle_qpoases_set_nwsr(C.c_int(nwsr_cap))
    st = np.array(st, dtype=np.float64)
    fs_plan = np.array(fs_plan, dtype=np.float64)
    fs_timing = np.ascontiguousarray(fs_timing, dtype=np.int32)
    traj = np.zeros((n_ticks, 6))
    wsr = C.c_int(0)
    L.oracle_forma_closed_loop.restype = C.c_int
    fails = L.oracle_forma_closed_loop(C.byref(p), fn, _p(st), _p(fs_plan), C.c_int(fs_plan.shape[0]),
                                       _p(fs_timing, C.c_int), C.c_int(len(fs_timing)), C.c_int(ds),
                                       C.c_int(n_ticks), C.c_int(push[0]), C.c_int(push[1]), C.c_int(push[2]),
                                       C.c_double(push[3]), C.c_double(push[4]), _p(traj), C.byref(wsr))
    return traj, fails, wsr.value, fs_plan
