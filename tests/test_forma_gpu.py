"""GPU parity of the formulation-A tick (canonical ISMPC with footsteps) against the CPU oracle."""
import numpy as np
import pytest

from quadruped_gait_generation_ismpc_b200 import abi, synth
from oracle import oracle as O
from parity import PRIMAL_TOL, primal_rel_err, active_set_mismatch, kkt_certificate

pytestmark = pytest.mark.gpu


def _advance(handle, inst, fs_timing, fs_plan, ticks):
    """Roll instance i forward ticks[i] ticks on the GPU (grouped by tick count) to get mid-gait states."""
    inst = inst.copy(); fs_plan = fs_plan.copy()
    for t in np.unique(ticks):
        if t == 0:
            continue
        sel = np.nonzero(ticks == t)[0]
        r = handle.forma_rollout(inst[sel], fs_timing, fs_plan, int(t), want_traj=False)
        assert (r["status"] & abi.ST_FAIL_MASK == 0).all()
        inst[sel] = r["inst"]
        for i in sel:   # only these instances' plan rows changed
            a = inst["plan_first_row"][i]; b = a + inst["n_fs"][i]
            fs_plan[a:b] = r["fs_plan"][a:b]
    return inst, fs_plan


def _compare(handle, model, inst, fs_timing, fs_plan, nthreads=8):
    g = handle.forma_solve_batch(inst, fs_timing, fs_plan)
    o = O.forma_batch(model, inst, fs_timing, fs_plan, nthreads=nthreads)
    ok = o["ret"] == 0
    assert ok.mean() > 0.9
    assert (g["out"]["status"][ok] & abi.ST_FAIL_MASK == 0).all()
    C, F = int(model["C"][0]), int(model["F"][0])
    # parity is per axis block [zd(C) | xf(F)]
    err = primal_rel_err(g["primal"][ok], o["primal"][ok])
    assert err.max() <= PRIMAL_TOL, "primal rel err %.3e" % err.max()
    assert np.abs(g["out"]["st"][ok] - o["out"]["st"][ok]).max() <= PRIMAL_TOL
    assert np.abs(g["out"]["pred_fs"][ok][:, :2 * F] - o["out"]["pred_fs"][ok][:, :2 * F]).max() <= PRIMAL_TOL
    mism, weak = active_set_mismatch(g["active"][ok], o["active"][ok], o["duals"][ok])
    assert mism.sum() == 0, "active set differs on %d rows (%d weak ignored)" % (mism.sum(), weak.sum())
    assert g["out"]["kkt_res"][ok].max() < 1e-8
    return g, o


def test_start_of_gait_trot(handle):
    model = abi.forma_model()
    handle.forma_set_model(model)
    inst, ft, plan = synth.forma_batch(64, gait="trot")
    _compare(handle, model, inst, ft, plan)


def test_mid_gait_trot_config2(handle):
    """Config 2 (formulation A): trot instances advanced to random gait phases, then one cold tick each."""
    model = abi.forma_model()
    handle.forma_set_model(model)
    n = 96
    inst, ft, plan = synth.forma_batch(n, gait="trot")
    rng = np.random.default_rng(11)
    ticks = rng.choice([0, 17, 49, 63, 98, 131, 207, 260], size=n)
    inst, plan = _advance(handle, inst, ft, plan, ticks)
    g, o = _compare(handle, model, inst, ft, plan)
    assert (np.abs(o["active"]).sum(axis=1) > 10).any(), "vacuous: no instance with a sizeable active set"


def test_mid_gait_walk_config3(handle):
    """Config 3: walking gait, Qf = 1e9, varied CoM height / step duration / ds."""
    model = abi.forma_model(q_foot=1e9)
    handle.forma_set_model(model)
    n = 96
    inst, ft, plan = synth.forma_batch(n, gait="walk", vary=True, ds=30, N_gait=108)
    rng = np.random.default_rng(12)
    ticks = rng.choice([0, 25, 61, 117, 180, 240], size=n)
    inst, plan = _advance(handle, inst, ft, plan, ticks)
    _compare(handle, model, inst, ft, plan)


def test_closed_loop_lockstep_with_push(handle):
    """Config 5 in small: GPU rollout vs the oracle run tick-by-tick from the GPU's previous state."""
    model = abi.forma_model()
    handle.forma_set_model(model)
    inst, ft, plan = synth.forma_batch(4, gait="trot", seed=3)
    push = synth.push_batch(4, seed=4)
    push["fs"] = 2
    T = 130
    r = handle.forma_rollout(inst, ft, plan, T, push=push)
    assert (r["status"] & abi.ST_FAIL_MASK == 0).all()
    cur, pl = inst.copy(), plan.copy()
    ct = np.zeros(4, dtype=int)
    for t in range(T):
        for i in range(4):
            if cur["fs_counter"][i] == push["fs"][i] and push["ct0"][i] <= ct[i] < push["ct1"][i]:
                cur["st"][i][1] += 0.01 * push["ax"][i]; cur["st"][i][4] += 0.01 * push["ay"][i]
        o = O.forma_batch(model, cur, ft, pl)
        assert (o["ret"] == 0).all()
        x = np.stack([o["out"]["st"][:, k] for k in (0, 3, 1, 4, 2, 5)], axis=1)
        assert np.abs(r["traj"][:, t] - x).max() <= 1e-6, "tick %d err %.3e" % (t, np.abs(r["traj"][:, t] - x).max())
        # continue from the GPU's state so errors do not compound
        for k, col in enumerate((0, 3, 1, 4, 2, 5)):
            cur["st"][:, col] = r["traj"][:, t, k]
        ct += 1
        for i in range(4):
            fc = cur["fs_counter"][i]
            if cur["j"][i] + 1 >= ft[cur["timing_first"][i] + fc]:
                fc += 1; cur["fs_counter"][i] = fc
                pred = np.array([o["out"]["pred_fs"][i][0], o["out"]["pred_fs"][i][3]])
                cur["cur_fs"][i] = pred; cur["fs_store"][i] = pred
                a = cur["plan_first_row"][i]; b = a + cur["n_fs"][i]
                pl[a:b] += pred - pl[a + fc - 1]
                cur["cl_first_ramp"][i] = 0; ct[i] = 0
        cur["j"] += 1
    assert np.array_equal(r["inst"]["fs_counter"], cur["fs_counter"])
    assert np.abs(r["fs_plan"] - pl).max() < 1e-6


def test_walking_fixture_closed_loop(handle):
    """The reference's own recorded output: CoM of walking/quad_walk_no_plots.m, phi=0, 10 cm steps
    (AMR_code_DART/MATLAB_trajectories/walking/phi0_10cm_50), first 200 ticks, 7 significant digits."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "matlab_fixtures.npz"))
    com = g["walk_phi0_com"]
    from quadruped_gait_generation_ismpc_b200 import plans
    model = abi.forma_model(q_foot=1e9)
    handle.forma_set_model(model)
    _, center = plans.walk_plan(phi=0.0)
    inst = np.zeros(1, dtype=abi.FORMA_INST)
    ft = np.arange(0, 2321, 50, dtype=np.int32)
    inst["st"][0] = [0.44, 0, 0.44, 0, 0, 0]; inst["cur_fs"][0] = center[0]; inst["fs_store"][0] = center[0]
    inst["height"] = 0.56; inst["wx"] = inst["wy"] = 0.02; inst["j"] = 1; inst["fs_counter"] = 1; inst["ds"] = 30
    inst["cl_first_ramp"] = 1; inst["n_timing"] = len(ft); inst["n_fs"] = center.shape[0]
    T = com.shape[0] - 1
    r = handle.forma_rollout(inst, ft, center, T)
    assert r["status"][0] & abi.ST_FAIL_MASK == 0
    err = np.abs(r["traj"][0, :T, :2] - com[1:T + 1, :2])
    assert err.max() < 5e-6, "max err vs MATLAB fixture %.3e" % err.max()   # quadprog tolerance + %e printing


def test_kinematic_rows_active(handle):
    """Footstep displacement bounds tight enough that kinematic rows enter the working set (bang.m:163-190)."""
    model = abi.forma_model(disp_forw=0.06, disp_forw_dummy=0.03, disp_L=0.05)
    handle.forma_set_model(model)
    n = 64
    inst, ft, plan = synth.forma_batch(n, gait="trot", seed=31)
    rng = np.random.default_rng(13)
    ticks = rng.choice([0, 17, 49, 63, 98, 131], size=n)
    inst, plan = _advance(handle, inst, ft, plan, ticks)
    g, o = _compare(handle, model, inst, ft, plan)
    C = int(model["C"][0])
    assert (o["active"][o["ret"] == 0][:, 2 * C:] != 0).any(), "vacuous: no kinematic row active"


def test_structured_solver_equals_dual_active_set(handle, monkeypatch):
    """The primal-dual active-set fast path and the dual active-set fallback are two exact methods: same answers."""
    model = abi.forma_model()
    handle.forma_set_model(model)
    n = 128
    inst, ft, plan = synth.forma_batch(n, gait="trot", seed=41)
    rng = np.random.default_rng(14)
    ticks = rng.choice([0, 9, 33, 49, 50, 77, 99, 100, 150, 222], size=n)
    inst, plan = _advance(handle, inst, ft, plan, ticks)
    a = handle.forma_solve_batch(inst, ft, plan)
    handle.set_option("forma_pdas", 0)
    try:
        b = handle.forma_solve_batch(inst, ft, plan)
    finally:
        handle.set_option("forma_pdas", 1)
    assert (b["out"]["status"] & abi.ST_GI_FALLBACK == 0).all()
    assert (a["out"]["status"] & abi.ST_FAIL_MASK == 0).all() and (b["out"]["status"] & abi.ST_FAIL_MASK == 0).all()
    assert (a["out"]["status"] & abi.ST_GI_FALLBACK != 0).mean() < 0.05, "fast path falls back too often"
    assert primal_rel_err(a["primal"], b["primal"]).max() <= 1e-7
    assert np.abs(a["out"]["st"] - b["out"]["st"]).max() <= 1e-8


def test_warm_started_rollout_equals_cold(handle, monkeypatch):
    """Closed loop with the working set carried from tick to tick == every tick solved from the empty set."""
    model = abi.forma_model()
    handle.forma_set_model(model)
    inst, ft, plan = synth.forma_batch(16, gait="trot", seed=51)
    push = synth.push_batch(16, seed=52)
    push["fs"] = 2
    w = handle.forma_rollout(inst, ft, plan, 160, push=push)
    handle.set_option("forma_warm", 0)
    try:
        c = handle.forma_rollout(inst, ft, plan, 160, push=push)
    finally:
        handle.set_option("forma_warm", 1)
    assert (w["status"] & abi.ST_FAIL_MASK == 0).all() and (c["status"] & abi.ST_FAIL_MASK == 0).all()
    assert np.abs(w["traj"] - c["traj"]).max() <= 1e-7
    assert np.array_equal(w["inst"]["fs_counter"], c["inst"]["fs_counter"])


@pytest.mark.parametrize("C,F,step", [(50, 3, 25), (37, 2, 20), (100, 5, 50), (100, 8, 50), (200, 3, 100), (400, 3, 200)])
def test_horizons_and_footstep_counts(handle, C, F, step):
    """configs[3] for formulation A (C = 50/100/200/400, ragged 37) and footstep counts other than the scripts' F = 3
    (F > 3 runs the kernel instantiation with the general saddle solve): mid-gait instances against the oracle.  The
    oracle's working-set budget is raised for the long horizons (nWSR = 300 is hit at C = 200)."""
    model = abi.forma_model(C=C, P=2 * C, F=F)
    handle.forma_set_model(model)
    n = 24 if C <= 200 else 8
    inst, ft, plan = synth.forma_batch(n, gait="trot", C=C, step=step, ds=max(2, step * 2 // 5), sim_ticks=24 * step,
                                       seed=70 + C + F)
    rng = np.random.default_rng(C + F)
    ticks = rng.choice([0, step // 3, step + 3, 2 * step + step // 2, 4 * step - 1], size=n)
    inst, plan = _advance(handle, inst, ft, plan, ticks)
    g = handle.forma_solve_batch(inst, ft, plan)
    o = O.forma_batch(model, inst, ft, plan, nthreads=8, nwsr_cap=4000)
    ok = o["ret"] == 0
    assert ok.mean() > 0.9
    assert (g["out"]["status"][ok] & abi.ST_FAIL_MASK == 0).all()
    # an optimality certificate that does not involve the oracle's SOLVER (only its dense builder): feasible,
    # stationary on the reported working set, multipliers of the right sign
    for i in np.nonzero(ok)[0]:
        it = inst[i]
        p = O.FormAParams(float(model["dt"][0]), float(np.sqrt(model["g_eta"][0] / it["height"])), float(it["wx"]), float(it["wy"]),
                          float(model["disp_forw"][0]), float(model["disp_forw_dummy"][0]), float(model["disp_L"][0]),
                          float(model["q_zdot"][0]), float(model["q_foot"][0]), C, 2 * C, F)
        tf, nt, a, nf = int(it["timing_first"]), int(it["n_timing"]), int(it["plan_first_row"]), int(it["n_fs"])
        Hd, gq, A, lb, ub = O.forma_build(p, it["st"], it["cur_fs"], it["fs_store"], int(it["j"]), int(it["fs_counter"]),
                                          ft[tf:tf + nt], int(it["ds"]), plan[a:a + nf], int(it["cl_first_ramp"]))
        feas, stat, wrong = kkt_certificate(Hd, gq, A, lb, ub, g["primal"][i], g["active"][i])
        assert feas <= 1e-9 and stat <= 1e-7 and wrong <= 1e-7, (i, feas, stat, wrong)
    err = primal_rel_err(g["primal"][ok], o["primal"][ok])
    if C >= 400:
        # At C = 400 qpOASES (Options::setToMPC, 200-350 working-set changes) stops with bound violations of ~2e-7 on
        # some instances (measured: its own A x leaves [lb, ub] by that much, this library's by 1e-16), which moves
        # its primal by up to 2.4e-6: the comparison is held to 1e-5 there and the certificate above carries the claim.
        assert err.max() <= 1e-5, "primal rel err %.3e" % err.max()
        return
    assert err.max() <= PRIMAL_TOL, "primal rel err %.3e" % err.max()
    assert np.abs(g["out"]["st"][ok] - o["out"]["st"][ok]).max() <= PRIMAL_TOL
    assert np.abs(g["out"]["pred_fs"][ok][:, :2 * F] - o["out"]["pred_fs"][ok][:, :2 * F]).max() <= PRIMAL_TOL
    mism, weak = active_set_mismatch(g["active"][ok], o["active"][ok], o["duals"][ok])
    assert mism.sum() == 0, "active set differs on %d rows (%d weak ignored)" % (mism.sum(), weak.sum())


def test_host_async_call_on_the_handles_stream(handle):
    """ISMPC_MEM_HOST_ASYNC + ismpc_handle_stream + ismpc_wait (the plumbing a host without CUDA headers uses) == the
    synchronous host-memory call; the buffers come from ismpc_host_alloc (pinned)."""
    import ctypes as C
    from quadruped_gait_generation_ismpc_b200 import binding
    model = abi.forma_model()
    handle.forma_set_model(model)
    n = 64
    inst, ft, plan = synth.forma_batch(n, gait="trot", seed=91)
    ref = handle.forma_solve_batch(inst, ft, plan)
    L = binding.lib()
    nV = 2 * (int(model["C"][0]) + int(model["F"][0]))
    ft = np.ascontiguousarray(ft, dtype=np.int32)
    sizes = [inst.nbytes, ft.nbytes, plan.nbytes, n * abi.FORMA_OUT.itemsize, n * nV * 8]
    ptrs = [L.ismpc_host_alloc(s) for s in sizes]
    assert all(ptrs)
    try:
        for p, a in zip(ptrs[:3], (inst, ft, plan)):
            C.memmove(p, a.ctypes.data, a.nbytes)
        st = handle.stream()
        handle.forma_solve_batch_raw(n, ptrs[0], ptrs[1], len(ft), ptrs[2], plan.shape[0], ptrs[3], primal=ptrs[4],
                                     mem=abi.MEM_HOST_ASYNC, stream=st)
        handle.wait(st)
        out = np.frombuffer(C.string_at(ptrs[3], sizes[3]), dtype=abi.FORMA_OUT)
        primal = np.frombuffer(C.string_at(ptrs[4], sizes[4]), dtype=np.float64).reshape(n, nV)
        assert out.tobytes() == ref["out"].tobytes()
        assert np.array_equal(primal, ref["primal"])
    finally:
        for p in ptrs:
            L.ismpc_host_free(p)


@pytest.mark.parametrize("gait", ["trot", "walk"])
def test_recorded_midgait_instances(handle, gait):
    """tests/golden/forma_midgait.npz: the kernel reproduces what it did when the fixture was recorded (same minimiser,
    same number of working-set iterations) -- the numbers tests/test_pdas_restatement.py reproduces in numpy on the CPU."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "forma_midgait.npz"))
    model, inst = gold[gait + "_model"], gold[gait + "_inst"]
    handle.forma_set_model(model)
    g = handle.forma_solve_batch(inst, gold[gait + "_fs_timing"], gold[gait + "_fs_plan"])
    assert (g["out"]["status"] & abi.ST_FAIL_MASK == 0).all()
    assert np.abs(g["primal"] - gold[gait + "_kernel_primal"]).max() <= 1e-9
    assert np.array_equal(g["active"], gold[gait + "_kernel_active"])
    assert np.array_equal(g["out"]["iters"], gold[gait + "_kernel_iters"])


def test_out_of_range_tables_are_flagged_not_read(handle):
    """Formulation A: records pointing outside the footstep-plan or timing tables get ISMPC_ST_QP_FAIL (tick and rollout)
    and leave the other instances' results as they are."""
    model = abi.forma_model()
    handle.forma_set_model(model)
    inst, ft, plan = synth.forma_batch(32, gait="trot", seed=93)
    ref = handle.forma_solve_batch(inst, ft, plan)
    bad = inst.copy()
    bad["plan_first_row"][2] = plan.shape[0] - 3
    bad["timing_first"][7] = len(ft) - 1
    bad["n_fs"][11] = 1 << 28
    g = handle.forma_solve_batch(bad, ft, plan)
    for i in (2, 7, 11):
        assert g["out"]["status"][i] & abi.ST_QP_FAIL
    ok = np.ones(32, bool); ok[[2, 7, 11]] = False
    assert np.array_equal(g["primal"][ok], ref["primal"][ok])
    r = handle.forma_rollout(bad, ft, plan, 12)
    for i in (2, 7, 11):
        assert r["status"][i] & abi.ST_QP_FAIL
    assert (r["status"][ok] & abi.ST_FAIL_MASK == 0).all()
