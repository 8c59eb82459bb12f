"""Status parity on the closed loops: where the GPU flags a failed tick, the reference's qpOASES
call (utils.cpp:121-130, whose return code utils.cpp:128 drops) rejects the same QP built from the same state, and
where the GPU solved, qpOASES solved.  The failing tick is located with the rollouts' per-tick status trace
(ismpc_formc_rollout_ex / ismpc_forma_rollout_ex2) and replayed from the GPU's own previous state, so nothing compounds.

Also here: the formulation-A closed loop with pushes at bench scale, checked against the oracle on sampled ticks."""
import numpy as np
import pytest

from quadruped_gait_generation_ismpc_b200 import abi, synth
from oracle import oracle as O
from parity import PRIMAL_TOL, primal_rel_err, active_set_mismatch, forma_certified_outliers

pytestmark = pytest.mark.gpu

FAIL_C = abi.ST_Z_FAIL | abi.ST_X_FAIL | abi.ST_Y_FAIL


def _walk_states_at(walk0, inst, plan, ticks_wanted, T):
    """Controller::update's bookkeeping (Controller.cpp:297-302 with the switch enabled, :503-504) replayed on the host:
    the walk state every instance has at the START of tick ticks_wanted[i] (after that tick's footstep switch)."""
    wk = walk0.copy()
    out = walk0.copy()
    first = inst["plan_first_row"].astype(np.int64)
    for t in range(T):
        fc = wk["footstep_counter"]
        can = (fc >= 0) & (fc < inst["n_steps"])
        tsw = plan[np.minimum(first + np.maximum(fc, 0), len(plan) - 1), 3]
        sw = can & (wk["sim_time"] >= tsw - 1.0)
        wk["control_iter"][sw] = 0; wk["mpc_iter"][sw] = 0
        wk["footstep_counter"][sw] += 1; wk["support_foot"][sw] = 1 - wk["support_foot"][sw]
        hit = ticks_wanted == t
        out[hit] = wk[hit]
        wk["control_iter"] += 1
        wk["mpc_iter"] = np.floor(wk["control_iter"] * 0.01 / 0.01).astype(np.int32)
        wk["sim_time"] += 1
    return out


def _state_before_tick(state0, traj, push, ticks):
    """State the rollout fed into tick ticks[i]: the previous tick's result (or the initial state) plus the push."""
    st = state0.copy()
    for i, t in enumerate(ticks):
        if t > 0:
            st["com_pos"][i] = traj[i, t - 1, :3]; st["com_vel"][i] = traj[i, t - 1, 3:]
        if push["ct0"][i] <= t < push["ct1"][i]:
            st["com_vel"][i, 0] += 0.01 * push["ax"][i]; st["com_vel"][i, 1] += 0.01 * push["ay"][i]
    return st


def test_formc_failed_closed_loop_ticks_fail_in_qpoases_too(handle):
    """configs[4]: 1,000 instances x 1,000 ticks with pushes.  Every instance with a failed tick: its FIRST failing tick,
    replayed as a single tick from the GPU's previous state, fails in the oracle on the same QP (z / x / y) -- and the
    tick before it, which the GPU solved, is solved by the oracle to 1e-6."""
    model = abi.formc_model()
    handle.formc_set_model(model)
    n, T = 1000, 1000
    state, walk, inst, plan = synth.formc_batch(n, seed=44, k0_cap=300)
    push = synth.push_batch(n, seed=45, formc=True)
    r = handle.formc_rollout(state, walk, inst, plan, T, push=push, want_trace=True)
    tr = r["trace"]
    assert np.array_equal(np.bitwise_or.reduce(tr, axis=1), r["status"]), "trace and accumulated status disagree"
    failed_tick = (tr & FAIL_C) != 0
    sel = np.nonzero(failed_tick.any(axis=1))[0]
    assert 1 <= len(sel) <= 40, "instances with a failed tick: %d (measured 26 of 1,000 on this input)" % len(sel)
    t_fail = failed_tick[sel].argmax(axis=1)
    for shift, expect_fail in ((0, True), (-1, False)):
        ticks = t_fail + shift
        keep = ticks >= 0
        s2, tk = sel[keep], ticks[keep]
        wk = _walk_states_at(walk[s2], inst[s2], plan, tk, T)
        st = _state_before_tick(state[s2], r["traj"][s2], push[s2], tk)
        g = handle.formc_solve_batch(st, wk, inst[s2], plan)
        # the single tick reproduces what the rollout recorded for that tick
        rec = tr[s2, tk]
        assert np.array_equal(g["out"]["status"] & FAIL_C, rec & FAIL_C)
        nxt = np.concatenate([g["out"]["next"]["com_pos"], g["out"]["next"]["com_vel"]], axis=1)
        assert np.abs(nxt - r["traj"][s2, tk]).max() <= 1e-12
        o = O.formc_batch(model, st, wk, inst[s2], plan, nthreads=8)
        zfail = o["ret"][:, 0] != 0
        assert np.array_equal((rec & abi.ST_Z_FAIL) != 0, zfail)
        for k, bit in ((1, abi.ST_X_FAIL), (2, abi.ST_Y_FAIL)):
            gk = (rec & bit) != 0
            assert np.array_equal(gk[~zfail], (o["ret"][:, k] != 0)[~zfail]), \
                "QP %d at shift %d: GPU %s vs oracle ret %s" % (k, shift, gk.astype(int), o["ret"][:, k])
        if expect_fail:
            assert ((rec & FAIL_C) != 0).all() and (o["ret"] != 0).any(axis=1).all()
        else:
            okk = (rec & FAIL_C) == 0
            assert okk.all() and (o["ret"][okk] == 0).all()
            err = primal_rel_err(g["primal"][okk].reshape(-1, 3, 100), o["primal"][okk].reshape(-1, 3, 100))
            assert err.max() <= PRIMAL_TOL
            mism, _ = active_set_mismatch(g["active"][okk], o["active"][okk], o["duals"][okk])
            assert mism.sum() == 0


def _forma_state_at(handle, inst, ft, plan, push, t):
    """Instance records and plans at the start of tick t of a pushed closed loop (rollout of t ticks from the start:
    deterministic, so this is the state the long rollout had), with that tick's push applied on the host
    (bang.m:104-114: ct counts the ticks since the last footstep switch)."""
    if t > 0:
        r = handle.forma_rollout(inst, ft, plan, int(t), push=push, want_traj=False)
        cur, pl = r["inst"], r["fs_plan"]
    else:
        cur, pl = inst.copy(), plan.copy()
    # ticks since the last switch: replay the timing
    n_pushed = 0
    for i in range(len(cur)):
        j, fsc, ct = int(inst["j"][i]), int(inst["fs_counter"][i]), 0
        tab = ft[inst["timing_first"][i]: inst["timing_first"][i] + inst["n_timing"][i]]
        for _ in range(int(t)):
            ct += 1
            if fsc + 1 <= len(tab) and j + 1 >= tab[fsc]:
                fsc += 1; ct = 0
            j += 1
        assert fsc == cur["fs_counter"][i] and j == cur["j"][i]
        if fsc == push["fs"][i] and push["ct0"][i] <= ct < push["ct1"][i]:
            cur["st"][i][1] += 0.01 * push["ax"][i]; cur["st"][i][4] += 0.01 * push["ay"][i]
            n_pushed += 1
    return cur, pl, n_pushed


def test_forma_closed_loop_bench_scale_sampled_ticks(handle):
    """The bench's formulation-A closed loop (trot, pushes, warm-started) on 64 instances x 250 ticks: on sampled ticks
    (inside and outside the push windows) a cold single tick from the rollout's own state equals the warm-started
    rollout's tick to 1e-9 and the oracle (qpOASES) to 1e-6 with the identical active set."""
    model = abi.forma_model()
    handle.forma_set_model(model)
    n, T = 64, 250
    inst, ft, plan = synth.forma_batch(n, gait="trot", seed=synth.SEED0 ^ 9)
    push = synth.push_batch(n)
    r = handle.forma_rollout_pred(inst, ft, plan, T, push=push, want_trace=True)
    assert (r["status"] & abi.ST_FAIL_MASK == 0).all()
    assert np.array_equal(np.bitwise_or.reduce(r["trace"].reshape(n, -1), axis=1), r["status"])
    checked_pushed = excused = 0
    for t in (0, 58, 103, 104, 110, 171, 249):
        cur, pl, n_pushed = _forma_state_at(handle, inst, ft, plan, push, t)
        checked_pushed += n_pushed
        g = handle.forma_solve_batch(cur, ft, pl)
        x = np.stack([g["out"]["st"][:, k] for k in (0, 3, 1, 4, 2, 5)], axis=1)
        assert np.abs(x - r["traj"][:, t]).max() <= 1e-9, "tick %d: cold tick vs warm rollout %.3e" % (t, np.abs(x - r["traj"][:, t]).max())
        o = O.forma_batch(model, cur, ft, pl, nthreads=8)
        ok = o["ret"] == 0
        assert ok.all(), "tick %d: oracle failed on %d instances the GPU solved" % (t, (~ok).sum())
        # over 1e-6 against qpOASES only where qpOASES itself stopped early, and then certified (parity.py)
        out = forma_certified_outliers(O, model, cur, ft, pl, g, o, ok, max_outliers=1)
        excused += len(out)
        keep = np.ones(n, bool); keep[out] = False
        assert primal_rel_err(g["primal"][keep], o["primal"][keep]).max() <= PRIMAL_TOL
        assert np.abs(g["out"]["st"][keep] - o["out"]["st"][keep]).max() <= PRIMAL_TOL
        mism, _ = active_set_mismatch(g["active"][keep], o["active"][keep], o["duals"][keep])
        assert mism.sum() == 0
    assert checked_pushed > 0, "vacuous: no sampled tick fell inside a push window"
    assert excused <= 2, "%d of %d sampled instance-ticks needed the certificate" % (excused, 7 * n)


def test_forma_infeasible_ticks_fail_in_qpoases_too(handle):
    """Pushes strong enough to make QP-1 infeasible on some instances (ISMPC_ST_QP_FAIL): the first failing tick of every
    such instance, replayed cold from the GPU's state, is rejected by qpOASES as well; instances the GPU solved on that
    tick are solved by qpOASES."""
    # footsteps may move by a few centimetres only (as in test_kinematic_rows_active), so a shove cannot be absorbed by
    # stepping wider; the push strength is raised until some -- not all -- instances become infeasible
    model = abi.forma_model(disp_forw=0.06, disp_forw_dummy=0.03, disp_L=0.05)
    handle.forma_set_model(model)
    n, T = 48, 140
    inst, ft, plan = synth.forma_batch(n, gait="trot", seed=61)
    base = synth.push_batch(n, seed=62)
    base["fs"] = 2
    for scale in (1.0, 2.0, 4.0, 8.0, 16.0, 32.0):
        push = base.copy()
        push["ax"] *= scale; push["ay"] *= scale
        r = handle.forma_rollout_pred(inst, ft, plan, T, push=push, want_trace=True)
        tr = (r["trace"] & abi.ST_QP_FAIL) != 0                   # n x T x 2
        failed = tr.any(axis=(1, 2))
        if 0 < failed.sum() < n:
            break
    assert 0 < failed.sum() < n, "want a mix of failed and solved instances, got %d of %d failed" % (failed.sum(), n)
    t_first = tr.any(axis=2).argmax(axis=1)
    for t in np.unique(t_first[failed]):
        sel = np.nonzero(failed & (t_first == t))[0]
        sel = np.concatenate([sel, np.nonzero(~failed)[0][:4]])      # plus a few instances that solved this tick
        cur, pl, _ = _forma_state_at(handle, inst[sel], ft, plan, push[sel], int(t))
        g = handle.forma_solve_batch(cur, ft, pl)
        gfail = (g["out"]["status"] & abi.ST_QP_FAIL) != 0
        assert np.array_equal(gfail, tr[sel, t].any(axis=1)), "cold tick and rollout disagree on which QPs fail"
        o = O.forma_batch(model, cur, ft, pl, nthreads=8)
        assert np.array_equal(gfail, o["ret"] != 0), "tick %d: GPU %s, oracle ret %s" % (t, gfail.astype(int), o["ret"])
        ok = ~gfail
        if ok.any():
            assert primal_rel_err(g["primal"][ok], o["primal"][ok]).max() <= PRIMAL_TOL
