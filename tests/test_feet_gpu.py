"""GPU: the stage after the hot path -- QP-1 closed loop -> real-foot placement -> export -- against the CPU oracle
and against the reference's own recorded foot trajectories (tests/golden/matlab_feet_fixtures.npz)."""
import os

import numpy as np
import pytest

from quadruped_gait_generation_ismpc_b200 import abi, export, plans, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "matlab_feet_fixtures.npz"))


def _instances(gait, phis, disp_A, step, ds, n_timing_ticks):
    n = len(phis)
    inst = np.zeros(n, dtype=abi.FORMA_INST); finst = np.zeros(n, dtype=abi.FEET_INST)
    ft = np.arange(0, n_timing_ticks, step, dtype=np.int32)
    centers, foots = [], []
    for i, phi in enumerate(phis):
        fp, c = (plans.trot_plan if gait == "trot" else plans.walk_plan)(phi=phi, disp_A=disp_A)
        inst["st"][i] = [c[0, 0], 0, c[0, 0], c[0, 1], 0, c[0, 1]]; inst["cur_fs"][i] = c[0]; inst["fs_store"][i] = c[0]
        inst["plan_first_row"][i] = sum(len(x) for x in centers); inst["n_fs"][i] = len(c)
        finst["plan_first_row"][i] = sum(len(x) for x in foots); finst["plan_rows"][i] = len(fp); finst["phi"][i] = phi
        centers.append(c); foots.append(fp)
    inst["height"] = 0.56; inst["wx"] = inst["wy"] = 0.02; inst["j"] = 1; inst["fs_counter"] = 1; inst["ds"] = ds
    inst["cl_first_ramp"] = 1; inst["n_timing"] = len(ft)
    finst["j"] = 1; finst["fs_counter"] = 1; finst["n_timing"] = len(ft)
    return inst, finst, ft, np.vstack(centers), np.ascontiguousarray(np.vstack(foots))


def _pipeline(handle, gait, model, inst, finst, ft, center, foot, T, n_steps, fixed, swing, push=None):
    handle.forma_set_model(model)
    r = handle.forma_rollout_pred(inst, ft, center, T, push=push)
    assert (r["status"] & abi.ST_FAIL_MASK == 0).all()
    fm = abi.feet_model(gait)
    foot2 = handle.feet_place_rollout(fm, finst, ft, r["pred"], foot)
    ex = handle.feet_export(fm, finst, foot2, n_steps, fixed, swing)
    return r, foot2, ex


@pytest.mark.parametrize("gait", ["walk", "trot"])
def test_feet_stage_matches_oracle(handle, gait):
    """Same predicted footsteps in, same foot plan and foot trajectories out (GPU kernels vs the C restatement)."""
    phis = [0.0, np.pi / 4, np.pi / 2, 0.3]
    if gait == "walk":
        inst, finst, ft, center, foot = _instances("walk", phis, 0.1, 50, 30, 2321)
        model, T, n_steps, fixed, swing = abi.forma_model(q_foot=1e9), 460, 12, 0, 50
    else:
        inst, finst, ft, center, foot = _instances("trot", phis, 0.12, 50, 20, 2321)
        model, T, n_steps, fixed, swing = abi.forma_model(), 330, 6, 20, 30
    push = synth.push_batch(len(phis), seed=77)       # a shove during the second step, so that footsteps get adapted
    push["fs"] = 2
    r, foot2, ex = _pipeline(handle, gait, model, inst, finst, ft, center, foot, T, n_steps, fixed, swing, push=push)
    fp = O.feet_params()
    for i, phi in enumerate(phis):
        a = finst["plan_first_row"][i]; b = a + finst["plan_rows"][i]
        ref = np.ascontiguousarray(foot[a:b].copy())
        j, fsc = 1, 1
        for t in range(T):
            if gait == "walk":
                O.feet_walk_tick(fp, fsc, fsc, r["pred"][i, t], ref)
            else:
                O.feet_trot_tick(fp, fsc, r["pred"][i, t], phi, ref)
            if fsc + 1 <= len(ft) and j + 1 >= ft[fsc]:
                fsc += 1
            j += 1
        assert np.abs(foot2[a:b] - ref).max() < 1e-12, "foot plan differs for phi=%g" % phi
        oe = O.feet_export(ref, n_steps, gait, fixed=fixed, swing=swing, step_duration=swing)
        for k in ("fl", "fr", "rl", "rr"):
            assert np.abs(ex[k][i] - oe[k]).max() < 1e-12, k
    assert np.abs(foot2 - foot).max() > 1e-4, "vacuous: the second stage changed nothing"


def test_walking_pipeline_matches_the_reference_files(handle):
    """All three recorded walking runs at once (phi = 0, pi/4, pi/2): CoM and the four feet, 2 000 samples each."""
    phis = [0.0, np.pi / 4, np.pi / 2]
    inst, finst, ft, center, foot = _instances("walk", phis, 0.1, 50, 30, 2321)
    r, foot2, ex = _pipeline(handle, "walk", abi.forma_model(q_foot=1e9), inst, finst, ft, center, foot, 460, 40, 0, 50)
    for i, key in enumerate(("walk_phi0", "walk_phipi4", "walk_phipi2")):
        pos, _ = export.com_rows(inst["st"][i], r["traj"][i], 0.56)
        assert np.abs(pos[:460] - GOLD[key + "_com"][:460]).max() < 5e-5
        for k in ("fl", "fr", "rl", "rr"):
            assert np.abs(ex[k][i] - GOLD["%s_%s" % (key, k)]).max() < 5e-5, (key, k)


def test_trotting_pipeline_matches_the_reference_files(handle):
    """quad_as_bip_no_plots.m, C = 160, 80-tick steps, the full 2 000 ticks: CoM and feet of the recorded runs."""
    model = abi.forma_model(C=160, P=320)
    for key, phi, dA, feet in (("trot_phi0", 0.0, 0.15, ("fr", "rl")), ("trot_phipi2", np.pi / 2, 0.15, ("fl", "fr", "rl", "rr")),
                               ("trot_phipi4_10cm", np.pi / 4, 0.10, ("fl", "fr", "rl", "rr"))):
        inst, finst, ft, center, foot = _instances("trot", [phi], dA, 80, 50, 3000)
        r, foot2, ex = _pipeline(handle, "trot", model, inst, finst, ft, center, foot, 2000, 25, 30, 50)
        pos, _ = export.com_rows(inst["st"][0], r["traj"][0], 0.56)
        assert np.abs(pos - GOLD[key + "_com"]).max() < 1e-5, key
        for k in feet:
            assert np.abs(ex[k][0] - GOLD["%s_%s" % (key, k)]).max() < 1e-5, (key, k)


@pytest.mark.parametrize("gait", ["trot", "walk"])
def test_plan_generators_on_device(handle, gait):
    """init_quadruped.m / init_quadruped2.m on the device == the host mirrors in plans.py (which the reference's
    recorded trajectories pin, see the pipeline tests above), for random step lengths and headings incl. clipped ones."""
    rng = np.random.default_rng(3)
    n = 64
    req = np.zeros(n, dtype=abi.PLAN_REQ)
    req["disp_A"] = rng.uniform(0.03, 0.7, n)                    # beyond 0.5 / 0.4 the step is clipped to the admissible region
    req["phi"] = rng.choice([0.0, np.pi / 4, np.pi / 2, 0.3, 1.2], size=n)
    req["disp_A"][:3] = [0.1, 0.15, 0.1]; req["phi"][:3] = [0.0, np.pi / 4, np.pi / 2]
    for N_gait in (100, 37):
        fp, ce = handle.plan_generate(abi.plan_model(gait, N_gait=N_gait), req)
        gen = plans.trot_plan if gait == "trot" else plans.walk_plan
        for i in range(n):
            rfp, rce = gen(N_gait=N_gait, disp_A=float(req["disp_A"][i]), phi=float(req["phi"][i]))
            assert fp[i].shape == rfp.shape and ce[i].shape[0] >= rce.shape[0] - 0
            assert np.abs(fp[i] - rfp).max() <= 1e-13, (gait, N_gait, i)
            m = min(len(ce[i]), len(rce))
            assert np.abs(ce[i][:m] - rce[:m]).max() <= 1e-12, (gait, N_gait, i)


@pytest.mark.parametrize("gait", ["trot", "walk"])
def test_out_of_range_feet_records_are_skipped_not_read(handle, gait):
    """Records whose plan rows or timing entries lie outside the tables handed to the call (or whose 1-based footstep
    counter is < 1) are skipped: the placement leaves their plan rows -- and everybody else's -- as they would be without
    them, the export writes zeros for them; nothing is read or written out of bounds."""
    phis = [0.0, np.pi / 4, np.pi / 2, 0.3, 0.0, 0.7]
    if gait == "walk":
        inst, finst, ft, center, foot = _instances("walk", phis, 0.1, 50, 30, 2321)
        model, T, n_steps, fixed, swing = abi.forma_model(q_foot=1e9), 230, 8, 0, 50
    else:
        inst, finst, ft, center, foot = _instances("trot", phis, 0.12, 50, 20, 2321)
        model, T, n_steps, fixed, swing = abi.forma_model(), 230, 5, 20, 30
    handle.forma_set_model(model)
    r = handle.forma_rollout_pred(inst, ft, center, T)
    fm = abi.feet_model(gait)
    good = handle.feet_place_rollout(fm, finst, ft, r["pred"], foot)
    ex_good = handle.feet_export(fm, finst, good, n_steps, fixed, swing)
    bad = finst.copy()
    bad["plan_first_row"][1] = foot.shape[0] - 3            # rows run off the end of the table
    bad["plan_first_row"][2] = -7                           # negative
    bad["timing_first"][3] = len(ft) - 2                    # timing entries run off the end
    bad["fs_counter"][4] = 0                                # 1-based counter below 1
    skipped = [1, 2, 3, 4]
    out = handle.feet_place_rollout(fm, bad, ft, r["pred"], foot)
    for i in range(len(phis)):
        a = finst["plan_first_row"][i]; b = a + finst["plan_rows"][i]
        if i in skipped:
            assert np.array_equal(out[a:b], foot[a:b]), "instance %d was skipped but its rows changed" % i
        else:
            assert np.array_equal(out[a:b], good[a:b]), "instance %d is valid but differs" % i
    ex = handle.feet_export(fm, bad, good, n_steps, fixed, swing)
    for k in ("fl", "fr", "rl", "rr"):
        for i in range(len(phis)):
            if i in (1, 2):                                  # the export reads plan rows only
                assert not ex[k][i].any()
            else:
                assert np.array_equal(ex[k][i], ex_good[k][i])
