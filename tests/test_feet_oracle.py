"""CPU: the oracle's restatement of the second stage (real-foot placement) and of the trajectory export, pinned
against the reference's OWN recorded outputs (AMR_code_DART/MATLAB_trajectories/**/foot_*.txt; excerpts committed by
tests/golden/make_feet_golden.py), and the wire format writers."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from quadruped_gait_generation_ismpc_b200 import export, plans

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "matlab_feet_fixtures.npz"))


def _params(C, Qf):
    p = O.FormAParams()
    p.dt, p.eta, p.wx, p.wy = 0.01, float(np.sqrt(9.8 / 0.56)), 0.02, 0.02
    p.disp_forw, p.disp_forw_dummy, p.disp_L, p.Qzdot, p.Qfoot, p.C, p.P, p.F = 0.5, 0.25, 0.4, 1.0, Qf, C, 2 * C, 3
    return p


@pytest.mark.parametrize("key,phi", [("walk_phi0", 0.0), ("walk_phipi4", np.pi / 4), ("walk_phipi2", np.pi / 2)])
def test_walking_feet_match_the_reference_files(key, phi):
    """quad_walk_no_plots.m end to end: QP-1 closed loop (460 ticks: the 8-phase counter never wraps there, so the
    second QP acts during the first eight steps only), second QP, export of all 2 000 samples of the four feet."""
    foot_plan, center = plans.walk_plan(phi=phi)
    foot_plan = np.ascontiguousarray(foot_plan)
    T = 460
    traj, fails, pred, fsc, _ = O.forma_closed_loop_pred(_params(100, 1e9), [center[0, 0], 0, center[0, 0], center[0, 1], 0, center[0, 1]],
                                                         center, np.arange(0, 2321, 50), 30, T, solver=O.SOLVER_PORT)
    assert fails == 0
    assert np.abs(traj[:T - 1, :2] - GOLD[key + "_com"][1:T, :2]).max() < 5e-5
    fp = O.feet_params()
    for j in range(T):
        O.feet_walk_tick(fp, int(fsc[j]), int(fsc[j]), pred[j], foot_plan)
    ex = O.feet_export(foot_plan, 40, "walk", step_duration=50)
    for k in ("fl", "fr", "rl", "rr"):
        assert np.abs(ex[k] - GOLD["%s_%s" % (key, k)]).max() < 5e-5, k     # quadprog tolerance + 7 printed digits


def test_trot_feet_first_steps_match_the_reference_file():
    """quad_as_bip_no_plots.m (C = 160, 80-tick steps): the first 3 steps' samples of foot_fr / foot_rl, which depend on
    the second QP of the first 240 ticks only (the full 2 000-tick check runs on the GPU, test_feet_gpu.py)."""
    phi = 0.0
    foot_plan, center = plans.trot_plan(phi=phi, disp_A=0.15)
    foot_plan = np.ascontiguousarray(foot_plan)
    T = 245
    traj, fails, pred, fsc, _ = O.forma_closed_loop_pred(_params(160, 1e7), [center[0, 0], 0, center[0, 0], center[0, 1], 0, center[0, 1]],
                                                         center, np.arange(0, 3000, 80), 50, T, solver=O.SOLVER_PORT)
    assert fails == 0
    fp = O.feet_params()
    for j in range(T):
        O.feet_trot_tick(fp, int(fsc[j]), pred[j], phi, foot_plan)
    ex = O.feet_export(foot_plan, 3, "trot", fixed=30, swing=50)
    for k in ("fr", "rl"):
        # row 4 of the plan is final only after step 3 ends: compare the samples that use rows 1..3 (2 steps)
        assert np.abs(ex[k][:160] - GOLD["trot_phi0_" + k][:160]).max() < 5e-6, k


def test_wire_format_round_trip(tmp_path):
    """fprintf('%d %d %d\\n', doubles): integers as integers, everything else as %e; first line of the reference's
    ComTrajectory file is '4.400000e-01 0 5.600000e-01'."""
    assert export.format_rows([[0.44, 0.0, 0.56]]) == "4.400000e-01 0 5.600000e-01\n"
    assert export.format_value(-3.0) == "-3" and export.format_value(1e-7) == "1.000000e-07"
    rng = np.random.default_rng(0)
    pos = rng.normal(size=(50, 3)); pos[:, 2] = 0.56
    vel = rng.normal(size=(50, 3)); vel[:, 2] = 0.0
    feet = {k: rng.normal(size=(80, 3)) for k in ("fl", "fr", "rl", "rr")}
    export.write_all(str(tmp_path), "walk_test", pos, vel, feet)
    back = export.read_rows(os.path.join(str(tmp_path), "foot_fl_walk_test.txt"))
    assert np.abs(back - feet["fl"]).max() < 1e-6 * np.abs(feet["fl"]).max()
    assert open(os.path.join(str(tmp_path), "ComVelocity_walk_test.txt")).readline().split()[2] == "0"


def test_cpp_writer_matches_python(tmp_path):
    """host/TrajectoryWriter.hpp produces byte-identical files."""
    import subprocess
    src = tmp_path / "w.cpp"
    src.write_text('''#include "TrajectoryWriter.hpp"
int main(int, char** argv) { double r[9] = {0.44, 0.0, 0.56, -1.25e-7, 3.0, 12345.678, 1e10, -0.0, 2.5};
  return ismpc_host::write_rows(argv[1], r, 3) ? 0 : 1; }''')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "w"
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(root, "quadruped_gait_generation_ismpc_b200", "host"),
                           str(src), "-o", str(exe)])
    out = tmp_path / "o.txt"
    subprocess.check_call([str(exe), str(out)])
    rows = [[0.44, 0.0, 0.56], [-1.25e-7, 3.0, 12345.678], [1e10, -0.0, 2.5]]
    assert out.read_text() == export.format_rows(rows)
