"""The C++ MPCSolver mirror (host/MPCSolver.hpp) driven like AMR_code_DART/Controller.cpp drives the reference class."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from quadruped_gait_generation_ismpc_b200 import abi, binding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_driver(d):
    exe = os.path.join(d, "shim_driver")
    libdir = os.path.dirname(binding.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "shim_driver.cpp"),
                           "-L" + libdir, "-lismpc_b200", "-Wl,-rpath," + libdir])
    return exe


def test_shim_compiles_and_fails_loudly_without_gpu():
    import torch
    with tempfile.TemporaryDirectory() as d:
        exe = _build_driver(d)
        if torch.cuda.is_available():
            pytest.skip("GPU present")
        r = subprocess.run([exe, "1"], capture_output=True, text=True)
        assert r.returncode != 0 and "ismpc_create" in r.stderr     # no CPU fallback behind the class


@pytest.mark.gpu
def test_shim_closed_loop_matches_oracle():
    """Controller-style loop through the C++ class == the CPU oracle (qpOASES) run in lock-step."""
    from oracle import oracle as O
    T = 30
    with tempfile.TemporaryDirectory() as d:
        exe = _build_driver(d)
        out = subprocess.check_output([exe, str(T)], text=True)
    traj = np.array([[float(x) for x in ln.split()] for ln in out.strip().splitlines()])
    assert traj.shape == (T, 7) and (traj[:, 6] == 0).all()
    model = abi.formc_model()
    state, walk, inst, plan = synth.reference_formc_instance()
    for k in range(T):
        walk["sim_time"] = k
        o = O.formc_batch(model, state, walk, inst, plan)
        assert (o["ret"] == 0).all()
        nxt = np.concatenate([o["out"]["next"]["com_pos"][0], o["out"]["next"]["com_vel"][0]])
        assert np.abs(traj[k, :6] - nxt).max() < 1e-6, "tick %d" % k
        # lock-step: continue from the GPU's state
        state["com_pos"][0] = traj[k, :3]; state["com_vel"][0] = traj[k, 3:6]
        walk["control_iter"] += 1
        walk["mpc_iter"] = int(np.floor(walk["control_iter"][0] * 0.01 / 0.01))


def _hostlib():
    import ctypes as C
    L = C.CDLL(os.path.join(os.path.dirname(binding.LIB_PATH), "libismpc_host.so"))
    L.ismpc_host_pipeline_create.restype = C.c_void_p
    L.ismpc_host_pipeline_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.ismpc_host_pipeline_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.ismpc_host_pipelines_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.ismpc_host_pipeline_destroy.argtypes = [C.c_void_p]
    L.ismpc_host_last_error.restype = C.c_char_p
    return L


def test_host_library_loads_and_fails_loudly_without_gpu():
    """lib/libismpc_host.so (the C++ serving loop, g++ only) exports its entry points; without a GPU the pipeline
    cannot be created and says why."""
    import torch
    L = _hostlib()
    for name in ("ismpc_host_pipeline_create", "ismpc_host_pipeline_run", "ismpc_host_pipeline_destroy",
                 "ismpc_host_pipeline_launches", "ismpc_host_last_error"):
        assert hasattr(L, name)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    model = abi.formc_model()
    plan = np.zeros((4, 4))
    assert not L.ismpc_host_pipeline_create(0, 8, 2, model.ctypes.data, 35, 10, plan.ctypes.data, 4)
    assert b"ismpc_create" in L.ismpc_host_last_error()


@pytest.mark.gpu
def test_cpp_serving_loop_equals_direct_calls(handle):
    """host/FormCPipeline.hpp: several ticks in flight on several handles give the records of one synchronous call each."""
    import ctypes as C
    L = _hostlib()
    n, depth, nb = 256, 3, 5
    model = abi.formc_model()
    batches = [synth.formc_batch(n, seed=60 + b, n_steps=40) for b in range(nb)]
    plans = np.concatenate([b[3] for b in batches])
    blocks, keep = [], []
    for b, (st, wk, ins, pl) in enumerate(batches):
        ins = ins.copy(); ins["plan_first_row"] += b * pl.shape[0]
        raw = np.concatenate([x.view(np.uint8).reshape(-1) for x in (st, wk, ins)]).copy()
        keep.append((st, wk, ins, raw)); blocks.append(raw.ctypes.data)
    arr = (C.c_void_p * nb)(*blocks)
    p = L.ismpc_host_pipeline_create(0, n, depth, model.ctypes.data, 35, 10, plans.ctypes.data, plans.shape[0])
    assert p, L.ismpc_host_last_error()
    try:
        steps = 2 * depth + 2
        out = np.zeros((depth, n), dtype=abi.FORMC_OUT)
        csum = C.c_longlong(0)
        assert L.ismpc_host_pipeline_run(p, 0, steps, arr, nb, C.byref(csum), out.ctypes.data) == 0, L.ismpc_host_last_error()
    finally:
        L.ismpc_host_pipeline_destroy(p)
    # two host threads, one pipeline each: thread t takes steps t, t+2, ...
    T = 2
    ps = [L.ismpc_host_pipeline_create(0, n, depth, model.ctypes.data, 35, 10, plans.ctypes.data, plans.shape[0]) for _ in range(T)]
    assert all(ps), L.ismpc_host_last_error()
    try:
        steps_t = 2 * T * depth + 1
        out_t = np.zeros((T, depth, n), dtype=abi.FORMC_OUT)
        assert L.ismpc_host_pipelines_run((C.c_void_p * T)(*ps), T, 0, steps_t, arr, nb, C.byref(csum), out_t.ctypes.data) == 0, \
            L.ismpc_host_last_error()
    finally:
        for q in ps:
            L.ismpc_host_pipeline_destroy(q)
    handle.formc_set_model(model)
    handle.formc_set_plan(plans)
    try:
        for k in range(steps - depth, steps):
            st, wk, ins, _ = keep[k % nb]
            ref = handle.formc_solve_batch(st, wk, ins, None, want_primal=False, want_active=False)
            assert out[k % depth].tobytes() == ref["out"].tobytes(), "step %d" % k
        for t in range(T):
            mine = list(range(t, steps_t, T))
            for j in range(len(mine) - depth, len(mine)):       # the last `depth` steps of thread t sit in slots j % depth
                st, wk, ins, _ = keep[mine[j] % nb]
                ref = handle.formc_solve_batch(st, wk, ins, None, want_primal=False, want_active=False)
                assert out_t[t, j % depth].tobytes() == ref["out"].tobytes(), "thread %d step %d" % (t, mine[j])
    finally:
        handle.formc_set_plan(None)


@pytest.mark.gpu
def test_cpp_serving_loop_packed_equals_direct_calls(handle):
    """The packed serving loop (one fleet per slot: constants and plans resident, one 128-byte record per instance and
    tick, read / written in place by the kernel): the records of the last tick of every slot equal a synchronous
    three-array call on the same values."""
    import ctypes as C
    import torch
    L = _hostlib()
    L.ismpc_host_pool_create.restype = C.c_void_p
    L.ismpc_host_pool_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.ismpc_host_pool_destroy.argtypes = [C.c_void_p]
    L.ismpc_host_pool_set_instances.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.ismpc_host_pool_run_packed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    n, T, D, R = 192, 2, 3, 2
    model = abi.formc_model()
    fleets = [synth.formc_batch(n, seed=90 + f, n_steps=40) for f in range(T * D)]
    plans = np.concatenate([f[3] for f in fleets])
    handle.formc_set_model(model)
    pool = L.ismpc_host_pool_create(0, n, T, D, model.ctypes.data, 35, 10, plans.ctypes.data, plans.shape[0])
    assert pool, L.ismpc_host_last_error()
    keep, ptrs, insts = [], [], []
    try:
        for slot, (st, wk, ins, pl) in enumerate(fleets):
            ins = ins.copy(); ins["plan_first_row"] += slot * pl.shape[0]
            insts.append(ins)
            assert L.ismpc_host_pool_set_instances(pool, slot // D, slot % D, ins.ctypes.data) == 0, L.ismpc_host_last_error()
            for j in range(R):                               # R successive ticks of the fleet's closed loop
                r = handle.formc_rollout(st, wk, ins, plans, j, want_traj=False) if j else dict(state=st, walk=wk)
                tk = abi.pack_ticks(r["state"], r["walk"])
                t = torch.from_numpy(tk.view(np.uint8).reshape(-1).copy()).pin_memory()
                keep.append((t, r["state"], r["walk"])); ptrs.append(t.data_ptr())
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        steps = T * D * 3                                     # every slot serves three ticks: blocks 0, 1, 0
        out = np.zeros((T, D, n), dtype=abi.FORMC_OUT)
        csum = C.c_longlong(0); el = C.c_double(0.0)
        assert L.ismpc_host_pool_run_packed(pool, 0, steps, arr, R, C.byref(csum), out.ctypes.data, C.byref(el)) == 0, \
            L.ismpc_host_last_error()
    finally:
        L.ismpc_host_pool_destroy(pool)
    for slot in range(T * D):
        _, st, wk = keep[slot * R + (2 % R)]
        ref = handle.formc_solve_batch(st, wk, insts[slot], plans, want_primal=False, want_active=False)
        assert out[slot // D, slot % D].tobytes() == ref["out"].tobytes(), "slot %d" % slot


def _build_pipeline_example(d):
    exe = os.path.join(d, "pipeline_example")
    libdir = os.path.dirname(binding.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-O1", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "pipeline_example.cpp"),
                           "-L" + libdir, "-lismpc_b200", "-Wl,-rpath," + libdir])
    return exe


def test_pipeline_example_compiles_with_plain_gpp():
    """host/FormCPipeline.hpp needs nothing but g++ and the C ABI (-Wall -Wextra -Werror); no GPU: loud failure."""
    import torch
    with tempfile.TemporaryDirectory() as d:
        exe = _build_pipeline_example(d)
        if torch.cuda.is_available():
            pytest.skip("GPU present")
        r = subprocess.run([exe, "2"], capture_output=True, text=True)
        assert r.returncode == 2 and "ismpc_create" in r.stderr


@pytest.mark.gpu
def test_pipeline_example_runs():
    """The C++ example steps 64 copies of the DART app's robot for a few ticks: the CoM stays at its height and starts
    moving along the plan like the single-instance closed loop of tests/test_formc_gpu.py."""
    with tempfile.TemporaryDirectory() as d:
        exe = _build_pipeline_example(d)
        out = subprocess.check_output([exe, "5"], text=True)
    rows = np.array([[float(x) for x in ln.split()] for ln in out.strip().splitlines()])
    assert rows.shape == (5, 4) and np.array_equal(rows[:, 0], np.arange(5))
    assert np.abs(rows[:, 3] - 0.69).max() < 1e-3 and np.isfinite(rows).all()


# ---- the drop-in header against the reference's OWN State / WalkState ("Fix the drop-in") -----------------
REF_DIR = "/root/reference/AMR_code_DART"
# AMR_code_DART/types.hpp:7-29 -- member names of `State`, in order (the stand-in is generated from this list where the
# reference tree is absent, e.g. on the GPU box; where it is present the reference's own header is used)
STATE_MEMBERS = ["comPos", "comVel", "comAcc", "zmpPos"] + \
    ["%s%sFoot%s" % (lr, bf, q) for lr, bf in (("left", "Back"), ("right", "Back"), ("left", "Front"), ("right", "Front"))
     for q in ("Pos", "Vel", "Acc")] + \
    ["torsoOrient", "leftBackFootOrient", "rightBackFootOrient", "leftFrontFootOrient", "rightFrontFootOrient"]


def _reference_headers(d):
    """types.hpp and parameters.cpp as the drop-in header will find them: the reference's own files where
    /root/reference exists (copied into the temporary build directory, never into the repository), generated stand-ins
    elsewhere; utils.cpp is always a stub (the real one needs HPIPM / BLASFEO / Eigen::Geometry)."""
    import shutil
    inc = os.path.join(d, "ref")
    os.makedirs(inc)
    have_ref = os.path.exists(os.path.join(REF_DIR, "types.hpp"))
    if have_ref:
        shutil.copy(os.path.join(REF_DIR, "types.hpp"), inc)
        shutil.copy(os.path.join(REF_DIR, "parameters.cpp"), inc)
    else:
        assert len(STATE_MEMBERS) == 21
        body = "".join("    Eigen::Vector3d %s;\n" % m for m in STATE_MEMBERS)
        meth = """
    inline Eigen::VectorXd getComPose() { Eigen::VectorXd p(6); p << torsoOrient, comPos; return p; }
    inline Eigen::VectorXd getSupportFootPose(bool s) { Eigen::VectorXd p(6); if (s == 0) p << leftBackFootOrient, leftBackFootPos; else p << rightBackFootOrient, rightBackFootPos; return p; }
    inline Eigen::VectorXd getFrontSwingFootPose(bool s) { Eigen::VectorXd p(6); if (s == 1) p << rightFrontFootOrient, rightFrontFootPos; else p << leftFrontFootOrient, leftFrontFootPos; return p; }
    inline Eigen::VectorXd getBackSwingFootPose(bool s) { Eigen::VectorXd p(6); if (s == 1) p << leftBackFootOrient, leftBackFootPos; else p << rightBackFootOrient, rightBackFootPos; return p; }
    inline Eigen::VectorXd getRelComPose(bool s) { return vvRel(getComPose(), getSupportFootPose(s)); }
    inline Eigen::VectorXd getRelFrontSwingFootPose(bool s) { return vvRel(getFrontSwingFootPose(s), getSupportFootPose(s)); }
    inline Eigen::VectorXd getRelBackSwingFootPose(bool s) { return vvRel(getBackSwingFootPose(s), getSupportFootPose(s)); }
"""
        open(os.path.join(inc, "types.hpp"), "w").write(
            '#pragma once\n#include <Eigen/Core>\n#include "utils.cpp"\nstruct State {\n' + body + meth + "};\n"
            "struct WalkState { bool supportFoot; double simulationTime; int mpcIter, controlIter, footstepCounter, indInitial; };\n")
        open(os.path.join(inc, "parameters.cpp"), "w").write(
            "#pragma once\n#include <math.h>\n"
            "const double mpcTimeStep = 0.01; const double controlTimeStep = 0.01; const double singleSupportDuration = 0.35;\n"
            "const double doubleSupportDuration = 0.1; const double predictionTime = 1.0; const double comTargetHeight = 0.69;\n"
            "const double footConstraintSquareWidth = 0.09; const double mass_hrp4 = 50.0; const double g = 9.81;\n"
            "const double eta = sqrt(g/comTargetHeight); const int N = round(predictionTime/mpcTimeStep);\n"
            "const int S = round(singleSupportDuration/mpcTimeStep); const int F = round(doubleSupportDuration/mpcTimeStep);\n")
    open(os.path.join(inc, "utils.cpp"), "w").write(
        "#pragma once\n#include <Eigen/Core>\n"
        "inline Eigen::VectorXd vvRel(Eigen::VectorXd v2, Eigen::VectorXd v1) { Eigen::VectorXd r(6);"
        " for (int i = 0; i < 6; ++i) r(i) = v2(i) - v1(i); return r; }\n")
    return inc, have_ref


def _build_dropin_driver(d):
    inc, have_ref = _reference_headers(d)
    exe = os.path.join(d, "dropin_driver")
    libdir = os.path.dirname(binding.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpp", "dropin_driver.cpp"),
                           "-I" + inc, "-I" + os.path.join(ROOT, "quadruped_gait_generation_ismpc_b200", "host", "dropin"),
                           "-I" + os.path.join(ROOT, "tests", "cpp", "eigen_stub"),
                           "-L" + libdir, "-lismpc_b200", "-Wl,-rpath," + libdir])
    return exe, have_ref


def test_dropin_header_compiles_against_the_references_state():
    """Controller.cpp's two call sites (`new MPCSolver(ftsp_and_time_ref)`, `desired = solver->solve(desired, walkState,
    ftsp_and_time_ref)`) compile against host/dropin/MPCSolver.hpp with `State` / `WalkState` from the reference's own
    types.hpp (in this container) or a stand-in with the same 21 members and getRel* methods (elsewhere).  No GPU: the
    constructor fails loudly."""
    import torch
    with tempfile.TemporaryDirectory() as d:
        exe, have_ref = _build_dropin_driver(d)
        if os.path.exists(REF_DIR):
            assert have_ref, "the reference tree is here but its types.hpp was not used"
        if torch.cuda.is_available():
            pytest.skip("GPU present")
        r = subprocess.run([exe, "1"], capture_output=True, text=True)
        assert r.returncode != 0 and "ismpc_create" in r.stderr


@pytest.mark.gpu
def test_dropin_closed_loop_carries_every_member_and_matches_oracle():
    """The drop-in class in the Controller's loop: CoM as the oracle's lock-step closed loop, every member of `State`
    that solve() does not own comes back untouched (MPCSolver.cpp:210,500), and the plan is uploaded again exactly when
    its contents change (once here), not every tick."""
    from oracle import oracle as O
    T = 24
    with tempfile.TemporaryDirectory() as d:
        exe, _ = _build_dropin_driver(d)
        out = subprocess.check_output([exe, str(T)], text=True).strip().splitlines()
    traj = np.array([[float(x) for x in ln.split()] for ln in out[:T]])
    assert traj.shape == (T, 7) and (traj[:, 6] == 0).all()
    assert out[T] == "carried 1" and out[T + 1] == "uploads 2" and out[T + 2] == "pose 6", out[T:]
    model = abi.formc_model()
    state, walk, inst, plan = synth.reference_formc_instance()
    for k in range(T):
        walk["sim_time"] = k
        o = O.formc_batch(model, state, walk, inst, plan)
        assert (o["ret"] == 0).all()
        nxt = np.concatenate([o["out"]["next"]["com_pos"][0], o["out"]["next"]["com_vel"][0]])
        assert np.abs(traj[k, :6] - nxt).max() < 1e-6, "tick %d" % k
        state["com_pos"][0] = traj[k, :3]; state["com_vel"][0] = traj[k, 3:6]
        walk["control_iter"] += 1
        walk["mpc_iter"] = int(np.floor(walk["control_iter"][0] * 0.01 / 0.01))
