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
