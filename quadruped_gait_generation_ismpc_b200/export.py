"""Trajectory files in the reference's wire format.

The MATLAB scripts write CoM position / velocity and the four foot trajectories with
`fprintf(file, '%d %d %d\\n', [x, y, z])` (trotting/quad_as_bip_no_plots.m:438-439,482-509;
walking/quad_walk_no_plots.m:507-513,563-613): MATLAB prints a double that is not integer-valued with `%e`
(7 significant digits) and an integer-valued one as an integer.  AMR_code_DART/Controller.cpp:147-281 reads
those files back (three numbers per line).  host/TrajectoryWriter.hpp is the C++ twin of this module.
"""
import numpy as np

FILES = {"com": "ComTrajectory_%s.txt", "vel": "ComVelocity_%s.txt", "fl": "foot_fl_%s.txt", "fr": "foot_fr_%s.txt",
         "rl": "foot_rl_%s.txt", "rr": "foot_rr_%s.txt"}


def format_value(v):
    """MATLAB's fprintf('%d', v) for a double."""
    v = float(v)
    if np.isfinite(v) and v == np.floor(v) and abs(v) < 2 ** 53:
        return "%d" % int(v)
    return "%e" % v


def format_rows(rows):
    return "".join("%s %s %s\n" % tuple(format_value(v) for v in r) for r in np.asarray(rows, dtype=np.float64))


def com_rows(state0, traj, height):
    """Rows of ComTrajectory_*.txt / ComVelocity_*.txt for one instance: the scripts print the state at the START of
    each tick (x_store(j), y_store(j), height) and the velocity after it (xd_store(j), yd_store(j), 0).
    state0 = (x, xd, xz, y, yd, yz) before the first tick; traj = (n_ticks x 6: x, y, xd, yd, xz, yz) after each tick."""
    traj = np.asarray(traj, dtype=np.float64)
    T = traj.shape[0]
    pos = np.empty((T, 3)); vel = np.empty((T, 3))
    pos[0, 0], pos[0, 1] = state0[0], state0[3]
    pos[1:, :2] = traj[:T - 1, :2]
    pos[:, 2] = height
    vel[:, :2] = traj[:, 2:4]
    vel[:, 2] = 0.0
    return pos, vel


def write_all(directory, tag, pos, vel, feet):
    """Writes the six files of one run; feet = dict(fl, fr, rl, rr) of (samples x 3) arrays."""
    import os
    os.makedirs(directory, exist_ok=True)
    data = dict(com=pos, vel=vel, **feet)
    for key, rows in data.items():
        with open(os.path.join(directory, FILES[key] % tag), "w") as f:
            f.write(format_rows(rows))


def read_rows(path):
    """What Controller.cpp:147-281 does: three numbers per line."""
    return np.array([[float(v) for v in ln.split()] for ln in open(path) if ln.strip()])
