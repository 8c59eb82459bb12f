"""Generates the committed golden fixtures.  Run HERE (the build container), where /root/reference exists:

    python tests/golden/make_golden.py

1. matlab_fixtures.npz  -- excerpts of the reference's own recorded outputs (the only golden vectors the
   reference holds for this path): closed-loop CoM position/velocity written by the
   MATLAB ISMPC scripts to AMR_code_DART/MATLAB_trajectories/**.txt ("%e", 7 significant digits).
   Identified configurations (by parameter scan against the first ticks):
     walking/phi0_10cm_50, phipi4_10cm_50 : quad_walk_no_plots.m  C=100 P=200 step=50 ds=30 Qf=1e9 disp_A=0.10
     trotting/phi0                        : quad_as_bip_no_plots.m C=160 P=320 step=80 ds=50 Qf=1e7 disp_A=0.15
2. oracle_formc.npz / oracle_forma.npz -- outputs of the reference's qpOASES (oracle/_ref, solveQP call form)
   on seeded synthetic instances, so the oracle port and the GPU can be checked where oracle/_ref is absent.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O                                   # noqa: E402
from quadruped_gait_generation_ismpc_b200 import abi, synth     # noqa: E402

REF = "/root/reference/AMR_code_DART/MATLAB_trajectories/"


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    fx = {}
    for key, d, tag, rows in (("walk_phi0", "walking/phi0_10cm_50/", "walk_phi0", 201),
                              ("walk_phipi4", "walking/phipi4_10cm_50/", "walk_phipi4", 201),
                              ("trot_phi0", "trotting/phi0/", "trot_phi0", 101)):
        fx[key + "_com"] = np.loadtxt(REF + d + "ComTrajectory_%s.txt" % tag)[:rows]
        p = REF + d + "ComVelocity_%s.txt" % tag
        if os.path.exists(p):
            fx[key + "_vel"] = np.loadtxt(p)[:rows]
    np.savez_compressed(os.path.join(HERE, "matlab_fixtures.npz"), **fx)

    # formulation C goldens: 48 seeded instances incl. active vertical rows and not-running instances
    model = abi.formc_model()
    parts = [synth.formc_batch(16, seed=1234), synth.formc_batch(16, seed=1235, z_spread=0.08),
             synth.formc_batch(16, seed=1236, vary_height=True, running_frac=0.5)]
    out = {}
    for k, (state, walk, inst, plan) in enumerate(parts):
        r = O.formc_batch(model, state, walk, inst, plan, solver=O.SOLVER_QPOASES, kind="ref")
        out.update({"state%d" % k: state, "walk%d" % k: walk, "inst%d" % k: inst, "plan%d" % k: plan,
                    "out%d" % k: r["out"], "primal%d" % k: r["primal"], "active%d" % k: r["active"].astype(np.int8),
                    "duals%d" % k: r["duals"], "ret%d" % k: r["ret"], "nwsr%d" % k: r["nwsr"]})
    np.savez_compressed(os.path.join(HERE, "oracle_formc.npz"), model=model, **out)

    # formulation A goldens: the probe-style closed loop (trot, bang.m parameters) sampled every 10th tick
    amodel = abi.forma_model()
    inst, ft, plan = synth.forma_batch(1, seed=99, gait="trot")
    p = O.FormAParams()
    p.dt, p.eta, p.wx, p.wy = 0.01, float(np.sqrt(9.8 / 0.56)), 0.02, 0.02
    p.disp_forw, p.disp_forw_dummy, p.disp_L, p.Qzdot, p.Qfoot, p.C, p.P, p.F = 0.5, 0.25, 0.4, 1.0, 1e7, 100, 200, 3
    cur, pl = inst.copy(), plan.copy()
    snaps_inst, snaps_plan, prim, act, dual, outs, rets = [], [], [], [], [], [], []
    for t in range(160):
        o = O.forma_batch(amodel, cur, ft, pl, solver=O.SOLVER_QPOASES, kind="ref")
        if t % 10 == 0 or t in (48, 49, 50, 51, 99, 100):
            snaps_inst.append(cur.copy()); snaps_plan.append(pl.copy()); prim.append(o["primal"][0])
            act.append(o["active"][0].astype(np.int8)); dual.append(o["duals"][0]); outs.append(o["out"].copy())
            rets.append(o["ret"][0])
        cur["st"][0] = o["out"]["st"][0]
        fc = cur["fs_counter"][0]
        if cur["j"][0] + 1 >= ft[fc]:
            fc += 1; cur["fs_counter"] = fc
            pred = np.array([o["out"]["pred_fs"][0][0], o["out"]["pred_fs"][0][3]])
            cur["cur_fs"][0] = pred; cur["fs_store"][0] = pred
            pl = pl + (pred - pl[fc - 1]); cur["cl_first_ramp"] = 0
        cur["j"] += 1
    np.savez_compressed(os.path.join(HERE, "oracle_forma.npz"), model=amodel, fs_timing=ft,
                        inst=np.concatenate(snaps_inst), plans=np.stack(snaps_plan), primal=np.stack(prim),
                        active=np.stack(act), duals=np.stack(dual), out=np.concatenate(outs), ret=np.array(rets))
    print("golden fixtures written")


if __name__ == "__main__":
    main()
