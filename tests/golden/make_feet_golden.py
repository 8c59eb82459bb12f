"""Generates tests/golden/matlab_feet_fixtures.npz.  Run HERE (the build container), where /root/reference exists:

    python tests/golden/make_feet_golden.py

The reference's own recorded outputs of the full pipeline (QP-1 closed loop -> second QP -> export):
AMR_code_DART/MATLAB_trajectories/{walking,trotting}/**/foot_{fl,fr,rl,rr}_*.txt and ComTrajectory_*.txt, 2 000 lines
each ("%e", 7 significant digits).  Identified generating configurations (parameter scan):
  walking/phi0_10cm_50, phipi4_10cm_50, phipi2_10cm_50 : quad_walk_no_plots.m  C=100 step=50 ds=30 Qf=1e9 disp_A=0.10
  trotting/phi0, phipi2                                : quad_as_bip_no_plots.m C=160 step=80 ds=50 Qf=1e7 disp_A=0.15
  trotting/phipi4/10cm                                 : same with disp_A=0.10  (first 2 000 lines: the file holds appended runs)
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/AMR_code_DART/MATLAB_trajectories/"
CASES = (("walk_phi0", "walking/phi0_10cm_50/", "walk_phi0"), ("walk_phipi4", "walking/phipi4_10cm_50/", "walk_phipi4"),
         ("walk_phipi2", "walking/phipi2_10cm_50/", "walk_phipi2"), ("trot_phi0", "trotting/phi0/", "trot_phi0"),
         ("trot_phipi2", "trotting/phipi2/", "trot_phipi2"), ("trot_phipi4_10cm", "trotting/phipi4/10cm/", "trot_phipi4"))


def main():
    fx = {}
    for key, d, tag in CASES:
        for name, fn in (("com", "ComTrajectory_%s.txt"), ("fl", "foot_fl_%s.txt"), ("fr", "foot_fr_%s.txt"),
                         ("rl", "foot_rl_%s.txt"), ("rr", "foot_rr_%s.txt")):
            p = REF + d + fn % tag
            if os.path.exists(p):
                fx["%s_%s" % (key, name)] = np.loadtxt(p)[:2000]
    np.savez_compressed(os.path.join(HERE, "matlab_feet_fixtures.npz"), **fx)
    print("written", sorted(fx))


if __name__ == "__main__":
    main()
