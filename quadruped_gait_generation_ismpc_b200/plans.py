"""Footstep-plan generators: host-side mirrors of the reference's MATLAB initialisation scripts.

  trot_plan  <- trotting/init_quadruped.m:4-184   (two-beat trot, quadruped as a virtual biped)
  walk_plan  <- walking/init_quadruped2.m:4-300   (four-beat walk, 8-phase cycle)

Both return (foot_plan, center): foot_plan is N_gait x 8 = [back_left, back_right, front_right,
front_left] (x, y each); center is N_gait x 2, the virtual-biped footstep = intersection of the two
diagonals of the support polygon (the scripts' symbolic solve() replaced by the closed-form 2x2 line
intersection).  `center` is what the ISMPC scripts use as fs_plan (quad_as_bip_bang.m:7).

controller_plan <- AMR_code_DART/Controller.cpp:89-97 (the C++ app's ftsp_and_time matrix).
"""
import math

import numpy as np


def _clip_step(disp_A, phi, disp_forw, disp_vertical):
    """init_quadruped.m:54-102 -- feasibility clipping of the nominal and first ('dummy') step."""
    x_p, y_p = disp_A * math.cos(phi), disp_A * math.sin(phi)
    x_d, y_d = x_p / 2, y_p / 2
    fd, vd = disp_forw / 2, disp_vertical / 2

    def clip(x, y, forw, vert):
        if y > vert or x > forw:
            if phi > math.atan(vert / forw):
                return vert * math.cos(phi) / math.sin(phi), vert
            return forw, forw * math.sin(phi) / math.cos(phi)
        return x, y

    x_d, y_d = clip(x_d, y_d, fd, vd)
    x_p, y_p = clip(x_p, y_p, disp_forw, disp_vertical)
    return x_p, y_p, x_d, y_d


def _diag_intersection(foot_row):
    """Intersection of line BL-FR with line BR-FL (init_quadruped.m:176-183).  polyfit(...,1) of two
    points is the line through them; vertical diagonals do not occur for a body of length disp_C > 0."""
    x1, y1, x2, y2, x3, y3, x4, y4 = foot_row  # BL, BR, FR, FL
    m1 = (y3 - y1) / (x3 - x1); q1 = y1 - m1 * x1
    m2 = (y4 - y2) / (x4 - x2); q2 = y2 - m2 * x2
    x = (q2 - q1) / (m1 - m2)
    return x, m1 * x + q1


def trot_plan(N_gait=100, disp_A=0.1, phi=0.0, disp_B=0.259394, disp_C=0.88,
              disp_forw=0.5, disp_i=0.4, disp_o=0.4):
    x_p, y_p, x_d, y_d = _clip_step(disp_A, phi, disp_forw, min(disp_i, disp_o))
    bl = np.tile([0.0, disp_B], (N_gait, 1)); br = np.tile([0.0, -disp_B], (N_gait, 1))
    fl = np.tile([disp_C, disp_B], (N_gait, 1)); fr = np.tile([disp_C, -disp_B], (N_gait, 1))
    # init_quadruped.m:113-117 (first half step), rows are 1-based in MATLAB
    bl[1, 0] = x_d; fr[1, 0] = disp_C + x_d
    bl[1, 1] = disp_B + y_d; fr[1, 1] = -disp_B + y_d
    for j in range(3, N_gait + 1):  # init_quadruped.m:120-149
        i = j - 1
        if j % 2 == 0:
            bl[i] = bl[i - 1] + [x_p, y_p]; fr[i] = fr[i - 1] + [x_p, y_p]
            br[i] = br[i - 1]; fl[i] = fl[i - 1]
        else:
            br[i] = br[i - 1] + [x_p, y_p]; fl[i] = fl[i - 1] + [x_p, y_p]
            bl[i] = bl[i - 1]; fr[i] = fr[i - 1]
    foot_plan = np.hstack([bl, br, fr, fl])
    center = np.zeros((N_gait, 2))
    center[0, 0] = disp_C / 2
    for k in range(1, N_gait):
        center[k] = _diag_intersection(foot_plan[k])
    return foot_plan, center


def walk_plan(N_gait=100, disp_A=0.1, phi=0.0, disp_B=0.259394, disp_C=0.88,
              disp_forw=0.5, disp_i=0.4, disp_o=0.4):
    x_p, y_p, x_d, y_d = _clip_step(disp_A, phi, disp_forw, min(disp_i, disp_o))
    rows = N_gait + 8  # the 8-phase loop writes up to row j+7 (MATLAB grows the arrays)
    bl = np.tile([0.0, disp_B], (rows, 1)); br = np.tile([0.0, -disp_B], (rows, 1))
    fl = np.tile([disp_C, disp_B], (rows, 1)); fr = np.tile([disp_C, -disp_B], (rows, 1))
    grown = N_gait
    # init_quadruped2.m:114-138 (dummy first half-cycle), 1-based rows 3..5
    fl[2, 0] = disp_C + x_d; fl[3, 0] = fl[2, 0]; fl[4, 0] = fl[2, 0]
    br[1, 0] = br[0, 0]; br[2, 0] = br[0, 0]; br[3, 0] = br[2, 0]; br[4, 0] = br[3, 0] + x_d
    fl[2, 1] = disp_B + y_d; fl[3, 1] = fl[2, 1]; fl[4, 1] = fl[2, 1]
    br[1, 1] = br[0, 1]; br[2, 1] = br[0, 1]; br[3, 1] = br[2, 1]; br[4, 1] = br[3, 1] + y_d
    for j in range(6, N_gait + 1, 8):  # init_quadruped2.m:141-219
        i = j - 1
        step = np.array([x_p, y_p])
        fr[i] = fr[i - 1]; fr[i + 1] = fr[i] + step
        for k in range(2, 8): fr[i + k] = fr[i + 1]
        bl[i] = bl[i - 1]; bl[i + 1] = bl[i]; bl[i + 2] = bl[i]; bl[i + 3] = bl[i + 2] + step
        for k in range(4, 8): bl[i + k] = bl[i + 3]
        fl[i] = fl[i - 1]
        for k in range(1, 5): fl[i + k] = fl[i]
        fl[i + 5] = fl[i + 4] + step; fl[i + 6] = fl[i + 5]; fl[i + 7] = fl[i + 5]
        br[i] = br[i - 1]
        for k in range(1, 7): br[i + k] = br[i]
        br[i + 7] = br[i + 6] + step
        grown = max(grown, j + 7)
    foot_plan = np.hstack([bl, br, fr, fl])[:grown]
    n = foot_plan.shape[0]
    center = np.zeros((max(n, N_gait), 2))
    center[0, 0] = disp_C / 2
    for j in range(1, N_gait - 4 + 1, 8):  # init_quadruped2.m:242-284
        for k in range(0, 8, 2):
            if j + k - 1 < n:
                center[j + k - 1] = _diag_intersection(foot_plan[j + k - 1])
        for k in (1, 3, 5, 7):
            if j + k - 1 < center.shape[0]:
                center[j + k - 1] = center[j + k - 2]
    return foot_plan, center[:max(n, N_gait)]


def controller_plan(n_steps=40, step_x=0.2, half_y=0.08, S=35, F=10):
    """AMR_code_DART/Controller.cpp:89-97: row 0 stays zero (the loop starts at 1); for i >= 1
    x=(i-1)*0.2, y=(-1)^(i-1)*0.08, z=0, t=(S+F)*i."""
    p = np.zeros((n_steps, 4))
    for i in range(1, n_steps):
        p[i, 0] = (i - 1) * step_x
        p[i, 1] = half_y * (1.0 if (i - 1) % 2 == 0 else -1.0)
        p[i, 3] = (S + F) * i
    return p
