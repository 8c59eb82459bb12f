"""Seeded synthetic ISMPC instances: the same arrays feed the CPU oracle and the GPU.

formc_batch  -- config 2: independent trot instances for MPCSolver::solve (formulation C)
forma_batch  -- configs 2/3: trot / walk instances for the canonical ISMPC (formulation A) at the
                start-of-gait state (tick j = 1); mid-gait states are produced by rolling these forward
"""
import numpy as np

from . import abi, plans

SEED0 = 0x15A9C0DE


def reference_formc_instance(n_ticks_time=0.0):
    """Config 1: the single instance of the DART app (parameters.cpp, Controller.cpp:89-97,112),
    with footstepCounter = 2 so that the ZMP box and the flight-phase rows are live."""
    state = np.zeros(1, dtype=abi.STATE)
    state["com_pos"][0] = [0.0, 0.0, 0.69]
    walk = np.zeros(1, dtype=abi.WALK)
    walk["sim_time"] = n_ticks_time
    walk["footstep_counter"] = 2
    inst = np.zeros(1, dtype=abi.FORMC_INST)
    inst["com_height"], inst["box_w"], inst["box_w_init"] = 0.69, 0.09, 2.0
    inst["S"], inst["F_ds"], inst["plan_first_row"], inst["n_steps"] = 35, 10, 0, 40
    return state, walk, inst, plans.controller_plan()


def formc_batch(n, seed=SEED0 ^ 2, N=100, n_steps=40, S=35, F_ds=10, vary_height=False, z_spread=0.01,
                running_frac=1.0, dcm_spread=0.035, k0_cap=800):
    """n randomised trot instances: step length L~U[0.05,0.25], half-width W~U[0.05,0.12] alternating,
    heading in {0, pi/4, pi/2}, per-step jitter N(0, 0.01^2), k0~U{0..k0max}, velocity U[-0.1,0.1]^2, position
    such that the divergent component is within dcm_spread of the value the footstep plan can stabilise
    (a CoM parked over one foot is infeasible for the 9 cm ZMP box), footstepCounter >= 2."""
    rng = np.random.default_rng(seed)
    per = S + F_ds
    state = np.zeros(n, dtype=abi.STATE)
    walk = np.zeros(n, dtype=abi.WALK)
    inst = np.zeros(n, dtype=abi.FORMC_INST)
    plan = np.zeros((n * n_steps, 4))
    k0max = n_steps * per - 2 * N - per
    assert k0max > 0, "plan too short for the horizon"
    for i in range(n):
        L = rng.uniform(0.05, 0.25); W = rng.uniform(0.05, 0.12)
        phi = rng.choice([0.0, np.pi / 4, np.pi / 2])
        c, s = np.cos(phi), np.sin(phi)
        p = np.zeros((n_steps, 4))
        for k in range(n_steps):
            fx, fy = k * L, W * (1 if k % 2 == 0 else -1)
            p[k, 0] = c * fx - s * fy + rng.normal(0, 0.01)
            p[k, 1] = s * fx + c * fy + rng.normal(0, 0.01)
            p[k, 3] = per * k
        plan[i * n_steps:(i + 1) * n_steps] = p
        k0 = int(rng.integers(0, min(k0_cap, k0max) + 1))
        step = k0 // per; r = k0 % per
        h = rng.uniform(0.45, 0.75) if vary_height else 0.69
        # A state the ISMPC can stabilise: the divergent component xi = c + cd/eta must match the discounted
        # future ZMP (stability row, MPCSolver.cpp:375-384, evaluated for u = box centres at nominal eta);
        # it is then perturbed by U[-dcm_spread, dcm_spread] (box capacity is ~0.044 for w = 0.09).
        eta = np.sqrt(9.81 / h); dt = 0.01
        t = k0 + np.arange(2 * N)
        si = t // per; ri = t % per
        a_ = p[np.minimum(si, n_steps - 1), :2]; b_ = p[np.minimum(si + 1, n_steps - 1), :2]
        ramp = np.where(ri < S, 0.0, (ri - S) / F_ds)[:, None]
        midw = a_ + (b_ - a_) * ramp
        midw[si >= n_steps - 1] = 0.0
        e1 = np.exp(eta * dt)
        avec = -np.exp(eta * dt * (N - 1 - np.arange(N))) * (e1 - 1.0)
        rhs = eta * dt * (np.exp(-eta * dt * np.arange(N)) @ midw[N:]) - avec @ midw[:N]
        xi0 = rhs / np.exp(eta * dt * N)
        vel = rng.uniform(-0.1, 0.1, 2)
        xi = xi0 + rng.uniform(-dcm_spread, dcm_spread, 2)
        pos = xi - vel / eta
        state["com_pos"][i] = [pos[0], pos[1], h + rng.uniform(-z_spread, z_spread)]
        state["com_vel"][i] = [vel[0], vel[1], rng.uniform(-0.05, 0.05)]
        walk["sim_time"][i] = k0
        walk["mpc_iter"][i] = r
        walk["control_iter"][i] = r
        walk["footstep_counter"][i] = (2 + step) if rng.uniform() < running_frac else int(rng.integers(0, 2))
        walk["support_foot"][i] = step % 2
        inst["com_height"][i] = h
        inst["box_w"][i] = 0.09; inst["box_w_init"][i] = 2.0
        inst["S"][i] = S; inst["F_ds"][i] = F_ds
        inst["plan_first_row"][i] = i * n_steps; inst["n_steps"][i] = n_steps
    return state, walk, inst, plan


def forma_batch(n, seed=SEED0 ^ 3, gait="trot", C=100, step=50, ds=20, sim_ticks=2000, vary=False,
                N_gait=100):
    """n formulation-A instances at the start-of-gait state (x = xz = disp_C/2 at rest on the first footstep,
    quad_as_bip_bang.m:44-52), each with its own plan: disp_A~U[0.05,0.15], heading in {0, pi/4, pi/2}.
    vary=True (config 3): h~U[0.45,0.75], step duration in {40,50,60}, ds in {20,30}.
    Returns inst, fs_timing (one table per distinct step duration, concatenated), fs_plan (n*N_gait x 2)."""
    rng = np.random.default_rng(seed)
    inst = np.zeros(n, dtype=abi.FORMA_INST)
    fs_plan = np.zeros((n * N_gait, 2))
    steps = [40, 50, 60] if vary else [step]
    tables, first = [], {}
    off = 0
    for s in steps:
        t = np.arange(0, sim_ticks + 12 * s + 1, s, dtype=np.int32)
        first[s] = (off, len(t)); off += len(t); tables.append(t)
    fs_timing = np.concatenate(tables)
    gen = plans.trot_plan if gait == "trot" else plans.walk_plan
    for i in range(n):
        disp_A = rng.uniform(0.05, 0.15)
        phi = float(rng.choice([0.0, np.pi / 4, np.pi / 2]))
        _, center = gen(N_gait=N_gait, disp_A=disp_A, phi=phi)
        fs_plan[i * N_gait:(i + 1) * N_gait] = center[:N_gait]
        s = int(rng.choice(steps))
        inst["st"][i] = [center[0, 0], 0.0, center[0, 0], center[0, 1], 0.0, center[0, 1]]
        inst["cur_fs"][i] = center[0]; inst["fs_store"][i] = center[0]
        inst["height"][i] = rng.uniform(0.45, 0.75) if vary else 0.56
        inst["wx"][i] = 0.02; inst["wy"][i] = 0.02
        inst["j"][i] = 1; inst["fs_counter"][i] = 1
        inst["ds"][i] = int(rng.choice([20, 30])) if vary else ds
        inst["cl_first_ramp"][i] = 1
        inst["timing_first"][i], inst["n_timing"][i] = first[s]
        inst["plan_first_row"][i] = i * N_gait; inst["n_fs"][i] = N_gait
    return inst, fs_timing, fs_plan


def push_batch(n, seed=SEED0 ^ 5, formc=False):
    """Config 5 pushes: velocity impulse for 14 consecutive ticks, a~U[0.3,1.0] m/s^2, direction uniform
    (trotting/quad_as_bip_no_plots.m:121-131).  formc: ticks counted from rollout start, start~U{100..800};
    form A: during footstep fs~U{2..8}, ct in [1,15)."""
    rng = np.random.default_rng(seed)
    push = np.zeros(n, dtype=abi.PUSH)
    a = rng.uniform(0.3, 1.0, n); th = rng.uniform(0, 2 * np.pi, n)
    push["ax"], push["ay"] = a * np.cos(th), a * np.sin(th)
    if formc:
        t0 = rng.integers(100, 801, n)
        push["ct0"], push["ct1"] = t0, t0 + 14
    else:
        push["fs"] = rng.integers(2, 9, n); push["ct0"] = 1; push["ct1"] = 15
    return push


def dense_qp_batch(shape, n, seed=SEED0 ^ 21):
    """DENSE inputs for the solveQP seam (ismpc_qp_solve_batch), in the two shapes the path produces, feasible by
    construction (bounds are laid around A v0 for a reference point v0; equality rows hold at v0).
    "formc_horizontal": nV = 100, nC = 101 -- stage 3 of MPCSolver::solve stacked as the reference would pass it to
        solveQP: H = I, g = -mid, A = [a'; I], a_i = -e^{eta dt (N-1-i)}(e^{eta dt} - 1) (nominal eta),
        box mid +- 0.045.
    "forma_stacked": nV = 206, nC = 208 -- the canonical ISMPC QP (quad_as_bip_bang.m:121-257): variables
        [zd_x(C); x_f(F); zd_y(C); y_f(F)], H = diag(1, Qf), rows = [stab_x; stab_y; ZMP_x(C); ZMP_y(C); kin_x(F); kin_y(F)],
        ZMP rows dt*tril(1) - mapping, mapping with a linear blend over ds samples before each footstep switch."""
    rng = np.random.default_rng(seed)
    if shape == "formc_horizontal":
        N, dt, eta = 100, 0.01, np.sqrt(9.81 / 0.69)
        a = -np.exp(eta * dt * (N - 1 - np.arange(N))) * (np.exp(eta * dt) - 1.0)
        H = np.broadcast_to(np.eye(N), (n, N, N)).copy()
        A = np.zeros((n, N + 1, N)); A[:, 0, :] = a; A[:, 1:, :] = np.eye(N)
        L = rng.uniform(0.05, 0.25, (n, 1)); k0 = rng.integers(0, 45, (n, 1))
        t = k0 + np.arange(N)[None, :]
        si, ri = t // 45, t % 45
        mid = L * (si + np.where(ri < 35, 0.0, (ri - 35) / 10.0))
        # the reference point sits on one side of its box over the first samples (where |a| is largest), so that the
        # minimum-norm correction nu*a saturates several box rows: working sets of 1 + a handful of rows, as on the path
        sgn = rng.choice([-1.0, 1.0], (n, 1)); width = rng.integers(18, 40, (n, 1))
        u0 = mid + 0.045 * sgn * (np.arange(N)[None, :] < width) * rng.uniform(0.8, 1.0, (n, N))
        lb = np.concatenate([(u0 @ a)[:, None], mid - 0.045], axis=1)
        ub = np.concatenate([(u0 @ a)[:, None], mid + 0.045], axis=1)
        return H, -mid, A, lb, ub
    C, F, dt, step, ds, Qf = 100, 3, 0.01, 50, 20, 1e7
    eta = np.sqrt(9.8 / 0.56)
    nv1 = C + F; nV = 2 * nv1; nC = nV + 2
    lam = np.exp(-eta * dt)
    stab = (1.0 / eta) * (1.0 - lam) / (1.0 - lam ** C) * np.exp(-eta * dt * np.arange(C)) - dt * np.exp(-eta * dt * C)
    P = dt * np.tril(np.ones((C, C)))
    H = np.zeros((n, nV, nV)); g = np.zeros((n, nV)); A = np.zeros((n, nC, nV)); lb = np.zeros((n, nC)); ub = np.zeros((n, nC))
    hd = np.concatenate([np.ones(C), Qf * np.ones(F), np.ones(C), Qf * np.ones(F)])
    D = np.eye(F) - np.eye(F, k=-1)
    for p in range(n):
        j0 = int(rng.integers(0, step))
        T = np.arange(1, F + 2) * step - j0                       # ticks until the next F+1 footstep switches
        mp = np.zeros((C, F + 1))
        for i in range(1, C + 1):
            pf = int((T <= i).sum())
            pf = min(pf, F)
            rem = (T[pf] - i) if pf <= F else ds + 1
            if rem > ds or pf == F:
                mp[i - 1, pf] = 1.0
            else:
                mp[i - 1, pf] = rem / ds; mp[i - 1, pf + 1] = 1.0 - rem / ds
        H[p][np.arange(nV), np.arange(nV)] = hd
        for ax in range(2):
            o = ax * nv1
            A[p, ax, o:o + C] = stab
            A[p, 2 + ax * C: 2 + (ax + 1) * C, o:o + C] = P
            A[p, 2 + ax * C: 2 + (ax + 1) * C, o + C:o + nv1] = -mp[:, 1:]
            A[p, 2 + 2 * C + ax * F: 2 + 2 * C + (ax + 1) * F, o + C:o + nv1] = D
        stepl = rng.uniform(0.05, 0.15) * np.array([np.cos(0.3), np.sin(0.3)])
        plan = np.arange(1, F + 1)[:, None] * stepl[None, :]
        v0 = np.concatenate([rng.normal(0, 0.2, C), plan[:, 0] + rng.normal(0, 0.01, F),
                             rng.normal(0, 0.2, C), plan[:, 1] + rng.normal(0, 0.01, F)])
        r0 = A[p] @ v0
        lb[p] = r0 - rng.uniform(0.0, 0.02, nC); ub[p] = r0 + rng.uniform(0.0, 0.02, nC)
        lb[p, :2] = ub[p, :2] = r0[:2]
        g[p] = np.concatenate([np.zeros(C), -Qf * plan[:, 0], np.zeros(C), -Qf * plan[:, 1]])
    return H, g, A, lb, ub
