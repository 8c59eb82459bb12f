"""BASELINE.json's full sizes on one GPU, through size-independent properties (the oracle would need minutes to hours
for them): 65,536 instances per call (configs[2]), 1,000 x 1,000 closed-loop ticks with pushes (configs[4]).

Properties: bit-for-bit determinism; permutation equivariance (no instance sees another); the contiguous 8-way shards of
SURVEY 8(e) solved one by one == the whole batch (what `bench.py --gpus 8` relies on); the solver's own KKT residual;
and oracle parity (primal 1e-6, identical active set) on a random sample of the same batch."""
import numpy as np
import pytest

from quadruped_gait_generation_ismpc_b200 import abi, sharding, synth
from oracle import oracle as O
from parity import PRIMAL_TOL, primal_rel_err, active_set_mismatch

pytestmark = pytest.mark.gpu

N_FULL = 65536
N_DISTINCT = 8192


def _tile_formc(seed):
    """65,536 form-C instances: 8,192 distinct randomised trot plans (configs[1] distribution), each shared by eight
    instances whose states differ (plans are read-only and may be shared, ismpc_b200.h)."""
    state, walk, inst, plan = synth.formc_batch(N_DISTINCT, seed=seed)
    rep = N_FULL // N_DISTINCT
    state, walk, inst = np.tile(state, rep), np.tile(walk, rep), np.tile(inst, rep)
    rng = np.random.default_rng(seed + 1)
    state["com_pos"][:, :2] += rng.uniform(-0.004, 0.004, (N_FULL, 2))
    state["com_vel"][:, :2] += rng.uniform(-0.02, 0.02, (N_FULL, 2))
    return state, walk, inst, plan


def test_formc_65536_instances(handle):
    model = abi.formc_model()
    handle.formc_set_model(model)
    state, walk, inst, plan = _tile_formc(41)
    g = handle.formc_solve_batch(state, walk, inst, plan)
    st = g["out"]["status"]
    failed = (st & (abi.ST_Z_FAIL | abi.ST_X_FAIL | abi.ST_Y_FAIL)) != 0
    assert failed.mean() < 0.01, "%d instances failed" % failed.sum()
    assert g["out"]["kkt_res"][~failed].max() < 1e-8
    assert np.isfinite(g["primal"][~failed]).all()
    # determinism
    g2 = handle.formc_solve_batch(state, walk, inst, plan)
    assert g["out"].tobytes() == g2["out"].tobytes()
    assert np.array_equal(g["primal"], g2["primal"]) and np.array_equal(g["active"], g2["active"])
    # permutation equivariance
    perm = np.random.default_rng(5).permutation(N_FULL)
    gp = handle.formc_solve_batch(state[perm], walk[perm], inst[perm], plan)
    assert gp["out"].tobytes() == g["out"][perm].tobytes()
    assert np.array_equal(gp["primal"], g["primal"][perm]) and np.array_equal(gp["active"], g["active"][perm])
    # the 8 contiguous shards of SURVEY 8(e), one call each
    for r in range(8):
        a, b = sharding.shard_range(N_FULL, r, 8)
        gs = handle.formc_solve_batch(state[a:b], walk[a:b], inst[a:b], plan, want_active=False)
        assert gs["out"].tobytes() == g["out"][a:b].tobytes(), "shard %d" % r
        assert np.array_equal(gs["primal"], g["primal"][a:b])
    # oracle parity on a sample of the same batch
    sel = np.sort(np.random.default_rng(6).choice(N_FULL, 256, replace=False))
    o = O.formc_batch(model, state[sel], walk[sel], inst[sel], plan, nthreads=8)
    ok = (o["ret"] == 0).all(axis=1) & (o["out"]["status"] & abi.ST_WINDOW == 0)
    assert ok.mean() > 0.9
    assert not failed[sel][ok].any()
    err = primal_rel_err(g["primal"][sel][ok].reshape(-1, 3, 100), o["primal"][ok].reshape(-1, 3, 100))
    assert err.max() <= PRIMAL_TOL, "primal rel err %.3e" % err.max()
    mism, _ = active_set_mismatch(g["active"][sel][ok], o["active"][ok], o["duals"][ok])
    assert mism.sum() == 0


def test_formc_packed_host_arrays_take_the_single_copy_path(handle):
    """state | walk | inst back to back in one allocation (one host->device copy inside the library) == three
    separate arrays."""
    handle.formc_set_model(abi.formc_model())
    state, walk, inst, plan = synth.formc_batch(1024, seed=43)
    a = handle.formc_solve_batch(state, walk, inst, plan)
    n = len(state)
    raw = np.concatenate([x.view(np.uint8).reshape(-1) for x in (state, walk, inst)]).copy()
    o1 = state.nbytes; o2 = o1 + walk.nbytes
    ps = raw[:o1].view(abi.STATE); pw = raw[o1:o2].view(abi.WALK); pi = raw[o2:].view(abi.FORMC_INST)
    assert pw.ctypes.data == ps.ctypes.data + n * abi.STATE.itemsize
    out = np.zeros(n, dtype=abi.FORMC_OUT); primal = np.zeros((n, 300)); active = np.zeros((n, 300), dtype=np.int8)
    handle.formc_solve_batch_raw(n, ps, pw, pi, plan, plan.shape[0], out, primal, active, mem=abi.MEM_HOST)
    assert out.tobytes() == a["out"].tobytes()
    assert np.array_equal(primal, a["primal"]) and np.array_equal(active, a["active"])


def test_formc_closed_loop_1000x1000_with_pushes(handle):
    """configs[4]: 1,000 instances x 1,000 ticks (10 s) with push disturbances.  Deterministic, equivariant, bounded
    where no tick failed, and ticks sampled along the run agree with the oracle started from the GPU's own previous
    state (errors must not compound into the check)."""
    model = abi.formc_model()
    handle.formc_set_model(model)
    n, T = 1000, 1000
    state, walk, inst, plan = synth.formc_batch(n, seed=44, k0_cap=300)
    push = synth.push_batch(n, seed=45, formc=True)
    r = handle.formc_rollout(state, walk, inst, plan, T, push=push)
    r2 = handle.formc_rollout(state, walk, inst, plan, T, push=push)
    assert np.array_equal(r["traj"], r2["traj"]) and np.array_equal(r["status"], r2["status"])
    perm = np.random.default_rng(7).permutation(n)
    rp = handle.formc_rollout(state[perm], walk[perm], inst[perm], plan, T, push=push[perm])
    assert np.array_equal(rp["traj"], r["traj"][perm])
    ok = (r["status"] & (abi.ST_Z_FAIL | abi.ST_X_FAIL | abi.ST_Y_FAIL)) == 0
    assert ok.mean() > 0.9, "%d instances had a failed tick" % (~ok).sum()
    assert np.isfinite(r["traj"][ok]).all()
    # the CoM stays within reach of its footstep plan for the whole run (the stability constraint at work)
    lo = plan.reshape(n, -1, 4)[:, :, :2].min(axis=1) - 0.5
    hi = plan.reshape(n, -1, 4)[:, :, :2].max(axis=1) + 0.5
    xy = r["traj"][ok][:, :, :2]
    assert (xy >= lo[ok][:, None, :]).all() and (xy <= hi[ok][:, None, :]).all()
    # no push, no failure: a second rollout started from tick 400's state continues the same trajectory bit for bit
    # is not available without the walk state at tick 400, so the lock-step check replays the bookkeeping on the host
    sel = np.nonzero(ok)[0][:24]
    st, wk = state[sel].copy(), walk[sel].copy()
    check_at = set(range(0, T, 97))
    per_tick_push = np.zeros((len(sel), 2))
    for t in range(T):
        for k, i in enumerate(sel):
            fc = wk["footstep_counter"][k]
            if fc < inst["n_steps"][i] and wk["sim_time"][k] >= plan[inst["plan_first_row"][i] + fc, 3] - 1:
                wk["control_iter"][k] = 0; wk["mpc_iter"][k] = 0
                wk["footstep_counter"][k] += 1; wk["support_foot"][k] = 1 - wk["support_foot"][k]
        if t in check_at:
            if t > 0:
                st["com_pos"] = r["traj"][sel, t - 1, :3]; st["com_vel"] = r["traj"][sel, t - 1, 3:]
            active_push = (push["ct0"][sel] <= t) & (t < push["ct1"][sel])
            per_tick_push[:, 0] = np.where(active_push, 0.01 * push["ax"][sel], 0.0)
            per_tick_push[:, 1] = np.where(active_push, 0.01 * push["ay"][sel], 0.0)
            stp = st.copy()
            stp["com_vel"][:, :2] += per_tick_push
            o = O.formc_batch(model, stp, wk, inst[sel], plan, nthreads=8)
            good = (o["ret"] == 0).all(axis=1)
            nxt = o["out"]["next"]
            e = max(np.abs(nxt["com_pos"][good] - r["traj"][sel, t, :3][good]).max(),
                    np.abs(nxt["com_vel"][good] - r["traj"][sel, t, 3:][good]).max())
            assert e <= PRIMAL_TOL, "tick %d: %.3e" % (t, e)
        wk["control_iter"] += 1
        wk["mpc_iter"] = np.floor(wk["control_iter"] * 0.01 / 0.01).astype(np.int32)
        wk["sim_time"] += 1


def _midgait_walk(handle, n_distinct, seed):
    inst, ft, plan = synth.forma_batch(n_distinct, gait="walk", vary=True, ds=30, N_gait=108, seed=seed)
    rng = np.random.default_rng(seed + 1)
    ticks = rng.choice([0, 25, 61, 117, 180, 240], size=n_distinct)
    inst = inst.copy(); plan = plan.copy()
    for t in np.unique(ticks):
        if t == 0:
            continue
        sel = np.nonzero(ticks == t)[0]
        r = handle.forma_rollout(inst[sel], ft, plan, int(t), want_traj=False)
        good = (r["status"] & abi.ST_FAIL_MASK) == 0
        assert good.mean() > 0.99
        inst[sel] = r["inst"]
        rows = (inst["plan_first_row"][sel][:, None] + np.arange(108)[None, :]).reshape(-1)
        plan[rows] = r["fs_plan"][rows]
    return inst, ft, plan


def test_forma_walk_65536_instances(handle):
    """configs[2]: 65,536 walking-gait instances with varied CoM height / step timing, one cold tick each."""
    model = abi.forma_model(q_foot=1e9)
    handle.forma_set_model(model)
    inst, ft, plan = _midgait_walk(handle, N_DISTINCT, 51)
    rep = N_FULL // N_DISTINCT
    inst = np.tile(inst, rep)
    rng = np.random.default_rng(52)
    d = rng.uniform(-0.003, 0.003, (N_FULL, 2))
    inst["st"][:, 0] += d[:, 0]; inst["st"][:, 3] += d[:, 1]
    inst["st"][:, 1] += rng.uniform(-0.02, 0.02, N_FULL); inst["st"][:, 4] += rng.uniform(-0.02, 0.02, N_FULL)
    g = handle.forma_solve_batch(inst, ft, plan)
    failed = (g["out"]["status"] & abi.ST_FAIL_MASK) != 0
    assert failed.mean() < 0.01, "%d instances failed" % failed.sum()
    assert g["out"]["kkt_res"][~failed].max() < 1e-8
    g2 = handle.forma_solve_batch(inst, ft, plan)
    assert g["out"].tobytes() == g2["out"].tobytes() and np.array_equal(g["primal"], g2["primal"])
    perm = np.random.default_rng(8).permutation(N_FULL)
    gp = handle.forma_solve_batch(inst[perm], ft, plan)
    assert np.array_equal(gp["primal"], g["primal"][perm]) and np.array_equal(gp["active"], g["active"][perm])
    for r in range(8):
        a, b = sharding.shard_range(N_FULL, r, 8)
        gs = handle.forma_solve_batch(inst[a:b], ft, plan)
        assert np.array_equal(gs["primal"], g["primal"][a:b]), "shard %d" % r
        assert np.array_equal(gs["active"], g["active"][a:b])
    sel = np.sort(np.random.default_rng(9).choice(N_FULL, 128, replace=False))
    o = O.forma_batch(model, inst[sel], ft, plan, nthreads=8)
    ok = (o["ret"] == 0) & ~failed[sel]
    assert ok.mean() > 0.9
    err = primal_rel_err(g["primal"][sel][ok], o["primal"][ok])
    assert err.max() <= PRIMAL_TOL, "primal rel err %.3e" % err.max()
    mism, _ = active_set_mismatch(g["active"][sel][ok], o["active"][ok], o["duals"][ok])
    assert mism.sum() == 0
