"""BASELINE.json configs[3]: horizon sweep N = 50/100/200/400, 1,024 instances, warp-per-QP vs CTA-per-QP vs
cluster-per-QP (formulation C, 3 QPs per instance-tick) and the formulation-A tick at the same horizons (C = N, P = 2N).
Times are per launch with BURST launches back to back between the CUDA events (hides the host-side launch cost).
usage: python tools/horizon_sweep.py [n] [out.json]"""
import json
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadruped_gait_generation_ismpc_b200 import abi, binding, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
small = [int(x) for x in os.environ.get("SWEEP_SMALL", "1,64").split(",") if x]
out_path = sys.argv[2] if len(sys.argv) > 2 else None
dev = torch.device("cuda", 0)
h = binding.Handle(0, max_batch=max(n, 1024))
stream = torch.cuda.current_stream().cuda_stream


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)


BURST = int(os.environ.get("SWEEP_BURST", "10"))


def timed(fn, reps=12, warm=3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for k in range(reps + warm):
        torch.cuda.synchronize()
        e0.record()
        for _ in range(BURST):
            fn()
        e1.record(); e1.synchronize()
        if k >= warm:
            ts.append(e0.elapsed_time(e1) * 1e3 / BURST)
    return statistics.median(ts)


rows = []
for N in (50, 100, 200, 400):
    steps = (2 * N + 900) // 45 + 3
    st, wk, ins, pl = synth.formc_batch(n, seed=N, N=N, n_steps=steps)
    h.formc_set_model(abi.formc_model(N=N))
    h.formc_prepare_gait(35, 10)
    d = [to_dev(x) for x in (st, wk, ins, pl)]
    out = torch.zeros(n * abi.FORMC_OUT.itemsize, dtype=torch.uint8, device=dev)
    ref = None
    h.set_option("formc_cluster_size", 0); h.set_option("formc_kernel", 2)
    for m in [n] + small:                # warp-per-QP (the default family)
        us = timed(lambda: h.formc_solve_batch_raw(m, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(),
                                                   pl.shape[0], out.data_ptr(), stream=stream))
        o = np.frombuffer(out.cpu().numpy().tobytes(), dtype=abi.FORMC_OUT).copy()
        rows.append({"formulation": "C", "N": N, "instances": m, "kernel": "warp_per_qp", "us_per_tick": us,
                     "qp_per_s": 3.0 * m / (us * 1e-6),
                     "failed": int((o["status"][:m] & (abi.ST_Z_FAIL | abi.ST_X_FAIL | abi.ST_Y_FAIL) != 0).sum())})
        print(rows[-1])
    h.set_option("formc_kernel", 0)
    for cs in (1, 2, 4, 8):
        h.set_option("formc_cluster_size", cs)
        us = timed(lambda: h.formc_solve_batch_raw(n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(),
                                                   pl.shape[0], out.data_ptr(), stream=stream))
        o = np.frombuffer(out.cpu().numpy().tobytes(), dtype=abi.FORMC_OUT).copy()
        if ref is None:
            ref = o
        same = bool(np.array_equal(o["next"]["com_pos"], ref["next"]["com_pos"]) and np.array_equal(o["status"], ref["status"]))
        rows.append({"formulation": "C", "N": N, "ctas_per_qp": cs, "us_per_tick": us, "qp_per_s": 3.0 * n / (us * 1e-6),
                     "identical_to_cta_per_qp": same,
                     "failed": int((o["status"] & (abi.ST_Z_FAIL | abi.ST_X_FAIL | abi.ST_Y_FAIL) != 0).sum())})
        print(rows[-1])
    for m in small:                      # latency regime: few instances, the GPU is not full
        for cs in (1, 2, 4, 8):
            h.set_option("formc_cluster_size", cs)
            us = timed(lambda: h.formc_solve_batch_raw(m, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(),
                                                       pl.shape[0], out.data_ptr(), stream=stream))
            rows.append({"formulation": "C", "N": N, "instances": m, "ctas_per_qp": cs, "us_per_tick": us})
            print(rows[-1])
    h.set_option("formc_cluster_size", 0)
for C in (50, 100, 200, 400):
    step = C // 2
    model = abi.forma_model(C=C, P=2 * C)
    h.forma_set_model(model)
    inst, ft, plan = synth.forma_batch(n, gait="trot", C=C, step=step, ds=max(2, step * 2 // 5), sim_ticks=20 * step)
    rng = np.random.default_rng(5)
    ticks = rng.choice([3, step // 3, step - 1, step + step // 4, 2 * step + 5, 3 * step - 2], size=n)
    for t in np.unique(ticks):
        sel = np.nonzero(ticks == t)[0]
        r = h.forma_rollout(inst[sel], ft, plan, int(t), want_traj=False)
        inst[sel] = r["inst"]
        for i in sel:
            a = inst["plan_first_row"][i]; b = a + inst["n_fs"][i]
            plan[a:b] = r["fs_plan"][a:b]
    d = [to_dev(x) for x in (inst, ft, plan)]
    out = torch.zeros(n * abi.FORMA_OUT.itemsize, dtype=torch.uint8, device=dev)
    us = timed(lambda: h.forma_solve_batch_raw(n, d[0].data_ptr(), d[1].data_ptr(), len(ft), d[2].data_ptr(), plan.shape[0],
                                               out.data_ptr(), stream=stream), reps=10, warm=2)
    o = np.frombuffer(out.cpu().numpy().tobytes(), dtype=abi.FORMA_OUT)
    rows.append({"formulation": "A", "N": C, "warps_per_qp": 2, "us_per_tick": us, "qp_per_s": n / (us * 1e-6),
                 "mean_iters": float(o["iters"].mean()), "failed": int((o["status"] & abi.ST_FAIL_MASK != 0).sum()),
                 "dual_active_set_fallbacks": int((o["status"] & abi.ST_GI_FALLBACK != 0).sum())})
    print(rows[-1])
if out_path:
    json.dump({"instances": n, "rows": rows}, open(out_path, "w"), indent=1)
