"""Form-A fast-path probe: times the tick with and without the instances that needed the dual active-set
fallback, and dumps those instances (gpurun_out/forma_fallback.npz) for offline analysis.
usage: python tools/forma_fallback_probe.py [n]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadruped_gait_generation_ismpc_b200 import abi, binding, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
gait = sys.argv[2] if len(sys.argv) > 2 else "trot"
dev = torch.device("cuda", 0)
h = binding.Handle(0, max_batch=max(n, 1024))
model = abi.forma_model(q_foot=1e9) if gait == "walk" else abi.forma_model()
h.forma_set_model(model)
if gait == "walk":
    inst, ft, plan = synth.forma_batch(n, gait="walk", vary=True, ds=30, N_gait=108)
else:
    inst, ft, plan = synth.forma_batch(n, gait="trot")
rng = np.random.default_rng(5)
ticks = rng.choice([3, 17, 36, 49, 63, 98, 131, 160, 207, 260], size=n)
for t in np.unique(ticks):
    sel = np.nonzero(ticks == t)[0]
    r = h.forma_rollout(inst[sel], ft, plan, int(t), want_traj=False)
    inst[sel] = r["inst"]
    for i in sel:
        a = inst["plan_first_row"][i]; b = a + inst["n_fs"][i]
        plan[a:b] = r["fs_plan"][a:b]


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)


def timed(sel_inst, label, reps=4):
    m = len(sel_inst)
    d = [to_dev(x) for x in (sel_inst, ft, plan)]
    out = torch.zeros(m * abi.FORMA_OUT.itemsize, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record()
        h.forma_solve_batch_raw(m, d[0].data_ptr(), d[1].data_ptr(), len(ft), d[2].data_ptr(), plan.shape[0],
                                out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3)
    o = np.frombuffer(out.cpu().numpy().tobytes(), dtype=abi.FORMA_OUT)
    it = np.sort(o["iters"])
    print("%s: n=%d best %.1f us (%.2f M QP/s), iters mean %.1f p50 %d p99 %d max %d, failed %d, fallback %d"
          % (label, m, best, m / best, it.mean(), it[m // 2], it[int(m * 0.99)], it[-1],
             (o["status"] & abi.ST_FAIL_MASK != 0).sum(), (o["status"] & abi.ST_GI_FALLBACK != 0).sum()))
    return o


o = timed(inst, "all")
fb = (o["status"] & abi.ST_GI_FALLBACK) != 0
if fb.any():
    os.makedirs("gpurun_out", exist_ok=True)
    idx = np.nonzero(fb)[0]
    np.savez("gpurun_out/forma_fallback_%s.npz" % gait, model=model, inst=inst[idx], fs_timing=ft,
             plans=np.stack([plan[inst["plan_first_row"][i]:inst["plan_first_row"][i] + inst["n_fs"][i]] for i in idx]))
    timed(inst[~fb], "without fallback instances")
for m in (1024, 4096):
    if m < n:
        sub = inst[~fb][:m]
        timed(sub, "first %d clean" % m)
