"""Dump the bench's cold mid-gait form-A tick workload (instances, iteration counts, solution) for offline analysis.
usage: python tools/forma_dump.py [n] [trot|walk]  ->  gpurun_out/forma_dump_<gait>.npz"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadruped_gait_generation_ismpc_b200 import abi, binding, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
gait = sys.argv[2] if len(sys.argv) > 2 else "trot"
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
g = h.forma_solve_batch(inst, ft, plan)
it = g["out"]["iters"]
print("iters mean %.1f p50 %d p90 %d p99 %d max %d" % (it.mean(), np.percentile(it, 50), np.percentile(it, 90),
                                                      np.percentile(it, 99), it.max()))
print("hist", np.bincount(np.minimum(it, 100))[:100].tolist())
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/forma_dump_%s.npz" % gait, model=model, inst=inst, fs_timing=ft, fs_plan=plan, ticks=ticks,
                    out=g["out"], primal=g["primal"], active=g["active"])
