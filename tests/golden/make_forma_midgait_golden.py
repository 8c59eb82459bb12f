"""Makes tests/golden/forma_midgait.npz: mid-gait formulation-A instances with what the CUDA kernel did on them.

Source: the dumps written ON A B200 by `python tools/forma_dump.py 1024 trot` and `python tools/forma_dump.py 2048 walk`
(gpurun_out/forma_dump_{trot,walk}.npz: the bench's cold mid-gait workloads, the kernel's iteration counts, primal and
working set).  This script keeps N_KEEP instances of each gait -- every fourth one plus the slowest ones -- with their own
plan rows and timing table compacted, so that the CPU tests can rebuild the QPs with the oracle's builder.
usage: python tests/golden/make_forma_midgait_golden.py   (from the repo root, after the two dumps)"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
N_KEEP = 24
out = {}
for gait in ("trot", "walk"):
    d = np.load(os.path.join(ROOT, "gpurun_out", "forma_dump_%s.npz" % gait))
    inst, ft, plan = d["inst"], d["fs_timing"], d["fs_plan"]
    iters = d["out"]["iters"]
    pick = list(range(0, 4 * (N_KEEP - 6), 4)) + list(np.argsort(iters)[-6:])
    pick = sorted(set(int(i) for i in pick))[:N_KEEP]
    sel = inst[pick].copy()
    plans, timings = [], []
    prow, trow = 0, 0
    for k, i in enumerate(pick):
        a, nf = int(inst["plan_first_row"][i]), int(inst["n_fs"][i])
        t0, nt = int(inst["timing_first"][i]), int(inst["n_timing"][i])
        plans.append(plan[a:a + nf]); timings.append(ft[t0:t0 + nt])
        sel["plan_first_row"][k] = prow; sel["timing_first"][k] = trow
        prow += nf; trow += nt
    out[gait + "_model"] = d["model"]
    out[gait + "_inst"] = sel
    out[gait + "_fs_plan"] = np.concatenate(plans)
    out[gait + "_fs_timing"] = np.concatenate(timings).astype(np.int32)
    out[gait + "_kernel_iters"] = iters[pick]
    out[gait + "_kernel_primal"] = d["primal"][pick]
    out[gait + "_kernel_active"] = d["active"][pick]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "forma_midgait.npz"), **out)
print({k: v.shape for k, v in out.items()})
