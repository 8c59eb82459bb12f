"""Driver for profiling the dense solveQP seam: python tools/prof_dense.py [shape] [n]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadruped_gait_generation_ismpc_b200 import abi, binding, synth  # noqa: E402
shape = sys.argv[1] if len(sys.argv) > 1 else "forma_stacked"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda", 0)
h = binding.Handle(0, max_batch=max(n, 1024))
H, g, A, lb, ub = synth.dense_qp_batch(shape, n)
nV, nC = H.shape[1], A.shape[1]
d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (H, g, A, lb, ub)]
dx = torch.zeros((n, nV), dtype=torch.float64, device=dev)
dst = torch.zeros(n, dtype=torch.int32, device=dev); dit = torch.zeros(n, dtype=torch.int32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(3):
    torch.cuda.synchronize(); e0.record()
    h.qp_solve_batch_raw(n, nV, nC, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(), dx.data_ptr(),
                         status=dst.data_ptr(), iters=dit.data_ptr(), mem=abi.MEM_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
    e1.record(); e1.synchronize()
    print("%s n=%d: %.3f ms, iters mean %.1f max %d, failed %d" % (shape, n, e0.elapsed_time(e1), dit.float().mean().item(), dit.max().item(), (dst != 0).sum().item()))
