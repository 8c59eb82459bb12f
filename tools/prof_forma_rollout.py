import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadruped_gait_generation_ismpc_b200 import abi, binding, synth
dev = torch.device("cuda", 0)
h = binding.Handle(0, max_batch=1024)
h.forma_set_model(abi.forma_model())
nr, T = 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 100
inst, ft, plan = synth.forma_batch(nr, gait="trot")
def to_dev(a): return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
d_ft = to_dev(ft); d_status = torch.zeros(nr, dtype=torch.int32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(3):
    d_inst, d_plan = to_dev(inst), to_dev(plan)
    torch.cuda.synchronize(); e0.record()
    h.forma_rollout_raw(nr, T, d_inst.data_ptr(), d_ft.data_ptr(), len(ft), d_plan.data_ptr(), plan.shape[0],
                        status=d_status.data_ptr(), mem=abi.MEM_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
    e1.record(); e1.synchronize()
    print("forma rollout %d x %d ticks: %.2f ms -> %.1f us per tick" % (nr, T, e0.elapsed_time(e1), e0.elapsed_time(e1) * 1e3 / T))
