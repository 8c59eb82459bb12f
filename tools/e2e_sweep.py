"""e2e serving loop (host/FormCPipeline.hpp over the C ABI) for a few thread x depth shapes and step counts.
usage: python tools/e2e_sweep.py [n]   (env ISMPC_HOST_DIRECT_OUT=0/1 picks the result path)"""
import ctypes as C
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadruped_gait_generation_ismpc_b200 import abi, binding, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
model = abi.formc_model(N=100)
batches = [synth.formc_batch(n, seed=100 + s, N=100) for s in range(8)]
plans = np.ascontiguousarray(np.concatenate([b[3] for b in batches]), dtype=np.float64)
pinned = []
row0 = 0
for st, wk, ins, pl in batches:
    ins = ins.copy(); ins["plan_first_row"] += row0; row0 += pl.shape[0]
    raw = np.concatenate([np.ascontiguousarray(a).view(np.uint8).reshape(-1) for a in (st, wk, ins)])
    pinned.append(torch.from_numpy(raw.copy()).pin_memory())
hostlib = C.CDLL(os.path.join(os.path.dirname(binding.LIB_PATH), "libismpc_host.so"))
hostlib.ismpc_host_pool_create.restype = C.c_void_p
hostlib.ismpc_host_pool_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
hostlib.ismpc_host_pool_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
hostlib.ismpc_host_pool_destroy.argtypes = [C.c_void_p]
hostlib.ismpc_host_pool_set_spin_us.argtypes = [C.c_void_p, C.c_int]
hostlib.ismpc_host_last_error.restype = C.c_char_p
hostlib.ismpc_host_pool_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
blocks = (C.c_void_p * len(pinned))(*[t.data_ptr() for t in pinned])
print("direct_out =", os.environ.get("ISMPC_HOST_DIRECT_OUT", "0"), "variant =", os.environ.get("ISMPC_FORMC_VARIANT", "0"), "n =", n)
# one synchronous call per tick (what a single Controller-style caller sees)
import time
h = binding.Handle(0, max_batch=n)
h.formc_set_model(model); h.formc_prepare_gait(35, 10); h.formc_set_plan(plans)
out_sync = torch.zeros(n * abi.FORMC_OUT.itemsize, dtype=torch.uint8).pin_memory()
o1 = batches[0][0].nbytes; o2 = o1 + batches[0][1].nbytes
stream = torch.cuda.current_stream().cuda_stream
def sync_step(k):
    p = pinned[k % len(pinned)].data_ptr()
    h.formc_solve_batch_raw(n, p, p + o1, p + o2, None, 0, out_sync.data_ptr(), mem=abi.MEM_HOST, stream=stream)
for k in range(20):
    sync_step(k)
ts = []
for r in range(5):
    t0 = time.perf_counter()
    for k in range(200):
        sync_step(k)
    ts.append((time.perf_counter() - t0) / 200)
print("synchronous ISMPC_MEM_HOST call from Python: %.1f us per tick" % (statistics.median(ts) * 1e6))
ins0 = batches[0][2].copy()
h.formc_set_instances(ins0)
tk0 = torch.from_numpy(abi.pack_ticks(batches[0][0], batches[0][1]).view(np.uint8).reshape(-1).copy()).pin_memory()
for zc in (1, 0):
    h.set_option("host_zero_copy", zc)
    for k in range(20):
        h.formc_solve_batch_packed_raw(n, tk0.data_ptr(), None, None, 0, out_sync.data_ptr(), mem=abi.MEM_HOST, stream=stream)
    ts = []
    for r in range(5):
        t0 = time.perf_counter()
        for k in range(200):
            h.formc_solve_batch_packed_raw(n, tk0.data_ptr(), None, None, 0, out_sync.data_ptr(), mem=abi.MEM_HOST, stream=stream)
        ts.append((time.perf_counter() - t0) / 200)
    print("synchronous packed call (host_zero_copy = %d) from Python: %.1f us per tick" % (zc, statistics.median(ts) * 1e6))
h.close()
hostlib.ismpc_host_pool_set_instances.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
hostlib.ismpc_host_pool_run_packed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
# packed loop: one fleet per slot (constants and plans resident), 4 successive ticks of its closed loop as tick records
hh = binding.Handle(0, max_batch=n); hh.formc_set_model(model); hh.formc_prepare_gait(35, 10)
fleet_ticks, fleet_inst = [], []
row0 = 0
for st, wk, ins, pl in batches:
    ins = ins.copy(); ins["plan_first_row"] += row0; row0 += pl.shape[0]
    fleet_inst.append(ins)
    tks = []
    for j in range(4):
        r = hh.formc_rollout(st, wk, ins, plans, j, want_traj=False) if j else dict(state=st, walk=wk)
        tks.append(torch.from_numpy(abi.pack_ticks(r["state"], r["walk"]).view(np.uint8).reshape(-1).copy()).pin_memory())
    fleet_ticks.append(tks)
hh.close()
for T, D in [(1, 8), (2, 4), (1, 4), (2, 2), (1, 2)]:
    pool = hostlib.ismpc_host_pool_create(0, n, T, D, model.ctypes.data, 35, 10, plans.ctypes.data, plans.shape[0])
    hostlib.ismpc_host_pool_set_spin_us(pool, 200000)
    ptrs = []
    for slot in range(T * D):
        f = slot % len(batches)
        assert hostlib.ismpc_host_pool_set_instances(pool, slot // D, slot % D, fleet_inst[f].ctypes.data) == 0
        ptrs += [t.data_ptr() for t in fleet_ticks[f]]
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    cs = C.c_longlong(0); el = C.c_double(0.0)
    hostlib.ismpc_host_pool_run_packed(pool, 0, 2 * T * D, arr, 4, C.byref(cs), None, None)
    res = []
    for K in (20, 200, 2000):
        ts = []
        for r in range(7):
            if hostlib.ismpc_host_pool_run_packed(pool, 0, K, arr, 4, C.byref(cs), None, C.byref(el)) != 0:
                raise RuntimeError(hostlib.ismpc_host_last_error().decode())
            ts.append(el.value)
        med = statistics.median(ts)
        w = C.c_double(0.0); sb = C.c_double(0.0)
        hostlib.ismpc_host_pool_stats(pool, C.byref(w), C.byref(sb))
        res.append("K=%d: %.2f us/step (%.0f M QP/s; host per step: submit %.2f us, wait %.2f us)" % (K, med / K * 1e6, 3.0 * n * K / med / 1e6, sb.value / K * 1e6, w.value / K * 1e6))
    print("PACKED T=%d D=%d  " % (T, D) + "  ".join(res), flush=True)
    hostlib.ismpc_host_pool_destroy(pool)
for T, D in [(2, 4)]:
    pool = hostlib.ismpc_host_pool_create(0, n, T, D, model.ctypes.data, 35, 10, plans.ctypes.data, plans.shape[0])
    if not pool:
        raise RuntimeError(hostlib.ismpc_host_last_error().decode())
    hostlib.ismpc_host_pool_set_spin_us(pool, 200000)
    cs = C.c_longlong(0); el = C.c_double(0.0)
    hostlib.ismpc_host_pool_run(pool, 0, 2 * T * D, blocks, len(pinned), C.byref(cs), None, None)
    res = []
    for K in (20, 200, 2000):
        ts = []
        for r in range(7):
            if hostlib.ismpc_host_pool_run(pool, 0, K, blocks, len(pinned), C.byref(cs), None, C.byref(el)) != 0:
                raise RuntimeError(hostlib.ismpc_host_last_error().decode())
            ts.append(el.value)
        med = statistics.median(ts)
        w = C.c_double(0.0); sb = C.c_double(0.0)
        hostlib.ismpc_host_pool_stats(pool, C.byref(w), C.byref(sb))
        res.append("K=%d: %.2f us/step (%.0f M QP/s; host per step: submit %.2f us, wait %.2f us)" % (K, med / K * 1e6, 3.0 * n * K / med / 1e6, sb.value / K * 1e6, w.value / K * 1e6))
    print("T=%d D=%d  " % (T, D) + "  ".join(res), flush=True)
    hostlib.ismpc_host_pool_destroy(pool)
