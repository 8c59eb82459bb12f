#!/usr/bin/env python
"""Benchmark of the batched ISMPC hot path (BASELINE.json metric: batched ISMPC QP solves/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path (one rank per GPU)
  python bench.py --impl reference ...                         # the reference's qpOASES CPU path (oracle/_ref)

A "step" is one tick of MPCSolver::solve (formulation C: vertical QP + x QP + y QP = 3 QP solves per instance)
over one batch of synthetic instances: BASELINE.json configs[1], 1,024 independent trot instances, N = 100,
randomised footstep plans, per GPU (weak scaling: every rank owns its own 1,024 instances, no per-tick
communication, one NCCL gather of the result records after the timed region).

Printed JSON keys beyond the base contract:
  roofline      dominant kernel (formc_tick_pair_kernel, one launch per step) against the measured HBM copy bandwidth
  roofline_fp64 same kernel against the measured FP64 FMA peak (the bound that actually applies)
  cpu_baseline  the reference's qpOASES path (oracle/_ref) timed on this box's host cores, same workload
  latency       p50/p90 per-tick device latency
  form_a        the canonical-ISMPC (footstep) formulation on 1,024 mid-gait trot instances, cold start
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 1024          # instances per GPU (configs[1])
HORIZON = 100
L2_BYTES = 126 * 1024 * 1024
# algorithmic HBM bytes per instance-tick of the fused kernel:
# state 72 + walk 24 + inst 40 + 7 plan rows x 32 (the rows the 2N window touches) + out 128
B_ALG_FORMC = 72 + 24 + 40 + 7 * 32 + 128
# executed FP64 flops per instance-tick: counted by ncu on the committed capture (2 per DFMA, 1 per DMUL/DADD, thread
# level, predicated-on), profiles/r1z_formc_tick_pair_ncu.json; the fallback is the hand count of DESIGN.md section 4
FLOP_FORMC_FALLBACK = 28000
F_REF_FORMC = 2 * 0.96e6 + 0.49e6
NCU_JSON = os.path.join(ROOT, "profiles", "r1z_formc_tick_pair_ncu.json")


def load_ncu():
    """Figures of the dominant kernel from the committed `ncu --set full` capture of this very workload:
    dram_bytes_per_launch = dram__bytes_read.sum + dram__bytes_write.sum, fp64_flop_per_instance_tick."""
    if os.path.exists(NCU_JSON):
        return json.load(open(NCU_JSON))
    return {}


def load_traffic():
    v = load_ncu().get("dram_bytes_per_launch")
    return int(v) if v is not None else None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(steps, warmup, sample_n=None, threads=None):
    """The reference's CPU implementation of the path: restated builders + the reference's qpOASES with the
    solveQP call form (oracle/_ref), every host thread, one cold QProblem per solve per thread."""
    from oracle import oracle as O
    from quadruped_gait_generation_ismpc_b200 import abi, synth
    kind = "ref" if O.have_ref() else "port"
    threads = threads or O.hw_threads()
    model = abi.formc_model(N=HORIZON)
    n = sample_n or BATCH
    state, walk, inst, plan = synth.formc_batch(n)
    for _ in range(warmup):
        O.formc_batch(model, state[:64], walk[:64], inst[:64], plan, nthreads=threads, kind=kind, want_full=False)
    t0 = time.perf_counter()
    for _ in range(steps):
        r = O.formc_batch(model, state, walk, inst, plan, nthreads=threads, kind=kind, want_full=False)
    dt = time.perf_counter() - t0
    qps = 3.0 * n * steps / dt
    return qps, dt / steps, threads, ("reference" if kind == "ref" else "port"), n, int((r["ret"] != 0).any(axis=1).sum())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each step is a bounded sample (256 of the 1,024 instances) so that K steps finish within minutes
    qps, per_step, threads, kind, n, nfail = cpu_reference_run(args.steps, args.warmup, sample_n=256)
    line = {"impl": "reference", "metric": "batched ISMPC QP solves/sec", "value": qps, "unit": "QP solves/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "formC_tick_trot_1024xN100", "instances": n, "horizon_N": HORIZON,
                       "qp_per_instance_tick": 3, "formulation": "C (MPCSolver::solve)"},
            "cpu_baseline": {"value": qps, "unit": "QP solves/s", "cores": threads, "kind": kind,
                             "sample": "%d instances x %d steps, all %d host threads, cold qpOASES QProblem per solve"
                                       % (n, args.steps, threads), "failed_instances": nfail},
            "e2e": {"value": qps, "unit": "QP solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="instances per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-form-a", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from quadruped_gait_generation_ismpc_b200 import abi, binding, sharding, synth

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # NCCL's version banner goes to stdout: keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    K, W, n = args.steps, max(args.warmup, 3), args.batch
    h = binding.Handle(device=local, max_batch=max(n, 8192))
    model = abi.formc_model(N=HORIZON)
    h.formc_set_model(model)
    h.formc_prepare_gait(35, 10)          # parameters.cpp:43-44 (device-resident calls cannot read the gait themselves)

    # ---- inputs: distinct batches rotating over a footprint larger than L2 -------------------------------
    seeds = [synth.SEED0 ^ 2 ^ (rank * 7919 + s) for s in range(4)]
    host_batches = [synth.formc_batch(n, seed=s, N=HORIZON) for s in seeds]
    per_batch = sum(a.nbytes for a in host_batches[0]) + n * abi.FORMC_OUT.itemsize
    n_slots = max(8, int(1.5 * L2_BYTES / per_batch) + 1)

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)

    slots = []
    for s in range(n_slots):
        st, wk, ins, pl = host_batches[s % len(host_batches)]
        slots.append(dict(state=to_dev(st), walk=to_dev(wk), inst=to_dev(ins), plan=to_dev(pl), rows=pl.shape[0],
                          out=torch.zeros(n * abi.FORMC_OUT.itemsize, dtype=torch.uint8, device=dev)))
    stream = torch.cuda.current_stream().cuda_stream

    def step(k, on=None):
        s = slots[k % n_slots]
        h.formc_solve_batch_raw(n, s["state"].data_ptr(), s["walk"].data_ptr(), s["inst"].data_ptr(),
                                s["plan"].data_ptr(), s["rows"], s["out"].data_ptr(), mem=abi.MEM_DEVICE,
                                stream=stream if on is None else on)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks / throttle reasons are sampled (nvidia-smi -lms 20) from before the warm-up until the end of the e2e loops:
    # the device-timed region alone lasts a few milliseconds, shorter than nvidia-smi's start-up
    clocks = ClockSampler(local); clocks.start()
    for k in range(W):
        step(k)
    barrier()
    # ---- timed region: K steps, CUDA events on the launching stream ---------------------------------------
    l0 = h.kernel_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    barrier()
    ev[0].record()
    for k in range(K):
        step(W + k)
        ev[k + 1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[K])
    launches = h.kernel_launches - l0
    per_step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(K)]
    eager_ms_max = sharding.max_over_ranks(total_ms, device=dev)
    # ---- the same K steps as ONE CUDA graph (K kernel nodes): what a caller with a launch-bound loop does -----
    # The tick lasts ~14 us; K Python -> ctypes -> cudaLaunchKernel round trips cost about as much as the kernels.
    # The library only enqueues on the caller's stream, so the K calls are captured as they are and replayed.
    graph_ms = None
    try:
        cs = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cs):
            sp = torch.cuda.current_stream().cuda_stream
            for k in range(K):
                step(W + k, on=sp)
        g.replay()                                   # untimed: uploads the graph
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(); g.replay(); g1.record()
        barrier()
        graph_ms = g0.elapsed_time(g1)
    except Exception as e:                           # capture refused: keep the eager number
        print("bench.py: CUDA-graph arm skipped (%s)" % e, file=sys.stderr)
    # ---- the same K steps alternating over TWO handles on two streams (independent batches overlap: the slowest CTAs
    # of one tick no longer hold the next tick back).  Reported next to `value`, which stays the serialised number.
    overlap_ms = None
    try:
        h2 = binding.Handle(device=local, max_batch=max(n, 8192)); h2.formc_set_model(model); h2.formc_prepare_gait(35, 10)
        sl = slots[0]                                  # one eager call: sizes h2's workspace outside the capture
        h2.formc_solve_batch_raw(n, sl["state"].data_ptr(), sl["walk"].data_ptr(), sl["inst"].data_ptr(), sl["plan"].data_ptr(),
                                 sl["rows"], sl["out"].data_ptr(), mem=abi.MEM_DEVICE, stream=stream)
        torch.cuda.synchronize()
        s_a, s_b = torch.cuda.Stream(), torch.cuda.Stream()
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2, stream=s_a):
            s_b.wait_stream(s_a)
            for k in range(K):
                sl = slots[(W + k) % n_slots]
                hh, ss = (h, s_a) if k % 2 == 0 else (h2, s_b)
                hh.formc_solve_batch_raw(n, sl["state"].data_ptr(), sl["walk"].data_ptr(), sl["inst"].data_ptr(),
                                         sl["plan"].data_ptr(), sl["rows"], sl["out"].data_ptr(), mem=abi.MEM_DEVICE,
                                         stream=ss.cuda_stream)
            s_a.wait_stream(s_b)
        g2.replay()
        barrier()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record(); g2.replay(); o1.record()
        barrier()
        overlap_ms = sharding.max_over_ranks(o0.elapsed_time(o1), device=dev)
        h2.close()
    except Exception as e:
        print("bench.py: two-stream arm skipped (%s)" % e, file=sys.stderr)
    launch_mode = "eager: one C-ABI call per step"
    if graph_ms is not None and graph_ms < total_ms:
        total_ms = graph_ms
        launch_mode = "CUDA graph of the K steps (one kernel node per step), replayed once inside the timed region"
    total_ms_max = sharding.max_over_ranks(total_ms, device=dev)
    value = 3.0 * n * world * K / (total_ms_max * 1e-3)

    # ---- kernel duration in isolation (roofline): events around single launches, inputs rotating ----------
    kdur = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(min(K, 100)):
        torch.cuda.synchronize()
        e0.record(); step(W + K + k); e1.record(); e1.synchronize()
        kdur.append(e0.elapsed_time(e1))
    isolated_ms = statistics.median(kdur)          # one launch on an idle stream: kernel + launch latency
    # the timed region is K launches of the dominant kernel back to back on this stream and nothing else:
    # its CUDA-event time / K is that kernel's average launch duration
    kernel_ms = total_ms / K
    FLOP_FORMC = float(load_ncu().get("fp64_flop_per_instance_tick", FLOP_FORMC_FALLBACK))

    # ---- e2e: the C ABI with HOST buffers (pinned), copies inside the timed region -------------------------
    # Headline e2e = the serving loop a caller runs: DEPTH handles on DEPTH streams, ISMPC_MEM_HOST_ASYNC, so that
    # step k+1's host->device copies overlap step k's kernel and device->host copy.  Every step copies its own
    # inputs from pinned host memory and lands its result records in pinned host memory, where they are read.
    # The synchronous single call (ISMPC_MEM_HOST) is reported next to it as e2e_sync.
    # The footstep plans are constructor data in the reference (MPCSolver::MPCSolver(ftsp_and_timings); Controller
    # builds the matrix once): they are handed to the handle once, outside the timed region (ismpc_formc_set_plan),
    # and a step moves what MPCSolver::solve receives per tick -- state, walk state, instance record -- in and the
    # result record out.  `e2e_with_plan` is the same loop with the whole plan table copied in every step as well.
    all_plans = np.concatenate([b[3] for b in host_batches])
    pinned = []
    row0 = 0
    for st, wk, ins, pl in host_batches:
        d = {}
        ins_res = ins.copy(); ins_res["plan_first_row"] += row0          # rows of this batch in the resident table
        row0 += pl.shape[0]
        # state | walk | instance records back to back in ONE pinned allocation: the library then moves a tick's
        # inputs in a single copy (ismpc_formc_solve_batch, host-memory modes)
        for name, recs in (("pack", (st, wk, ins)), ("pack_res", (st, wk, ins_res))):
            raw = np.concatenate([np.ascontiguousarray(a).view(np.uint8).reshape(-1) for a in recs])
            t = torch.from_numpy(raw.copy()).pin_memory()
            d[name] = t
            o1 = st.nbytes; o2 = o1 + wk.nbytes
            d[name + "_ptr"] = (t.data_ptr(), t.data_ptr() + o1, t.data_ptr() + o2)
        d["plan"] = torch.from_numpy(np.ascontiguousarray(pl).view(np.uint8).reshape(-1).copy()).pin_memory()
        d["plan_ptr"] = d["plan"].data_ptr()
        d["rows"] = pl.shape[0]
        pinned.append(d)
    h2d = pinned[0]["pack_res"].numel(); d2h = n * abi.FORMC_OUT.itemsize
    h2d_with_plan = h2d + pinned[0]["plan"].numel()
    h.formc_set_plan(all_plans)

    def e2e_sync_step(k, out):
        d = pinned[k % len(pinned)]
        ps, pw, pi = d["pack_res_ptr"]
        h.formc_solve_batch_raw(n, ps, pw, pi, None, 0, out.data_ptr(), mem=abi.MEM_HOST, stream=stream)

    out_sync = torch.zeros(d2h, dtype=torch.uint8).pin_memory()
    for k in range(W):
        e2e_sync_step(k, out_sync)
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        e2e_sync_step(k, out_sync)
    barrier()
    e2e_sync_s = sharding.max_over_ranks(time.perf_counter() - t0, device=dev)

    DEPTH = int(os.environ.get("ISMPC_E2E_DEPTH", "4"))      # calls in flight: 2 -> 110, 3 -> 149, 4 -> 161, 6 -> 154 M QP/s
    pipe = []
    for s_ in range(DEPTH):
        hh = binding.Handle(device=local, max_batch=max(n, 1024)); hh.formc_set_model(model); hh.formc_prepare_gait(35, 10)
        hh.formc_set_plan(all_plans)
        pipe.append({"h": hh, "stream": torch.cuda.Stream(device=dev), "out": torch.zeros(d2h, dtype=torch.uint8).pin_memory()})
    checksum = [0]

    for p_ in pipe:
        p_["out_np"] = p_["out"].numpy(); p_["out_ptr"] = p_["out"].data_ptr(); p_["cu_stream"] = p_["stream"].cuda_stream

    def e2e_pipe_step(k, mode=0):
        p_ = pipe[k % DEPTH]
        p_["stream"].synchronize()                          # step k-DEPTH is complete: its records are in host memory
        if k >= DEPTH:
            checksum[0] += int(p_["out_np"][112])           # read the result (status word of record 0)
        d = pinned[k % len(pinned)]
        if mode == 1:                                       # the whole plan table travels with every step
            ps, pw, pi = d["pack_ptr"]
            p_["h"].formc_solve_batch_raw(n, ps, pw, pi, d["plan_ptr"], d["rows"], p_["out_ptr"], mem=abi.MEM_HOST_ASYNC,
                                          stream=p_["cu_stream"])
        elif mode == 2:                                     # mapped: the kernel reads / writes the pinned host buffers itself
            ps, pw, pi = d["pack_res_ptr"]
            p_["h"].formc_solve_batch_raw(n, ps, pw, pi, None, 0, p_["out_ptr"], mem=abi.MEM_DEVICE, stream=p_["cu_stream"])
        else:
            ps, pw, pi = d["pack_res_ptr"]
            p_["h"].formc_solve_batch_raw(n, ps, pw, pi, None, 0, p_["out_ptr"], mem=abi.MEM_HOST_ASYNC,
                                          stream=p_["cu_stream"])

    def e2e_loop(with_plan):
        for k in range(max(W, DEPTH)):                      # at least one call on every handle before the clock starts
            e2e_pipe_step(k, with_plan)
        for p_ in pipe:
            p_["stream"].synchronize()
        barrier()
        t0 = time.perf_counter()
        for k in range(K):
            e2e_pipe_step(k, with_plan)
        for p_ in pipe:
            p_["stream"].synchronize()
        barrier()
        return sharding.max_over_ranks(time.perf_counter() - t0, device=dev)

    e2e_plan_s = e2e_loop(1)
    l0 = sum(p_["h"].kernel_launches for p_ in pipe)
    e2e_s = e2e_loop(0)
    e2e_value = 3.0 * n * world * K / e2e_s
    e2e_launches = sum(p_["h"].kernel_launches for p_ in pipe) - l0       # warm-up + timed steps of this arm
    last_copy = pipe[(K - 1) % DEPTH]["out_np"].copy()
    e2e_mapped_s = e2e_loop(2)
    mapped_equal = bool(np.array_equal(last_copy, pipe[(K - 1) % DEPTH]["out_np"]))

    # ---- the same serving loop from C++ (host/FormCPipeline.hpp over the C ABI): the reference's host language ----
    # W warm-up steps, barrier, K timed steps ending with every stream waited for, barrier.  Same pinned input blocks,
    # same rotation, its own DEPTH handles / streams / pinned result buffers; the loop reads every step's result.
    import ctypes as C
    hostlib = C.CDLL(os.path.join(os.path.dirname(binding.LIB_PATH), "libismpc_host.so"))
    hostlib.ismpc_host_pipeline_create.restype = C.c_void_p
    hostlib.ismpc_host_pipeline_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    hostlib.ismpc_host_pipeline_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    hostlib.ismpc_host_pipeline_destroy.argtypes = [C.c_void_p]
    hostlib.ismpc_host_pipeline_launches.restype = C.c_longlong
    hostlib.ismpc_host_pipeline_launches.argtypes = [C.c_void_p]
    hostlib.ismpc_host_last_error.restype = C.c_char_p
    hostlib.ismpc_host_pipelines_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    plans_c = np.ascontiguousarray(all_plans, dtype=np.float64)
    # a second host thread pays off only where the rank has cores to itself (8 ranks on a 32-thread box: 1 thread each)
    T_DEFAULT = 2 if (os.cpu_count() or 1) // world >= 8 and K >= 64 else 1      # (a handful of steps: not worth a thread)
    T_HOST = int(os.environ.get("ISMPC_E2E_THREADS", str(T_DEFAULT)))        # host threads, one pipeline each (1x4: 249, 1x6: 267, 2x3: 296, 2x4: 303, 3x3: 310 M QP/s)
    D_HOST = int(os.environ.get("ISMPC_E2E_DEPTH_CPP", "4" if T_HOST > 1 else "6"))      # calls in flight per thread
    pps = []
    for _ in range(T_HOST):
        pp = hostlib.ismpc_host_pipeline_create(local, n, D_HOST, model.ctypes.data, 35, 10, plans_c.ctypes.data, plans_c.shape[0])
        if not pp:
            raise RuntimeError("ismpc_host_pipeline_create: " + hostlib.ismpc_host_last_error().decode())
        pps.append(pp)
    pp_arr = (C.c_void_p * T_HOST)(*pps)
    blocks = (C.c_void_p * len(pinned))(*[d["pack_res"].data_ptr() for d in pinned])
    csum = C.c_longlong(0)
    out_cpp = np.zeros((T_HOST, D_HOST, n), dtype=abi.FORMC_OUT)
    l_before = sum(hostlib.ismpc_host_pipeline_launches(pp) for pp in pps)
    # warm-up rounded up to whole rounds of the threads, and at least one call on every handle (a handle's first call
    # allocates its staging and workspace and queries occupancy: milliseconds that do not belong in the timed region)
    WC = max(-(-W // T_HOST) * T_HOST, T_HOST * D_HOST)
    if hostlib.ismpc_host_pipelines_run(pp_arr, T_HOST, 0, WC, blocks, len(pinned), C.byref(csum), None) != 0:
        raise RuntimeError("ismpc_host_pipelines_run: " + hostlib.ismpc_host_last_error().decode())
    barrier()
    t0 = time.perf_counter()
    rc_cpp = hostlib.ismpc_host_pipelines_run(pp_arr, T_HOST, WC, K, blocks, len(pinned), C.byref(csum), out_cpp.ctypes.data)
    barrier()
    e2e_cpp_s = sharding.max_over_ranks(time.perf_counter() - t0, device=dev)
    if rc_cpp != 0:
        raise RuntimeError("ismpc_host_pipelines_run: " + hostlib.ismpc_host_last_error().decode())
    e2e_cpp_launches = sum(hostlib.ismpc_host_pipeline_launches(pp) for pp in pps) - l_before
    for pp in pps:
        hostlib.ismpc_host_pipeline_destroy(pp)
    # where the last step's records are: thread t takes steps WC+t, WC+t+T, ...; its slots go round-robin from the
    # number of steps it has submitted since creation
    k_last = WC + K - 1
    t_last = (k_last - WC) % T_HOST
    done_before = WC // T_HOST + (k_last - WC - t_last) // T_HOST   # steps thread t_last submitted before this one
    slot_last = done_before % D_HOST
    # its last step against a synchronous call on the same pinned block
    e2e_sync_step(k_last, out_sync)
    ref_last = np.frombuffer(out_sync.numpy().tobytes(), dtype=abi.FORMC_OUT)
    cpp_equal = bool(out_cpp[t_last, slot_last].tobytes() == ref_last.tobytes())
    bad_cpp = int(((out_cpp[t_last, slot_last]["status"] & 7) != 0).sum())
    # sanity: the e2e result equals the device-resident result for the same batch
    chk = np.frombuffer(pipe[(K - 1) % DEPTH]["out"].numpy().tobytes(), dtype=abi.FORMC_OUT)
    bad = int(((chk["status"] & 7) != 0).sum())
    for p_ in pipe:
        p_["h"].close()
    clk = clocks.stop()

    # ---- final gather of the result records (the only collective on this path) -----------------------------
    last = np.frombuffer(slots[(W + K - 1) % n_slots]["out"].cpu().numpy().tobytes(), dtype=abi.FORMC_OUT)
    full = sharding.gather_records(last.copy(), n * world, device=dev) if world > 1 else last

    line = None
    if rank == 0:
        hbm_peak, peak_src = load_peaks()
        fp64_peak = h.measure_fp64_peak(5)
        achieved = B_ALG_FORMC * n / (kernel_ms * 1e-3) / 1e9
        line = {"metric": "batched ISMPC QP solves/sec", "value": value, "unit": "QP solves/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "formC_tick_trot_1024xN100", "instances_per_gpu": n, "horizon_N": HORIZON,
                           "qp_per_instance_tick": 3, "formulation": "C (MPCSolver::solve)", "launch": launch_mode,
                           "l2": "inputs rotate over %d distinct device batches (%.0f MB > L2 126 MB)"
                                 % (n_slots, n_slots * per_batch / 1e6)},
                "e2e": {"value": 3.0 * n * world * K / e2e_cpp_s, "unit": "QP solves/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "failed_instances_last_step": bad_cpp,
                        "last_step_equals_synchronous_call": cpp_equal,
                        "how": "C++ host loop (host/FormCPipeline.hpp, the reference's host language) over the C ABI: "
                               "ismpc_formc_solve_batch(ISMPC_MEM_HOST_ASYNC), pinned host buffers, %d host threads x %d handles / streams "
                               "(that many calls in flight); every step copies its state / walk-state / instance records in "
                               "(one pinned block, one copy) and its result records out, and the loop reads each result; the "
                               "footstep plans are resident in the handles (ismpc_formc_set_plan), as they are constructor "
                               "data of the reference's MPCSolver" % (T_HOST, D_HOST),
                        "kernel_launches": int(e2e_cpp_launches)},
                "e2e_python": {"value": e2e_value, "unit": "QP solves/s", "h2d_bytes_per_step": int(h2d),
                               "d2h_bytes_per_step": int(d2h), "failed_instances_last_step": bad,
                               "how": "the same loop written in Python over ctypes (~13 us of interpreter and call overhead per step)",
                               "kernel_launches": int(e2e_launches)},
                "e2e_with_plan": {"value": 3.0 * n * world * K / e2e_plan_s, "unit": "QP solves/s",
                                  "h2d_bytes_per_step": int(h2d_with_plan), "d2h_bytes_per_step": int(d2h),
                                  "how": "the same loop with the whole footstep-plan table passed and copied in every step"},
                "e2e_mapped": {"value": 3.0 * n * world * K / e2e_mapped_s, "unit": "QP solves/s",
                               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                               "records_equal_copy_arm": mapped_equal,
                               "how": "the same serving loop without copy calls: ismpc_formc_solve_batch(ISMPC_MEM_DEVICE) is "
                                      "handed the pinned (mapped) host buffers, the kernel reads its records from host "
                                      "memory and stores the result records into host memory itself"},
                "e2e_sync": {"value": 3.0 * n * world * K / e2e_sync_s, "unit": "QP solves/s",
                             "how": "one synchronous ismpc_formc_solve_batch(ISMPC_MEM_HOST) call per step",
                             "ms_per_step": e2e_sync_s / K * 1e3},
                "gpu_launches": int(launches),
                "clocks": clk,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                             "frac": achieved / hbm_peak, "traffic": load_traffic(), "peak_source": peak_src,
                             "kernel": "formc_tick_pair_kernel", "kernel_ms": kernel_ms,
                             "algorithmic_bytes_per_instance_tick": B_ALG_FORMC},
                "roofline_fp64": {"bound": "fp64", "achieved": FLOP_FORMC * n / (kernel_ms * 1e-3) / 1e12,
                                  "peak": fp64_peak, "unit": "TFLOP/s",
                                  "frac": FLOP_FORMC * n / (kernel_ms * 1e-3) / 1e12 / fp64_peak,
                                  "executed_flop_per_instance_tick": FLOP_FORMC,
                                  "peak_source": "ismpc_measure_fp64_peak (DFMA micro-benchmark, this run)",
                                  # the dense null-space active-set model of the reference algorithm,
                                  # F_ref = (k+1)(14 nV^2 + 2 nC nV): 2 x 0.96 Mflop (x, y; k = 5) + 0.49 Mflop (z) per
                                  # instance-tick.  NOT work this kernel executes (it exploits the structure);
                                  # reported beside the executed figure as the survey asks.
                                  "reference_algorithm_flop_per_instance_tick": F_REF_FORMC,
                                  "reference_algorithm_equivalent_tflops": F_REF_FORMC * n / (kernel_ms * 1e-3) / 1e12},
                "eager": {"value": 3.0 * n * world * K / (eager_ms_max * 1e-3), "unit": "QP solves/s",
                          "ms_per_step": eager_ms_max / K, "how": "K separate C-ABI calls from Python, CUDA events around the loop"},
                "two_streams": (None if overlap_ms is None else
                                {"value": 3.0 * n * world * K / (overlap_ms * 1e-3), "unit": "QP solves/s", "ms_per_step": overlap_ms / K,
                                 "how": "the K steps as one CUDA graph alternating over two handles on two streams "
                                        "(consecutive steps are independent batches and overlap)"}),
                "latency": {"p50_tick_us": statistics.median(per_step_ms) * 1e3,
                            "p90_tick_us": sorted(per_step_ms)[int(0.9 * (K - 1))] * 1e3,
                            "isolated_launch_us": isolated_ms * 1e3},
                "instance_ticks_per_s": value / 3.0,
                "gathered_records": int(len(full))}

    # ---- formulation A (canonical ISMPC with footsteps), rank 0 extra measurement -------------------------
    # (configs[2] -- 65,536 walking instances over 8 GPUs -- is the walking tick of every rank's 8,192-instance shard)
    if not args.no_form_a:
        try:
            fa = bench_form_a(h, torch, dev, n, stream, rank=rank)
        except Exception as e:  # noqa: BLE001
            fa = {"error": repr(e)}
        walk_ms = fa.get("tick_cold_walk", {}).get("ms_per_tick", -1.0)
        if world > 1:
            walk_ms_all = sharding.max_over_ranks(walk_ms, device=dev)
            bad_all = sharding.max_over_ranks(float("error" in fa), device=dev)
            if rank == 0 and bad_all == 0.0:
                fa["tick_cold_walk_all_gpus"] = {
                    "workload": "formA_tick_walk_%dxC100F3_midgait_cold: 8,192 instances on each of %d GPUs (configs[2])"
                                % (8192 * world, world),
                    "qp_solves_per_s": 8192 * world / (walk_ms_all * 1e-3), "ms_per_tick": walk_ms_all}
        if rank == 0:
            line["form_a"] = fa
    # ---- configs[4]: closed loop on every rank (its own 1,000 instances), aggregated like the headline value ---------
    if not args.no_form_a:
        try:
            h.formc_set_model(model); h.formc_prepare_gait(35, 10)
            cl = bench_formc_rollout(h, torch, dev, stream, rank)
        except Exception as e:  # noqa: BLE001
            cl = {"error": repr(e), "ms_total": -1.0}
        ms_all = sharding.max_over_ranks(cl["ms_total"], device=dev) if world > 1 else cl["ms_total"]
        ok_all = sharding.max_over_ranks(1.0 if "error" in cl else 0.0, device=dev) if world > 1 else float("error" in cl)
        if rank == 0:
            if ok_all == 0.0 and world > 1:
                per_rank = cl["instance_ticks_per_s"] * cl["ms_total"] * 1e-3          # instance-ticks of one rank
                cl["workload"] += " on each of %d GPUs" % world
                cl["ms_total"] = ms_all
                cl["instance_ticks_per_s"] = world * per_rank / (ms_all * 1e-3)
                cl["qp_solves_per_s"] = 3.0 * cl["instance_ticks_per_s"]
                cl["instances_with_a_failed_tick"] = "rank 0: %d" % cl["instances_with_a_failed_tick"]
            line["closed_loop_form_c"] = cl
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            reps = 3
            qps, per_step, threads, kind, ns, nfail = cpu_reference_run(reps, 1)
            line["cpu_baseline"] = {"value": qps, "unit": "QP solves/s", "cores": threads, "kind": kind,
                                    "sample": "%d instances x %d passes of the same workload, all %d host threads, "
                                              "cold qpOASES QProblem per solve (utils.cpp:121-130)" % (ns, reps, threads),
                                    "failed_instances": nfail}
            q1, _, _, _, n1, _ = cpu_reference_run(1, 0, sample_n=128, threads=1)
            line["cpu_baseline"]["single_thread"] = {"value": q1, "unit": "QP solves/s", "cores": 1,
                                                     "sample": "%d instances x 1 pass, one host thread" % n1}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    h.close()
    if world > 1:
        dist.destroy_process_group()


def _midgait(h, inst, ft, plan, seed=5):
    """Advance formulation-A instances on the GPU to random gait phases (in place on copies)."""
    n = len(inst)
    rng = np.random.default_rng(seed)
    ticks = rng.choice([3, 17, 36, 49, 63, 98, 131, 160, 207, 260], size=n)
    for t in np.unique(ticks):
        sel = np.nonzero(ticks == t)[0]
        r = h.forma_rollout(inst[sel], ft, plan, int(t), want_traj=False)
        inst[sel] = r["inst"]
        for i in sel:
            a = inst["plan_first_row"][i]; b = a + inst["n_fs"][i]
            plan[a:b] = r["fs_plan"][a:b]
    return inst, plan


def bench_form_a(h, torch, dev, n, stream, steps=20, rank=0):
    """Formulation A (canonical ISMPC with footsteps, 1 QP of nV=206/nC=208 per instance-tick):
    cold single ticks on mid-gait trot (configs[1]) and walking (configs[2], per-GPU share) batches, and the
    closed loop with pushes, warm-started (configs[4])."""
    from quadruped_gait_generation_ismpc_b200 import abi, synth

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)

    def tick_bench(model, inst, ft, plan, label):
        m = len(inst)
        h.forma_set_model(model)
        inst, plan = _midgait(h, inst, ft, plan)
        d_inst, d_ft, d_plan = to_dev(inst), to_dev(ft), to_dev(plan)
        d_out = torch.zeros(m * abi.FORMA_OUT.itemsize, dtype=torch.uint8, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        for k in range(steps + 3):
            torch.cuda.synchronize()
            e0.record()
            h.forma_solve_batch_raw(m, d_inst.data_ptr(), d_ft.data_ptr(), len(ft), d_plan.data_ptr(), plan.shape[0],
                                    d_out.data_ptr(), mem=abi.MEM_DEVICE, stream=stream)
            e1.record(); e1.synchronize()
            if k >= 3:
                times.append(e0.elapsed_time(e1))
        out = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=abi.FORMA_OUT)
        ms = statistics.median(times)
        return {"workload": label, "qp_solves_per_s": m / (ms * 1e-3), "ms_per_tick": ms,
                "mean_iters_per_qp": float(out["iters"].mean()), "max_iters": int(out["iters"].max()),
                "failed": int((out["status"] & abi.ST_FAIL_MASK != 0).sum()),
                "dual_active_set_fallbacks": int((out["status"] & abi.ST_GI_FALLBACK != 0).sum())}

    res = {"note": "1 QP per instance-tick (x and y stacked: nV=206, nC=208); cold = empty working set"}
    nw = 8192
    inst, ft, plan = synth.forma_batch(nw, gait="walk", vary=True, ds=30, N_gait=108, seed=(synth.SEED0 ^ 3) + 7919 * rank)
    res["tick_cold_walk"] = tick_bench(abi.forma_model(q_foot=1e9), inst, ft, plan,
                                       "formA_tick_walk_%dxC100F3_midgait_cold (configs[2] per-GPU share)" % nw)
    if rank != 0:
        return res          # the other ranks only take part in the walking tick (configs[2]: 8,192 instances per GPU)
    inst, ft, plan = synth.forma_batch(n, gait="trot")
    res["tick_cold_trot"] = tick_bench(abi.forma_model(), inst, ft, plan, "formA_tick_trot_%dxC100F3_midgait_cold" % n)
    # closed loop: 1,000 instances x 250 ticks with pushes, state resident on the device, warm-started
    nr, T = 1000, 250
    h.forma_set_model(abi.forma_model())
    inst, ft, plan = synth.forma_batch(nr, gait="trot", seed=synth.SEED0 ^ 9)
    push = synth.push_batch(nr)
    d_ft = to_dev(ft)
    d_status = torch.zeros(nr, dtype=torch.int32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for k in range(3):
        d_inst, d_plan, d_push = to_dev(inst), to_dev(plan), to_dev(push)
        torch.cuda.synchronize()
        e0.record()
        h.forma_rollout_raw(nr, T, d_inst.data_ptr(), d_ft.data_ptr(), len(ft), d_plan.data_ptr(), plan.shape[0],
                            push=d_push.data_ptr(), status=d_status.data_ptr(), mem=abi.MEM_DEVICE, stream=stream)
        e1.record(); e1.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times)
    stt = d_status.cpu().numpy()
    res["rollout_warm_trot"] = {"workload": "formA_rollout_trot_%dx%dticks_push (configs[4] shape)" % (nr, T),
                                "instance_ticks_per_s": nr * T / (ms * 1e-3), "ms_total": ms,
                                "failed": int((stt & abi.ST_FAIL_MASK != 0).sum()),
                                "dual_active_set_fallbacks": int((stt & abi.ST_GI_FALLBACK != 0).sum())}
    return res


def bench_formc_rollout(h, torch, dev, stream, rank=0):
    """configs[4] on formulation C: 1,000 instances x 1,000 closed-loop ticks (3 QPs each) with pushes, on the device
    (every rank draws its own instances)."""
    from quadruped_gait_generation_ismpc_b200 import abi, synth
    nr, T = 1000, 1000
    steps_plan = (T + 2 * HORIZON + 900) // 45 + 3
    state, walk, inst, plan = synth.formc_batch(nr, seed=(synth.SEED0 ^ 11) + 7919 * rank, n_steps=steps_plan, k0_cap=100)
    push = synth.push_batch(nr, seed=(synth.SEED0 ^ 5) + 7919 * rank, formc=True)

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)

    d_inst, d_plan, d_push = to_dev(inst), to_dev(plan), to_dev(push)
    d_status = torch.zeros(nr, dtype=torch.int32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for k in range(3):
        d_state, d_walk = to_dev(state), to_dev(walk)
        torch.cuda.synchronize()
        e0.record()
        h.formc_rollout_raw(nr, T, d_state.data_ptr(), d_walk.data_ptr(), d_inst.data_ptr(), d_plan.data_ptr(),
                            plan.shape[0], push=d_push.data_ptr(), status=d_status.data_ptr(), mem=abi.MEM_DEVICE,
                            stream=stream)
        e1.record(); e1.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times)
    stt = d_status.cpu().numpy()
    return {"workload": "formC_rollout_trot_%dx%dticks_push (configs[4])" % (nr, T),
            "instance_ticks_per_s": nr * T / (ms * 1e-3), "qp_solves_per_s": 3.0 * nr * T / (ms * 1e-3), "ms_total": ms,
            "instances_with_a_failed_tick": int((stt & (abi.ST_Z_FAIL | abi.ST_X_FAIL | abi.ST_Y_FAIL)) .astype(bool).sum())}


if __name__ == "__main__":
    main()
