"""Host-side logic: sharding across ranks (gloo, world_size 2 -- the N>1 path without GPUs), synthetic batches."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from quadruped_gait_generation_ismpc_b200 import abi, sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition():
    for n in (0, 1, 7, 1024, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharding.shard_sizes(n, world)


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(n, rank, world)
    out = np.zeros(hi - lo, dtype=abi.FORMC_OUT)
    ids = np.arange(lo, hi)
    out["fz0"] = ids * 1.5; out["status"] = ids % 7; out["next"]["com_pos"][:, 1] = -ids
    full = sharding.gather_records(out, n)
    t = sharding.max_over_ranks(10.0 + rank)
    if rank == 0:
        ok = (np.array_equal(full["fz0"], np.arange(n) * 1.5) and np.array_equal(full["status"], np.arange(n) % 7)
              and np.array_equal(full["next"]["com_pos"][:, 1], -np.arange(n)) and t == 10.0 + world - 1)
        q.put(bool(ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 1024])
def test_gather_world2_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n % 10
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_synth_batches_are_deterministic_and_in_window():
    a = synth.formc_batch(32, seed=3); b = synth.formc_batch(32, seed=3)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    state, walk, inst, plan = a
    per = inst["S"] + inst["F_ds"]
    assert ((walk["sim_time"] + 200) <= inst["n_steps"] * per).all()
    ai, ft, fp = synth.forma_batch(8, seed=4, vary=True)
    assert (ai["j"] == 1).all() and fp.shape == (800, 2) and (ai["timing_first"] + ai["n_timing"] <= len(ft)).all()


def test_bench_cpu_pinning_degrades_gracefully():
    """bench.py pins a rank's host threads next to its GPU (sysfs + nvidia-smi); on a box without a GPU it reports the
    error and leaves the affinity alone; whatever it does, the process keeps at least one CPU."""
    import os
    import bench
    before = os.sched_getaffinity(0)
    try:
        r = bench.pin_to_gpu_numa_node(0, 1)
        assert isinstance(r, dict)
        after = os.sched_getaffinity(0)
        assert len(after) >= 1
        if "error" in r:
            assert after == before
        else:
            assert set(r["cpus_of_this_rank"]) == after and after <= before
    finally:
        os.sched_setaffinity(0, before)
