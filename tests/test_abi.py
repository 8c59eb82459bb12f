"""The C ABI: struct layouts match the numpy mirrors, the library loads and exports every declared symbol,
and the product fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from quadruped_gait_generation_ismpc_b200 import abi, binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ismpc_b200.h")


def test_struct_layouts_match_header():
    names = list(abi.DTYPES)
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ismpc_b200.h"', 'int main(void){']
    for nme in names:
        lines.append('printf("%s %%zu", sizeof(%s));' % (nme, nme))
        for f in abi.DTYPES[nme].names:
            lines.append('printf(" %s=%%zu", offsetof(%s, %s));' % (f, nme, f))
        lines.append('printf("\\n");')
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c"); exe = os.path.join(d, "t")
        open(src, "w").write("\n".join(lines))
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    for line in out.strip().splitlines():
        tok = line.split()
        nme, size = tok[0], int(tok[1])
        dt = abi.DTYPES[nme]
        assert dt.itemsize == size == abi.SIZES[nme], nme
        for kv in tok[2:]:
            f, off = kv.split("=")
            assert dt.fields[f][1] == int(off), (nme, f)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(binding.LIB_PATH):
        binding.build()
    declared = set(re.findall(r"\b(ismpc_[a-z_0-9]+)\s*\(", open(HEADER).read()))
    assert declared == set(binding.EXPORTS)
    L = C.CDLL(binding.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), "missing export %s" % s
    assert b"sm_100a" in binding.lib().ismpc_version()


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(binding.IsmpcError):
        binding.Handle(device=0, max_batch=8)


def test_product_never_touches_the_oracle():
    """No file of the shipped package or its CUDA sources references the oracle."""
    pkg = os.path.join(ROOT, "quadruped_gait_generation_ismpc_b200")
    for dp, _, fs in os.walk(pkg):
        if os.sep + "lib" in dp:
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                bad = re.search(r"(from|import)\s+oracle|oracle[/\\]|libismpc_oracle|ismpc_oracle\.h", txt)
                assert bad is None, "%s uses the oracle: %s" % (f, bad.group(0))
    libs = subprocess.check_output(["ldd", binding.LIB_PATH], text=True) if os.path.exists(binding.LIB_PATH) else ""
    assert "oracle" not in libs


def test_plumbing_entry_points_reject_bad_arguments():
    """ismpc_handle_stream / ismpc_wait / ismpc_host_alloc (hosts without CUDA headers): NULL handle and empty
    allocations are refused without touching the device; without a GPU the pinned allocation fails with NULL."""
    L = binding.lib()
    assert L.ismpc_handle_stream(None) is None
    assert L.ismpc_wait(None, None) != 0
    assert L.ismpc_host_alloc(0) is None
    L.ismpc_host_free(None)                                  # a no-op, like free(NULL)
    import torch
    if not torch.cuda.is_available():
        assert L.ismpc_host_alloc(4096) is None


def test_packed_entry_points_reject_bad_arguments_without_touching_the_device():
    """ismpc_formc_set_instances / ismpc_formc_solve_batch_packed: a NULL handle is refused before anything else happens
    (the same on a box without a GPU); the packed record is 128 bytes, state first, walk state at byte 72."""
    L = binding.lib()
    tick = np.zeros(4, dtype=abi.FORMC_TICK); out = np.zeros(4, dtype=abi.FORMC_OUT); inst = np.zeros(4, dtype=abi.FORMC_INST)
    assert L.ismpc_formc_set_instances(None, inst.ctypes.data, 4, abi.MEM_HOST) != 0
    assert L.ismpc_formc_solve_batch_packed(None, 4, tick.ctypes.data, None, None, 0, out.ctypes.data, None, None, abi.MEM_HOST, None) != 0
    assert abi.FORMC_TICK.itemsize == 128 and abi.FORMC_TICK.fields["state"][1] == 0 and abi.FORMC_TICK.fields["walk"][1] == 72
    st = np.zeros(3, dtype=abi.STATE); wk = np.zeros(3, dtype=abi.WALK)
    st["com_pos"] = np.arange(9).reshape(3, 3); wk["mpc_iter"] = [7, 8, 9]; wk["sim_time"] = [0.5, 1.5, 2.5]
    t = abi.pack_ticks(st, wk)
    assert t["state"].tobytes() == st.tobytes() and t["walk"].tobytes() == wk.tobytes() and not t["reserved"].any()


def test_host_library_uses_only_the_c_abi():
    """lib/libismpc_host.so (the C++ serving loop) links against the product library and the C++ runtime only: no CUDA
    runtime, no torch -- it is what a C++ caller of the reference's class would compile with plain g++."""
    host = os.path.join(os.path.dirname(binding.LIB_PATH), "libismpc_host.so")
    assert os.path.exists(host)
    needed = subprocess.check_output(["readelf", "-d", host], text=True)
    libs = re.findall(r"NEEDED.*\[(.*?)\]", needed)
    assert "libismpc_b200.so" in libs
    assert not [x for x in libs if "cuda" in x.lower() or "torch" in x.lower()], libs
