"""Summarise one kernel of an ncu report (--set full) into the JSON bench.py reads and a metrics text file.
usage: python tools/ncu_summary.py report.ncu-rep kernel-regex instances out.json out_metrics.txt"""
import csv
import io
import json
import re
import subprocess
import sys

rep, kre, n_inst, out_json, out_txt = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, val = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, val)}


def num(name, default=None):
    if name not in m:
        return default
    try:
        return float(m[name][0].replace(",", ""))
    except ValueError:
        return default


def to_bytes(name):
    v, u = m[name]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u.split("/")[0]]
    return float(v.replace(",", "")) * scale


# FP64 work actually executed: predicated-on THREAD instructions per opcode from the SASS page of the same capture
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
shdr = next(r for r in srows if r and r[0] == "Address")
i_src, i_thr = shdr.index("Source"), shdr.index("Predicated-On Thread Instructions Executed")
dfma = dmul = dadd = 0.0
for r in srows:
    if not r or not r[0].startswith("0x"):
        continue
    t = r[i_src].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c = float(r[i_thr])
    if op == "DFMA":
        dfma += c
    elif op == "DMUL":
        dmul += c
    elif op == "DADD":
        dadd += c
flop = 2 * dfma + dmul + dadd
dur_v, dur_u = m["gpu__time_duration.sum"]
dur_us = float(dur_v) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[dur_u]
summary = {
    "kernel": m["Kernel Name"][0] if "Kernel Name" in m else kre,
    "source": "ncu --set full --clock-control none (cold caches, serialised), " + rep.split("/")[-1],
    "instances": n_inst,
    "gpu_time_duration_us": dur_us,
    "dram_bytes_read": to_bytes("dram__bytes_read.sum"),
    "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
    "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
    "fp64_thread_inst": {"dfma": dfma, "dmul": dmul, "dadd": dadd},
    "fp64_flop_per_launch": flop,
    "fp64_flop_per_instance_tick": flop / n_inst,
    "warp_inst_executed_per_instance": num("smsp__inst_executed.sum", 0.0) / n_inst,
    "registers_per_thread": num("launch__registers_per_thread"),
    "shared_mem_per_block_bytes": to_bytes("launch__shared_mem_per_block") if "launch__shared_mem_per_block" in m else None,
    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "pipe_fp64_cycles_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "icc_hit_rate_pct": num("sm__icc_request_hit_rate.pct"),
    "warp_stall_samples": {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): num(k) for k in hdr
                           if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and num(k)},
}
json.dump(summary, open(out_json, "w"), indent=1)
keep = re.compile(r"^(gpu__time|dram__bytes|launch__|sm__cycles|sm__inst_executed|smsp__inst_executed\.sum|smsp__issue_active|"
                  r"sm__pipe_fp64|smsp__sass_thread_inst_executed_op_d|lts__t_sectors_srcunit_tex_op_read\.sum|sm__icc|"
                  r"smsp__pcsamp_warps_issue_stalled|l1tex__t_sector_hit_rate|sm__warps_active|smsp__average_warp_latency)")
with open(out_txt, "w") as f:
    for h, u, v in zip(hdr, units, val):
        if keep.match(h):
            f.write("%-90s %s %s\n" % (h, v, u))
print(json.dumps(summary, indent=1))
