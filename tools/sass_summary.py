"""SASS evidence of the product library: per kernel, the instruction families that matter on this path -- 1-D TMA bulk
copies (UBLKCP) with their mbarrier waits (SYNCS), FP64 arithmetic (DFMA / DMUL / DADD), FP64 tensor-core MMA (DMMA), the
reciprocal seed (MUFU.RCP64H), warp reductions (REDUX), shuffles, programmatic-dependent-launch control (ACQBULK / ... as
emitted for griddepcontrol), local-memory traffic (LDL / STL = spills).
usage: python tools/sass_summary.py [lib.so] > profiles/rN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "quadruped_gait_generation_ismpc_b200", "lib", "libismpc_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
fams = ["UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "DMMA", "MUFU.RCP64H", "REDUX", "SHFL", "LDS", "STS", "LDG", "STG", "LDL", "STL",
        "BAR", "ATOMG", "RED", "HMMA", "UTCMMA", "CALL"]
print("library: %s\narchitectures in the fatbin: %s\n" % (os.path.basename(lib), ", ".join(arch)))
print("%-58s %7s " % ("kernel", "instr") + " ".join("%6s" % f[:6] for f in fams))
tot = collections.Counter()
for part in re.split(r"\n\s+Function : ", txt)[1:]:
    name = part.split("\n", 1)[0].strip()
    ops = collections.Counter()
    n = 0
    for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", part):
        op = m.group(1)
        n += 1
        for f in fams:
            if op == f or op.startswith(f + "."):
                ops[f] += 1
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = re.sub(r"\(.*", "", dem).replace("ismpc::", "")
    print("%-58s %7d " % (dem[:58], n) + " ".join("%6d" % ops[f] for f in fams))
    tot.update(ops); tot["_n"] += n
print("%-58s %7d " % ("TOTAL", tot["_n"]) + " ".join("%6d" % tot[f] for f in fams))
print("\nprogrammatic dependent launch: PREEXIT (griddepcontrol.launch_dependents) x %d, ACQBULK (griddepcontrol.wait) x %d"
      % (len(re.findall(r"\bPREEXIT\b", txt)), len(re.findall(r"\bACQBULK\b", txt))))
print("no tensor-core MMA other than DMMA is expected: there is no low-precision contraction on this path (DESIGN.md section 2)")
