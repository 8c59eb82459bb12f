#!/bin/bash
# One GPU box, one call: the parity suite, the bench lines and the ncu evidence of a round.  Run through tools/gpu.sh:
#   tools/gpu.sh --timeout 1500 -- 'bash tools/profile_round.sh r2f'
# Everything lands in gpurun_out/<tag>_*; tools/profile_collect.sh <tag> then writes the summaries into profiles/.
# (ncu runs come after the un-profiled run of the same command; numbers printed under ncu are never bench values.)
tag=${1:-rX}
o=gpurun_out
python -m pytest tests -q -m gpu > $o/${tag}_tests.log 2>&1; tail -2 $o/${tag}_tests.log
python __graft_entry__.py smoke > $o/${tag}_smoke.log 2>&1; tail -1 $o/${tag}_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference.json 2> $o/${tag}_bench_reference.err
python bench.py --steps 20 --warmup 5 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; tail -2 $o/${tag}_bench.err
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $o/${tag}_b_small.json 2> $o/${tag}_b_small.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $o/${tag}_ncu_b.log 2>&1
ISMPC_VARIANT=16 python tools/prof_tick.py formc 1024 3 > $o/${tag}_pt16.log 2>&1 && \
ISMPC_VARIANT=16 ncu --set full --clock-control none --import-source on -k regex:formc_tick_warp_kernel -s 20 -c 1 -o $o/${tag}_formc_warp16 \
    python tools/prof_tick.py formc 1024 3 > $o/${tag}_ncu1.log 2>&1
ISMPC_VARIANT=16 python tools/prof_tick.py formcp 1024 3 > $o/${tag}_ptp.log 2>&1 && \
ISMPC_VARIANT=16 ncu --set full --clock-control none --import-source on -k regex:formc_tick_warp_kernel -s 20 -c 1 -o $o/${tag}_formc_warp16_packed \
    python tools/prof_tick.py formcp 1024 3 > $o/${tag}_ncu2.log 2>&1
python tools/prof_tick.py formc 1024 3 > $o/${tag}_pt2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:formc_tick_pair_kernel -s 20 -c 1 -o $o/${tag}_formc_pair \
    python tools/prof_tick.py formc 1024 3 > $o/${tag}_ncu3.log 2>&1
python tools/prof_tick.py forma 1024 4 > $o/${tag}_pta.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:forma_tick_kernel -s 2 -c 1 -o $o/${tag}_forma_tick \
    python tools/prof_tick.py forma 1024 4 > $o/${tag}_ncu4.log 2>&1
python tools/prof_tick.py forma 8192 3 >> $o/${tag}_pta.log 2>&1
cat $o/${tag}_pt16.log $o/${tag}_ptp.log $o/${tag}_pt2.log $o/${tag}_pta.log | grep -v "^$" | tail -20
