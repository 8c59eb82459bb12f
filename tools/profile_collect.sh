#!/bin/bash
# Turns what tools/profile_round.sh left in gpurun_out/<tag>_* into the tracked summaries under profiles/ (run here, after the call).
tag=${1:-rX}
o=gpurun_out; p=profiles
cp $o/${tag}_bench.json $p/${tag}_bench.json
cp $o/${tag}_bench_reference.json $p/${tag}_bench_reference.json
cp $o/${tag}_launches_bench.csv $p/${tag}_launches_bench.csv
tail -3 $o/${tag}_tests.log > $p/${tag}_gpu_tests.txt
python tools/ncu_summary.py $o/${tag}_formc_warp16.ncu-rep formc_tick_warp_kernel 1024 $p/${tag}_formc_tick_warp16_ncu.json $p/${tag}_formc_tick_warp16_ncu_metrics.txt
python tools/ncu_lines.py $o/${tag}_formc_warp16.ncu-rep formc_tick_warp_kernel 40 > $p/${tag}_formc_tick_warp16_stall_lines.txt
python tools/ncu_summary.py $o/${tag}_formc_warp16_packed.ncu-rep formc_tick_warp_kernel 1024 $p/${tag}_formc_tick_warp16_packed_ncu.json $p/${tag}_formc_tick_warp16_packed_ncu_metrics.txt
python tools/ncu_summary.py $o/${tag}_formc_pair.ncu-rep formc_tick_pair_kernel 1024 $p/${tag}_formc_tick_pair_ncu.json $p/${tag}_formc_tick_pair_ncu_metrics.txt
python tools/ncu_lines.py $o/${tag}_formc_pair.ncu-rep formc_tick_pair_kernel 40 > $p/${tag}_formc_tick_pair_stall_lines.txt
python tools/ncu_summary.py $o/${tag}_forma_tick.ncu-rep forma_tick_kernel 1024 $p/${tag}_forma_tick_ncu.json $p/${tag}_forma_tick_ncu_metrics.txt
python tools/ncu_lines.py $o/${tag}_forma_tick.ncu-rep forma_tick_kernel 40 > $p/${tag}_forma_tick_stall_lines.txt
python tools/sass_summary.py > $p/${tag}_sass_summary.txt
ls -la $p/${tag}_*
