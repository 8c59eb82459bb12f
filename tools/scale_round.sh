#!/bin/bash
# The contract's scaling run on one 8-GPU box: bench.py at N = 1, 2, 4, 8 (torchrun for N > 1) + the multi-GPU tests.
#   tools/gpu.sh --gpus 8 --timeout 900 -- 'bash tools/scale_round.sh r2f'
tag=${1:-rX}; o=gpurun_out; extra=${2:---no-extras}
python -m pytest tests/test_multigpu.py -q -m gpu > $o/${tag}_mgtests.log 2>&1; tail -2 $o/${tag}_mgtests.log
for N in 1 2 4 8; do
  if [ $N = 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 5 $extra > $o/${tag}_bench_${N}gpu.json 2> $o/${tag}_bench_${N}gpu.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 $extra > $o/${tag}_bench_${N}gpu.json 2> $o/${tag}_bench_${N}gpu.err; fi
done
