#!/bin/bash
# build everything in-tree, then run a command on a GPU box:  tools/gpu.sh [--gpus N] [--timeout S] -- 'cmd'
set -e
cd "$(dirname "$0")/.."
make -s -C quadruped_gait_generation_ismpc_b200/csrc -j8 2>&1 | grep -E "error|warning: v|Error" || true
make -s -C oracle -j8 port ref 2>&1 | tail -2 || true
exec /usr/local/graft/bin/gpurun "$@"
