#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-r2l}
O=gpurun_out; mkdir -p $O
for th in 0 8 16; do
  RTGRFF_HOST_THREADS=$th timeout 300 python scripts/stage_bench.py --quick > $O/${T}_stage_t$th.json 2> $O/${T}_stage_t$th.err
done
