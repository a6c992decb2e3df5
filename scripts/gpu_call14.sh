#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-r2n}
O=gpurun_out; mkdir -p $O
for n in 1 2 3 4 8; do RTGRFF_FORK_STREAMS=$n timeout 300 python scripts/gpu_probe.py c4all > $O/${T}_c4all_fork$n.log 2>&1; done
