#!/bin/bash
# A/B of the working tree's library against build/libs/lib_<variant>.so on the whole bench call (development aid)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-ab}; shift
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/gpu_probe.py c4all > $O/${T}_c4all_head.log 2>&1
for v in "$@"; do RTGRFF_LIB=$PWD/build/libs/lib_$v.so timeout 300 python scripts/gpu_probe.py c4all > $O/${T}_c4all_$v.log 2>&1; done
