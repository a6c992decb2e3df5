#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-ab}; shift
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/gpu_probe.py c4all > $O/${T}_c4all_head.log 2>&1
for v in "$@"; do RTGRFF_LIB=$PWD/build/libs/lib_$v.so timeout 300 python scripts/gpu_probe.py c4all > $O/${T}_c4all_$v.log 2>&1; done
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
