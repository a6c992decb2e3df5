#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-r2f}
O=gpurun_out; mkdir -p $O
for v in head minb7 minb6 minb5; do
  if [ $v = head ]; then L=; else L=$PWD/build/libs/lib_$v.so; fi
  RTGRFF_LIB=$L timeout 300 python scripts/gpu_probe.py c4all > $O/${T}_c4all_$v.log 2>&1
done
du -sh $O
