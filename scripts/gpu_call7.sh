#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-r2g}
O=gpurun_out; mkdir -p $O
RTGRFF_CARVEOUT=-1 timeout 300 python scripts/gpu_probe.py c4all > $O/${T}_c4all_nocarve.log 2>&1
for v in head prefetch unrollq b32; do
  if [ $v = head ]; then L=; else L=$PWD/build/libs/lib_$v.so; fi
  RTGRFF_LIB=$L timeout 300 python scripts/gpu_probe.py c4all > $O/${T}_c4all_$v.log 2>&1
done
du -sh $O
