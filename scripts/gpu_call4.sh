#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-r2d}
O=gpurun_out; mkdir -p $O
for v in head nopack f16; do
  if [ $v = head ]; then L=; else L=$PWD/build/libs/lib_$v.so; fi
  RTGRFF_LIB=$L timeout 300 python scripts/gpu_probe.py c4all > $O/${T}_c4all_$v.log 2>&1
done
RTGRFF_TILE=4x8 timeout 300 python scripts/gpu_probe.py c4freq > $O/${T}_probe_head.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
du -sh $O
