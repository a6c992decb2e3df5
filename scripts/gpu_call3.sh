#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-r2c}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
for v in head f1 minb9 f1minb9; do
  if [ $v = head ]; then L=; else L=$PWD/build/libs/lib_$v.so; fi
  RTGRFF_LIB=$L timeout 300 python scripts/gpu_probe.py c4all > $O/${T}_c4all_$v.log 2>&1
  RTGRFF_LIB=$L RTGRFF_TILE=4x8 timeout 300 python scripts/gpu_probe.py c4freq > $O/${T}_probe_$v.log 2>&1
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "rc=$?" >> $O/${T}_bench.err
du -sh $O
