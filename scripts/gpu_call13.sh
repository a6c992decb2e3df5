#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-r2m}
O=gpurun_out; mkdir -p $O
RTGRFF_CELL_CUBE=0 timeout 300 python scripts/gpu_probe.py c4all > $O/${T}_c4all_nodecube.log 2>&1
timeout 300 python scripts/gpu_probe.py c4all > $O/${T}_c4all_cellcube.log 2>&1
RTGRFF_CELL_CUBE=0 timeout 300 python scripts/stage_bench.py --quick > $O/${T}_stage_nodecube.json 2> $O/${T}_stage_nodecube.err
timeout 300 python scripts/stage_bench.py --quick > $O/${T}_stage_cellcube.json 2> $O/${T}_stage_cellcube.err
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
