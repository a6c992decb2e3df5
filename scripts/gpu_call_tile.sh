#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-r2s}
O=gpurun_out; mkdir -p $O
for t in 4x8 8x4 2x16 16x2 32x1 1x32 8x8 4x16; do RTGRFF_TILE=$t timeout 300 python scripts/gpu_probe.py c4all > $O/${T}_c4all_tile$t.log 2>&1; done
