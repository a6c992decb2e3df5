#!/bin/bash
# What the driver does at round end on one GPU: the gpu tests, smoke(), the bench line at 20/5 with its wall time.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-r2z}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "rc=$?" >> $O/${T}_smoke.log
S=$(date +%s)
timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "rc=$? wall=$(( $(date +%s) - S )) s" >> $O/${T}_bench.err
tail -2 $O/${T}_pytest.log; tail -2 $O/${T}_smoke.log; tail -1 $O/${T}_bench.err; cut -c1-300 $O/${T}_bench.json
