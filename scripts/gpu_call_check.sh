#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-r2j}
O=gpurun_out; mkdir -p $O
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $O/${T}_smoke.log 2>&1; echo "rc=$?" >> $O/${T}_smoke.log
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "rc=$?" >> $O/${T}_bench.err
( time python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
for i in 1 2; do timeout 300 python scripts/stage_bench.py --quick > $O/${T}_stage$i.json 2> $O/${T}_stage$i.err; done
du -sh $O
