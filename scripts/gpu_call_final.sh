#!/bin/bash
# Final single-GPU evidence run of a round: tests, bench (both arms), launch list, ncu captures.  Outputs < 64 MiB.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-r2h}
O=gpurun_out; mkdir -p $O
M="smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread"
./build/microbench > $O/${T}_microbench.json 2> $O/${T}_microbench.err
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "rc=$?" >> $O/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
B="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline"
timeout 600 $B > $O/${T}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/${T}_launches.csv $B > $O/${T}_ncu_launches.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:render_map -s 24 -c 8 --csv --log-file $O/${T}_c4_counters.csv $B > $O/${T}_ncu_c4_counters.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_map -s 24 -c 1 -o $O/${T}_render_map_1500MHz $B > $O/${T}_ncu_full.log 2>&1
B5="python bench.py --config c5 --steps 1 --warmup 3 --no-extras --no-cpu-baseline"
timeout 600 $B5 > $O/${T}_plain_c5.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none -k regex:render_map -s 48 -c 16 --csv --log-file $O/${T}_c5_counters.csv $B5 > $O/${T}_ncu_c5_counters.log 2>&1
find $O -type f -size +45M -delete
du -sh $O; ls -la $O | tail -30
