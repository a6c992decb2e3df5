#!/bin/bash
# GPU call: tests, microbench, A/B probes, bench, ncu.  Outputs in gpurun_out/ (kept under 64 MiB: the merge limit).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-r2b}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/${T}_gpu.txt 2>&1
nproc >> $O/${T}_gpu.txt; free -g | head -2 >> $O/${T}_gpu.txt
./build/microbench > $O/${T}_microbench.json 2> $O/${T}_microbench.err
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
for v in head minb10 minb12 b128; do
  if [ $v = head ]; then L=; else L=$PWD/build/libs/lib_$v.so; fi
  RTGRFF_LIB=$L RTGRFF_TILE=4x8 timeout 300 python scripts/gpu_probe.py c4freq > $O/${T}_probe_$v.log 2>&1
done
RTGRFF_GRFF64=1 RTGRFF_TILE=4x8 timeout 300 python scripts/gpu_probe.py c4freq > $O/${T}_probe_grff64.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "rc=$?" >> $O/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
timeout 600 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/${T}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/${T}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_map -s 3 -c 1 -o $O/${T}_render_map \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/${T}_ncu_full.log 2>&1
timeout 300 python scripts/stage_bench.py --quick > $O/${T}_stage.json 2> $O/${T}_stage.err
for c in c1 c2 c3; do
  case $c in c1) K=sample_paths;; c2) K=grff_slice;; c3) K="emission_rays|trace_rays";; esac
  timeout 200 python scripts/ncu_stage.py $c > $O/${T}_ncu_$c.plain 2>&1 && \
  timeout 600 ncu --set full --clock-control none -k "regex:$K" -s 1 -c 2 -o $O/${T}_stage_$c \
      python scripts/ncu_stage.py $c > $O/${T}_ncu_$c.log 2>&1
  ncu -i $O/${T}_stage_$c.ncu-rep --page raw --csv > $O/${T}_stage_$c.raw.csv 2>/dev/null && rm -f $O/${T}_stage_$c.ncu-rep
done
find $O -type f -size +40M -delete
du -sh $O; ls -la $O | tail -40
