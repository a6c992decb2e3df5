#!/bin/bash
# GPU call 1 of round 2: tests, microbench, A/B probes, bench, ncu.  Everything lands in gpurun_out/.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/r2a_gpu.txt 2>&1
nproc >> $O/r2a_gpu.txt; free -g | head -2 >> $O/r2a_gpu.txt
./build/microbench > $O/r2a_microbench.json 2> $O/r2a_microbench.err
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2a_pytest.log
for v in head minb10 minb12 b128; do
  if [ $v = head ]; then L=; else L=$PWD/build/libs/lib_$v.so; fi
  RTGRFF_LIB=$L RTGRFF_TILE=4x8 timeout 300 python scripts/gpu_probe.py c4freq > $O/r2a_probe_$v.log 2>&1
done
RTGRFF_GRFF64=1 RTGRFF_TILE=4x8 timeout 300 python scripts/gpu_probe.py c4freq > $O/r2a_probe_grff64.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "rc=$?" >> $O/r2a_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2a_bench_ref.json 2> $O/r2a_bench_ref.err
# ncu: launch list of the bench command, then one full capture of the map kernel and of the stage kernels
timeout 600 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/r2a_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2a_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/r2a_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_map -s 3 -c 1 -o $O/r2a_render_map \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/r2a_ncu_full.log 2>&1
timeout 300 python scripts/stage_bench.py --quick > $O/r2a_stage.json 2> $O/r2a_stage.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:sample_paths|grff_slice|emission_rays" -c 40 -o $O/r2a_stage_kernels \
    python scripts/stage_bench.py --quick > $O/r2a_ncu_stage.log 2>&1
ls -la $O | tail -30
