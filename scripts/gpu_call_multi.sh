#!/bin/bash
# multi-GPU validation: the driver's launch line for N ranks, CUDA arm then reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}; T=${2:-r2e}; STEPS=${3:-3}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/${T}_gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps $STEPS --warmup 3 > $O/${T}_bench_${N}gpu.json 2> $O/${T}_bench_${N}gpu.err; echo "rc=$?" >> $O/${T}_bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/${T}_ref_${N}gpu.json 2> $O/${T}_ref_${N}gpu.err; echo "rc=$?" >> $O/${T}_ref_${N}gpu.err
tail -c 600 $O/${T}_bench_${N}gpu.err; du -sh $O
