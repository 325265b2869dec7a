#!/bin/bash
# multi-GPU bench exactly as the driver launches it (+ the NCCL variant of the collective for comparison)
N=${1:-2}
OUT=gpurun_out/multi$N
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv | tee $OUT/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep -v "^Grouping\|^W0\|^\*\*\*\|OMP_NUM" | tail -5 | tee $OUT/bench.log
DKS_BENCH_NCCL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep -v "^Grouping\|^W0\|^\*\*\*\|OMP_NUM" | tail -2 | tee $OUT/bench_nccl.log
if [ "${2:-}" = "ref" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -3 | tee $OUT/bench_ref.log
fi
