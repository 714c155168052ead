#!/bin/bash
# Multi-GPU visit: bench.py under torchrun with the slab + c4 legs, the slab cross-check, and the gloo-free NCCL tests.
#   gpurun --gpus N -- bash scripts/gpu_multi.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
echo "== bench c3 + slab + c4, N=$N"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3 | tee gpurun_out/bench_n$N.json | cut -c1-3000
echo "== slab_check N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    scripts/slab_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -12 | tee gpurun_out/slab_check_n$N.txt
