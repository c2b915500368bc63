#!/bin/bash
# Multi-GPU experiments that were still open at the end of round 1 (configs[3] full vertices gather), NG GPUs (default 4):
#   NCCL gather with the persistent kernels leaving SMs free (--sm-limit) and NCCL's CTAs UNCAPPED -- never measured cleanly
#   (the one round-1 run had NCCL_MAX_CTAS=16 set by mistake); copy-engine pushes; statistics-only gather for reference.
# usage (on the GPU box): NG=4 bash tools/gpu_multi_gpu_sweep.sh <tag>
NG=${NG:-4}; OUT=gpurun_out; mkdir -p $OUT; port=29800
run() { port=$((port+1)); HP3D_NCCL_CTAS=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $port bench.py --gpus $NG --steps 10 --warmup 3 --profile "$@" 2>/dev/null | tail -1 | sed "s/^/[$*] /"; }
{
  run --transport nccl --sm-limit 0
  run --transport nccl --sm-limit 132
  run --transport nccl --sm-limit 116
  run --transport p2p
  run --gather stats
} | tee $OUT/${1:-r02}_${NG}gpu_sweep.txt
