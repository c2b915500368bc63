#!/bin/bash
# NCCL vertices gather (own communicator) at NG GPUs: gather granularity, NVLS on/off, channel counts. ms per step.
NG=${NG:-4}; OUT=gpurun_out; mkdir -p $OUT; port=29900
run() { port=$((port+1)); env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $port bench.py --gpus $NG --steps 10 --warmup 3 --profile --transport nccl $EXTRA 2>&1 | grep profile_ms | tail -1 | sed "s/^/[$* $EXTRA] /"; }
{
  EXTRA="" run A=0
  EXTRA="--vchunks 1" run A=0
  EXTRA="--vchunks 2" run A=0
  EXTRA="--vchunks 8" run A=0
  EXTRA="" run NCCL_NVLS_ENABLE=0
  EXTRA="" run NCCL_MAX_CTAS=8
  EXTRA="" run NCCL_MIN_CTAS=32
} | tee $OUT/${1:-r02}_${NG}gpu_sweep2.txt
