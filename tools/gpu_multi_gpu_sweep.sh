#!/bin/bash
# Multi-GPU transport experiments for the configs[3] full vertices gather, NG GPUs (default 4): ms per step of
#   the NVLS multicast-store kernel (hp3d_peer_push_multicast), the unicast peer-store kernel, copy-engine pushes and NCCL's
#   all-gather (all on a communicator of their own); statistics-only gather for reference (the no-communication floor).
# usage (on the GPU box): NG=4 bash tools/gpu_multi_gpu_sweep.sh <tag>
NG=${NG:-4}; OUT=gpurun_out; mkdir -p $OUT; port=29800
run() { port=$((port+1)); HP3D_NCCL_CTAS=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $port bench.py --gpus $NG --steps 10 --warmup 3 --profile "$@" 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$\|NCCL version" | tail -2 | sed "s/^/[$*] /"; }
{
  run --transport p2p --push mc
  run --transport p2p --push mc --push-ctas 74
  run --transport p2p --push kernel
  run --transport p2p --push ce
  run --transport nccl
  run --transport nccl --sm-limit 132
  run --gather stats
} | tee $OUT/${1:-r02}_${NG}gpu_sweep.txt
