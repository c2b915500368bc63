#!/bin/bash
# 2-GPU bench (configs[3] full vertices gather): copy-engine pushes over symmetric memory vs NCCL all-gather
OUT=gpurun_out; mkdir -p $OUT
port=29600
for tr in p2p nccl; do
  port=$((port+1))
  HP3D_NCCL_CTAS=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port $port bench.py --gpus ${NG:-2} --steps 10 --warmup 3 --profile --transport $tr 2>$OUT/sweep_$tr.err | tail -1 | sed "s/^/transport=$tr /"
  grep -v "^\*\|OMP_NUM" $OUT/sweep_$tr.err | tail -6
done | tee $OUT/${1:-r01v}_2gpu_sweep.txt
