#!/bin/bash
# LBS kernel variants (HP3D_LBS_MODE) timed alone + GPU tests of the SMPL / sampler paths + sampler occupancy capture.
TAG=${1:-r01x}
OUT=gpurun_out; mkdir -p $OUT
for m in 0 1 2 3 0 1 2 3; do HP3D_LBS_MODE=$m python tools/bench_lbs.py; done 2>&1 | grep -v Warning | tee $OUT/${TAG}_lbs_sweep.jsonl
for m in 1 2; do HP3D_LBS_MODE=$m timeout 300 python -m pytest tests/test_gpu_smpl.py -m gpu -x -q 2>&1 | tail -2; done
timeout 300 python -m pytest tests/test_gpu_sampler.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/sampler_stress.py 2>&1 | tail -4 | tee $OUT/${TAG}_sampler_stress.jsonl
timeout 300 ncu --set full --clock-control none -k regex:'mf_sample' -c 3 -o $OUT/${TAG}_sampler python tools/sampler_stress.py > $OUT/${TAG}_ncu_sampler.log 2>&1
ncu -i $OUT/${TAG}_sampler.ncu-rep --page raw --csv > $OUT/${TAG}_sampler_raw.csv 2>/dev/null
