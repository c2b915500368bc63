#!/bin/bash
# One gpurun call: GPU tests, bench line (+ reference arm), ncu launch list, ncu --set full of the fused SMPL kernel (traffic).
# Usage (from the repo root, on the GPU box): bash tools/gpu_round.sh <tag>
TAG=${1:-r02x}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-300 $OUT/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --profile --steps 1 > $OUT/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'smpl_fused_kernel' -c 2 -o $OUT/${TAG}_full python bench.py --profile --steps 1 > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/${TAG}_full_raw.csv $OUT/${TAG}_ncu_full_summary.csv
python tools/launch_breakdown.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_breakdown.md
ls -la $OUT
