#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list of the same bench command, one
# ncu --set full capture of the dominant kernel.  Usage: gpurun -- 'bash scripts/gpu_round.sh <tag>'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $OUT/${TAG}_bench_line.json 2> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench_line.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference_line.json 2>> $OUT/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/${TAG}_launches_bench_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/${TAG}_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lmc_spec_kernel -s 3 -c 1 \
  -f -o $OUT/${TAG}_spec_cfg2 python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_full.log 2>&1
# other BASELINE configs (kernel-only) and the ncu capture of the Wang-Landau kernel (config 4)
timeout 300 python scripts/config_bench.py 3 4 5 > $OUT/${TAG}_configs_3_4_5.jsonl 2> $OUT/${TAG}_configs.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lmc_run_kernel -s 2 -c 1 \
  -f -o $OUT/${TAG}_cfg4 python scripts/config_bench.py 4 > $OUT/${TAG}_cfg4_ncu.log 2>&1
ls -la $OUT | tail -20
#   ncu --set full --clock-control none --import-source on -k regex:lmc_run_kernel -s 2 -c 1 -f -o gpurun_out/cfg5 python scripts/config_bench.py 5
