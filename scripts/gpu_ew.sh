#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== config 3 4 5 (defaults)"; timeout 900 python scripts/config_bench.py 3 4 5 2>&1 | grep config | tee gpurun_out/configs_$1.jsonl
echo "== config 5 gather"; LMC_EWALD_FIELD=0 timeout 900 python scripts/config_bench.py 5 2>&1 | tail -1
echo "== config 5 field"; LMC_EWALD_FIELD=1 timeout 900 python scripts/config_bench.py 5 2>&1 | tail -1
} > gpurun_out/ew_$1.log 2>&1
cat gpurun_out/ew_$1.log
