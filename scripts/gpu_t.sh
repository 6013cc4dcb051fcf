#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python scripts/config_bench.py 3 4 5 2>&1 | grep config | tee gpurun_out/configs_$1.jsonl
} > gpurun_out/t_$1.log 2>&1
cat gpurun_out/t_$1.log
