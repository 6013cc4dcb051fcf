#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== config 5 gather / field (templated EWMODE)"
LMC_EWALD_FIELD=0 timeout 900 python scripts/config_bench.py 5 2>&1 | tail -1
LMC_EWALD_FIELD=1 timeout 900 python scripts/config_bench.py 5 2>&1 | tail -1
echo "== config 2 hot classic (T=5000)"
GS=32 TEMP=5000 SWEEPS=20 timeout 300 python scripts/perf_probe.py 2>&1 | tail -1
} > gpurun_out/t_$1.log 2>&1
cat gpurun_out/t_$1.log
