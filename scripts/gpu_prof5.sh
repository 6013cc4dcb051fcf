#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lmc_run_kernel -s 2 -c 1 -f -o gpurun_out/r01c_cfg5 python scripts/config_bench.py 5 > gpurun_out/r01c_cfg5_ncu.log 2>&1
tail -3 gpurun_out/r01c_cfg5_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lmc_run_kernel -s 2 -c 1 -f -o gpurun_out/r01c_cfg4 python scripts/config_bench.py 4 > gpurun_out/r01c_cfg4_ncu.log 2>&1
tail -3 gpurun_out/r01c_cfg4_ncu.log
ls -la gpurun_out/*.ncu-rep
