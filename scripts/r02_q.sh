#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:lmc_run_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/r02q_cfg5 -f python scripts/prof_cfg.py 5 1 5 > gpurun_out/r02q_ncu5.log 2>&1; tail -1 gpurun_out/r02q_ncu5.log
