#!/bin/bash
mkdir -p gpurun_out
LMC_SPEC_ENV=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:lmc_spec_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r02v_cfg2_env -f python scripts/prof_cfg.py 2 8 5 > gpurun_out/r02v_ncu2.log 2>&1; tail -2 gpurun_out/r02v_ncu2.log
