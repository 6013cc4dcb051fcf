#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "wang or config4" > gpurun_out/r02g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log
tail -6 gpurun_out/r02g_pytest.log
for v in 0 1 4; do LMC_WL2=$v python scripts/prof_cfg.py 4 40 3; done
LMC_WL2=4 timeout 600 ncu --set full --import-source on --clock-control none -k regex:lmc_wl3 --launch-skip 2 --launch-count 1 -o gpurun_out/r02g_wl3 -f python scripts/prof_cfg.py 4 4 3 > gpurun_out/r02g_ncu.log 2>&1
tail -1 gpurun_out/r02g_ncu.log
