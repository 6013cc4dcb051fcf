#!/bin/bash
mkdir -p gpurun_out
for v in 1 3; do
  LMC_WL2=$v timeout 600 ncu --set full --import-source on --clock-control none -k regex:lmc_wl2 --launch-skip 2 --launch-count 1 -o gpurun_out/r02f_wl2_ne$v -f python scripts/prof_cfg.py 4 4 3 > gpurun_out/r02f_ncu_ne$v.log 2>&1
  tail -1 gpurun_out/r02f_ncu_ne$v.log
done
timeout 600 python -m pytest tests -m gpu -q -x -k "ewald or config5 or config3" 2>&1 | tail -3
