#!/bin/bash
mkdir -p gpurun_out
echo "TF spec"; timeout 600 python scripts/prof_cfg.py 5 1 25
timeout 900 ncu --set full --import-source on --clock-control none -k regex:lmc_spec_tf_kernel --launch-skip 20 --launch-count 1 -o gpurun_out/r02y_cfg5_tf -f python scripts/prof_cfg.py 5 1 22 > gpurun_out/r02y_ncu5.log 2>&1; tail -2 gpurun_out/r02y_ncu5.log
timeout 1500 python -m pytest tests -m gpu -q -x -k "table_flip or config5" 2>&1 | tail -2
