#!/bin/bash
# table-flip kernel: parity tests, equilibrated sweep time, one ncu capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "table_flip or config5" 2>&1 | tail -2
python scripts/prof_cfg.py 5 1 25
bash scripts/ncu_cfg.sh 5 lmc_spec_tf_kernel 20 1 r02y_cfg5_tf
