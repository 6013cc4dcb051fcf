#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "table_flip or config5 or tableflip" > gpurun_out/r02x_pytest.log 2>&1; tail -15 gpurun_out/r02x_pytest.log
echo "TF spec"; timeout 600 python scripts/prof_cfg.py 5 1 6
echo "TF classic"; LMC_SPEC_TF=0 timeout 600 python scripts/prof_cfg.py 5 1 6
