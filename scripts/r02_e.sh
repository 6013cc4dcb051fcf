#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
tail -6 gpurun_out/r02e_pytest.log
for v in 0 1 3; do LMC_WL2=$v python scripts/prof_cfg.py 4 40 3; done
for c in 3 5; do python scripts/e2e_breakdown.py $c 2>&1 | tail -1 | tee -a gpurun_out/r02e_breakdown.jsonl; done
