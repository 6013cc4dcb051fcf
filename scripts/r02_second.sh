#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -8 gpurun_out/r02b_pytest.log
for c in 2 3 5; do python scripts/e2e_breakdown.py $c 2>&1 | tail -1 | tee -a gpurun_out/r02b_breakdown.jsonl; done
