#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02w_pytest.log 2>&1; tail -3 gpurun_out/r02w_pytest.log
for e in 0 1; do
echo "ENV_DEFAULT=$e cfg2"; LMC_SPEC_ENV_DEFAULT=$e python scripts/prof_cfg.py 2 8 5
echo "ENV_DEFAULT=$e cfg3"; LMC_SPEC_ENV_DEFAULT=$e python scripts/prof_cfg.py 3 8 5
done
