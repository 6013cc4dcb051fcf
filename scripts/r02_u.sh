#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "speculative or ewald or config2 or config3" > gpurun_out/r02u_pytest.log 2>&1; tail -3 gpurun_out/r02u_pytest.log
for e in 1 0; do
echo "ENV=$e cfg2"; LMC_SPEC_ENV=$e python scripts/prof_cfg.py 2 8 5
echo "ENV=$e cfg3"; LMC_SPEC_ENV=$e python scripts/prof_cfg.py 3 8 5
done
LMC_SPEC_ENV=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:lmc_spec_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r02v_cfg2_env -f python scripts/prof_cfg.py 2 8 5 > gpurun_out/r02v_ncu2.log 2>&1; tail -2 gpurun_out/r02v_ncu2.log
