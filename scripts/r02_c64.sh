#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "speculative or config2 or smoke or canonical_swap or api" 2>&1 | tail -3
echo "c64"; python scripts/prof_cfg.py 2 8 5
echo "gather"; LMC_SPEC_C64=0 python scripts/prof_cfg.py 2 8 5
bash scripts/ncu_cfg.sh 2 lmc_spec_c64_kernel 3 8 r02_c64_cfg2
