#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02t_pytest.log 2>&1; tail -3 gpurun_out/r02t_pytest.log
python scripts/prof_cfg.py 5 1 5
LMC_COMPACT_TABLES_OFF=1 python scripts/prof_cfg.py 5 1 5
python scripts/prof_cfg.py 3 8 5
LMC_COMPACT_TABLES_OFF=1 python scripts/prof_cfg.py 3 8 5
