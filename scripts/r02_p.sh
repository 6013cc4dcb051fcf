#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "table_flip or config5 or tableflip or anneal" 2>&1 | tail -3
python scripts/prof_cfg.py 5 1 5
