#!/bin/bash
for g in 0 8 16; do echo "LMC_GROUP_SIZE=$g"; LMC_GROUP_SIZE=$g python scripts/prof_cfg.py 5 1 5 2>&1 | tail -1; done
