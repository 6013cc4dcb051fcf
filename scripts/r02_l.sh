#!/bin/bash
for lib in liblmc.so liblmc_tf3.so liblmc_tf4.so; do echo $lib; LMC_LIBRARY=$PWD/smol_b200/_lib/$lib python scripts/prof_cfg.py 5 1 5 | tail -1; done
LMC_LIBRARY=$PWD/smol_b200/_lib/liblmc_tf3.so PROF_WALKERS=32768 python scripts/prof_cfg.py 5 1 3 | tail -1
PROF_WALKERS=32768 python scripts/prof_cfg.py 5 1 3 | tail -1
