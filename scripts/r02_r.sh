#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "ewald or semigrand or table_flip or config5 or config3 or composite or multistep or biased" > gpurun_out/r02r_pytest.log 2>&1
head -60 gpurun_out/r02r_pytest.log
