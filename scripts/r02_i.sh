#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -15 gpurun_out/r02i_pytest.log
