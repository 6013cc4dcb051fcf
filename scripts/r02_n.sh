#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "edge_cases" > gpurun_out/r02n_pytest.log 2>&1; tail -30 gpurun_out/r02n_pytest.log
