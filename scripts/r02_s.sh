#!/bin/bash
mkdir -p gpurun_out
LMC_COMPACT_TABLES_OFF=1 timeout 600 python -m pytest tests -m gpu -q -x -k "canonical_ewald_swap" > gpurun_out/r02s_off.log 2>&1; tail -3 gpurun_out/r02s_off.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k "canonical_ewald_swap and classic" > gpurun_out/r02s_san.log 2>&1; grep -m1 -A25 "Invalid\|misaligned\|ERROR" gpurun_out/r02s_san.log | head -60
