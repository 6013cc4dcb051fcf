#!/bin/bash
# quick A/B of the speculative kernel on config 2: lanes per step, merge level
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k speculative 2>&1 | tail -3
for sg in 4 1; do for mg in 2; do
  echo "== LMC_SPEC_SG=$sg LMC_SPEC_MERGE=$mg"
  LMC_SPEC_SG=$sg LMC_SPEC_MERGE=$mg GS=-1 SWEEPS=40 timeout 300 python scripts/perf_probe.py 2>&1 | tail -1
done; done
echo "== hot (T=5000) sg4 / sg2 / classic"
LMC_SPEC_SG=4 GS=-1 TEMP=5000 SWEEPS=20 timeout 300 python scripts/perf_probe.py 2>&1 | tail -1
LMC_SPEC_SG=1 GS=-1 TEMP=5000 SWEEPS=20 timeout 300 python scripts/perf_probe.py 2>&1 | tail -1
GS=32 TEMP=5000 SWEEPS=20 timeout 300 python scripts/perf_probe.py 2>&1 | tail -1
} > gpurun_out/ab_$1.log 2>&1
cat gpurun_out/ab_$1.log
{
echo "== T=2000 / 3000 sg 1 vs 4"
for t in 1500 2000 3000; do for sg in 1 4; do
LMC_SPEC_SG=$sg GS=-1 TEMP=$t SWEEPS=20 timeout 300 python scripts/perf_probe.py 2>&1 | tail -1
done; done
} >> gpurun_out/ab_$1.log 2>&1
tail -8 gpurun_out/ab_$1.log
