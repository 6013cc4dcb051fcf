#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "speculative" 2>&1 | tail -3
for l in 0 1; do for t in 1000 1500 2000; do
echo "== LISTS=$l T=$t"; LMC_SPEC_LISTS=$l GS=-1 TEMP=$t SWEEPS=40 timeout 300 python scripts/perf_probe.py 2>&1 | tail -1
done; done
timeout 300 python bench.py --no-cpu --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench value', d['value'], 'e2e', d['e2e']['value'])"
} > gpurun_out/t_$1.log 2>&1
cat gpurun_out/t_$1.log
