#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench value', d['value'], 'e2e', d['e2e'])"
timeout 300 python scripts/e2e_profile.py 2>&1 | tail -25
} > gpurun_out/t_$1.log 2>&1
cat gpurun_out/t_$1.log
