#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "api or table_flip or config5" 2>&1 | tail -3
python scripts/prof_cfg.py 5 1 5
python bench.py --config 2 --no-cpu --no-curve --sustain-seconds 0 | python -c "import sys,json;d=json.loads(sys.stdin.read());print('cfg2', d['value'], d['e2e']['value'], d['e2e']['value']/d['value'])"
python bench.py --config 4 --no-cpu --sustain-seconds 0 | python -c "import sys,json;d=json.loads(sys.stdin.read());print('cfg4', d['value'], d['e2e']['value'], d['e2e']['value']/d['value'])"
