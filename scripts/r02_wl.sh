#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "wang or config4 or multistep or edge" > gpurun_out/r02c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -25 gpurun_out/r02c_pytest.log
for v in 0 1 3; do
  LMC_WL2=$v timeout 300 python bench.py --config 4 --no-cpu --steps 6 2> gpurun_out/r02c_cfg4_wl$v.err | tee gpurun_out/r02c_cfg4_wl$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('WL2=$v', d['value'], d['ms_per_step'], d['config']['acceptance_ratio'], d['config']['wang_landau'])"
  tail -2 gpurun_out/r02c_cfg4_wl$v.err
done
