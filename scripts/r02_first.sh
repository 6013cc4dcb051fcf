#!/bin/bash
# round 2, first GPU pass: the whole GPU suite, then one bench line per BASELINE config
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
for c in 2 3 4 5; do
  timeout 900 python bench.py --config $c --cpu-seconds 6 > gpurun_out/r02a_bench_cfg$c.json 2> gpurun_out/r02a_bench_cfg$c.err
  echo "bench cfg$c rc=$?"
  tail -c 600 gpurun_out/r02a_bench_cfg$c.json
  tail -3 gpurun_out/r02a_bench_cfg$c.err
done
