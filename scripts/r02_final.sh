#!/bin/bash
# final pass of the round: whole GPU suite, smoke, one bench line per config, launch list of the default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log; tail -4 gpurun_out/r02f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for c in 2 3 4 5; do
  timeout 900 python bench.py --config $c > gpurun_out/r02f_bench_cfg$c.json 2> gpurun_out/r02f_bench_cfg$c.err
  echo "bench cfg$c rc=$?"
done
timeout 600 python bench.py --impl reference > gpurun_out/r02f_bench_reference.json 2>/dev/null; echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_bench_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-curve --sustain-seconds 0 > gpurun_out/r02f_launches.log 2>&1; echo "launch list rc=$?"
