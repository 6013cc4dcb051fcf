#!/bin/bash
# last pass of the round: whole GPU suite, smoke, the default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log; tail -3 gpurun_out/r02f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_cfg2.json 2> gpurun_out/r02f_bench_cfg2.err; echo "bench rc=$?"
