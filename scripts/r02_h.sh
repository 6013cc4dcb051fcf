#!/bin/bash
# convergence run of config 4, bench lines of all configs, ncu captures of the dominant kernels
mkdir -p gpurun_out
python scripts/wl_convergence.py 8e7 1048576 > gpurun_out/r02h_cfg4_convergence.json 2> gpurun_out/r02h_conv.err
tail -25 gpurun_out/r02h_cfg4_convergence.json
for c in 2 3 4 5; do
  timeout 900 python bench.py --config $c --cpu-seconds 8 > gpurun_out/r02h_bench_cfg$c.json 2> gpurun_out/r02h_bench_cfg$c.err
  echo "bench cfg$c rc=$?"; tail -c 300 gpurun_out/r02h_bench_cfg$c.json; echo
done
# ncu --set full of the kernel each config spends its time in (the last of 6 launches: equilibrated walkers)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lmc_spec --launch-skip 5 --launch-count 1 -o gpurun_out/r02h_cfg2 -f python scripts/prof_cfg.py 2 8 6 > gpurun_out/r02h_ncu2.log 2>&1; tail -1 gpurun_out/r02h_ncu2.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lmc_spec --launch-skip 5 --launch-count 1 -o gpurun_out/r02h_cfg3 -f python scripts/prof_cfg.py 3 8 6 > gpurun_out/r02h_ncu3.log 2>&1; tail -1 gpurun_out/r02h_ncu3.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lmc_wl3 --launch-skip 2 --launch-count 1 -o gpurun_out/r02h_cfg4 -f python scripts/prof_cfg.py 4 10 3 > gpurun_out/r02h_ncu4.log 2>&1; tail -1 gpurun_out/r02h_ncu4.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:lmc_run_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/r02h_cfg5 -f python scripts/prof_cfg.py 5 1 5 > gpurun_out/r02h_ncu5.log 2>&1; tail -1 gpurun_out/r02h_ncu5.log
