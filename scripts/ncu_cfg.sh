#!/bin/bash
# One ncu --set full capture of the kernel a BASELINE config runs (gpurun -- 'bash scripts/ncu_cfg.sh 5 lmc_spec_tf_kernel 20 1'):
#   $1 config (2..5), $2 kernel name regex, $3 launches to skip (equilibration), $4 sampling intervals per launch, $5 tag
# then here: python scripts/ncu_summary.py gpurun_out/<tag>.ncu-rep <tag> <attempted steps per launch>
#            python scripts/ncu_lines.py gpurun_out/<tag>.ncu-rep <attempted steps per launch> 40
CFG=${1:-5}; KRN=${2:-lmc_spec_tf_kernel}; SKIP=${3:-20}; SPL=${4:-1}; TAG=${5:-ncu_cfg$CFG}
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:$KRN --launch-skip $SKIP --launch-count 1 \
  -o gpurun_out/$TAG -f python scripts/prof_cfg.py $CFG $SPL $((SKIP + 2)) > gpurun_out/$TAG.log 2>&1
tail -2 gpurun_out/$TAG.log
