#!/bin/bash
# config 5 (north star) on the GPUs of this box; weak line + strong line; config 2 scaling check; 2-rank sharding test
N=$1
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests -m gpu -q -k "sharded or not_current" > gpurun_out/r02m_pytest_2gpu.log 2>&1; tail -3 gpurun_out/r02m_pytest_2gpu.log
fi
timeout 900 python bench.py --gpus $N --config 5 --steps 6 --no-cpu > gpurun_out/r02m_cfg5_gpus$N.json 2> gpurun_out/r02m_cfg5_gpus$N.err
echo "cfg5 x$N rc=$?"; tail -c 700 gpurun_out/r02m_cfg5_gpus$N.json; echo; tail -3 gpurun_out/r02m_cfg5_gpus$N.err
timeout 600 python bench.py --gpus $N --config 2 --no-cpu > gpurun_out/r02m_cfg2_gpus$N.json 2> gpurun_out/r02m_cfg2_gpus$N.err
echo "cfg2 x$N rc=$?"; tail -c 400 gpurun_out/r02m_cfg2_gpus$N.json; echo
if [ "$N" = "2" ]; then
  for ch in 1 2; do
    NCCL_MAX_NCHANNELS=$ch NCCL_MIN_NCHANNELS=1 timeout 600 python bench.py --gpus $N --config 2 --no-cpu > gpurun_out/r02m_cfg2_gpus${N}_ch$ch.json 2>/dev/null
    python -c "import json;d=json.loads(open('gpurun_out/r02m_cfg2_gpus${N}_ch$ch.json').read().strip().splitlines()[-1]);print('channels $ch', d['value'], d['ms_per_step'], d['gather_tail_ms'])"
  done
fi
