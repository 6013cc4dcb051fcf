#!/bin/bash
# config 5 (north star) on N GPUs of this box: weak line (4096 walkers per GPU) + strong line (32768 walkers in total)
N=$1
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests -m gpu -q -k "sharded or not_current" > gpurun_out/r02g_pytest_2gpu.log 2>&1; tail -3 gpurun_out/r02g_pytest_2gpu.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config 5 --steps 6 --no-cpu > gpurun_out/r02g_cfg5_gpus$N.json 2> gpurun_out/r02g_cfg5_gpus$N.err
echo "cfg5 x$N rc=$?"; tail -c 900 gpurun_out/r02g_cfg5_gpus$N.json; echo; tail -3 gpurun_out/r02g_cfg5_gpus$N.err
