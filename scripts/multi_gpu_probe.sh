#!/bin/bash
# Multi-GPU checks on a box with N visible GPUs (run from the repo root: gpurun --gpus N -- 'bash scripts/multi_gpu_probe.sh'):
# the mcd_create_multi tests, the raw host->device ceiling with and without NUMA-local pinned buffers, and the
# end-to-end throughput of one process driving all GPUs through one multi-GPU context.  Every step has its own timeout.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
(lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; free -g) > gpurun_out/lscpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multi_ctx.py -x -q -m gpu 2>&1 | tail -5
for mode in default numa; do
  for n in 1 2 4 8; do
    [ $n -le $N ] || continue
    timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      scripts/h2d_probe.py $mode 2>/dev/null | grep probe | tee -a gpurun_out/h2d_probe.jsonl
  done
done
timeout 200 python scripts/multi_e2e.py 100000 2>&1 | grep probe | tee -a gpurun_out/h2d_probe.jsonl
