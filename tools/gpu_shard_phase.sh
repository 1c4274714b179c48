#!/bin/bash
# phase timing of a sharded registration (2 GPUs) next to the unsharded one; run under gpurun --gpus 2
mkdir -p gpurun_out
CFG=${1:-5}
LS_PHASE_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench_shard.py --config $CFG --steps 1 --warmup 1 --no-parity > gpurun_out/phase2.json 2> gpurun_out/phase2.err
LS_PHASE_TIMING=1 timeout 300 python bench_shard.py --config $CFG --steps 1 --warmup 1 --no-parity > gpurun_out/phase1.json 2> gpurun_out/phase1.err
grep "it " gpurun_out/phase2.err | tail -120 | head -40
echo ---- single
grep "it " gpurun_out/phase1.err | tail -50 | head -20
