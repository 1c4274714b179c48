#!/bin/bash
# usage: tools/gpu_shard.sh N [config] [steps]   (run under gpurun --gpus N)
N=${1:-2}; CFG=${2:-5}; STEPS=${3:-20}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench_shard.py --config $CFG --steps $STEPS --warmup 3 > gpurun_out/shard_cfg${CFG}_${N}gpu.json 2> gpurun_out/shard_cfg${CFG}_${N}gpu.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench_shard.py --config $CFG --steps $STEPS --warmup 3 > gpurun_out/shard_cfg${CFG}_${N}gpu.json 2> gpurun_out/shard_cfg${CFG}_${N}gpu.err
fi
echo "rc $?"; tail -c 1500 gpurun_out/shard_cfg${CFG}_${N}gpu.json; tail -5 gpurun_out/shard_cfg${CFG}_${N}gpu.err | cut -c1-300
