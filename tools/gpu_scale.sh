#!/bin/bash
# multi-GPU runs (one process per GPU, NCCL): default bench (config 2, weak scaling) and config 3 (trajectory)
mkdir -p gpurun_out
N=${1:-2}
P=$((29500 + N))
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/r2_scale_cfg2_n$N.json 2> gpurun_out/r2_scale_cfg2_n$N.err; echo "cfg2 n=$N rc $?"; tail -2 gpurun_out/r2_scale_cfg2_n$N.err | cut -c1-300
cut -c1-400 gpurun_out/r2_scale_cfg2_n$N.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+40)) bench.py --gpus $N --config 3 --steps ${STEPS3:-200} > gpurun_out/r2_scale_cfg3_n$N.json 2> gpurun_out/r2_scale_cfg3_n$N.err; echo "cfg3 n=$N rc $?"; tail -2 gpurun_out/r2_scale_cfg3_n$N.err | cut -c1-300
cut -c1-900 gpurun_out/r2_scale_cfg3_n$N.json
