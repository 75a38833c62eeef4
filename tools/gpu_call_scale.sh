#!/bin/bash
# bench.py on N GPUs (collective KV gather, then the gather fused into the attention launch); no correctness sweep (tools/gpu_call_sp.sh)
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29612 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 1200 gpurun_out/bench_n$N.json
SAB_SP_FUSED_GATHER=1 timeout 600 $TR --master-port 29613 bench.py --gpus $N --no-sweep > gpurun_out/bench_n${N}_fused.json 2> gpurun_out/bench_n${N}_fused.err; tail -c 600 gpurun_out/bench_n${N}_fused.json; tail -3 gpurun_out/bench_n${N}_fused.err
