#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SAB_TEST_FUSED_GATHER=1 timeout 600 $TR --master-port 29611 tests/sp_check.py > gpurun_out/sp_check_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/sp_check_n$N.log; grep -E "rank 0|SP_CHECK|rc=|Error|error" gpurun_out/sp_check_n$N.log | tail -30
timeout 600 $TR --master-port 29612 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 1500 gpurun_out/bench_n$N.json
SAB_SP_FUSED_GATHER=1 timeout 600 $TR --master-port 29613 bench.py --gpus $N --no-sweep > gpurun_out/bench_n${N}_fused.json 2> gpurun_out/bench_n${N}_fused.err; tail -c 700 gpurun_out/bench_n${N}_fused.json; tail -3 gpurun_out/bench_n${N}_fused.err
