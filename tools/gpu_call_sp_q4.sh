#!/bin/bash
# sequence parallel with attn_q4.cu: tests/sp_check.py with the kernel forced (its shapes are small: the default dispatch would keep
# attn_alt.cu), then bench.py --gpus N with the default dispatch (S = 32768: csrc/attn.cu prefer_q4 picks attn_q4.cu)
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SAB_ATTN_KERNEL=q4 timeout 400 $TR --master-port 29611 tests/sp_check.py > gpurun_out/sp_check_q4_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/sp_check_q4_n$N.log; grep -E "rank 0|rc=" gpurun_out/sp_check_q4_n$N.log | tail -12
timeout 400 $TR --master-port 29612 bench.py --gpus $N --no-sweep > gpurun_out/bench_n${N}_q4.json 2> gpurun_out/bench_n${N}_q4.err; tail -c 900 gpurun_out/bench_n${N}_q4.json
