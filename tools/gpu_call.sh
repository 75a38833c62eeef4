#!/bin/bash
export PYTHONUNBUFFERED=1
L=$PWD/sageattention_b200/lib
mkdir -p gpurun_out
SAB_ATTN_KERNEL=alt SAB_LIB_PATH=$L/libsab_alt_tau4.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:sage_attn -s 2 -c 1 -f -o gpurun_out/alt4_attn python tools/run_attn_once.py 2 32 8192 128 0 per_thread 3 > gpurun_out/ncu_alt4.log 2>&1; tail -3 gpurun_out/ncu_alt4.log
SAB_LIB_PATH=$L/libsab_lz_a_p8.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:sage_attn -s 2 -c 1 -f -o gpurun_out/lzp8w_attn python tools/run_attn_once.py 2 32 8192 128 0 per_warp 3 > gpurun_out/ncu_lzp8w.log 2>&1; tail -3 gpurun_out/ncu_lzp8w.log
