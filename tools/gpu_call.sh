#!/bin/bash
export PYTHONUNBUFFERED=1
L=$PWD/sageattention_b200/lib
mkdir -p gpurun_out
for v in t2_base t2_p8 t2_a_p8 t2_a_p6 t2_ad_p12 t2_d_p8; do SAB_LIB_PATH=$L/libsab_$v.so timeout 200 python tools/perf_kernel.py short2 > gpurun_out/perf_$v.log 2>&1; echo "$v: $(tail -1 gpurun_out/perf_$v.log | cut -d' ' -f2-)"; done
SAB_ATTN_KERNEL=alt SAB_LIB_PATH=$L/libsab_alt_tau4.so timeout 200 python tools/perf_kernel.py short2 > gpurun_out/perf_alt_tau4.log 2>&1; echo "alt_tau4: $(tail -1 gpurun_out/perf_alt_tau4.log | cut -d' ' -f2-)"
