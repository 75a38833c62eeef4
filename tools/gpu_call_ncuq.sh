#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
for v in product "$@"; do
  L=""; [ "$v" != product ] && L="sageattention_b200/lib/libsab_$v.so"
  SAB_LIB_PATH=$L timeout 600 ncu --set full --clock-control none --import-source on -k regex:"quant_int8_kernel|v_quant_transpose_kernel|channel_stats_stage1" -s 5 -c 5 -o gpurun_out/ncu_quant_$v -f python tools/frontend_once.py > gpurun_out/ncu_quant_$v.log 2>&1; tail -2 gpurun_out/ncu_quant_$v.log
done
