#!/bin/bash
# attn_q4.cu (one CTA per SM, four softmax warpgroups) against attn_alt.cu over the sequence length; optional variants of the library
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
SAB_ATTN_KERNEL=q4 timeout 300 python tools/first_run_check.py > gpurun_out/q4_check.log 2>&1; echo "q4 check rc=$? $(tail -1 gpurun_out/q4_check.log)"
for k in alt q4; do
  SAB_ATTN_KERNEL=$k timeout 600 python tools/perf_kernel.py long > gpurun_out/perf_long_$k.log 2>&1; echo "$k: $(tail -1 gpurun_out/perf_long_$k.log)"
done
for v in "$@"; do
  SAB_ATTN_KERNEL=q4 SAB_LIB_PATH=sageattention_b200/lib/libsab_$v.so timeout 600 python tools/perf_kernel.py long > gpurun_out/perf_long_$v.log 2>&1; echo "$v: $(tail -1 gpurun_out/perf_long_$v.log)"
done
SAB_ATTN_KERNEL=q4 timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "attention_vs_oracle or full_size_config1 or api_behaviour or real_reference" > gpurun_out/q4_tests.log 2>&1; tail -3 gpurun_out/q4_tests.log
