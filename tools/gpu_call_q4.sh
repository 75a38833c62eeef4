#!/bin/bash
# first contact of the 4-warpgroup kernel: watchdog build first (bounded mbarrier waits), then correctness and timing on the product build
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
export SAB_ATTN_KERNEL=q4
SAB_LIB_PATH=sageattention_b200/lib/libsab_wdq4.so timeout 300 python tools/first_run_check.py > gpurun_out/q4_wd.log 2>&1; echo "wd rc=$?"; grep -c "mbarrier timeout" gpurun_out/q4_wd.log; tail -4 gpurun_out/q4_wd.log | cut -c1-250
if grep -q "mbarrier timeout\|Error\|error" gpurun_out/q4_wd.log; then echo "watchdog build failed: stop"; exit 0; fi
SAB_Q4_CTAS=3 SAB_LIB_PATH=sageattention_b200/lib/libsab_wdq4.so timeout 300 python tools/first_run_check.py > gpurun_out/q4_wd3.log 2>&1; echo "wd (3 CTAs) rc=$?"; grep -c "mbarrier timeout" gpurun_out/q4_wd3.log; tail -2 gpurun_out/q4_wd3.log | cut -c1-250
if grep -q "mbarrier timeout\|Error\|error" gpurun_out/q4_wd3.log; then echo "watchdog build (3 CTAs) failed: stop"; exit 0; fi
SAB_Q4_CTAS=3 timeout 300 python tools/first_run_check.py > gpurun_out/q4_check3.log 2>&1; echo "check (3 CTAs) rc=$?"; tail -1 gpurun_out/q4_check3.log | cut -c1-250
timeout 300 python tools/first_run_check.py > gpurun_out/q4_check.log 2>&1; echo "check rc=$?"; tail -3 gpurun_out/q4_check.log | cut -c1-250
timeout 300 python tools/perf_kernel.py short2 > gpurun_out/perf_q4.log 2>&1; echo "q4: $(tail -1 gpurun_out/perf_q4.log)"
SAB_ATTN_KERNEL=alt timeout 300 python tools/perf_kernel.py short2 > gpurun_out/perf_alt.log 2>&1; echo "alt: $(tail -1 gpurun_out/perf_alt.log)"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "attention_vs_oracle or full_size_config1 or api_behaviour or real_reference" > gpurun_out/q4_tests.log 2>&1; tail -5 gpurun_out/q4_tests.log
