#!/bin/bash
# front-end (quantiser) timing, kernel timing and the GPU test suite on the product build (+ optional variants)
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 300 python tools/frontend_timing.py > gpurun_out/frontend_product.log 2>&1; echo "product:"; tail -4 gpurun_out/frontend_product.log | cut -c1-330
for v in "$@"; do
  SAB_LIB_PATH=sageattention_b200/lib/libsab_$v.so timeout 300 python tools/frontend_timing.py > gpurun_out/frontend_$v.log 2>&1; echo "$v:"; tail -4 gpurun_out/frontend_$v.log | cut -c1-330
done
timeout 300 python tools/perf_kernel.py > gpurun_out/perf_product.log 2>&1; tail -1 gpurun_out/perf_product.log
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; tail -8 gpurun_out/gpu_tests.log
