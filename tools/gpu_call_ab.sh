#!/bin/bash
# A/B of two builds of the library (product vs lib/libsab_<variant>.so) with tools/perf_kernel.py, then the GPU test suite
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
for rep in 1 2; do
  timeout 300 python tools/perf_kernel.py > gpurun_out/perf_product_$rep.log 2>&1; echo "product: $(tail -1 gpurun_out/perf_product_$rep.log)"
  for v in "$@"; do
    SAB_LIB_PATH=sageattention_b200/lib/libsab_$v.so timeout 300 python tools/perf_kernel.py > gpurun_out/perf_${v}_$rep.log 2>&1; echo "$v: $(tail -1 gpurun_out/perf_${v}_$rep.log)"
  done
done
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; tail -8 gpurun_out/gpu_tests.log
