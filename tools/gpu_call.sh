#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 300 python tools/first_run_check.py > gpurun_out/first_run_check.log 2>&1; tail -2 gpurun_out/first_run_check.log
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; tail -8 gpurun_out/gpu_tests.log
timeout 300 python tools/perf_kernel.py > gpurun_out/perf_product.log 2>&1; tail -1 gpurun_out/perf_product.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 400 gpurun_out/bench_ref.json
timeout 1200 python tools/profile_round.py capture r02 > gpurun_out/profile_capture.log 2>&1; tail -5 gpurun_out/profile_capture.log
