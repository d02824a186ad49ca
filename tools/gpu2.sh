#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_perf_pedersen.py > gpurun_out/quick_perf_pedersen.txt 2>&1
cat gpurun_out/quick_perf_pedersen.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pedersen.py -x -q -k "commitment or two_to_one" > gpurun_out/sanitizer_pedersen.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer_pedersen.log
tail -5 gpurun_out/sanitizer_pedersen.log
