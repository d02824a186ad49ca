#!/bin/bash
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "not full_size and not large_batch and not linearity and not dist and not node_for_node" > gpurun_out/sanitizer_r1c_memcheck.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer_r1c_memcheck.log
tail -6 gpurun_out/sanitizer_r1c_memcheck.log
