#!/bin/bash
# round 1, after the fp_sqr carry fix: whole GPU suite, perf probe, contract bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r1b_tests.log 2>&1
tail -14 gpurun_out/r1b_tests.log
timeout 200 python tools/quick_perf.py > gpurun_out/r1b_quick_perf.txt 2>&1
cat gpurun_out/r1b_quick_perf.txt
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r1b_bench_n1.json 2> gpurun_out/r1b_bench_n1.err
tail -c 1500 gpurun_out/r1b_bench_n1.json
