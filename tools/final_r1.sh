#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "not full_size and not large_batch and not linearity and not dist" > gpurun_out/sanitizer_r1_memcheck.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer_r1_memcheck.log
tail -4 gpurun_out/sanitizer_r1_memcheck.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pedersen_hash_gather -s 1 -c 1 -f -o gpurun_out/prof_r1_pedersen_gather python tools/ncu_target.py bls pedersen 18 > gpurun_out/ncu_ped2.log 2>&1
timeout 200 python tools/quick_perf_pedersen.py > gpurun_out/quick_perf_pedersen.txt 2>&1
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err
tail -c 600 gpurun_out/bench_r1_n1.json
