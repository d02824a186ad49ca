#!/bin/bash
# First GPU session: tests, microbenchmark, quick perf, ncu captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 120 ./tools/ubench_int > gpurun_out/ubench_int.txt 2>&1
timeout 300 python tools/quick_perf.py > gpurun_out/quick_perf.txt 2>&1
cat gpurun_out/quick_perf.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_poseidon_crh -s 2 -c 2 -o gpurun_out/prof_r1_crh python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1_first.csv python tools/ncu_target.py merkle > gpurun_out/ncu_launch.log 2>&1
ls -la gpurun_out
