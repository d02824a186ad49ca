#!/bin/bash
# folded-shift squaring reduction: parity (Poseidon tests incl. crafted operands; 2^22 node-for-node) + perf probe
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_poseidon.py tests/test_gpu_merkle.py -m gpu -x -q -k "not 24" > gpurun_out/exp_fold_tests.log 2>&1
tail -3 gpurun_out/exp_fold_tests.log
timeout 200 python tools/quick_perf.py > gpurun_out/exp_fold_perf.txt 2>&1
cat gpurun_out/exp_fold_perf.txt
