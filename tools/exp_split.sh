#!/bin/bash
# round-1 experiment: merged round loop vs dedicated partial-round loop (CPB_POS_SPLIT), one GPU, plus the new parity tests
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_poseidon.py tests/test_gpu_bowe_hopwood.py tests/test_gpu_merkle.py tests/test_gpu_pedersen.py -m gpu -x -q -k "sponge or bowe or multiproof or commitment or pedersen_merkle" > gpurun_out/exp_tests.log 2>&1
tail -3 gpurun_out/exp_tests.log
for v in libcpb200.so libcpb200_split.so libcpb200_split_roll.so; do
  echo "== $v" >> gpurun_out/exp_split.txt
  CPB_LIB_NAME=$v timeout 200 python tools/quick_perf.py >> gpurun_out/exp_split.txt 2>&1
done
cat gpurun_out/exp_split.txt
