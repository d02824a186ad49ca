#!/bin/bash
# last run of round 1 on one GPU: whole GPU suite (with the full-size Pedersen pass) and the Pedersen kernel captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r1d_tests.log 2>&1
tail -11 gpurun_out/r1d_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pedersen_hash_gather -s 1 -c 1 -f -o gpurun_out/prof_r1_pedersen_gather python tools/ncu_target.py bls pedersen 18 > gpurun_out/ncu_ped2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_pedersen_hash<" -s 1 -c 1 -f -o gpurun_out/prof_r1_pedersen_smem8 python tools/ncu_target.py bls pedersen8 18 > gpurun_out/ncu_ped8.log 2>&1
tail -3 gpurun_out/ncu_ped2.log gpurun_out/ncu_ped8.log
ls -la gpurun_out/*.ncu-rep | tail -5
