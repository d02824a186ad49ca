#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_poseidon.py tests/test_gpu_merkle.py -m gpu -x -q -k "team or crafted or build_matches or full_size_trees and not 24 and not 22" > gpurun_out/exp_team2_tests.log 2>&1
tail -3 gpurun_out/exp_team2_tests.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_poseidon.py -m gpu -x -q -k "team" > gpurun_out/exp_team2_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/exp_team2_racecheck.log
tail -5 gpurun_out/exp_team2_racecheck.log
timeout 120 python tools/exp_team.py > gpurun_out/exp_team2.txt 2>&1
cat gpurun_out/exp_team2.txt
