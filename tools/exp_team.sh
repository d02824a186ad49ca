#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp_team.txt
for tm in 0 2048 4096 8192 16384 32768; do
  CPB_TEAM_MAX=$tm timeout 120 python tools/exp_team.py >> gpurun_out/exp_team.txt 2>&1
done
for s in 1 4 16; do
  CPB_MERKLE_STREAMS=$s timeout 120 python tools/exp_team.py >> gpurun_out/exp_team.txt 2>&1
done
cat gpurun_out/exp_team.txt
