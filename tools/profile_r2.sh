#!/bin/bash
# Round-2 profile captures (run under gpurun on ONE GPU): `gpurun -- tools/profile_r2.sh`.  Outputs land in gpurun_out/;
# the summaries under profiles/r2_* are made on the CPU box with tools/ncu_summary.py / tools/ncu_opmix.py.
mkdir -p gpurun_out
# every launch of the default bench command with its device time (cold-cache, serialised: compare SHARES with the bench line)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/r2_bench_under_ncu.log 2>&1
# dominant kernels, full sets
for f in bn254 bls; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_poseidon_crh -s 1 -c 1 -f -o gpurun_out/prof_r2_crh_${f} \
      python tools/ncu_target.py $f compress 20 > gpurun_out/r2_ncu_${f}.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pedersen_hash -s 1 -c 1 -f -o gpurun_out/prof_r2_pedersen \
    python tools/ncu_target.py bls pedersen 18 > gpurun_out/r2_ncu_ped.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_poseidon_tree_top -s 1 -c 1 -f -o gpurun_out/prof_r2_tree_top \
    python tools/ncu_target.py bn254 top 13 > gpurun_out/r2_ncu_top.log 2>&1
