#!/bin/bash
# Round-1 profile captures (run under gpurun on ONE GPU).  Outputs land in gpurun_out/; summaries are made
# on the CPU box with tools/ncu_summary.py / tools/ncu_opmix.py and committed under profiles/.
mkdir -p gpurun_out
# every launch of the default bench command with its device time (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_bench.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
# dominant kernels, full sets
for f in bn254 bls; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_poseidon_crh -s 1 -c 1 -f -o gpurun_out/prof_r1_crh_${f}_final \
      python tools/ncu_target.py $f compress 20 > gpurun_out/ncu_${f}.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pedersen_hash -s 1 -c 1 -f -o gpurun_out/prof_r1_pedersen_final \
    python tools/ncu_target.py bls pedersen 18 > gpurun_out/ncu_ped.log 2>&1
# the real bench numbers (not under a profiler), with clocks sampled by bench.py itself
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err
timeout 300 python bench.py --steps 5 --warmup 3 --workload 'merkle_2^20_poseidon_bls12_381' > gpurun_out/bench_r1_n1_bls20.json 2>> gpurun_out/bench_r1_n1.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2>> gpurun_out/bench_r1_n1.err
timeout 100 ./tools/ubench_int > gpurun_out/ubench_int.txt 2>&1
timeout 200 python tools/quick_perf.py > gpurun_out/quick_perf.txt 2>&1
timeout 200 python tools/quick_perf_pedersen.py > gpurun_out/quick_perf_pedersen.txt 2>&1
ls -la gpurun_out | tail -20
