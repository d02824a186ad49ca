#!/bin/bash
# round 1 closing run on ONE GPU: whole GPU suite, every bench line, launch list and full captures of the dominant kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r1c_tests.log 2>&1
tail -10 gpurun_out/r1c_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1c.err
timeout 300 python bench.py --steps 5 --warmup 3 --workload 'merkle_2^20_poseidon_bls12_381' > gpurun_out/bench_r1_n1_bls20.json 2>> gpurun_out/bench_r1c.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2>> gpurun_out/bench_r1c.err
timeout 100 python bench.py --workload 'pedersen_crh_2^20_jubjub' --steps 5 --warmup 3 > gpurun_out/bench_r1_pedersen.json 2>> gpurun_out/bench_r1c.err
timeout 100 python bench.py --workload 'mixed_merkle_2^22' --steps 5 --warmup 3 > gpurun_out/bench_r1_mixed.json 2>> gpurun_out/bench_r1c.err
python3 -c "
import json
for f in ('bench_r1_n1','bench_r1_n1_bls20','bench_r1_reference','bench_r1_pedersen','bench_r1_mixed'):
    j=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, j['value'], j['unit'], j.get('ms_per_step'), (j.get('e2e') or {}).get('value'), (j.get('integer_pipe') or {}).get('frac'))
"
timeout 200 python tools/quick_perf.py > gpurun_out/quick_perf.txt 2>&1; cat gpurun_out/quick_perf.txt
timeout 200 python tools/quick_perf_pedersen.py > gpurun_out/quick_perf_pedersen.txt 2>&1; tail -6 gpurun_out/quick_perf_pedersen.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_bench.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
for f in bn254 bls; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_poseidon_crh -s 1 -c 1 -f -o gpurun_out/prof_r1_crh_${f}_final \
      python tools/ncu_target.py $f compress 20 > gpurun_out/ncu_${f}.log 2>&1
done
ls -la gpurun_out | tail -12
