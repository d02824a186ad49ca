#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err
timeout 300 python bench.py --steps 5 --warmup 3 --workload 'merkle_2^20_poseidon_bls12_381' > gpurun_out/bench_r1_n1_bls20.json 2>> gpurun_out/bench_r1_n1.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2>> gpurun_out/bench_r1_n1.err
timeout 100 python bench.py --workload 'pedersen_crh_2^20_jubjub' --steps 5 --warmup 3 > gpurun_out/bench_r1_pedersen.json 2>> gpurun_out/bench_r1_n1.err
timeout 100 python bench.py --workload 'mixed_merkle_2^22' --steps 5 --warmup 3 > gpurun_out/bench_r1_mixed.json 2>> gpurun_out/bench_r1_n1.err
timeout 200 python tools/quick_perf.py > gpurun_out/quick_perf.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_poseidon_compress_team -s 2 -c 1 -f -o gpurun_out/prof_r1_team python tools/ncu_target.py bls compress 11 > gpurun_out/ncu_team.log 2>&1
python3 -c "
import json
for f in ('bench_r1_n1','bench_r1_n1_bls20','bench_r1_reference','bench_r1_pedersen','bench_r1_mixed'):
    j=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, j['value'], j['unit'], j.get('ms_per_step'), (j.get('e2e') or {}).get('value'))
"
cat gpurun_out/quick_perf.txt
