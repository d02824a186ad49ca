#!/bin/bash
# Round-2 profile captures (run under gpurun on ONE GPU): `gpurun -- tools/profile_r2.sh`.  Outputs land in gpurun_out/.
# The .ncu-rep files are summarised ON the box (tools/ncu_summary.py: raw page; tools/ncu_opmix.py: source page) and then
# deleted: gpurun brings back at most 64 MiB and one capture with sources is ~30 MiB.
mkdir -p gpurun_out
# every launch of the default bench command with its device time (cold-cache, serialised: compare SHARES with the bench line)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/r2_bench_under_ncu.log 2>&1
cap() {   # name kernel-regex target-args...
  name=$1; regex=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex -s 1 -c 1 -f -o gpurun_out/prof_r2_$name \
      python tools/ncu_target.py "$@" > gpurun_out/r2_ncu_$name.log 2>&1
  { echo "# ncu --set full --clock-control none, B200, round 2: $name (python tools/ncu_target.py $*)"; python tools/ncu_summary.py gpurun_out/prof_r2_$name.ncu-rep;
    echo; echo "# dynamic SASS opcode mix (ncu source page)"; python tools/ncu_opmix.py gpurun_out/prof_r2_$name.ncu-rep; } > gpurun_out/r2_ncu_$name.txt 2>&1
  rm -f gpurun_out/prof_r2_$name.ncu-rep
}
cap crh_bn254 k_poseidon_crh bn254 compress 20
cap crh_bls k_poseidon_crh bls compress 20
cap pedersen_gather k_pedersen_hash bls pedersen 18
cap tree_top k_poseidon_tree_top bn254 top 13
