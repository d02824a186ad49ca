#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp_v3.txt
for v in libcpb200.so libcpb200_sbox5.so libcpb200_allsplit.so; do
  echo "== $v" >> gpurun_out/exp_v3.txt
  CPB_LIB_NAME=$v timeout 200 python tools/quick_perf.py 2>&1 | grep -v "2^16" >> gpurun_out/exp_v3.txt
done
cat gpurun_out/exp_v3.txt
