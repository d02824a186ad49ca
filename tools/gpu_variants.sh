#!/bin/bash
# Compare launch-bounds variants of the Poseidon kernels (development probe).
for v in "" _mb6 _mb8; do
  echo "== variant libcpb200$v.so"
  CPB_LIB_NAME=libcpb200$v.so timeout 300 python tools/quick_perf.py 2>&1 | grep -E "2\^22|merkle"
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
