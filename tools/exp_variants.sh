#!/bin/bash
# Development: compare library variants built with CPB_LIB_NAME=libcpb200_<tag>.so (crypto_primitives_b200/_build.py) on the GPU.
# usage: tools/exp_variants.sh out.txt tag1 tag2 ...   ("base" = the default library)
out=$1; shift
: > $out
for tag in "$@"; do
  if [ "$tag" = base ]; then lib=libcpb200.so; else lib=libcpb200_$tag.so; fi
  echo "=== $tag ($lib)" >> $out
  CPB_LIB_NAME=$lib timeout 300 python -m pytest tests/test_gpu_poseidon.py -x -q -k "crh_matches_oracle or crafted or reference_kat" 2>&1 | tail -1 >> $out
  CPB_LIB_NAME=$lib timeout 300 python tools/quick_perf.py >> $out 2>&1
done
