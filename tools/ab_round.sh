#!/bin/bash
# usage (on the GPU box): tools/ab_round.sh out_tag tag [tag ...]  -- parity of every tag's build (poseidon / merkle / commit
# tests through VX_B200_LIB), then tools/ab_lib.sh over all tags; log under gpurun_out/<out_tag>_ab.log
out=$1; shift
mkdir -p gpurun_out
( for t in "$@"; do
    echo "parity $t:"; VX_B200_LIB=$PWD/vectorx_b200/libvectorx_b200.$t.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
      -k "${PARITY_K:-poseidon or hash_no_pad or merkle or commit_from_values or golden or field or ntt}" 2>&1 | tail -2
  done
  STEPS=${STEPS:-10} tools/ab_lib.sh "$@" ) 2>&1 | tee gpurun_out/${out}_ab.log
