#!/bin/bash
# usage (on the GPU box): tools/ab_lib.sh [tag ...]   -- bench phase times of the shipped library and of A/B builds
# (vectorx_b200/libvectorx_b200.<tag>.so, made with `make -C vectorx_b200/csrc VARIANT=tag EXTRA=...`)
cd "$(dirname "$0")/.."
for tag in "" "$@"; do
  lib=vectorx_b200/libvectorx_b200${tag:+.$tag}.so
  VX_B200_LIB=$PWD/$lib python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-prove ${BENCH_ARGS} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
p=d['roofline']['phase_ms']
print('${tag:-shipped}', 'value %.1f e2e %.1f ms %.3f |' % (d['value'], d['e2e']['value'], d['ms_per_step']), ' '.join('%s %.3f' % (k, v) for k, v in p.items()))
"
done
