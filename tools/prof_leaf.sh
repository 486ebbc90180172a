#!/bin/bash
# usage: tools/prof_leaf.sh <tag> [variant]   -- runs on the GPU box: parity subset, bench phase times, ncu --set full of the leaf kernel
cd /root/repo
tag=$1; v=${2:-0}
/usr/local/graft/bin/gpurun --timeout 900 -- "bash tools/ab.sh $v; VX_POSEIDON_VARIANT=$v timeout 500 ncu --set full --clock-control none --import-source on -k regex:leaf_hash_kernel -s 4 -c 1 -o gpurun_out/leaf_$tag -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c.log 2>&1" 2>&1 | grep -E "passed|failed|value|error|GPU-minutes"
ncu -i gpurun_out/leaf_$tag.ncu-rep --page raw --csv > gpurun_out/leaf_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/leaf_$tag.ncu-rep --page source --csv > gpurun_out/leaf_${tag}_src.csv 2>/dev/null
python tools/profile_digest.py kernel gpurun_out/leaf_${tag}_raw.csv gpurun_out/leaf_${tag}_src.csv gpurun_out/leaf_${tag}.md 8912896
grep -E "time_duration|pipe_alu|fmaheavy|issue_active|inst_executed.sum|math_pipe|not_selected|wait_per|dispatch" gpurun_out/leaf_${tag}.md
sed -n '/opcode histogram/,$p' gpurun_out/leaf_${tag}.md | sed -n 4,18p
