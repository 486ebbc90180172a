#!/bin/bash
# usage (here, no GPU): tools/prof_round.sh <tag>   -- one gpurun call: launch list of the bench command, `ncu --set full` of the
# leaf-hash kernel and of the two LDE passes; the .ncu-rep files come back in gpurun_out/ and are digested into profiles/
cd "$(dirname "$0")/.."
tag=$1
/usr/local/graft/bin/gpurun --timeout 1500 -- "
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-prove > gpurun_out/l.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:leaf_hash_kernel -s 3 -c 1 -o gpurun_out/${tag}_leaf -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-prove > gpurun_out/c.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'ntt_pass256_kernel<1|ntt_final4096_kernel<0' -s 6 -c 2 -o gpurun_out/${tag}_ntt -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-prove > gpurun_out/n.log 2>&1
tail -2 gpurun_out/c.log | cut -c1-200" 2>&1 | tail -5
for k in leaf ntt; do
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page source --csv > gpurun_out/${tag}_${k}_src.csv 2>/dev/null
done
