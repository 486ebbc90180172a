#!/bin/bash
# Runs ON the GPU box: everything the driver runs at round end (GPU tests, smoke, both bench arms) plus the ncu
# launch list of the bench command and one full capture of the dominant kernel.  Outputs under gpurun_out/<tag>_*.
tag=${1:-val}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py 2>gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/${tag}_bench_ref.err | tail -1 > gpurun_out/${tag}_bench_reference.json
python - <<PY
import json
for f in ("gpurun_out/${tag}_bench.json", "gpurun_out/${tag}_bench_reference.json"):
    try:
        d = json.loads(open(f).read())
        print(f, "value", round(d["value"], 1), d["unit"], "e2e", d.get("e2e", {}).get("value"), "ms", d.get("ms_per_step"),
              (d.get("roofline") or {}).get("phase_ms"), "frac", (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "FAILED", e, open(f).read()[-800:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-prove > gpurun_out/${tag}_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:leaf_hash_kernel -s 4 -c 1 -o gpurun_out/${tag}_leaf -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-prove > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
