#!/bin/bash
# usage (on the GPU box): tools/jit_round.sh tag -- parity of the run-time compiled quotient kernel, then whole-proof phase
# timings at 2^16 rows with the gate program interpreted / compiled (two register budgets); logs under gpurun_out/<tag>_*
tag=${1:-jit}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_prover.py -x -q -m gpu -k "circuit_digest or zs_partial or quotient or full_proof" ) 2>&1 | tail -6
( time VX_TEST_JIT_ALL=1 timeout 600 python -m pytest tests/test_gpu_prover.py -x -q -m gpu -k "all_gate_kinds_compiled" ) 2>&1 | tail -5
for v in "0 0" "1 0" "1 4"; do
  set -- $v
  VX_PROVE_BITS=16 VX_PROVE_COMPILE=$1 VX_JIT_MINB=$2 timeout 600 python -m pytest tests/test_gpu_prove_timing.py -x -q -m gpu -k prove_verifies 2>&1 | tail -2
  cp gpurun_out/prove_timing.json gpurun_out/${tag}_prove_timing_c$1_m$2.json
  python - <<PY
import json
d = json.load(open("gpurun_out/prove_timing.json"))
print("compile=$1 minb=$2", "prove_ms %.2f" % d["prove_ms"], "quotient %.3f" % d["phase_ms"]["compute quotient polys"], "compile_s", d.get("gate_compile_s"))
PY
done
