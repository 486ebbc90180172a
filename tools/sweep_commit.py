#!/usr/bin/env python
"""BASELINE.json config 5 (commit sweep 2^18-2^24 rows x 64-512 columns, rate_bits 1-3) and the STARK shapes of config 3 on
ONE GPU: `vx_commit_from_values` with the values resident in HBM, CUDA events on the library's stream, 3 timed commits after
2 warm-ups.  One JSON line per shape (Melem/s of n*c input elements, phase times); shapes that do not fit the GPU are
reported as skipped.  usage (on the GPU box): python tools/sweep_commit.py > gpurun_out/sweep.jsonl"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorx_b200 as vx                                              # noqa: E402
from vectorx_b200._lib import check, load, vp                          # noqa: E402

SHAPES = [(18, 64, 3), (18, 512, 1), (18, 1271, 1), (18, 2502, 1), (20, 128, 2), (20, 256, 2), (20, 512, 1), (22, 64, 3),
          (22, 128, 1), (22, 256, 2), (24, 64, 1), (24, 64, 3), (24, 128, 1)]
P = 0xFFFFFFFF00000001


def main():
    ctx = vx.Context(0)
    lib = load()
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
    free_b, _ = torch.cuda.mem_get_info(0)
    for log_n, c, rate in SHAPES:
        n = 1 << log_n
        need = 8 * n * c * (3 + (1 << rate)) + 64 * (n << rate)       # values + stage + coeffs + LDE + digests
        if need > 0.9 * free_b:
            print(json.dumps({"shape": [log_n, c, rate], "skipped": f"needs {need / 1e9:.0f} GB"}), flush=True)
            continue
        g = torch.Generator(device="cuda").manual_seed(log_n * 1000 + c)
        vals = torch.randint(0, 2**62, (c, n), dtype=torch.int64, device="cuda", generator=g)     # any u64 < p is canonical
        times, phases = [], {}
        for it in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h = vp()
            e0.record(stream)
            check(lib.vx_commit_from_values(ctx.handle, vals.data_ptr(), c, log_n, rate, 4, ctypes.byref(h)), "commit")
            e1.record(stream)
            torch.cuda.synchronize()
            if it >= 2:
                times.append(e0.elapsed_time(e1))
                for k, v in ctx.phase_ms().items():
                    phases[k] = phases.get(k, 0.0) + v / 3
            lib.vx_batch_free(h)
        ms = sum(times) / len(times)
        print(json.dumps({"shape": [log_n, c, rate], "rows": n, "cols": c, "rate_bits": rate, "ms_per_commit": ms,
                          "melem_per_s": n * c / ms / 1e3, "phase_ms": phases,
                          "lde_gb": 8 * (n << rate) * c / 1e9}), flush=True)
        del vals
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
