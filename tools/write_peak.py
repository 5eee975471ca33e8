"""Measures the device's write-bandwidth ceilings (GB/s) for the mechanisms the stiffness kernels can use."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import gemlab_b200 as g

ctx = g.Context(0)
nbytes = int(float(sys.argv[1]) * 1e9) if len(sys.argv) > 1 else 36_864_000_000
buf = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
res = {}
for w in (8, 12, 24, 32, 48, 96, 192, 384, 768):
    for mode in (1, 3):
        res[f"bulk_{w}warps_mode{mode}"] = round(ctx.hbm_write_peak(buf, "bulk", w, mode))
for mode in (0, 1, 2):
    res[f"stg128_mode{mode}"] = round(ctx.hbm_write_peak(buf, "stg", mode=mode))
res["memset"] = ctx.hbm_write_peak(buf, "memset")
# copy (read + write) for comparison with MEASURED_PEAKS.json
a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda")
b = torch.empty_like(a)
best = 0
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best = max(best, 2 * a.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
res["torch_copy_rw"] = best
print(json.dumps(res))
