"""Achieved HBM bandwidth of the S1 kernel (b200_bpr: the one stage that streams per-scenario arrays
from HBM - 9 bytes per link) on a batch far larger than L2.  python tools/measure_stages.py [B]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import b200env

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
dev = torch.device("cuda", 0)
eng = b200env.AssignmentEngine(b200env.Network(b200env.load_sioux_falls()), dev, "fp32")
flow = torch.rand((B, eng.E), device=dev) * 20000.0
dmg = (torch.rand((B, eng.E), device=dev) < 0.3).to(torch.uint8)
t = torch.empty_like(flow)
for _ in range(3):
    eng.bpr(flow, dmg, out=t)
torch.cuda.synchronize()
reps = 10
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(reps):
    eng.bpr(flow, dmg, out=t)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / reps
nbytes = B * eng.E * 9
peak = 6454.6
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
print(json.dumps({"kernel": "bpr_kernel<float,4>", "scenarios": B, "links": B * eng.E, "bytes": nbytes, "ms": ms,
                  "GBps": nbytes / ms / 1e6, "peak_GBps": peak, "frac_of_measured_peak": nbytes / ms / 1e6 / peak}))
