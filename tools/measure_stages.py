"""Achieved HBM bandwidth of the four STAGE kernels behind b200_bpr / b200_sssp / b200_aon /
b200_update_reduce on Sioux Falls, on batches far larger than L2 (the fused kernels keep this state on
chip; the stage kernels stream it).  Bytes per scenario = what each call must read and write once:
  S1 bpr     9E                      (flow 4E + mask E in, t 4E out)
  S2 sssp    4E + 16 O N             (t in; dist f64 + pred + pedge out)
  S3 aon     4 O N + 4E + 8          (pedge in; aux + unassigned out)
  S4 update  21E + 24                (flow, aux, t 12E + mask E in; flow, t 8E + tstt, gap, status out)
python tools/measure_stages.py [B_bpr] [B_other]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import b200env

B1 = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
B2 = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
dev = torch.device("cuda", 0)
eng = b200env.AssignmentEngine(b200env.Network(b200env.load_sioux_falls()), dev, "fp32")
E, O, N = eng.E, eng.O, eng.N
peak = 6454.6
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def report(kernel, B, nbytes, ms):
    print(json.dumps({"kernel": kernel, "scenarios": B, "bytes": nbytes, "ms": ms, "GBps": nbytes / ms / 1e6,
                      "peak_GBps": peak, "frac_of_measured_peak": nbytes / ms / 1e6 / peak}), flush=True)


flow = torch.rand((B1, E), device=dev) * 20000.0
dmg = (torch.rand((B1, E), device=dev) < 0.3).to(torch.uint8)
t = torch.empty_like(flow)
report("bpr_kernel<float,4>", B1, B1 * E * 9, timed(lambda: eng.bpr(flow, dmg, out=t)))
del flow, dmg, t
torch.cuda.empty_cache()

flow = torch.rand((B2, E), device=dev) * 20000.0
dmg = (torch.rand((B2, E), device=dev) < 0.3).to(torch.uint8)
t = eng.bpr(flow, dmg)
res = {}
report("sssp_kernel<float>", B2, B2 * (4 * E + 16 * O * N), timed(lambda: res.__setitem__("s", eng.sssp(t)), reps=3))
dist, pred, pedge = res["s"]
del dist, pred
report("aon_kernel<float>", B2, B2 * (4 * O * N + 4 * E + 8), timed(lambda: res.__setitem__("a", eng.aon(pedge)), reps=3))
aux, un = res["a"]
status = torch.zeros(B2, dtype=torch.int32, device=dev)
report("update_kernel<float>", B2, B2 * (21 * E + 24),
       timed(lambda: eng.update_reduce("fw", 3, flow, aux, t, dmg, un, None, status), reps=5))
