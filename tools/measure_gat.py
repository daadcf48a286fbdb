"""Achieved HBM bandwidth of the GAT edge-softmax + aggregation kernels on Sioux Falls
(per graph and layer, read xw once + write out once = 2*N*H*C*4 bytes; the
attention terms add (2N + E+N)*H*4 in and (E+N)*H*4 out).  python tools/measure_gat.py [B] [H] [C]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import b200env

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
H = int(sys.argv[2]) if len(sys.argv) > 2 else 4
C = int(sys.argv[3]) if len(sys.argv) > 3 else 256
dev = torch.device("cuda", 0)
net = b200env.Network(b200env.load_sioux_falls())
N, E = net.num_nodes, net.num_edges
xw = torch.randn(B, N, H, C, device=dev, requires_grad=True)
a_src = torch.randn(B, N, H, device=dev, requires_grad=True)
a_dst = torch.randn(B, N, H, device=dev, requires_grad=True)
a_edge = torch.randn(B, E + N, H, device=dev, requires_grad=True)
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


with torch.no_grad():
    ms_f = timed(lambda: b200env.gat_edge_softmax_aggregate(net, xw, a_src, a_dst, a_edge))
fwd_bytes = B * (2 * N * H * C * 4 + (2 * N + 2 * (E + N)) * H * 4)
out, alpha = b200env.gat_edge_softmax_aggregate(net, xw, a_src, a_dst, a_edge)
g = torch.randn_like(out)
ms_fb = timed(lambda: torch.autograd.grad(b200env.gat_edge_softmax_aggregate(net, xw, a_src, a_dst, a_edge)[0], (xw, a_src, a_dst, a_edge), g))
bwd_bytes = B * (3 * N * H * C * 4 + (2 * N + 2 * (E + N)) * H * 4 * 2)  # read xw, grad_out; write grad_xw
ms_b = ms_fb - ms_f
print(json.dumps({"graphs": B, "H": H, "C": C, "fwd_ms": ms_f, "fwd_GBps": fwd_bytes / ms_f / 1e6,
                  "fwd_frac_of_measured_peak": fwd_bytes / ms_f / 1e6 / peak, "bwd_ms": ms_b,
                  "bwd_GBps": bwd_bytes / ms_b / 1e6, "bwd_frac_of_measured_peak": bwd_bytes / ms_b / 1e6 / peak}))
