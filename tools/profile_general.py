"""One fused assignment on a synthetic network (for ncu):  python tools/profile_general.py an|cs [B] [K]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import b200env

name = sys.argv[1] if len(sys.argv) > 1 else "an"
B = int(sys.argv[2]) if len(sys.argv) > 2 else (512 if name == "an" else 148)
K = int(sys.argv[3]) if len(sys.argv) > 3 else 8
g = b200env.synthetic.anaheim_size(0) if name == "an" else b200env.synthetic.chicago_sketch_size(0)
dev = torch.device("cuda", 0)
eng = b200env.AssignmentEngine(b200env.Network(g), dev, "fp32")
rng = np.random.default_rng(0)
masks = np.zeros((B, eng.E), np.uint8)
for b in range(B):
    masks[b, rng.choice(eng.E, int(eng.E * 0.3), replace=False)] = 1
dmg = torch.from_numpy(masks).to(dev)
for _ in range(2):
    flow = torch.zeros((B, eng.E), dtype=eng.real, device=dev)
    out = eng.assign(dmg, flow, K, "fw")
torch.cuda.synchronize()
print(eng.kernel_name, float(out["tstt"].mean()))
