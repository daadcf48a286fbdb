"""Time of the state-feature kernel (S5: Brandes betweenness + node / link features, reference
get_state :751-819) per scenario.  python tools/measure_features.py [sf an cs]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import b200env

CASES = {"sf": (b200env.load_sioux_falls, 4096), "an": (lambda: b200env.synthetic.anaheim_size(0), 1024),
         "cs": (lambda: b200env.synthetic.chicago_sketch_size(0), 256)}
dev = torch.device("cuda", 0)
for name in (sys.argv[1:] or ["sf", "an", "cs"]):
    make, B = CASES[name]
    eng = b200env.AssignmentEngine(b200env.Network(make()), dev, "fp32")
    rng = np.random.default_rng(0)
    masks = np.zeros((B, eng.E), np.uint8)
    for b in range(B):
        masks[b, rng.choice(eng.E, int(eng.E * 0.3), replace=False)] = 1
    dmg = torch.from_numpy(masks).to(dev)
    flow = torch.rand((B, eng.E), device=dev) * 1000.0
    tstt = torch.full((B,), 100.0, dtype=torch.float64, device=dev)
    for _ in range(2):
        eng.state_features(flow, dmg, dmg, tstt, 1e-3)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        eng.state_features(flow, dmg, dmg, tstt, 1e-3)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 3
    print(json.dumps({"network": name, "N": eng.N, "E": eng.E, "scenarios": B, "ms": ms, "states_per_s": B / ms * 1e3}))
