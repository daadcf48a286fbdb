"""Sioux Falls fixed-step assignment rate against the batch size (wave quantisation of the specialised
kernels).  Run once as is and once with B200_STATIC_NO_GRP=1 to compare the grouped kernel with one warp
per scenario.  Usage: python tools/sweep_static.py [B ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200env  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [592, 2368, 3552, 4096, 4144, 8192, 16384, 65536]
    dev = torch.device("cuda", 0)
    eng = b200env.AssignmentEngine(b200env.Network(b200env.load_sioux_falls()), dev, "fp32")
    rng = np.random.default_rng(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for B in sizes:
        masks = np.zeros((B, eng.E), np.uint8)
        for b in range(B):
            masks[b, rng.choice(eng.E, size=22, replace=False)] = 1
        dmg = torch.from_numpy(masks).to(dev)
        flow = torch.zeros((B, eng.E), dtype=torch.float32, device=dev)
        ts = []
        for i in range(23):
            flow.zero_()
            flush.fill_(i & 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.assign(dmg, flow, 20, "fw")
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts[3:]))
        print(f"B={B:6d}  {ms:8.4f} ms  {B / ms / 1e3:8.3f} M assignments/s  config={eng.last_config}", flush=True)


if __name__ == "__main__":
    main()
