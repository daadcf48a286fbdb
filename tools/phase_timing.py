"""Per-phase SM clocks of the general fused kernel (tuning build).  Build and run on the GPU box:
  B200_LIB=/tmp/libphase.so B200_NVCC_EXTRA=-DB200_PHASE_TIMING python tools/phase_timing.py an:128 cs:32
Phases (CTA 0, thread 0, summed over the K iterations of its first scenario... of all its scenarios):
0 zero loads + barrier, 1 customization, 2 warp 0's origins, 3 wait for the other warps, 4 cluster barrier,
5 cluster combine, 6 flow update + costs."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import b200env

NAMES = ["zero+bar", "customize", "own origins", "wait warps", "cluster bar", "cluster combine", "update"]
CASES = {"an": (lambda: b200env.synthetic.anaheim_size(0), 20, "fw"), "cs": (lambda: b200env.synthetic.chicago_sketch_size(0), 20, "msa")}


def main():
    b200env.build_native()
    lib = b200env._native.lib()
    dev = torch.device("cuda", 0)
    if os.environ.get("PHASE_SKIP_AUX"):
        lib.b200_debug_flags(1)
    for arg in sys.argv[1:]:
        name, B = arg.split(":")
        B = int(B)
        make, K, method = CASES[name]
        eng = b200env.AssignmentEngine(b200env.Network(make()), dev, "fp32")
        rng = np.random.default_rng(0)
        masks = np.zeros((B, eng.E), np.uint8)
        for b in range(B):
            masks[b, rng.choice(eng.E, max(1, int(eng.E * 0.3)), replace=False)] = 1
        dmg = torch.from_numpy(masks).to(dev)
        flow = torch.zeros((B, eng.E), dtype=torch.float32, device=dev)
        eng.assign(dmg, flow, K, method)
        out = (ctypes.c_ulonglong * 16)()
        lib.b200_debug_phase_clocks(out, 1)
        flow.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        eng.assign(dmg, flow, K, method)
        e.record()
        torch.cuda.synchronize()
        lib.b200_debug_phase_clocks(out, 1)
        tot = sum(out[:7])
        print(f"{name} B={B} {s.elapsed_time(e):.3f} ms config={eng.last_config}  CTA 0: {tot / 1.965e6:.3f} ms of SM clocks")
        for n, v in zip(NAMES, out):
            print(f"   {n:16s} {v / K / 1.965e3:9.2f} us/iteration  {100.0 * v / max(tot, 1):5.1f} %")
        print(f"      chain walk: {out[13] / K:.1f} loop iterations per iteration of warp 0 (its origins together), "
              f"{out[14] / max(out[13], 1):.1f} active lanes on average, {out[12] / max(out[13], 1):.0f} clocks per loop iteration")
        sub = list(out)[8:13]
        for n, v in zip(["init + up phase", "down sweep", "check + select", "scatter/count/leaves", "chain walk"], sub):
            print(f"      origin solve: {n:22s} {100.0 * v / max(sum(sub), 1):5.1f} %  ({v / K / 1.965e3:8.2f} us/iteration)")


if __name__ == "__main__":
    main()
