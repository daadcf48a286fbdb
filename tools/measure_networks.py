"""Device-timed throughput of the fused assignment on the named network sizes (not a bench line;
numbers quoted in DESIGN.md).  python tools/measure_networks.py [sf an cs]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import b200env

CASES = {
    "sf": (b200env.load_sioux_falls, 4096, 20, "fw"),
    "an": (lambda: b200env.synthetic.anaheim_size(0), 1024, 20, "fw"),
    "cs": (lambda: b200env.synthetic.chicago_sketch_size(0), 256, 20, "msa"),
}


def main():
    dev = torch.device("cuda", 0)
    for arg in (sys.argv[1:] or ["sf", "an", "cs"]):   # "cs" or "cs:32" (scenarios per launch)
        name = arg.split(":")[0]
        make, B, K, method = CASES[name]
        if ":" in arg:
            B = int(arg.split(":")[1])
        g = make()
        net = b200env.Network(g)
        for precision in ("fp32", "fp64"):
            eng = b200env.AssignmentEngine(net, dev, precision)
            rng = np.random.default_rng(0)
            masks = np.zeros((B, eng.E), np.uint8)
            k = max(1, int(eng.E * 0.3))
            for b in range(B):
                masks[b, rng.choice(eng.E, k, replace=False)] = 1
            dmg = torch.from_numpy(masks).to(dev)
            flow = torch.zeros((B, eng.E), dtype=eng.real, device=dev)
            reps = 3 if name != "sf" else 10
            eng.assign(dmg, flow, K, method)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(reps):
                flow.zero_()
                out = eng.assign(dmg, flow, K, method)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / reps
            bytes_per = K * ((33 if precision == "fp32" else 65) * eng.E + 8 * eng.O * eng.N)
            print(json.dumps({"network": name, "kernel": eng.kernel_name, "config": eng.last_config, "cluster_env": os.environ.get("B200_CLUSTER", "auto"), "N": eng.N, "E": eng.E, "O": eng.O, "B": B, "K": K, "method": method,
                              "precision": precision, "ms": round(ms, 3), "assignments_per_s": round(B / ms * 1e3, 1),
                              "algorithmic_GBps": round(bytes_per * B / ms / 1e6, 2),
                              "median_gap": float(out["gap"].median())}), flush=True)


if __name__ == "__main__":
    main()
