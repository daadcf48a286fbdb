import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, b200env
E = b200env.EdgeData
rng = np.random.default_rng(8)
n = 70
links = {(i, i + 1) for i in range(1, n - 1)} | {(i + 1, i) for i in range(1, n - 1) if i % 3}
links |= {(int(a), int(b)) for a, b in rng.integers(1, n, size=(60, 2)) if a != b}
links = sorted(links)
g = b200env.GraphData(num_nodes=n, edges=[E(u, v, 900.0, float(1 + (u * 7 + v) % 5), float(1 + (u * 7 + v) % 5), 0.15, 4.0) for u, v in links],
                      od_demand={(int(o), int(d)): float(10 * (1 + (o + d) % 4)) for o in (1, 5, 9, 30, 44) for d in (2, 17, 33, 52, n)})
rng2 = np.random.default_rng(5)
masks = np.zeros((6, len(links)), np.uint8)
for b in range(6):
    masks[b, rng2.choice(len(links), max(1, int(len(links) * 0.2)), replace=False)] = 1
eng = b200env.AssignmentEngine(b200env.Network(g), "cuda:0", "fp32", unassigned_penalty=1e4)
dev = torch.device("cuda:0")
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
out = eng.assign(torch.from_numpy(masks).to(dev), torch.zeros((6, len(links)), device=dev), K, "msa")
torch.cuda.synchronize()
print("OK", eng.last_config, out["tstt"].cpu().numpy()[:3], out["unassigned"].cpu().numpy()[:3], dict(E=len(links)))
