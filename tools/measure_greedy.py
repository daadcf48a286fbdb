"""One-step greedy baseline (reference src/baselines/__init__.py:35-69, 22.9 s per Sioux Falls episode
on the reference's CPU path): whole episodes for B scenarios at once.  python tools/measure_greedy.py [B]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import b200env

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
env = b200env.BatchedRepairEnv(b200env.load_sioux_falls(), B, assignment_method="msa", assignment_iters=20,
                               device="cuda:0", seeds=list(range(B)))
bl = b200env.baselines
pol = lambda s: bl.select_greedy_one_step(env)  # noqa: E731
bl.run_episode(env, pol, max_steps=2)
torch.cuda.synchronize()
a0 = env.engine.launches
t0 = time.perf_counter()
res = bl.run_episode(env, pol)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
k = int(res["steps"].max())
cand = B * sum(range(1, k + 1))
print(json.dumps({"scenarios": B, "steps_per_episode": k, "episode_wall_s": wall, "episodes_per_s": B / wall,
                  "candidate_assignments": cand + B * (k + 1), "assignments_per_s": (cand + B * (k + 1)) / wall,
                  "mean_tstt_auc": float(res["tstt_auc"].mean()), "kernel_launches": env.engine.launches - a0}))
