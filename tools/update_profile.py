"""Where a learner update spends its GPU time (torch profiler, kernel totals) and how much of its wall clock
the GPU is busy at all (launch-bound or not).  python tools/update_profile.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import b200env

b200env.build_native()
tr = b200env.rl.BatchedTrainer(b200env.load_sioux_falls(), dict(num_envs=1024, batch_start=256, buffer_size=100_000), device="cuda:0")
tr.run(12)
agent = tr.agent
batch, idx, w = tr.replay.sample(256, generator=tr.gen)
for _ in range(3):
    agent.update(batch, weights=w, alpha_max=2.5)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    agent.update(batch, weights=w, alpha_max=2.5)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 5 * 1e3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        agent.update(batch, weights=w, alpha_max=2.5)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
gpu_ms = sum(e.self_device_time_total for e in ev) / 5 / 1e3
n_k = sum(e.count for e in ev if e.self_device_time_total > 0) / 5
print(f"update wall {wall:.2f} ms, GPU busy {gpu_ms:.2f} ms, {n_k:.0f} GPU kernels/copies per update")
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=18, max_name_column_width=60))
st = tr.state
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        agent.select_action(st.node_features, st.edge_features, st.action_mask, generator=tr.gen)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
print(f"select_action GPU busy {sum(e.self_device_time_total for e in ev) / 5 / 1e3:.2f} ms, "
      f"{sum(e.count for e in ev if e.self_device_time_total > 0) / 5:.0f} kernels")
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=8, max_name_column_width=60))
