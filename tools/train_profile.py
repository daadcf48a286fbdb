"""Where a lock-step training step spends its time (CUDA-synchronised wall clock per phase)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import b200env

b200env.build_native()
tr = b200env.rl.BatchedTrainer(b200env.load_sioux_falls(), dict(num_envs=1024, batch_start=256, buffer_size=100_000), device="cuda:0")
tr.run(24)   # past the CUDA-graph capture of the learner update (its 4th call)
env, agent = tr.env, tr.agent


def timed(fn, n=10):
    fn()   # one untimed call: lazy initialisation, graph capture
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, out


st = tr.state
ms, out = timed(lambda: agent.select_action(st.node_features, st.edge_features, st.action_mask, generator=tr.gen))
print(f"select_action      {ms:7.3f} ms")
acts = out.action.to(torch.int32)
ms, _ = timed(lambda: env.compute_flow_assignment())
print(f"assignment (cfw60) {ms:7.3f} ms")
ms, _ = timed(lambda: env.get_state())
print(f"get_state          {ms:7.3f} ms")
ms, _ = timed(lambda: env.step(torch.zeros_like(acts)))
print(f"env.step (invalid) {ms:7.3f} ms")
batch, idx, w = tr.replay.sample(256, generator=tr.gen)
ms, _ = timed(lambda: tr.replay.sample(256, generator=tr.gen))
print(f"replay.sample      {ms:7.3f} ms")
ms, _ = timed(lambda: agent.update(batch, weights=w, alpha_max=2.5), n=5)
print(f"agent.update       {ms:7.3f} ms")
ms, _ = timed(lambda: tr.step(), n=12)
print(f"trainer.step       {ms:7.3f} ms")
