"""Host-side (CPU tensor) tests of the device replay: the same torch code runs on CUDA in
production; here it is exercised on CPU tensors against the reference's sum-tree semantics
(src/train.py:27-91)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def _replay_cls():
    return importlib.import_module("sac-gat-her_transportationrl_b200.rl.replay").DeviceReplay


def _batch(n, N=5, E=7, base=0):
    g = torch.Generator().manual_seed(base)
    return {
        "node_x": torch.rand(n, N, 4, generator=g), "edge_attr": torch.rand(n, E, 6, generator=g),
        "action_mask": (torch.rand(n, E, generator=g) > 0.5).to(torch.uint8),
        "next_node_x": torch.rand(n, N, 4, generator=g), "next_edge_attr": torch.rand(n, E, 6, generator=g),
        "next_action_mask": (torch.rand(n, E, generator=g) > 0.5).to(torch.uint8),
        "action": torch.arange(base, base + n), "reward": torch.arange(n).float(), "done": torch.zeros(n),
        "prev_tstt": torch.full((n,), 10.0, dtype=torch.float64), "next_tstt": torch.full((n,), 9.0, dtype=torch.float64),
        "initial_tstt": torch.full((n,), 12.0, dtype=torch.float64),
    }


def test_ring_buffer_and_max_priority_for_new_items():
    R = _replay_cls()(8, 5, 7, device="cpu")
    s0 = R.add_batch(_batch(5))
    assert s0.tolist() == [0, 1, 2, 3, 4] and R.size == 5 and R.ptr == 5
    # new items carry (max_priority + eps) ** alpha, reference :50-54
    np.testing.assert_allclose(R.priority[:5].numpy(), (1.0 + 1e-6) ** 0.6, rtol=1e-6)
    s1 = R.add_batch(_batch(5, base=100))
    assert s1.tolist() == [5, 6, 7, 0, 1] and R.size == 8 and R.ptr == 2
    assert R.store["action"].tolist() == [103, 104, 2, 3, 4, 100, 101, 102]
    with pytest.raises(ValueError):
        _replay_cls()(4, 5, 7, device="cpu").sample(2)


def test_proportional_sampling_weights_and_priority_update():
    R = _replay_cls()(16, 5, 7, alpha=0.6, beta=0.4, device="cpu")
    R.add_batch(_batch(10))
    td = torch.tensor([0.0, 1.0, 2.0, 4.0, 8.0, 0.5, 0.25, 3.0, 6.0, 10.0])
    R.update_priorities(torch.arange(10), td)
    p = (td.numpy() + 1e-6) ** 0.6
    np.testing.assert_allclose(R.priority[:10].numpy(), p, rtol=1e-6)
    assert R.max_priority == pytest.approx(10.0 + 1e-6)
    g = torch.Generator().manual_seed(0)
    counts = np.zeros(10)
    for _ in range(200):
        batch, idx, w = R.sample(64, generator=g)
        counts += np.bincount(idx.numpy(), minlength=10)
        probs = p[idx.numpy()] / p.sum()
        expect = (10 * probs) ** (-0.4)
        np.testing.assert_allclose(w.numpy(), expect / expect.max(), rtol=1e-5)
        assert torch.equal(batch["action"], R.store["action"][idx])
    freq = counts / counts.sum()
    np.testing.assert_allclose(freq, p / p.sum(), atol=0.01)


def test_hindsight_relabel_overwrites_last_feature_column_like_the_reference():
    R = _replay_cls()
    b = _batch(64)

    def reward_fn(prev, nxt, init, next_mask, goal):
        left = (goal.bool() & next_mask.bool()).sum(dim=1)
        return (prev - nxt) * 2.0, (left == 0)

    out = R.relabel_hindsight(b, 1.0, reward_fn, reward_scale=0.5, generator=torch.Generator().manual_seed(1))
    goal = (1 - b["next_action_mask"]).float()
    assert torch.equal(out["edge_attr"][:, :, -1], goal) and torch.equal(out["next_edge_attr"][:, :, -1], goal)
    assert torch.equal(out["edge_attr"][:, :, :-1], b["edge_attr"][:, :, :-1])
    np.testing.assert_allclose(out["reward"].numpy(), 1.0)         # (10 - 9) * 2 * 0.5
    assert torch.all(out["done"] == 1.0)                           # goal = already repaired links -> complete
    same = R.relabel_hindsight(b, 0.0, reward_fn)
    assert same is b
