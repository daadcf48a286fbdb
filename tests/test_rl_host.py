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
    # new items carry (max_priority + eps) ** alpha and raise the running maximum by eps each, reference :50-54
    np.testing.assert_allclose(R.priority[:5].numpy(), (1.0 + 1e-6 * np.arange(1, 6)) ** 0.6, rtol=1e-6)
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


# ---------------------------------------------------------------------------------------------
# DeviceReplay against the reference's own sum-tree ReplayBuffer (src/train.py:27-91)
# ---------------------------------------------------------------------------------------------
class _TranscribedReplay:
    """Transcription of the semantics of the reference's ``ReplayBuffer`` (src/train.py:27-91) without
    the tree: new items enter at the running maximum priority, ``p = (|x| + eps) ** alpha``, sampling
    probability ``p / sum p``, weights ``(size * prob) ** -beta`` divided by their maximum, priority
    updates applied in order.  Used when the reference source is not on this machine."""

    def __init__(self, capacity, alpha=0.6, beta=0.4, eps=1e-6):
        self.capacity, self.alpha, self.beta, self.eps = capacity, alpha, beta, eps
        self.p = np.zeros(capacity, np.float32)
        self.max_priority, self.ptr, self.size = 1.0, 0, 0

    def add(self, item, priority=None):
        priority = float(abs(self.max_priority if priority is None else priority) + self.eps)
        self.max_priority = max(self.max_priority, priority)
        self.p[self.ptr] = np.float32(0) + np.float32(priority ** self.alpha - 0.0)
        self.ptr = (self.ptr + 1) % self.capacity
        self.size = min(self.size + 1, self.capacity)

    def update_priorities(self, indices, td_errors):
        for i, err in zip(indices, td_errors):
            priority = float(abs(err) + self.eps)
            self.max_priority = max(self.max_priority, priority)
            self.p[int(i)] = priority ** self.alpha

    def leaf_priorities(self):
        return self.p.copy()


def _reference_replay_cls():
    """The reference's ReplayBuffer class, compiled from ITS OWN source (the class body only: the module
    imports torch_geometric, which is not installed), or None when the source is not on this machine."""
    import ast
    from dataclasses import dataclass
    from typing import List
    for root in (os.path.join(REPO, "oracle", "_ref"), "/root/reference"):
        path = os.path.join(root, "src", "train.py")
        if os.path.exists(path):
            tree = ast.parse(open(path).read())
            node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ReplayBuffer")
            ns = {"np": np, "dataclass": dataclass, "List": List}
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
            return ns["ReplayBuffer"]
    return None


def _leaves(ref):
    return ref.tree[ref.capacity:ref.capacity + ref.capacity].copy() if hasattr(ref, "tree") else ref.leaf_priorities()


@pytest.mark.parametrize("which", ["reference", "transcription"])
def test_device_replay_matches_reference_sum_tree_semantics(which):
    cls = _reference_replay_cls() if which == "reference" else _TranscribedReplay
    if cls is None:
        pytest.skip("reference source (oracle/_ref or /root/reference) is not on this machine")
    cap, alpha, beta, eps = 64, 0.6, 0.4, 1e-6
    ref = cls(capacity=cap, alpha=alpha, beta=beta, eps=eps)
    R = _replay_cls()(cap, 5, 7, alpha=alpha, beta=beta, eps=eps, device="cpu")
    rng = np.random.default_rng(0)
    # 1) default-priority inserts: every one enters at the running maximum and raises it by eps
    for i in range(40):
        ref.add(("item", i))
    R.add_batch(_batch(40))
    np.testing.assert_allclose(R.priority.numpy()[:40], _leaves(ref)[:40], rtol=2e-7, atol=1e-7)
    assert R.max_priority == pytest.approx(ref.max_priority, rel=1e-12)
    # 2) priority updates incl. an index drawn twice (last occurrence wins) and a new maximum
    idx = np.array([3, 7, 3, 12, 39, 7, 0], np.int64)
    td = np.array([0.5, -2.0, 4.0, 0.0, 9.5, 1.25, -0.01])
    ref.update_priorities(idx, td)
    R.update_priorities(torch.from_numpy(idx), torch.from_numpy(td))
    np.testing.assert_allclose(R.priority.numpy()[:40], _leaves(ref)[:40], rtol=2e-7, atol=1e-7)
    assert R.priority[3].item() == pytest.approx((4.0 + eps) ** alpha, rel=1e-6)       # not 0.5: the later draw
    assert R.max_priority == pytest.approx(ref.max_priority, rel=1e-12) and ref.max_priority == pytest.approx(9.5 + eps)
    # 3) inserts after the update enter at the NEW maximum; the ring wraps
    for i in range(30):
        ref.add(("late", i))
    R.add_batch(_batch(30, base=500))
    np.testing.assert_allclose(R.priority.numpy(), _leaves(ref), rtol=2e-7, atol=1e-7)   # the tree accumulates fp32 deltas
    assert (R.ptr, R.size) == (ref.ptr, ref.size) == (6, 64)
    # 4) explicit priorities on insert (the workers' TD estimates, reference :51-54)
    pr = rng.random(5) * 3
    for x in pr:
        ref.add(("explicit",), priority=float(x))
    R.add_batch(_batch(5, base=900), priority=torch.from_numpy(pr))
    np.testing.assert_allclose(R.priority.numpy(), _leaves(ref), rtol=2e-7, atol=1e-7)   # the tree accumulates fp32 deltas
    # 5) importance weights: exactly (size * p / total) ** -beta / max over the batch
    g = torch.Generator().manual_seed(5)
    _, sidx, w = R.sample(256, generator=g)
    leaves = _leaves(ref).astype(np.float64)
    probs = leaves[sidx.numpy()] / leaves.sum()
    expect = (ref.size * probs) ** (-beta)
    np.testing.assert_allclose(w.numpy(), expect / expect.max(), rtol=1e-5)
    # 6) sampling distribution: chi-square of 200k draws against p / sum p (and against the reference's own
    #    tree descent when it is the real class)
    g = torch.Generator().manual_seed(11)
    draws = torch.cat([R.sample(4096, generator=g)[1] for _ in range(50)]).numpy()
    counts = np.bincount(draws, minlength=cap).astype(np.float64)
    expected = leaves / leaves.sum() * draws.size
    chi2 = float(((counts - expected) ** 2 / expected).sum())
    assert chi2 < 120.0, chi2          # 63 degrees of freedom: P(chi2 > 120) ~ 2e-5
    if which == "reference":
        np.random.seed(3)
        ref_draws = np.concatenate([ref.sample(2048)[1] for _ in range(20)])
        rc = np.bincount(ref_draws, minlength=cap).astype(np.float64)
        chi2_ref = float(((rc - expected * ref_draws.size / draws.size) ** 2 / (expected * ref_draws.size / draws.size)).sum())
        assert chi2_ref < 120.0, chi2_ref


def test_transcription_equals_reference_class_on_a_random_operation_sequence():
    cls = _reference_replay_cls()
    if cls is None:
        pytest.skip("reference source is not on this machine")
    a, b = cls(capacity=32), _TranscribedReplay(32)
    rng = np.random.default_rng(1)
    for step in range(300):
        if rng.random() < 0.6:
            pr = None if rng.random() < 0.5 else float(rng.normal() * 3)
            a.add(step, pr); b.add(step, pr)
        elif a.size:
            idx = rng.integers(0, a.size, 6)
            td = rng.normal(size=6) * 5
            a.update_priorities(idx, td); b.update_priorities(idx, td)
        np.testing.assert_allclose(_leaves(a), _leaves(b), rtol=3e-6, atol=1e-7)   # the tree accumulates fp32 deltas
        assert a.max_priority == b.max_priority and (a.ptr, a.size) == (b.ptr, b.size)


# ---------------------------------------------------------------------------------------------
# checkpoint compatibility with the reference's shipped model file (src/rl/sac.py:265-287)
# ---------------------------------------------------------------------------------------------
def _reference_checkpoint():
    for root in (os.path.join(REPO, "oracle", "_ref"), "/root/reference"):
        path = os.path.join(root, "history-data", "outputs1", "model_best.pt")
        if os.path.exists(path):
            return path
    return None


def test_reference_checkpoint_loads_into_discrete_sac():
    """`history-data/outputs1/model_best.pt` (hidden 64, embed 64, node_in 3, written by an older revision
    of the reference: no LayerNorm keys yet) loads with strict=False; every `encoder.layers.*`,
    `encoder.input_proj.*` and `edge_mlp.*` tensor of the file is consumed, and the only keys our modules
    miss are the LayerNorms that revision did not have."""
    path = _reference_checkpoint()
    if path is None:
        pytest.skip("reference checkpoint (oracle/_ref) is not on this machine")
    import b200env
    sac_mod = importlib.import_module("sac-gat-her_transportationrl_b200.rl.sac")
    state = torch.load(path, map_location="cpu", weights_only=True)
    dims = dict(node_in=state["actor"]["encoder.layers.0.lin.weight"].shape[1],
                edge_in=state["actor"]["encoder.layers.0.lin_edge.weight"].shape[1],
                hidden=state["actor"]["encoder.layers.0.att_src"].shape[2],
                embed=state["actor"]["encoder.layers.2.att_src"].shape[2])
    assert dims == dict(node_in=3, edge_in=6, hidden=64, embed=64)
    net = b200env.Network(b200env.load_sioux_falls())
    agent = sac_mod.DiscreteSAC(net, dims["node_in"], dims["edge_in"], dims["hidden"], dims["embed"], num_layers=3,
                                share_critic_encoder=False, device="cpu")
    with pytest.raises(RuntimeError):
        agent.load(path)                       # strict: the LayerNorm keys are missing in that revision
    report = agent.load(path, strict=False)
    for name in ("actor", "critic1", "critic2", "target1", "target2"):
        assert report[name]["unexpected"] == [], name                       # every tensor of the file was consumed
        missing = report[name]["missing"]
        assert missing and all(k.startswith(("node_norm.", "edge_norm.", "encoder.norms.")) for k in missing), missing
        own = getattr(agent, name).state_dict()
        for key, ref in state[name].items():
            assert torch.equal(own[key], ref), (name, key)
    assert float(agent.log_alpha) == pytest.approx(float(state["log_alpha"]))
    # our own save / load keeps the reference's file layout
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        agent.save(os.path.join(d, "m.pt"))
        again = torch.load(os.path.join(d, "m.pt"), map_location="cpu", weights_only=True)
        assert set(again) == set(state)
        assert set(state["actor"]) <= set(again["actor"])
