"""GPU tests of the learner side: actor/critic on the CUDA GAT kernel, SAC update, checkpoint
layout, the batched trainer loop (BASELINE.json config 5 at toy size)."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rl():
    return importlib.import_module("sac-gat-her_transportationrl_b200.rl")


def test_actor_critic_shapes_masking_and_reference_state_dict_layout(b200env, native_lib, sf_graph):
    rl = _rl()
    net = b200env.Network(sf_graph)
    dev = torch.device("cuda", 0)
    actor = rl.Actor(net, 4, 6, 32, 16, num_layers=3).to(dev)
    critic = rl.Critic(net, 4, 6, 32, 16, num_layers=3).to(dev)
    B = 5
    x, ea = torch.rand(B, 24, 4, device=dev), torch.rand(B, 76, 6, device=dev)
    mask = (torch.rand(B, 76, device=dev) > 0.6).float()
    mask[0] = 0  # no valid action: uniform, like PyG's softmax over an all -1e9 row
    logits, probs, attn = actor(x, ea, mask, return_attention=True)
    assert probs.shape == (B, 76) and attn.shape == (B, 76 + 24, 1)
    torch.testing.assert_close(probs.sum(dim=1), torch.ones(B, device=dev))
    assert torch.all(probs[1:][mask[1:] <= 0] < 1e-20)
    torch.testing.assert_close(probs[0], torch.full((76,), 1 / 76, device=dev))
    assert critic(x, ea).shape == (B, 76)
    keys = set(actor.state_dict())
    # the reference's (PyG GATConv) parameter names, src/rl/sac.py:23-33 + gat_encoder.py:20-31
    for k in ("node_norm.weight", "edge_norm.bias", "encoder.layers.0.lin.weight", "encoder.layers.0.lin_edge.weight",
              "encoder.layers.0.att_src", "encoder.layers.0.att_dst", "encoder.layers.0.att_edge",
              "encoder.layers.2.bias", "encoder.input_proj.weight", "encoder.norms.2.weight", "edge_mlp.0.weight",
              "edge_mlp.2.bias"):
        assert k in keys, k
    assert actor.state_dict()["encoder.layers.0.lin.weight"].shape == (4 * 32, 4)
    assert actor.state_dict()["edge_mlp.0.weight"].shape == (32, 16 * 4 + 6)


def _dense_reference_losses(agent, batch, weights):
    """The reference's loss formulas (src/rl/sac.py:190-224) written per graph with python loops."""
    with torch.no_grad():
        f = lambda k: batch[k].float()  # noqa: E731
        _, nprobs, _ = agent.actor(f("next_node_x"), f("next_edge_attr"), f("next_action_mask"))
        qn = torch.min(agent.target1(f("next_node_x"), f("next_edge_attr")), agent.target2(f("next_node_x"), f("next_edge_attr")))
        q1 = agent.critic1(f("node_x"), f("edge_attr"))
        q2 = agent.critic2(f("node_x"), f("edge_attr"))
        _, probs, _ = agent.actor(f("node_x"), f("edge_attr"), f("action_mask"))
        crit, act = 0.0, 0.0
        B = q1.shape[0]
        for b in range(B):
            v = sum(float(nprobs[b, e]) * (float(qn[b, e]) - float(agent.alpha) * np.log(float(nprobs[b, e]) + 1e-8))
                    for e in range(q1.shape[1]))
            tgt = float(batch["reward"][b]) + (1.0 - float(batch["done"][b])) * agent.gamma * v
            a = int(batch["action"][b])
            crit += float(weights[b]) * ((float(q1[b, a]) - tgt) ** 2 + (float(q2[b, a]) - tgt) ** 2)
            qa = torch.min(q1[b], q2[b])
            act += sum(float(probs[b, e]) * (float(agent.alpha) * np.log(float(probs[b, e]) + 1e-8) - float(qa[e]))
                       for e in range(q1.shape[1]))
        return crit / B, act / B


def test_sac_update_matches_per_graph_formulas_and_checkpoint_roundtrip(b200env, native_lib, sf_graph, tmp_path):
    rl = _rl()
    net = b200env.Network(sf_graph)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    agent = rl.DiscreteSAC(net, 4, 6, 32, 16, num_layers=3, lr=1e-3, grad_clip=1.0, share_critic_encoder=False, device=dev)
    B = 6
    g = torch.Generator(device=dev).manual_seed(3)
    batch = {
        "node_x": torch.rand(B, 24, 4, device=dev, generator=g), "edge_attr": torch.rand(B, 76, 6, device=dev, generator=g),
        "action_mask": (torch.rand(B, 76, device=dev, generator=g) > 0.5).to(torch.uint8),
        "next_node_x": torch.rand(B, 24, 4, device=dev, generator=g),
        "next_edge_attr": torch.rand(B, 76, 6, device=dev, generator=g),
        "next_action_mask": (torch.rand(B, 76, device=dev, generator=g) > 0.5).to(torch.uint8),
        "action": torch.randint(0, 76, (B,), device=dev, generator=g), "reward": torch.rand(B, device=dev, generator=g),
        "done": torch.tensor([0, 0, 1, 0, 1, 0], device=dev).float(),
    }
    w = torch.rand(B, device=dev, generator=g) + 0.5
    crit_ref, act_ref = _dense_reference_losses(agent, batch, w)
    before = [p.detach().clone() for p in agent.actor.parameters()]
    out = agent.update(batch, weights=w, alpha_max=2.5)
    assert float(out["critic_loss"]) == pytest.approx(crit_ref, rel=2e-4, abs=1e-6)
    assert float(out["actor_loss"]) == pytest.approx(act_ref, rel=2e-4, abs=1e-6)
    assert out["td_errors"].shape == (B,) and torch.all(torch.isfinite(out["td_errors"]))
    assert any(not torch.equal(a, b) for a, b in zip(before, agent.actor.parameters()))
    for s_, t_ in ((agent.critic1, agent.target1),):  # polyak: target moved a little towards the critic
        d = sum(float((p - tp).abs().sum()) for p, tp in zip(s_.parameters(), t_.parameters()))
        assert d > 0
    path = tmp_path / "model.pt"
    agent.save(str(path))
    state = torch.load(str(path), weights_only=True)
    assert set(state) == {"actor", "critic1", "critic2", "target1", "target2", "log_alpha"}  # reference :265-276
    other = rl.DiscreteSAC(net, 4, 6, 32, 16, num_layers=3, share_critic_encoder=False, device=dev)
    other.load(str(path), map_location=dev)
    for a, b in zip(agent.actor.parameters(), other.actor.parameters()):
        assert torch.equal(a, b)
    assert float(other.log_alpha) == pytest.approx(float(agent.log_alpha))


def test_sac_update_cuda_graph_replay_equals_eager(b200env, native_lib, sf_graph):
    """``cuda_graph=True``: three eager updates, then the four stages are captured and replayed.  Same batches into
    an eager twin -> same losses and parameters (Adam with device-side step counts rounds its bias corrections
    differently, hence the tolerance); a batch of another size falls back to eager."""
    rl = _rl()
    net = b200env.Network(sf_graph)
    dev = torch.device("cuda", 0)
    agents = []
    for graphed in (False, True):
        torch.manual_seed(5)
        agents.append(rl.DiscreteSAC(net, 4, 6, 32, 16, num_layers=3, lr=1e-3, grad_clip=1.0, share_critic_encoder=False,
                                     device=dev, cuda_graph=graphed))
    eager, graphed = agents
    for a, b in zip(eager.actor.parameters(), graphed.actor.parameters()):
        assert torch.equal(a, b)
    g = torch.Generator(device=dev).manual_seed(11)

    def make(B):
        return {
            "node_x": torch.rand(B, 24, 4, device=dev, generator=g), "edge_attr": torch.rand(B, 76, 6, device=dev, generator=g),
            "action_mask": (torch.rand(B, 76, device=dev, generator=g) > 0.4).to(torch.uint8),
            "next_node_x": torch.rand(B, 24, 4, device=dev, generator=g),
            "next_edge_attr": torch.rand(B, 76, 6, device=dev, generator=g),
            "next_action_mask": (torch.rand(B, 76, device=dev, generator=g) > 0.4).to(torch.uint8),
            "action": torch.randint(0, 76, (B,), device=dev, generator=g), "reward": torch.rand(B, device=dev, generator=g),
            "done": (torch.rand(B, device=dev, generator=g) > 0.8).float(),
        }, torch.rand(B, device=dev, generator=g) + 0.5

    for it in range(8):
        batch, w = make(16 if it != 6 else 9)   # update 6 has another shape: eager on both
        oe = eager.update(batch, weights=w, alpha_max=2.5)
        og = graphed.update(batch, weights=w, alpha_max=2.5)
        for k in ("critic_loss", "actor_loss", "alpha_loss", "alpha", "q_mean"):
            assert float(og[k]) == pytest.approx(float(oe[k]), rel=2e-3, abs=2e-5), (it, k)
        assert torch.allclose(og["td_errors"], oe["td_errors"], rtol=2e-3, atol=2e-5), it
    assert graphed._graph is not None and len(graphed._graph["graphs"]) == 4 and eager._graph is None
    # Adam's normalised step turns a rounding-level gradient difference on a near-zero gradient (the float atomics of
    # index_add_ already differ from run to run) into a difference of up to one step size (lr) per update for that
    # element; the losses above are the real check, the parameters must agree to within those few step sizes
    for (n, a), b in zip(eager.actor.named_parameters(), graphed.actor.parameters()):
        assert float((a - b).abs().max()) < 8 * 2e-3, n
        if a.numel() >= 256:
            assert float((a - b).norm() / a.norm()) < 0.05, n
    for a, b in zip(eager.target1.parameters(), graphed.target1.parameters()):
        assert torch.allclose(a, b, rtol=0, atol=2e-4)    # polyak with tau 0.005: 0.5 % of the critics' difference


def test_batched_trainer_runs_episodes_updates_and_resets(b200env, native_lib, sf_graph):
    rl = _rl()
    cfg = dict(num_envs=32, assignment_iters=5, assignment_method="msa", buffer_size=4096, batch_start=64, update_every=2,
               batch_size=32, hidden_dim=32, embed_dim=16, max_steps=6, her_ratio=0.5, fixed_damage=False,
               damaged_ratio=0.1, reward_mode="log_delta", reward_clip=0.0, reward_beta=10.0, reward_gamma=0.1)
    tr = rl.BatchedTrainer(sf_graph, cfg, device="cuda:0")
    res = tr.run(12)
    assert res["env_steps_per_s"] > 0 and res["updates_per_s"] > 0 and res["episodes"] >= 32
    assert tr.replay.size == 12 * 32
    assert all(np.isfinite(v) for v in res["losses"].values())
    assert int(tr.ep_steps.max()) <= 6
    # a reset scenario is a fresh episode: damage restored to 7 links, flows re-assigned
    assert torch.all(tr.env.is_damaged.sum(dim=1) <= 7) and torch.all(tr.env.tstt > 0)


def test_evaluate_policies_on_common_scenarios(b200env, native_lib, sf_graph):
    """eval.py-style comparison: trained-policy wrapper, greedy and a simple heuristic on the same
    damage draws; greedy must not end worse than 'first damaged link' on average."""
    rl = _rl()
    bl = importlib.import_module("sac-gat-her_transportationrl_b200.baselines")
    B = 12
    env = b200env.BatchedRepairEnv(sf_graph, B, assignment_method="msa", assignment_iters=8, device="cuda:0",
                                   seeds=list(range(B)), damaged_ratio=0.08)
    agent = rl.DiscreteSAC(env.network, 4, 6, 32, 16, device="cuda:0")
    pols = {"actor": bl.make_actor_policy(agent), "greedy": lambda s: bl.select_greedy_one_step(env),
            "first": lambda s: torch.argmax(s.action_mask, dim=1).to(torch.int32)}
    table = bl.evaluate(env, pols, max_steps=20)
    assert set(table) == {"actor", "greedy", "first"}
    k = max(1, int(76 * 0.08))
    assert all(t["steps"] == k for t in table.values())            # masked policies always repair a damaged link
    assert table["greedy"]["tstt_auc"] <= table["first"]["tstt_auc"] * (1 + 1e-9)


# ---------------------------------------------------------------------------------------------
# the batched trainer on two NCCL ranks (BASELINE.json config 5; reference src/train.py:138-214, :730-913)
# ---------------------------------------------------------------------------------------------
def _nccl_trainer_worker(rank, world, port, out_dir):
    import os
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, repo)
    import torch.distributed as dist
    import b200env
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cfg = dict(num_envs=32, assignment_iters=8, assignment_method="cfw", batch_start=64, buffer_size=4096, batch_size=32,
                   update_every=2, hidden_dim=16, embed_dim=16, fixed_damage=False, max_steps=6, seed=5)
        tr = b200env.rl.BatchedTrainer(b200env.load_sioux_falls(), cfg, device=f"cuda:{rank}")
        res = tr.run(12)
        R = tr.replay
        state = {"masks0": tr.env.is_damaged.cpu(), "size": R.size, "ptr": R.ptr,
                 "store": {k: v[:R.size].cpu() for k, v in R.store.items()}, "priority": R.priority[:R.size].cpu(),
                 "max_priority": R.max_priority, "updates": tr.stats["updates"],
                 "actor": {k: v.cpu() for k, v in tr.agent.actor.state_dict().items()},
                 "critic1": {k: v.cpu() for k, v in tr.agent.critic1.state_dict().items()},
                 "log_alpha": tr.agent.log_alpha.detach().cpu(), "res": res}
        torch.save(state, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_batched_trainer_two_nccl_ranks_stay_in_lock_step(native_lib, tmp_path):
    """Two ranks, NCCL: the per-step transition all-gather gives both replays the same contents in the same
    order, the all-gathered priority updates keep the priorities equal, the gradient all-reduce keeps the
    weights equal after every update - while each rank rolls out its OWN scenarios (rank-offset seeds)."""
    import torch.multiprocessing as mp
    port = 29000 + (int(torch.randint(0, 2000, ()).item()))
    mp.spawn(_nccl_trainer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = (torch.load(tmp_path / f"rank{r}.pt", weights_only=False) for r in (0, 1))
    assert a["updates"] == b["updates"] > 0 and a["size"] == b["size"] == 2 * 32 * 12 and a["ptr"] == b["ptr"]
    for k in a["store"]:
        assert torch.equal(a["store"][k], b["store"][k]), k                      # identical replay contents
    assert torch.equal(a["priority"], b["priority"]) and a["max_priority"] == b["max_priority"]
    for mod in ("actor", "critic1"):
        for k in a[mod]:
            torch.testing.assert_close(a[mod][k], b[mod][k], rtol=0, atol=0, msg=f"{mod}.{k}")   # identical weights
    assert torch.equal(a["log_alpha"], b["log_alpha"])
    assert not torch.equal(a["masks0"], b["masks0"])                           # different scenarios per rank
    # rank 0's half of every gathered step comes first: the first 32 stored actions are rank 0's own
    assert a["res"]["world_size"] == 2 and a["res"]["nccl_time_share"] >= 0.0
