"""GPU tests of the drop-in env API: RepairEnv facade and BatchedRepairEnv against goldens
recorded from the reference env (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL_F32 = 1e-4


def _env(b200env, graph, **kw):
    return b200env.RepairEnv(graph, device="cuda:0", **kw)


def test_facade_reset_goldens(b200env, native_lib, sf_graph, golden):
    """SURVEY.md §4 known answers: seed=0 K=20 msa/fw/cfw reset + first-step TSTT and reward."""
    G = golden("sf_traces.npz")
    for method in ("msa", "fw", "cfw"):
        env = _env(b200env, sf_graph, seed=0, assignment_method=method, assignment_iters=20)
        key = f"s0_{method}"
        assert np.array_equal(np.flatnonzero(env.is_damaged > 0), G[f"{key}_dmg"])  # same RNG + retry rule
        assert abs(env.tstt - float(G[f"{key}_reset_tstt"])) / env.tstt <= RTOL_F32
        assert env.initial_tstt == env.tstt
        a = int(G[f"{key}_dmg"][0])
        state, reward, done, info = env.step(a)
        assert abs(info["tstt"] - float(G[f"{key}_step_tstt"])) / info["tstt"] <= RTOL_F32
        np.testing.assert_allclose(reward, float(G[f"{key}_step_reward"]), rtol=1e-3, atol=1e-6)
        assert done == bool(G[f"{key}_step_done"])
        assert state.node_features.shape == (24, 4) and state.edge_features.shape == (76, 6)
        assert state.action_mask[a] == 0.0 and env.capacities[a] == env.initial_capacities[a]


def test_facade_seeds_table(b200env, native_lib, sf_graph, golden):
    G = golden("sf_kat.npz")
    for method in ("msa", "fw"):
        got = [_env(b200env, sf_graph, seed=s, assignment_method=method).tstt for s in range(1, 6)]
        np.testing.assert_allclose(got, G[f"seeds1to5_{method}_K20_tstt"], rtol=RTOL_F32)


def test_facade_episode_curve_rewards_done(b200env, native_lib, sf_graph, golden):
    """Full repair episodes incl. one invalid action (reward -1, no assignment)."""
    G = golden("sf_episode.npz")
    for method, K, seed in (("fw", 50, 43), ("msa", 20, 7)):
        key = f"{method}_K{K}_s{seed}"
        env = _env(b200env, sf_graph, seed=seed, assignment_method=method, assignment_iters=K, unassigned_penalty=1e4)
        dmg = G[f"{key}_dmg"]
        assert np.array_equal(np.flatnonzero(env.is_damaged > 0), dmg)
        curve, rew, dones = [env.tstt], [], []
        for i, a in enumerate(dmg):
            if i == 3:
                _, r, d, info = env.step(int(dmg[0]))
                assert r == -1.0 and d is False
                curve.append(info["tstt"]); rew.append(r); dones.append(d)
            _, r, d, info = env.step(int(a))
            curve.append(info["tstt"]); rew.append(r); dones.append(d)
        np.testing.assert_allclose(curve, G[f"{key}_tstt"], rtol=RTOL_F32)
        np.testing.assert_allclose(rew, G[f"{key}_reward"], rtol=2e-3, atol=2e-5)
        assert dones == list(G[f"{key}_done"]) and dones[-1] is True
        np.testing.assert_allclose(env.flow, G[f"{key}_final_flow"], rtol=RTOL_F32, atol=1e-2)


def test_facade_state_features(b200env, native_lib, sf_graph, golden):
    G = golden("sf_features.npz")
    for seed in range(6):
        env = _env(b200env, sf_graph, seed=seed, assignment_method="fw", assignment_iters=20)
        st = env.get_state()
        np.testing.assert_allclose(st.node_features, G[f"s{seed}_node0"], rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(st.edge_features, G[f"s{seed}_edge0"], rtol=1e-4, atol=2e-6)
        for a in G[f"s{seed}_acts"]:
            st, _, _, _ = env.step(int(a))
        np.testing.assert_allclose(st.node_features, G[f"s{seed}_node3"], rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(st.edge_features, G[f"s{seed}_edge3"], rtol=1e-4, atol=2e-6)
        assert np.array_equal(st.action_mask, G[f"s{seed}_mask3"])
        assert abs(st.log_tstt - float(G[f"s{seed}_logtstt3"])) <= 1e-5
        assert st.edge_index is env.edge_index and st.goal_mask is not env.goal_mask
    np.testing.assert_allclose(env.betweenness_vec, G["static_betweenness"], rtol=1e-5, atol=1e-7)


def test_facade_errors_and_reward_modes(b200env, native_lib, sf_graph, golden):
    env = _env(b200env, sf_graph, seed=0)
    with pytest.raises(ValueError):
        env.step(76)
    with pytest.raises(ValueError):
        env.step(-1)
    env.assignment_iters = 0
    with pytest.raises(ValueError):
        env.compute_flow_assignment()
    with pytest.raises(NotImplementedError):
        _env(b200env, sf_graph, assignment_method="gp")
    G = golden("sf_kat.npz")
    env.initial_tstt = float(G["reward_initial_tstt"])
    gm = env.goal_mask
    for mode, row in zip(G["reward_modes"], G["reward_rows"]):
        prev, curr, complete, clip, expect = row
        dm = np.zeros_like(gm) if complete else gm
        got = env.compute_reward_with_goal(prev, curr, gm, dm, alpha=1.0, beta=10.0, gamma=0.1, mode=str(mode), clip=clip)
        assert got == pytest.approx(expect, rel=1e-12, abs=1e-12)


def test_facade_greedy_style_attribute_poking(b200env, native_lib, sf_graph):
    """The greedy baseline edits is_damaged/capacities, calls compute_flow_assignment and restores
    flow/tstt (reference src/baselines/__init__.py:43-67)."""
    env = _env(b200env, sf_graph, seed=3, assignment_method="msa")
    base_flow, base_tstt, base_dmg = env.flow.copy(), env.tstt, env.is_damaged.copy()
    cand = np.flatnonzero(env.is_damaged > 0)[:3]
    seen = []
    for a in cand:
        env.is_damaged[a] = 0.0
        env.capacities[a] = env.initial_capacities[a]
        env.compute_flow_assignment()
        seen.append(env.tstt)
        env.is_damaged[:] = base_dmg
        env.flow = base_flow.copy()
        env.tstt = base_tstt
    # replay through step() on a twin env: same numbers
    for a, expect in zip(cand, seen):
        twin = _env(b200env, sf_graph, seed=3, assignment_method="msa")
        _, _, _, info = twin.step(int(a))
        assert info["tstt"] == expect


def test_batched_env_matches_facade(b200env, native_lib, sf_graph):
    B = 16
    benv = b200env.BatchedRepairEnv(sf_graph, B, assignment_method="fw", assignment_iters=20, device="cuda:0",
                                    seeds=list(range(100, 100 + B)), reward_mode="log_delta")
    st = benv.reset()
    singles = [_env(b200env, sf_graph, seed=100 + b, assignment_method="fw", assignment_iters=20) for b in range(B)]
    assert np.array_equal(benv.is_damaged.cpu().numpy(), np.stack([(e.is_damaged > 0).astype(np.uint8) for e in singles]))
    assert np.array_equal(benv.tstt.cpu().numpy(), np.array([e.tstt for e in singles]))
    for _ in range(3):
        acts = []
        for b, e in enumerate(singles):
            dm = np.flatnonzero(e.is_damaged > 0)
            acts.append(int(dm[0]) if b % 4 else 0)  # every 4th env takes link 0 (maybe invalid)
        st, reward, done, info = benv.step(torch.tensor(acts, dtype=torch.int32))
        for b, e in enumerate(singles):
            s1, r1, d1, i1 = e.step(acts[b])
            assert info["tstt"][b].item() == i1["tstt"]
            assert reward[b].item() == pytest.approx(r1, rel=1e-12, abs=1e-12)
            assert bool(done[b]) == d1
            np.testing.assert_array_equal(st.node_features[b].cpu().numpy(), s1.node_features)
            np.testing.assert_array_equal(st.edge_features[b].cpu().numpy(), s1.edge_features)
    with pytest.raises(ValueError):
        benv.step(torch.full((B,), 76, dtype=torch.int32))


def test_host_entry_point_roundtrip(b200env, native_lib, sf_graph, golden):
    G = golden("sf_batch4096_fw_k20.npz")
    B = 512
    masks = np.zeros((B, 76), np.uint8)
    np.put_along_axis(masks, G["dmg"][:B].astype(np.int64), 1, axis=1)
    benv = b200env.BatchedRepairEnv(sf_graph, B, assignment_method="fw", assignment_iters=20, device="cuda:0")
    pinned = torch.from_numpy(masks).pin_memory()
    tstt = benv.assign_host(pinned.numpy())
    np.testing.assert_allclose(tstt, G["tstt_reset"][:B], rtol=RTOL_F32)


def test_batched_greedy_matches_per_candidate_oracle(b200env, native_lib, sf_graph, sf_oracle):
    """One-step greedy (reference src/baselines/__init__.py:35-69): every (scenario, damaged link)
    clone is one row of a single fused launch; the chosen link must be the argmin (first minimum)
    of the per-candidate TSTTs the C oracle computes from the same warm-start flows."""
    from importlib import import_module
    bl = import_module("sac-gat-her_transportationrl_b200.baselines")
    B = 6
    benv = b200env.BatchedRepairEnv(sf_graph, B, assignment_method="msa", assignment_iters=20, device="cuda:0",
                                    seeds=list(range(40, 40 + B)))
    benv.reset()
    flow_before, dmg_before = benv.flow.clone(), benv.is_damaged.clone()
    acts = bl.select_greedy_one_step(benv).cpu().numpy()
    assert torch.equal(benv.flow, flow_before) and torch.equal(benv.is_damaged, dmg_before)  # env untouched
    dmg, flow = benv.is_damaged.cpu().numpy(), benv.flow.cpu().numpy()
    for b in range(B):
        best, best_t = -1, float("inf")
        for a in np.flatnonzero(dmg[b]):
            m = dmg[b].copy()
            m[a] = 0
            t = sf_oracle.assign(m, flow[b], 20, "msa")["tstt"]
            if t < best_t:
                best, best_t = int(a), t
        assert acts[b] == best


def test_batched_episode_runner_and_simple_policies(b200env, native_lib, sf_graph):
    from importlib import import_module
    bl = import_module("sac-gat-her_transportationrl_b200.baselines")
    B = 8
    benv = b200env.BatchedRepairEnv(sf_graph, B, assignment_method="msa", assignment_iters=10, device="cuda:0",
                                    seeds=list(range(B)), damaged_ratio=0.1)
    first_damaged = lambda st: torch.argmax(st.action_mask, dim=1).to(torch.int32)  # noqa: E731
    res = bl.run_episode(benv, first_damaged, max_steps=50)
    k = max(1, int(76 * 0.1))
    assert torch.all(res["steps"] == k)  # every action repairs one damaged link
    assert res["tstt_curve"].shape[0] == k and torch.all(torch.isfinite(res["tstt_curve"]))
    # same episode, one scenario at a time through the facade
    for b in (0, 5):
        e = b200env.RepairEnv(sf_graph, seed=b, assignment_method="msa", assignment_iters=10, damaged_ratio=0.1,
                              device="cuda:0")
        st, curve, total = e.get_state(), [], 0.0
        done = False
        while not done:
            a = int(np.argmax(st.action_mask))
            st, r, done, info = e.step(a)
            curve.append(info["tstt"])
            total += r
        np.testing.assert_array_equal(res["tstt_curve"][:, b].cpu().numpy(), np.array(curve))
        assert res["reward"][b].item() == pytest.approx(total, rel=1e-12)
        assert res["tstt_auc"][b].item() == pytest.approx(float(np.trapezoid(curve)), rel=1e-12)
    g = torch.Generator(device="cuda:0").manual_seed(0)
    st = benv.reset()
    a = bl.select_random(st, generator=g)
    assert torch.all(st.action_mask[torch.arange(B), a.long()] > 0)
    for pol in (bl.select_max_vc, bl.select_max_flow):
        assert pol(st).shape == (B,)
    stuck = bl.run_episode(benv, lambda s: torch.zeros(B, dtype=torch.int32, device="cuda:0"), max_steps=3)
    assert stuck["tstt_curve"].shape[0] == 3  # a policy that never finishes stops at max_steps


@pytest.mark.parametrize("which", ["an", "cs"])
def test_state_features_large_networks_vs_oracle_betweenness(b200env, native_lib, which):
    """`features_kernel` beyond Sioux Falls: node feature 0 (Brandes betweenness of the undamaged
    subgraph / its maximum, reference src/env/repair_env.py:753-770) and the remaining node / link
    features against the C oracle's `orc_betweenness` and the reference's formulas (:771-807) for 8
    damage draws on the Anaheim-size network (3 on the Chicago-Sketch-size one)."""
    from oracle.c_oracle import COracle
    g = b200env.synthetic.anaheim_size(0) if which == "an" else b200env.synthetic.chicago_sketch_size(0)
    orc = COracle.from_graph_data(g)
    net = b200env.Network(g)
    eng = b200env.AssignmentEngine(net, "cuda:0", "fp32")
    B = 8 if which == "an" else 3
    E, N = eng.E, eng.N
    rng = np.random.default_rng(31)
    dmg = np.zeros((B, E), np.uint8)
    for b in range(B):
        dmg[b, rng.choice(E, int(E * 0.3), replace=False)] = 1
    dmg[0] = 0                                                     # the undamaged network too
    flow = (rng.random((B, E)) * 9000.0).astype(np.float32)
    goal = dmg.copy()
    goal[:, ::7] = 0
    tstt = rng.random(B) * 500.0 + 2.0
    d = torch.device("cuda", 0)
    node, edge, raw = eng.state_features(torch.from_numpy(flow).to(d), torch.from_numpy(dmg).to(d),
                                         torch.from_numpy(goal).to(d), torch.from_numpy(tstt).to(d), 1e-3,
                                         return_raw_betweenness=True)
    node, edge, raw = node.cpu().numpy(), edge.cpu().numpy(), raw.cpu().numpy()
    row, col = net.edge_index
    cap0, t0 = net.initial_capacities, net.t0
    for b in range(B):
        bw = orc.betweenness(dmg[b])
        np.testing.assert_allclose(raw[b], bw, rtol=2e-5, atol=1e-9)
        mx = bw.max()
        np.testing.assert_allclose(node[b, :, 0], bw / mx if mx > 0 else bw, rtol=2e-5, atol=1e-7)
        # remaining_ratio, avg_flow_norm, log10 TSTT (reference :771-790)
        cap = np.where(dmg[b] > 0, np.float32(1e-3), cap0)
        remaining = float(np.sum(goal[b].astype(np.float32) * dmg[b])) / max(float(goal[b].sum()), 1.0)
        avg_flow = float(np.mean(flow[b][dmg[b] == 0])) / max(net.total_demand / max(E, 1), 1.0)
        np.testing.assert_allclose(node[b, :, 1], np.float32(remaining), rtol=1e-6)
        np.testing.assert_allclose(node[b, :, 2], np.float32(avg_flow), rtol=1e-5)
        np.testing.assert_allclose(node[b, :, 3], np.float32(np.log10(max(tstt[b], 1.0))), rtol=1e-6)
        # link features (reference :792-807)
        exp_e = np.stack([
            np.log10(t0 + 1.0) / np.log10(net.max_t0 + 1.0),
            np.log10(cap + 1.0) / np.log10(net.max_capacity + 1.0),
            np.where(dmg[b] > 0, 0.0, np.clip(np.log1p(flow[b] / np.maximum(cap, 1e-6)), 0.0, 10.0)),
            dmg[b].astype(np.float32), goal[b].astype(np.float32),
            np.arange(E, dtype=np.float32) / max(E - 1, 1)], axis=1)
        np.testing.assert_allclose(edge[b], exp_e, rtol=1e-4, atol=2e-6)


def test_device_damage_sampler_rule_determinism_and_distribution(b200env, native_lib, sf_graph):
    """`b200_sample_damage` (reset's damage draw on the device, reference :167-190): every row gets exactly
    `count` damaged links; the undamaged links are strongly connected on the nodes they touch whenever the
    sampler says it found such a draw within 50 tries; same seed + counters => same masks; `select` leaves
    the other rows untouched; and the per-link damage frequencies agree with the host sampler (numpy
    Generator + the reference's retry rule) - same rule, same distribution, different random stream."""
    B = 4096
    env = b200env.BatchedRepairEnv(sf_graph, B, device="cuda:0", assignment_iters=2, seeds=list(range(B)), sampler_seed=77)
    net = env.network
    env.sample_damage_device(0.3)
    m1, tries = env.is_damaged.cpu().numpy().copy(), env._sampler_tries.cpu().numpy().copy()
    assert np.all(m1.sum(axis=1) == 22) and m1.max() == 1
    assert tries.min() >= 1 and np.median(tries) > 2          # ~13 % of the draws are accepted
    for b in range(0, B, 64):
        if tries[b] <= 50:
            assert net.active_subgraph_strongly_connected(m1[b].astype(np.float32)), b
    assert np.unique(m1, axis=0).shape[0] > B * 0.99          # the rows are different draws
    # determinism: a twin env with the same sampler seed reproduces the stream
    twin = b200env.BatchedRepairEnv(sf_graph, B, device="cuda:0", assignment_iters=2, seeds=list(range(B)), sampler_seed=77)
    twin.sample_damage_device(0.3)
    assert torch.equal(twin.is_damaged, env.is_damaged)
    # select: only the chosen rows are redrawn, and their stream moves on
    sel = torch.zeros(B, dtype=torch.uint8, device="cuda:0")
    sel[::3] = 1
    counters0 = env._sampler_counters.clone()
    env.sample_damage_device(0.3, sel)
    m2 = env.is_damaged.cpu().numpy()
    keep = sel.cpu().numpy() == 0
    assert np.array_equal(m2[keep], m1[keep]) and not np.array_equal(m2[~keep], m1[~keep])
    assert torch.equal(env._sampler_counters[1::3], counters0[1::3]) and bool((env._sampler_counters[::3] > counters0[::3]).all())
    # distribution: per-link damage frequency vs 3000 host draws of the reference rule
    rng = np.random.default_rng(5)
    host = np.zeros(76)
    n_host = 3000
    for _ in range(n_host):
        host[net.sample_damage(rng, 0.3)] += 1
    f_dev, f_host = m1.mean(axis=0), host / n_host
    se = np.sqrt(f_host * (1 - f_host) * (1 / n_host + 1 / B))
    assert np.abs(f_dev - f_host).max() < 5 * se.max(), (np.abs(f_dev - f_host).max(), se.max())
    assert f_dev.std() > 0.01                                  # connectivity makes some links much rarer than others


def test_reset_where_runs_on_the_device_and_fixed_damage_semantics(b200env, native_lib, sf_graph):
    """`reset_where` redraws, reassigns and re-bases only the selected scenarios, without touching the others;
    under `fixed_damage` with a seed every env of every rank shares ONE scenario (the reference's meaning,
    :79-81 and src/train.py:160-161), without a seed every env keeps the first draw of its own generator."""
    B = 64
    env = b200env.BatchedRepairEnv(sf_graph, B, device="cuda:0", assignment_method="fw", assignment_iters=20,
                                   seeds=list(range(300, 300 + B)))
    env.reset()
    for _ in range(2):
        env.step(torch.argmax(env.is_damaged, dim=1).to(torch.int32))
    before = {k: getattr(env, k).clone() for k in ("is_damaged", "goal_mask", "flow", "tstt", "initial_tstt")}
    mask = torch.zeros(B, dtype=torch.bool, device="cuda:0")
    mask[5::7] = True
    env.reset_where(mask)
    keep = ~mask
    for k, v in before.items():
        assert torch.equal(getattr(env, k)[keep], v[keep]), k
    assert torch.all(env.is_damaged[mask].sum(dim=1) == 22) and torch.equal(env.goal_mask[mask], env.is_damaged[mask])
    assert torch.equal(env.initial_tstt[mask], env.tstt[mask]) and not torch.equal(env.is_damaged[mask], before["is_damaged"][mask])
    # the re-based scenarios carry a genuine cold-start assignment of their new damage
    from oracle.c_oracle import COracle
    orc = COracle.from_graph_data(sf_graph)
    for b in torch.nonzero(mask).flatten().tolist()[:3]:
        o = orc.assign(env.is_damaged[b].cpu().numpy(), None, 20, "fw")
        assert env.tstt[b].item() == o["tstt"] and np.array_equal(env.flow[b].cpu().numpy(), o["flow"])
    # fixed damage with a seed: one shared scenario, identical on "another rank" (different env seeds)
    a = b200env.BatchedRepairEnv(sf_graph, 8, device="cuda:0", assignment_iters=2, fixed_damage=True, fixed_damage_seed=42,
                                 seeds=list(range(8)))
    b_ = b200env.BatchedRepairEnv(sf_graph, 8, device="cuda:0", assignment_iters=2, fixed_damage=True, fixed_damage_seed=42,
                                  seeds=list(range(8, 16)))
    a.reset(); b_.reset()
    ref = b200env.RepairEnv(sf_graph, device="cuda:0", assignment_iters=2, fixed_damage=True, fixed_damage_seed=42, seed=0)
    assert torch.equal(a.is_damaged, b_.is_damaged) and torch.all(a.is_damaged == a.is_damaged[0])
    assert np.array_equal(a.is_damaged[0].cpu().numpy(), (ref.is_damaged > 0).astype(np.uint8))
    a.step(torch.argmax(a.is_damaged, dim=1).to(torch.int32))
    a.reset_where(torch.ones(8, dtype=torch.bool, device="cuda:0"))
    assert torch.equal(a.is_damaged, b_.is_damaged)                      # the cached scenario again
    # fixed damage without a seed: per-env first draws, different between envs and between "ranks"
    c = b200env.BatchedRepairEnv(sf_graph, 8, device="cuda:0", assignment_iters=2, fixed_damage=True, seeds=list(range(8)))
    d = b200env.BatchedRepairEnv(sf_graph, 8, device="cuda:0", assignment_iters=2, fixed_damage=True, seeds=list(range(8, 16)))
    c.reset(); d.reset()
    assert not torch.equal(c.is_damaged, d.is_damaged) and not torch.all(c.is_damaged == c.is_damaged[0])
    first = c.is_damaged.clone()
    c.reset()
    assert torch.equal(c.is_damaged, first)


def test_feature_kernels_agree_on_small_networks(b200env, native_lib, sf_graph, monkeypatch):
    """Networks up to 32 nodes take the lane-per-source feature kernel (one warp per scenario);
    B200_FEATURES_GENERIC routes them through the warp-per-source kernel of the larger networks.  Both
    must give the same betweenness / features (summation orders differ: 1e-6), also on a lattice with
    dead-end-free one-way structure removed links, an undamaged and a heavily damaged network."""
    from oracle.c_oracle import COracle
    for g in (sf_graph, b200env.synthetic.small_grid(4, 5, 6, seed=3)):
        orc = COracle.from_graph_data(g)
        E = len(g.edges)
        rng = np.random.default_rng(17)
        B = 40
        dmg = (rng.random((B, E)) < 0.3).astype(np.uint8)
        dmg[0] = 0
        dmg[1] = (rng.random(E) < 0.8).astype(np.uint8)          # most links gone: several unreachable pairs
        flow = (rng.random((B, E)) * 5000).astype(np.float32)
        tstt = rng.random(B) * 300 + 1
        d = torch.device("cuda", 0)
        args = lambda: (torch.from_numpy(flow).to(d), torch.from_numpy(dmg).to(d), torch.from_numpy(dmg).to(d),  # noqa: E731
                        torch.from_numpy(tstt).to(d), 1e-3)
        small = b200env.AssignmentEngine(b200env.Network(g), d, "fp32").state_features(*args(), return_raw_betweenness=True)
        monkeypatch.setenv("B200_FEATURES_GENERIC", "1")
        generic = b200env.AssignmentEngine(b200env.Network(g), d, "fp32").state_features(*args(), return_raw_betweenness=True)
        monkeypatch.delenv("B200_FEATURES_GENERIC")
        for a, b_ in zip(small, generic):
            np.testing.assert_allclose(a.cpu().numpy(), b_.cpu().numpy(), rtol=2e-6, atol=1e-7)
        for b in (0, 1, 7, 39):
            np.testing.assert_allclose(small[2][b].cpu().numpy(), orc.betweenness(dmg[b]), rtol=2e-5, atol=1e-9)
