"""CPU tests: the oracle (numpy/scipy port and C restatement) against golden vectors recorded
from the reference itself.  This is what pins the oracle (the reference has no
tests of its own for this path)."""
import numpy as np
import pytest

from oracle import ref_port as rp

METHODS = ("msa", "fw", "cfw")


@pytest.fixture(scope="module")
def net(sf_graph):
    return rp.PortNetwork.from_graph_data(sf_graph)


def _mask(idx, E=76):
    m = np.zeros(E, np.float32)
    m[idx] = 1.0
    return m


def _cap(net, dmg):
    cap = net.cap0.copy()
    cap[dmg > 0] = 1e-3
    return cap


@pytest.mark.parametrize("method", METHODS)
def test_port_reproduces_reference_traces(net, golden, method):
    """Same numpy/scipy calls in the same order: per-iteration costs, trees, AON loads, flows.
    Trees/AON are integer-valued and must match exactly; costs/flows/TSTT to 1e-6 (bit-equal on a
    host with the same numpy SIMD dispatch as tests/golden/meta.json)."""
    G = golden("sf_traces.npz")
    for seed in (0, 1, 2):
        key = f"s{seed}_{method}"
        dmg = _mask(G[f"{key}_dmg"])
        tr = {}
        flow, tstt, un = rp.assign(net, _cap(net, dmg), dmg, None, 20, method, trace=tr)
        assert np.array_equal(np.stack(tr["pred"]), G[f"{key}_reset_pred"])
        assert np.array_equal(np.stack(tr["aux"]), G[f"{key}_reset_aux"])
        np.testing.assert_allclose(np.stack(tr["t"]), G[f"{key}_reset_t"], rtol=1e-6)
        np.testing.assert_allclose(flow, G[f"{key}_reset_flow"][-1], rtol=1e-6, atol=1e-3)
        assert tstt == pytest.approx(float(G[f"{key}_reset_tstt"]), rel=1e-6)
        assert un == 0.0
        # warm-start step
        dmg2 = dmg.copy()
        dmg2[G[f"{key}_dmg"][0]] = 0.0
        flow2, tstt2, _ = rp.assign(net, _cap(net, dmg2), dmg2, flow, 20, method)
        assert tstt2 == pytest.approx(float(G[f"{key}_step_tstt"]), rel=1e-6)


@pytest.mark.parametrize("method", METHODS)
def test_c_oracle_reproduces_reference_traces(sf_oracle, golden, method):
    G = golden("sf_traces.npz")
    for seed in (0, 1, 2):
        key = f"s{seed}_{method}"
        dmg = _mask(G[f"{key}_dmg"]).astype(np.uint8)
        r = sf_oracle.assign(dmg, None, 20, method, trace=True)
        assert np.array_equal(r["pred"], G[f"{key}_reset_pred"].astype(np.int32)), "trees (every iteration)"
        assert np.array_equal(r["aux"], G[f"{key}_reset_aux"]), "AON link loads (every iteration)"
        np.testing.assert_allclose(r["flow"], G[f"{key}_reset_flow"][-1], rtol=1e-6, atol=1e-3)
        assert r["tstt"] == pytest.approx(float(G[f"{key}_reset_tstt"]), rel=1e-6)
        r64 = sf_oracle.assign(dmg, None, 20, method, dtype=np.float64)
        assert r64["tstt"] == pytest.approx(float(G[f"{key}_reset_tstt"]), rel=1e-6 if method != "cfw" else 5e-6)
        # warm start from the reference's flow
        dmg2 = dmg.copy()
        dmg2[G[f"{key}_dmg"][0]] = 0
        r2 = sf_oracle.assign(dmg2, G[f"{key}_reset_flow"][-1], 20, method, trace=True)
        assert np.array_equal(r2["pred"], G[f"{key}_step_pred"].astype(np.int32))
        assert r2["tstt"] == pytest.approx(float(G[f"{key}_step_tstt"]), rel=1e-6)


def test_oracles_known_answers_undamaged(sf_oracle, net, golden):
    G = golden("sf_kat.npz")
    zero = np.zeros(76, np.uint8)
    for method in METHODS:
        for K in (20, 50, 60, 200):
            ref = float(G[f"undamaged_{method}_K{K}_tstt"])
            r = sf_oracle.assign(zero, None, K, method)
            assert r["tstt"] == pytest.approx(ref, rel=2e-6), (method, K)
            np.testing.assert_allclose(r["flow"], G[f"undamaged_{method}_K{K}_flow"], rtol=1e-5, atol=1e-2)
    # SURVEY.md §4 table
    assert float(G["undamaged_msa_K20_tstt"]) == pytest.approx(22.582355796, rel=1e-9)
    assert float(G["undamaged_fw_K200_tstt"]) == pytest.approx(20.727071547, rel=1e-9)
    _, tstt, _ = rp.assign(net, net.cap0.copy(), np.zeros(76, np.float32), None, 20, "msa")
    assert tstt == pytest.approx(float(G["undamaged_msa_K20_tstt"]), rel=1e-7)


def test_c_oracle_batch_subset_vs_reference(sf_oracle, golden):
    G = golden("sf_batch4096_fw_k20.npz")
    for b in range(0, 512, 8):
        dmg = _mask(G["dmg"][b]).astype(np.uint8)
        r = sf_oracle.assign(dmg, None, 20, "fw")
        assert r["tstt"] == pytest.approx(float(G["tstt_reset"][b]), rel=1e-6)
        np.testing.assert_allclose(r["flow"], G["flow_reset"][b], rtol=1e-6, atol=1e-3)
        dmg[G["dmg"][b][0]] = 0
        r2 = sf_oracle.assign(dmg, r["flow"], 20, "fw")
        assert r2["tstt"] == pytest.approx(float(G["tstt_step"][b]), rel=1e-6)


def test_c_oracle_anaheim_and_chicago_size(b200env, golden):
    from oracle.c_oracle import COracle
    G = golden("an_traces.npz")
    g = b200env.synthetic.anaheim_size(0)
    orc = COracle.from_graph_data(g)
    for key in ("s0_fw", "s1_msa"):
        dmg = _mask(G[f"{key}_dmg"], 914).astype(np.uint8)
        r = orc.assign(dmg, None, 5, key.split("_")[1], trace=True)
        assert np.array_equal(r["pred"], G[f"{key}_pred"].astype(np.int32))
        assert np.array_equal(r["aux"], G[f"{key}_aux"])
        assert r["tstt"] == pytest.approx(float(G[f"{key}_tstt"]), rel=1e-6)
    C = golden("cs_k2.npz")
    gc = b200env.synthetic.chicago_sketch_size(0)
    assert len(gc.od_demand) == 140374 or abs(len(gc.od_demand) - 140000) < 2000
    orc = COracle.from_graph_data(gc)
    r = orc.assign(_mask(C["dmg"], 2950).astype(np.uint8), None, 2, "msa", trace=True)
    assert np.array_equal(r["pred"][:, :8, :], C["pred"].astype(np.int32))
    assert np.array_equal(r["aux"], C["aux"])
    assert r["tstt"] == pytest.approx(float(C["tstt"]), rel=1e-6)


def test_pairwise_sum_is_numpy_order(sf_oracle):
    rng = np.random.default_rng(0)
    for n in (0, 3, 8, 76, 77, 128, 129, 914, 2950):
        for _ in range(20):
            a = (rng.random(n) * 1e5).astype(np.float32)
            assert np.float32(sf_oracle.pairwise_sum(a)) == np.sum(a)
            a64 = rng.random(n) * 1e5
            assert sf_oracle.pairwise_sum(a64) == np.sum(a64)


def test_port_env_episode_and_features(net, golden):
    G = golden("sf_episode.npz")
    key = "msa_K20_s7"
    env = rp.PortEnv(net, assignment_method="msa", assignment_iters=20, unassigned_penalty=1e4, seed=7)
    dmg = G[f"{key}_dmg"]
    assert np.array_equal(np.flatnonzero(env.is_damaged > 0), dmg)  # sampler = reference RNG + retry
    curve, rew, dones = [env.tstt], [], []
    for i, a in enumerate(dmg):
        if i == 3:
            r, d = env.step(int(dmg[0]))
            curve.append(env.tstt); rew.append(r); dones.append(d)
        r, d = env.step(int(a))
        curve.append(env.tstt); rew.append(r); dones.append(d)
    np.testing.assert_allclose(curve, G[f"{key}_tstt"], rtol=1e-6)
    np.testing.assert_allclose(rew, G[f"{key}_reward"], rtol=1e-5, atol=1e-9)
    assert dones == list(G[f"{key}_done"])
    F = golden("sf_features.npz")
    env = rp.PortEnv(net, assignment_method="fw", assignment_iters=20, seed=2)
    node, edge, _ = env.features()
    np.testing.assert_allclose(node, F["s2_node0"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(edge, F["s2_edge0"], rtol=1e-5, atol=1e-6)


def test_betweenness_restatements_agree(sf_oracle, net, golden):
    F = golden("sf_features.npz")
    bw0 = rp.betweenness_unweighted(24, net.row, net.col, np.ones(76, bool))
    np.testing.assert_allclose(bw0.astype(np.float32), F["static_betweenness"], rtol=1e-6)
    np.testing.assert_allclose(sf_oracle.betweenness(np.zeros(76, np.uint8)), bw0, rtol=1e-12)
    dmg = F["s3_dmg0"]
    a = rp.betweenness_unweighted(24, net.row, net.col, dmg == 0)
    np.testing.assert_allclose(sf_oracle.betweenness(dmg.astype(np.uint8)), a, rtol=1e-12, atol=1e-15)
    nodes = F["s3_node0"][:, 0]
    np.testing.assert_allclose((a / a.max()).astype(np.float32), nodes, rtol=1e-5, atol=1e-7)


def test_reward_modes(net, golden):
    G = golden("sf_kat.npz")
    gm = np.ones(76, np.float32)
    for mode, row in zip(G["reward_modes"], G["reward_rows"]):
        prev, curr, complete, clip, expect = row
        dm = np.zeros(76, np.float32) if complete else gm
        got = rp.reward_value(prev, curr, gm, dm, float(G["reward_initial_tstt"]), 1.0, 10.0, 0.1, str(mode), clip)
        assert got == pytest.approx(expect, rel=1e-12, abs=1e-12)


def test_oracle_edge_cases(sf_oracle):
    with pytest.raises(ValueError):
        sf_oracle.assign(np.zeros(76, np.uint8), None, 0, "msa")
    all_dmg = np.ones(76, np.uint8)
    r = sf_oracle.assign(all_dmg, None, 3, "msa")
    assert np.isfinite(r["tstt"]) and r["unassigned"] == 0.0  # damaged links stay traversable at 1e6
    r = sf_oracle.assign(np.zeros(76, np.uint8), None, 10, "fw_ls", dtype=np.float64)
    r_fw = sf_oracle.assign(np.zeros(76, np.uint8), None, 10, "fw", dtype=np.float64)
    assert r["gap"] < r_fw["gap"]
