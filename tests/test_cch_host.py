"""Host-side check of the contraction-hierarchy planner behind the general fused kernel
(csrc/b200_cch.cu): `b200_cch_host_sssp` replays on the CPU exactly the schedule the kernel runs
(customization rows, up phase along the elimination tree, down-sweep batches incl. split destinations);
its distances must equal the oracle's fp64 Dijkstra (scipy semantics, reference
src/env/repair_env.py:485-489) BIT FOR BIT when link costs are fp32 values, because every path sum is
then exact in fp64 and independent of the association order."""
import ctypes as C

import numpy as np
import pytest


def _host_sssp(lib, orc, cost, origin):
    dist = np.empty(orc.N, np.float64)
    stats = (C.c_int64 * 8)()
    cost = np.ascontiguousarray(cost, np.float64)
    rc = lib.b200_cch_host_sssp(orc.N, orc.E, orc.row.ctypes.data, orc.col.ctypes.data, cost.ctypes.data, int(origin),
                                dist.ctypes.data, stats)
    assert rc == 0
    return dist, list(stats)


@pytest.mark.parametrize("which", ["sf", "an", "cs"])
def test_hierarchy_schedule_reproduces_dijkstra_distances(b200env, native_lib, which):
    from oracle.c_oracle import COracle
    g = {"sf": b200env.load_sioux_falls, "an": lambda: b200env.synthetic.anaheim_size(0),
         "cs": lambda: b200env.synthetic.chicago_sketch_size(0)}[which]()
    orc = COracle.from_graph_data(g)
    rng = np.random.default_rng(3)
    n_cost, n_orig = (4, 6) if which != "cs" else (2, 3)
    for c in range(n_cost):
        # fp32-valued costs: free-flow times (integers: ties everywhere) or BPR-like values, 30 % at 1e6
        base = orc.t0.astype(np.float64) if c == 0 else (orc.t0 * (1 + rng.random(orc.E) * 7).astype(np.float32)).astype(np.float64)
        cost = base.copy()
        if c > 0:
            cost[rng.choice(orc.E, int(orc.E * 0.3), replace=False)] = 1e6
        origins = rng.choice(orc.origins, n_orig, replace=False)
        ref_dist, _, _ = orc.sssp(cost, origins)
        for i, o in enumerate(origins):
            dist, stats = _host_sssp(native_lib, orc, cost, o)
            assert np.array_equal(dist, ref_dist[i]), f"{which}: cost set {c}, origin {o}"
    # the plan itself: every link is a filled edge, fill stays small on road-like networks
    assert stats[0] >= orc.E and stats[0] <= 4 * orc.E
    assert stats[2] + stats[3] + 2 < 65535


def test_hierarchy_handles_unreachable_nodes_and_one_way_links(b200env, native_lib):
    from oracle.c_oracle import COracle
    E = b200env.EdgeData
    # a one-way chain 1->2->3->4 with a dead-end spur 5->3 and an isolated sink 6 reached from 4 only
    links = [(1, 2, 2.0), (2, 3, 3.0), (3, 4, 1.0), (5, 3, 4.0), (4, 6, 2.0), (2, 1, 2.0)]
    g = b200env.GraphData(num_nodes=6, edges=[E(u, v, 1000.0, t, t, 0.15, 4.0) for u, v, t in links],
                          od_demand={(1, 4): 10.0})
    orc = COracle.from_graph_data(g)
    cost = orc.t0.astype(np.float64)
    for o in range(6):
        ref, _, _ = orc.sssp(cost, [o])
        dist, _ = _host_sssp(native_lib, orc, cost, o)
        assert np.array_equal(dist, ref[0]), o


def test_hierarchy_on_random_digraphs_with_one_way_links(b200env, native_lib):
    """One-way links make the directed fill sparser than the undirected one: the up phase must follow the
    directed edges to higher ranks, not the elimination tree (regression: a 70-node network with a one-way
    chain gave labels that were too long)."""
    from oracle.c_oracle import COracle
    E = b200env.EdgeData
    for seed, n, extra in ((8, 70, 60), (1, 40, 90), (2, 120, 100), (3, 25, 200)):
        rng = np.random.default_rng(seed)
        links = {(i, i + 1) for i in range(1, n - 1)} | {(i + 1, i) for i in range(1, n - 1) if i % 3}
        links |= {(int(a), int(b)) for a, b in rng.integers(1, n, size=(extra, 2)) if a != b}
        links = sorted(links)
        g = b200env.GraphData(num_nodes=n, edges=[E(u, v, 900.0, float(1 + (u * 7 + v) % 5), float(1 + (u * 7 + v) % 5), 0.15, 4.0)
                                                   for u, v in links], od_demand={(1, 2): 1.0})
        orc = COracle.from_graph_data(g)
        cost = orc.t0.astype(np.float64)
        cost[rng.choice(orc.E, orc.E // 5, replace=False)] = 1e6
        ref, _, _ = orc.sssp(cost, np.arange(n, dtype=np.int32))
        for o in range(n):
            dist, _ = _host_sssp(native_lib, orc, cost, o)
            assert np.array_equal(dist, ref[o]), (seed, o)


@pytest.mark.parametrize("force_d", ["2", "3", "8"])
def test_hierarchy_dense_graph_wide_destinations(b200env, native_lib, monkeypatch, force_d):
    """A dense digraph: the first eliminated nodes have ~100 higher-ranked in-neighbours.  With D forced to 2 that is
    more than the 32 x D edges a lane group of one batch can take (partial slots + a second stage), with D = 3 / 8 it
    is a group of 32 / 16 lanes reduced inside one batch; every variant must reproduce Dijkstra bit for bit."""
    from oracle.c_oracle import COracle
    E = b200env.EdgeData
    n = 110
    rng = np.random.default_rng(5)
    links = sorted({(int(a), int(b)) for a, b in rng.integers(1, n + 1, size=(9000, 2)) if a != b})
    g = b200env.GraphData(num_nodes=n, edges=[E(u, v, 900.0, float(1 + (u * 13 + v * 7) % 9), float(1 + (u * 13 + v * 7) % 9), 0.15, 4.0)
                                               for u, v in links], od_demand={(1, 2): 1.0})
    orc = COracle.from_graph_data(g)
    cost = orc.t0.astype(np.float64)
    cost[rng.choice(orc.E, orc.E // 4, replace=False)] = 1e6
    monkeypatch.setenv("B200_CCH_D", force_d)
    origins = np.array([0, 57, n - 1], dtype=np.int32)   # every call plans the hierarchy again: keep it to three
    ref, _, _ = orc.sssp(cost, origins)
    for i, o in enumerate(origins):
        dist, stats = _host_sssp(native_lib, orc, cost, o)
        assert np.array_equal(dist, ref[i]), (force_d, o)
    assert stats[7] == int(force_d)
    if force_d == "2":
        assert stats[4] > n + 1, "expected partial slots: in-degree beyond 32 x D"
