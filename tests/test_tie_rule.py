"""Property test: scipy's Dijkstra predecessor choice under exact ties
is 'tight in-link with the smallest dist[u], then the largest u' - the rule the oracle and the
CUDA kernel implement.  Runs against whatever scipy is installed."""
import numpy as np
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import dijkstra

from oracle.c_oracle import COracle


def _check(N, row, col, w, sources):
    w32 = w.astype(np.float32)
    dist, pred = dijkstra(csr_matrix((w32, (row, col)), shape=(N, N)), directed=True, indices=sources,
                          return_predecessors=True)
    orc = COracle(N, row, col, np.ones(len(row)), np.ones(len(row)), [(int(s), int((s + 1) % N), 1.0) for s in sources])
    d2, p2, _ = orc.sssp(w32, origins=np.asarray(sources, np.int32))
    assert np.array_equal(dist, d2)
    assert np.array_equal(pred, p2)
    w64 = w32.astype(np.float64)
    ties = 0
    for k, o in enumerate(sources):
        d = dist[k]
        tight = np.isfinite(d[row]) & (d[row] + w64 == d[col]) & (col != o)
        cnt = np.bincount(col[tight], minlength=N)
        ties += int((cnt > 1).sum())
    return ties


def test_sioux_falls_free_flow_with_damage(sf_graph):
    row = np.array([e.u - 1 for e in sf_graph.edges])
    col = np.array([e.v - 1 for e in sf_graph.edges])
    t0 = np.array([e.t0 for e in sf_graph.edges])
    rng = np.random.default_rng(0)
    ties = 0
    for _ in range(40):
        w = t0.copy()
        w[rng.choice(76, 22, replace=False)] = 1e6
        ties += _check(24, row, col, w, list(range(24)))
    assert ties > 100  # the case really exercises ties


def test_integer_lattices_and_random_digraphs(b200env):
    rng = np.random.default_rng(1)
    ties = 0
    for seed in range(4):
        g = b200env.synthetic.make_lattice_network(8, 9, 72, 110, 6, 1.0, seed)
        row = np.array([e.u - 1 for e in g.edges]); col = np.array([e.v - 1 for e in g.edges])
        ties += _check(72, row, col, np.array([e.t0 for e in g.edges]), list(range(0, 72, 5)))
    for _ in range(10):
        N = int(rng.integers(20, 200))
        pairs = set()
        while len(pairs) < 4 * N:
            u, v = rng.integers(0, N, 2)
            if u != v:
                pairs.add((int(u), int(v)))
        pairs = sorted(pairs)
        row = np.array([p[0] for p in pairs]); col = np.array([p[1] for p in pairs])
        w = rng.choice([0.5, 1.0, 2.0, 1e6], size=len(pairs))
        ties += _check(N, row, col, w, list(range(0, N, 7)))
    assert ties > 500
