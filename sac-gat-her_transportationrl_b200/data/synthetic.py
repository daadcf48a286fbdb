"""Seeded synthetic road networks of the sizes BASELINE.json names.

The reference ships only Sioux Falls (reference: src/data/tntp_download.py:23-43); the larger
benchmark networks are generated here.  Construction rules (they matter for parity):

* nodes are cells of an ``rows x cols`` lattice, optionally thinned to ``num_nodes`` while staying
  connected; links are undirected lattice edges (a random spanning tree first, then random extra
  lattice edges) emitted in both directions -> strongly connected, no parallel links, no self
  loops (scipy's ``csr_matrix`` would *sum* parallel links, reference: src/env/repair_env.py:488);
* links are sorted by (u, v) and node ids are 1-based, like a TNTP file;
* free-flow times are integers in 1..10 (same value both ways) so that cold-start iteration 0
  has many exact shortest-path ties and exercises the tie rule;
* demands are integer multiples of 10 -> every all-or-nothing link sum is exact in fp32/fp64,
  so scatter-add order cannot change results (SURVEY.md §7 hard part 9).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .tntp_parser import EdgeData, GraphData


def _lattice_edges(rows: int, cols: int, alive: np.ndarray) -> List[Tuple[int, int]]:
    out = []
    for r in range(rows):
        for c in range(cols):
            a = r * cols + c
            if not alive[a]:
                continue
            if c + 1 < cols and alive[a + 1]:
                out.append((a, a + 1))
            if r + 1 < rows and alive[a + cols]:
                out.append((a, a + cols))
    return out


def _connected(n_cells: int, alive: np.ndarray, edges: List[Tuple[int, int]]) -> bool:
    adj = [[] for _ in range(n_cells)]
    for a, b in edges:
        adj[a].append(b)
        adj[b].append(a)
    start = int(np.flatnonzero(alive)[0])
    seen = np.zeros(n_cells, dtype=bool)
    seen[start] = True
    stack = [start]
    while stack:
        a = stack.pop()
        for b in adj[a]:
            if not seen[b]:
                seen[b] = True
                stack.append(b)
    return bool(seen[alive].all())


def make_lattice_network(
    rows: int,
    cols: int,
    num_nodes: int,
    num_undirected_links: int,
    num_zones: int,
    od_keep_prob: float = 1.0,
    seed: int = 0,
) -> GraphData:
    rng = np.random.default_rng(seed)
    n_cells = rows * cols
    alive = np.ones(n_cells, dtype=bool)
    # thin the lattice one cell at a time, keeping it connected
    to_remove = n_cells - num_nodes
    if to_remove < 0:
        raise ValueError("num_nodes exceeds lattice size")
    order = rng.permutation(n_cells)
    k = 0
    while to_remove > 0:
        if k >= len(order):
            raise RuntimeError("could not thin the lattice to the requested size")
        cell = int(order[k])
        k += 1
        alive[cell] = False
        if _connected(n_cells, alive, _lattice_edges(rows, cols, alive)):
            to_remove -= 1
        else:
            alive[cell] = True
    cand = _lattice_edges(rows, cols, alive)
    if num_undirected_links > len(cand) or num_undirected_links < num_nodes - 1:
        raise ValueError(f"link count must be in [{num_nodes - 1}, {len(cand)}]")
    # random spanning tree (Kruskal on a shuffled edge list), then random extra edges
    perm = rng.permutation(len(cand))
    parent = np.arange(n_cells)

    def find(x: int) -> int:
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = int(parent[x])
        return x

    chosen, rest = [], []
    for i in perm:
        a, b = cand[int(i)]
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[ra] = rb
            chosen.append((a, b))
        else:
            rest.append((a, b))
    chosen += rest[: num_undirected_links - len(chosen)]
    # compact ids (1-based), emit both directions, sort by (u, v)
    new_id = np.full(n_cells, -1, dtype=np.int64)
    new_id[alive] = np.arange(1, num_nodes + 1)
    t0_und = rng.integers(1, 11, size=len(chosen))
    cap_und = np.round(rng.uniform(2000.0, 26000.0, size=len(chosen)), 3)
    rows_out = []
    for (a, b), t0, cap in zip(chosen, t0_und, cap_und):
        u, v = int(new_id[a]), int(new_id[b])
        rows_out.append((u, v, float(cap), float(t0)))
        rows_out.append((v, u, float(cap), float(t0)))
    rows_out.sort(key=lambda r: (r[0], r[1]))
    edges = [EdgeData(u=u, v=v, capacity=cap, t0=t0, length=t0, b=0.15, power=4.0) for u, v, cap, t0 in rows_out]
    zones = np.sort(rng.choice(num_nodes, size=num_zones, replace=False)) + 1
    od = {}
    for o in zones:
        for d in zones:
            if o == d:
                continue
            keep = od_keep_prob >= 1.0 or rng.random() < od_keep_prob
            q = 10.0 * float(rng.integers(1, 40))
            if keep:
                od[(int(o), int(d))] = q
    return GraphData(num_nodes=num_nodes, edges=edges, od_demand=od)


def anaheim_size(seed: int = 0) -> GraphData:
    """416 nodes / 914 directed links / 38 zones / 1 406 OD (config 3)."""
    return make_lattice_network(16, 26, 416, 457, 38, 1.0, seed)


def chicago_sketch_size(seed: int = 0) -> GraphData:
    """933 nodes / 2 950 directed links / 387 zones / ~140 k OD (config 4)."""
    return make_lattice_network(31, 31, 933, 1475, 387, 0.94, seed)


def small_grid(rows: int = 3, cols: int = 4, num_zones: int = 5, seed: int = 0) -> GraphData:
    """Tiny network for pure-Python-speed tests."""
    n = rows * cols
    max_links = rows * (cols - 1) + cols * (rows - 1)
    return make_lattice_network(rows, cols, n, max(n - 1, int(0.85 * max_links)), num_zones, 1.0, seed)
