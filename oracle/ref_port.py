"""ORACLE (test infrastructure, not product code) - numpy/scipy restatement of the reference env.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module.  It restates, with the reference's own third-party numerics (numpy fp32 ufuncs and
``scipy.sparse.csgraph.dijkstra``), the algorithm of

* ``compute_travel_time``        reference: src/env/repair_env.py:667-677
* ``_all_or_nothing`` (scipy)    reference: src/env/repair_env.py:481-503, 707-722
* ``compute_flow_assignment``    reference: src/env/repair_env.py:299-345   (msa / fw / cfw)
* ``compute_tstt``               reference: src/env/repair_env.py:724-735
* ``reset`` damage sampling      reference: src/env/repair_env.py:167-205
* ``step`` / reward / done       reference: src/env/repair_env.py:207-237, 244-294
* ``get_state`` features         reference: src/env/repair_env.py:751-819

Parity pin: ``tests/golden/*.npz`` were produced by importing the reference itself in the build
container (``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py`` checks this module
against them.  Because the arithmetic is the same numpy/scipy calls in the same order, the match
is bit-for-bit on a host with the same numpy SIMD dispatch and <= 1e-6 relative elsewhere
(numpy's fp32 ``power`` is not correctly rounded and differs by an ulp between AVX512/SVML and
scalar builds, SURVEY.md §7 hard part 2).

Everything takes plain arrays so the module does not depend on the product package.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import connected_components, dijkstra

NO_PRED = -9999


@dataclass
class PortNetwork:
    """Static network in the layout the reference builds in ``RepairEnv.__init__`` (:85-97)."""

    num_nodes: int
    row: np.ndarray  # int64 [E] 0-based tail node
    col: np.ndarray  # int64 [E] 0-based head node
    t0: np.ndarray  # float32 [E]
    cap0: np.ndarray  # float32 [E]
    od: List[Tuple[int, int, float]]  # (origin0, dest0, demand) in trip-table order
    total_demand: float = 0.0
    edge_of: Dict[Tuple[int, int], int] = field(default_factory=dict)
    od_by_origin: Dict[int, List[Tuple[int, float]]] = field(default_factory=dict)

    @classmethod
    def from_graph_data(cls, g) -> "PortNetwork":
        row = np.array([e.u - 1 for e in g.edges], dtype=np.int64)
        col = np.array([e.v - 1 for e in g.edges], dtype=np.int64)
        net = cls(
            num_nodes=int(g.num_nodes),
            row=row,
            col=col,
            t0=np.array([e.t0 for e in g.edges], dtype=np.float32),
            cap0=np.array([e.capacity for e in g.edges], dtype=np.float32),
            od=[(o - 1, d - 1, float(q)) for (o, d), q in g.od_demand.items()],
        )
        net.total_demand = float(np.sum([q for (_, _, q) in net.od]))
        net.edge_of = {(int(u), int(v)): i for i, (u, v) in enumerate(zip(row, col))}
        for o, d, q in net.od:
            net.od_by_origin.setdefault(o, []).append((d, q))
        return net

    @property
    def num_edges(self) -> int:
        return int(self.row.shape[0])


# ---------------------------------------------------------------------------------------------
# stages
# ---------------------------------------------------------------------------------------------

def bpr_times(net: PortNetwork, flow, cap, dmg, alpha=0.15, beta=4.0) -> np.ndarray:
    """fp32 BPR sweep, op order of the reference (:670-676)."""
    f = np.asarray(flow, dtype=np.float32)
    c = np.maximum(cap, 1e-6)
    vc = np.clip(f / c, 0.0, 10.0)
    t = net.t0 * (1.0 + alpha * (vc ** beta))
    t = t.astype(np.float32)
    t[dmg > 0.5] = 1e6
    return t


def shortest_path_trees(net: PortNetwork, t: np.ndarray):
    """All-sources Dijkstra exactly as the reference calls it (:485-489)."""
    graph = csr_matrix((t.copy(), (net.row, net.col)), shape=(net.num_nodes, net.num_nodes))
    dist, pred = dijkstra(graph, directed=True, indices=range(net.num_nodes), return_predecessors=True)
    return dist, pred


def aon_load(net: PortNetwork, pred: np.ndarray, dtype=np.float32):
    """Per-OD walk dest -> origin along ``pred`` (:490-503, :707-722)."""
    aux = np.zeros(net.num_edges, dtype=dtype)
    unassigned = 0.0
    for o, dests in net.od_by_origin.items():
        prow = pred[o]
        for d, q in dests:
            if d == o or prow[d] < 0:
                unassigned += q
                continue
            cur, hops = d, []
            while cur != o and cur != NO_PRED:
                hops.append(cur)
                cur = int(prow[cur])
            if cur != o:
                unassigned += q
                continue
            hops.append(o)
            for head, tail in zip(hops[:-1], hops[1:]):
                aux[net.edge_of[(tail, head)]] += q
    return aux, unassigned


def tstt_value(net: PortNetwork, flow, t, unassigned=0.0, penalty=2e7) -> float:
    base = float(np.sum(np.asarray(flow, np.float32) * np.asarray(t, np.float32)))
    total = max(net.total_demand, 1.0)
    out = base / total
    if unassigned > 0:
        out += float(penalty) * (float(unassigned) / total)
    return out


def assign(
    net: PortNetwork,
    cap: np.ndarray,
    dmg: np.ndarray,
    flow0: Optional[np.ndarray],
    iters: int,
    method: str = "msa",
    alpha: float = 0.15,
    beta: float = 4.0,
    penalty: float = 2e7,
    trace: Optional[dict] = None,
):
    """K fixed-step iterations (:299-345).  Returns (flow fp32, tstt, unassigned)."""
    if iters <= 0:
        raise ValueError("assignment_iters must be > 0 to update TSTT.")
    flow = np.zeros(net.num_edges, np.float32) if flow0 is None else np.asarray(flow0, np.float32).copy()
    t = bpr_times(net, flow, cap, dmg, alpha, beta)
    d_prev = None
    unassigned = 0.0
    for it in range(iters):
        _, pred = shortest_path_trees(net, t)
        aux, unassigned = aon_load(net, pred)
        if trace is not None:
            trace.setdefault("t", []).append(t.copy())
            trace.setdefault("pred", []).append(pred.copy())
            trace.setdefault("aux", []).append(aux.copy())
        d_fw = aux - flow
        if method == "cfw":
            if d_prev is None:
                direction = d_fw
            else:
                num = float(np.dot(d_fw, d_fw - d_prev))
                den = float(np.dot(d_prev, d_prev)) + 1e-12
                direction = d_fw + max(0.0, num / den) * d_prev
            lam = 2.0 / (it + 2.0)
            flow = np.maximum(flow + lam * direction, 0.0)
            d_prev = direction
        else:
            lam = 2.0 / (it + 2.0) if method == "fw" else 1.0 / (it + 1.0)
            flow = (1 - lam) * flow + lam * aux
        if np.any(np.isnan(flow)):
            flow = np.nan_to_num(flow)
        t = bpr_times(net, flow, cap, dmg, alpha, beta)
        if trace is not None:
            trace.setdefault("flow", []).append(flow.copy())
    return flow, tstt_value(net, flow, t, unassigned, penalty), unassigned


# ---------------------------------------------------------------------------------------------
# env glue: damage sampling, reward, features
# ---------------------------------------------------------------------------------------------

def active_subgraph_strongly_connected(net: PortNetwork, dmg_mask: np.ndarray) -> bool:
    """``nx.is_strongly_connected(nx_graph.edge_subgraph(active))`` (:178-186): only nodes that
    still touch an active link belong to the subgraph."""
    act = np.flatnonzero(np.asarray(dmg_mask) == 0)
    if act.size == 0:
        return False
    nodes = np.unique(np.concatenate([net.row[act], net.col[act]]))
    remap = np.full(net.num_nodes, -1, dtype=np.int64)
    remap[nodes] = np.arange(nodes.size)
    g = csr_matrix((np.ones(act.size), (remap[net.row[act]], remap[net.col[act]])), shape=(nodes.size, nodes.size))
    ncomp, _ = connected_components(g, directed=True, connection="strong")
    return ncomp == 1


def sample_damage(net: PortNetwork, rng: np.random.Generator, ratio: float = 0.3, max_retries: int = 50):
    """Damage draw with the reference's retry rule (:168-190).  Returns sorted? no - raw indices."""
    E = net.num_edges
    count = max(1, int(E * ratio))
    chosen = None
    for _ in range(max_retries):
        cand = rng.choice(E, size=count, replace=False)
        mask = np.zeros(E, dtype=np.float32)
        mask[cand] = 1.0
        if active_subgraph_strongly_connected(net, mask):
            chosen = cand
            break
    if chosen is None:
        chosen = rng.choice(E, size=count, replace=False)
    return chosen


def goal_complete(goal_mask, damaged_mask) -> bool:
    return bool(np.sum(goal_mask * damaged_mask) == 0.0)


def reward_value(prev, curr, goal_mask, damaged_mask, initial_tstt, alpha=1.0, beta=10.0, gamma=0.1,
                 mode="delta", clip=0.0) -> float:
    """Five reward modes of the reference (:244-291)."""
    if mode in ("minimize_tstt", "rel_improve"):
        base = initial_tstt if initial_tstt is not None else prev
        ratio = curr / max(base, 1.0)
        if mode == "minimize_tstt":
            r = -alpha * ratio
        else:
            r = alpha * (((prev - curr) / max(base, 1.0)) * 100.0) - 1.0 * ratio
        r = r + (beta if goal_complete(goal_mask, damaged_mask) else 0.0)
        if clip and clip > 0:
            r = float(np.clip(r, -clip, clip))
        return r
    if mode == "neg_tstt":
        delta = -curr
    elif mode == "log_delta":
        delta = np.log10(max(prev, 1.0)) - np.log10(max(curr, 1.0))
    else:
        delta = prev - curr
    r = alpha * delta + (beta if goal_complete(goal_mask, damaged_mask) else 0.0) - gamma
    if clip and clip > 0:
        r = float(np.clip(r, -clip, clip))
    return r


def betweenness_unweighted(num_nodes: int, row, col, active: np.ndarray) -> np.ndarray:
    """Brandes betweenness (unweighted, directed, endpoints excluded, normalised by
    1/((n-1)(n-2)) over the n nodes of the edge-induced active subgraph) - the quantity
    ``nx.betweenness_centrality(subgraph, normalized=True)`` returns (:753-761).  Nodes outside
    the subgraph get 0."""
    act = np.flatnonzero(active)
    out = np.zeros(num_nodes, dtype=np.float64)
    if act.size == 0:
        return out
    nodes = np.unique(np.concatenate([np.asarray(row)[act], np.asarray(col)[act]]))
    adj: Dict[int, List[int]] = {int(v): [] for v in nodes}
    for e in act:
        adj[int(row[e])].append(int(col[e]))
    for s in nodes:
        s = int(s)
        sigma = {int(v): 0.0 for v in nodes}
        dist = {int(v): -1 for v in nodes}
        preds: Dict[int, List[int]] = {int(v): [] for v in nodes}
        sigma[s], dist[s] = 1.0, 0
        order, queue, head = [], [s], 0
        while head < len(queue):
            v = queue[head]
            head += 1
            order.append(v)
            for w in adj[v]:
                if dist[w] < 0:
                    dist[w] = dist[v] + 1
                    queue.append(w)
                if dist[w] == dist[v] + 1:
                    sigma[w] += sigma[v]
                    preds[w].append(v)
        delta = {int(v): 0.0 for v in nodes}
        for w in reversed(order):
            coeff = (1.0 + delta[w]) / sigma[w]
            for v in preds[w]:
                delta[v] += sigma[v] * coeff
            if w != s:
                out[w] += delta[w]
    n = nodes.size
    if n > 2:
        out *= 1.0 / ((n - 1) * (n - 2))
    return out


def state_features(net: PortNetwork, flow, cap, dmg, goal_mask, tstt):
    """Node [N,4] and link [E,6] features of ``get_state`` (:751-809), fp32."""
    E, N = net.num_edges, net.num_nodes
    dmg = np.asarray(dmg, np.float32)
    bw = betweenness_unweighted(N, net.row, net.col, dmg == 0).astype(np.float32)
    bw_max = float(np.max(bw)) if bw.size else 0.0
    if bw_max > 0:
        bw = bw / bw_max
    raw_vc = np.zeros(E, dtype=np.float32)
    for i in range(E):
        raw_vc[i] = flow[i] / max(cap[i], 1e-6)
    vc = np.clip(np.log1p(np.where(dmg > 0, 0.0, raw_vc)), 0.0, 10.0)
    goal_total = float(np.sum(goal_mask))
    remaining_ratio = float(np.sum(goal_mask * dmg)) / max(goal_total, 1.0)
    avg_flow = float(np.mean(flow[dmg == 0])) if np.sum(dmg == 0) > 0 else 0.0
    avg_flow_norm = avg_flow / max(net.total_demand / max(E, 1), 1.0)
    log_tstt = float(np.log10(max(tstt, 1.0))) if tstt is not None else 0.0
    node = np.stack(
        [bw, np.full(N, remaining_ratio, np.float32), np.full(N, avg_flow_norm, np.float32),
         np.full(N, log_tstt, np.float32)], axis=1)
    max_t0, max_cap = float(np.max(net.t0)), float(np.max(net.cap0))
    edge_id_norm = np.arange(E, dtype=np.float32) / max(E - 1, 1)
    t0_norm = np.log10(net.t0 + 1.0) / np.log10(max_t0 + 1.0)
    cap_norm = np.log10(cap + 1.0) / np.log10(max_cap + 1.0)
    edge = np.stack([t0_norm.astype(np.float32), cap_norm.astype(np.float32), vc, dmg,
                     np.asarray(goal_mask, np.float32), edge_id_norm], axis=1)
    return node, edge, log_tstt


class PortEnv:
    """Single-scenario env with the reference's reset/step semantics, used as a CPU checker and
    as the timed ``--impl reference`` arm of bench.py (kind "port")."""

    def __init__(self, net: PortNetwork, damaged_ratio=0.3, bpr_alpha=0.15, bpr_beta=4.0, assignment_iters=20,
                 assignment_method="msa", reward_mode="log_delta", reward_alpha=1.0, reward_beta=10.0,
                 reward_gamma=0.1, reward_clip=0.0, capacity_damage=1e-3, unassigned_penalty=2e7, seed=0):
        self.net = net
        self.alpha, self.beta = bpr_alpha, bpr_beta
        self.iters, self.method = assignment_iters, assignment_method.lower()
        self.reward_mode, self.reward_alpha, self.reward_beta = reward_mode, reward_alpha, reward_beta
        self.reward_gamma, self.reward_clip = reward_gamma, reward_clip
        self.capacity_damage, self.penalty = capacity_damage, unassigned_penalty
        self.rng = np.random.default_rng(seed)
        self.initial_tstt = None
        self.reset(damaged_ratio)

    def reset(self, damaged_ratio=0.3, damaged_indices=None):
        E = self.net.num_edges
        idx = sample_damage(self.net, self.rng, damaged_ratio) if damaged_indices is None else damaged_indices
        self.is_damaged = np.zeros(E, np.float32)
        self.is_damaged[idx] = 1.0
        self.capacities = self.net.cap0.copy()
        self.capacities[idx] = self.capacity_damage
        self.goal_mask = self.is_damaged.copy()
        self.flow = np.zeros(E, np.float32)
        self.assign()
        self.initial_tstt = self.tstt
        return self.tstt

    def assign(self, trace=None):
        self.flow, self.tstt, self.unassigned = assign(
            self.net, self.capacities, self.is_damaged, self.flow, self.iters, self.method, self.alpha, self.beta,
            self.penalty, trace)
        return self.tstt

    def step(self, a: int):
        if a < 0 or a >= self.net.num_edges:
            raise ValueError(f"action_edge_id {a} out of range (0..{self.net.num_edges - 1})")
        if self.is_damaged[a] == 0:
            return -1.0, False
        prev = self.tstt
        self.is_damaged[a] = 0.0
        self.capacities[a] = self.net.cap0[a]
        self.assign()
        r = reward_value(prev, self.tstt, self.goal_mask, self.is_damaged, self.initial_tstt, self.reward_alpha,
                         self.reward_beta, self.reward_gamma, self.reward_mode, self.reward_clip)
        return r, goal_complete(self.goal_mask, self.is_damaged)

    def features(self):
        return state_features(self.net, self.flow, self.capacities, self.is_damaged, self.goal_mask, self.tstt)
