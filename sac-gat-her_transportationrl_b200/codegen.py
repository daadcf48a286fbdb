"""Topology headers for the network-specialised assignment kernel (csrc/b200_assign_static.cuh).

DATA ONLY: node order, CSC in-link lists and link ids of one network as a constexpr struct; the
kernel template that consumes it is hand-written.  Used for the built-in Sioux Falls header
(tools/gen_static_topology.py) and for on-demand plugins (_native.specialize).

Internal node numbering = position in the Gauss-Seidel sweep order.  The order is the ascending
free-flow distance from the root that needs the fewest alternating sweeps (all origins at once,
what the kernel does) on the free-flow costs and five deterministic damage/jitter patterns.
In-links of a node are listed by ascending ORIGINAL tail node id (the scipy tie rule picks the
largest tail among equal candidates).
"""
from __future__ import annotations

import numpy as np


def _sweeps(order, inl, row, w, origins, cap=400):
    N = len(order)
    d = np.full((len(origins), N), np.inf)
    for i, o in enumerate(origins):
        d[i, o] = 0.0
    s = 0
    while True:
        changed = False
        for v in (order if s % 2 == 0 else order[::-1]):
            for e in inl[v]:
                c = d[:, row[e]] + w[e]
                m = c < d[:, v]
                if m.any():
                    d[m, v] = c[m]
                    changed = True
        s += 1
        if not changed or s > cap:
            return s


def choose_order(N, row, col, t0, origins):
    E = len(row)
    inl = [[] for _ in range(N)]
    for e in range(E):
        inl[col[e]].append(e)
    rng = np.random.default_rng(12345)
    patterns = [t0.astype(np.float64)]
    for _ in range(5):
        w = t0.astype(np.float64) * (1.0 + 0.5 * rng.random(E))
        w[rng.random(E) < 0.3] = 1e6
        patterns.append(w)
    best = None
    for root in range(N):
        dist = np.full(N, np.inf)
        dist[root] = 0.0
        done = np.zeros(N, bool)
        for _ in range(N):
            cand = np.where(done, np.inf, dist)
            u = int(np.argmin(cand))
            if not np.isfinite(cand[u]):
                break
            done[u] = True
            for e in range(E):  # undirected free-flow distances
                a, b = int(row[e]), int(col[e])
                if a == u and dist[u] + t0[e] < dist[b]:
                    dist[b] = dist[u] + t0[e]
                if b == u and dist[u] + t0[e] < dist[a]:
                    dist[a] = dist[u] + t0[e]
        order = [int(x) for x in np.argsort(dist, kind="stable")]
        cost = sum(_sweeps(order, inl, row, w, origins) for w in patterns)
        if best is None or cost < best[0]:
            best = (cost, order, root)
    return best[1], best[2], best[0]


def _cch_eliminate(N, tails, heads, pick):
    """Eliminate the nodes in the order `pick(alive, adj)` yields; returns (order, filled edges, lower
    triangles {(u, x): [v, ..]}): eliminating v adds the shortcut (u, x) for every in-neighbour u and
    out-neighbour x of v that are still alive."""
    out = [set() for _ in range(N)]
    inn = [set() for _ in range(N)]
    for u, v in zip(tails, heads):
        out[u].add(v)
        inn[v].add(u)
    adj = [out[i] | inn[i] for i in range(N)]
    alive = set(range(N))
    order, tris = [], {}
    while alive:
        v = pick(alive, adj)
        alive.discard(v)
        order.append(v)
        for u in sorted(inn[v] & alive):
            for x in sorted(out[v] & alive):
                if u != x:
                    tris.setdefault((u, x), []).append(v)
                    out[u].add(x)
                    inn[x].add(u)
                    adj[u].add(x)
                    adj[x].add(u)
    edges = [(u, x) for u in range(N) for x in sorted(out[u])]
    return order, edges, tris


def _cch_levels(edges, tris):
    """Customization level of every filled edge: 0 = nothing to do, L = all lower edges are < L."""
    level = {}

    def lvl(e):
        if e not in level:
            level[e] = 0 if e not in tris else 1 + max(max(lvl((e[0], v)), lvl((v, e[1]))) for v in tris[e])
        return level[e]

    for e in edges:
        lvl(e)
    return level


def _cch_batches(edges, tris, level, key, width=32):
    """Customization batches: edges of one level, most triangles first, `width` per batch."""
    batches = []
    for L in range(1, max(level.values(), default=0) + 1):
        es = sorted((e for e in edges if level[e] == L), key=lambda e: (-len(tris[e]), key(e)))
        batches += [es[c:c + width] for c in range(0, len(es), width)]
    return batches


def cch_plan(N, tails, heads, trials: int = 1500):
    """Customizable-contraction-hierarchy plan of one small topology (internal node indices).

    With all link costs exactly summable in fp64 (see csrc/b200_assign_static.cuh) the shortest
    distances from an origin over the filled graph are the result of ONE sweep over the in-edges from
    lower-ranked tails in elimination order followed by ONE sweep over those from higher-ranked tails
    in reverse order - instead of Gauss-Seidel sweeps to a fixed point.  The elimination order is a
    minimum-degree order with seeded random tie-breaking, the best of `trials` by the instruction
    estimate of what the kernel executes per scenario and iteration: 4 per filled edge (two sweeps)
    plus 8 + 9 T per customization batch (T = most triangles of an edge in the batch).
    Returns (order, filled edge list, lower triangles per edge).
    """
    import random

    rnd = random.Random(20240229)

    def pick(alive, adj):
        md = min(len(adj[v] & alive) for v in alive)
        slack = 1 if rnd.random() < 0.2 else 0
        return rnd.choice(sorted(v for v in alive if len(adj[v] & alive) <= md + slack))

    best = None
    for _ in range(max(trials, 1)):
        order, edges, tris = _cch_eliminate(N, tails, heads, pick)
        batches = _cch_batches(edges, tris, _cch_levels(edges, tris), key=lambda e: e)
        cost = 4 * len(edges) + sum(8 + 9 * len(tris[b[0]]) for b in batches)
        if best is None or cost < best[0]:
            best = (cost, order, edges, tris)
    return best[1], best[2], best[3]


def _strongly_connected(N, tails, heads) -> bool:
    for a, b in ((tails, heads), (heads, tails)):
        nxt = [[] for _ in range(N)]
        for u, v in zip(a, b):
            nxt[u].append(v)
        seen, stack = {0}, [0]
        while stack:
            for v in nxt[stack.pop()]:
                if v not in seen:
                    seen.add(v)
                    stack.append(v)
        if len(seen) != N:
            return False
    return True


def emit_cch(N, koff, src) -> str:
    """constexpr CCH tables for the header (see cch_plan): W slot layout, sweep schedules and the
    table-driven customization batches."""
    heads = [i for i in range(N) for _ in range(koff[i], koff[i + 1])]
    tails = list(src)
    order, edges, tris = cch_plan(N, tails, heads)
    rank = [0] * N
    for r, v in enumerate(order):
        rank[v] = r
    # W slots: for every node in elimination order its in-edges from lower-ranked tails (up group),
    # then for every node its in-edges from higher-ranked tails (down group); groups are padded to an
    # even slot count so a group is ceil(n/2) aligned 16-byte loads
    slot_of, csrc = {}, []
    up_off, up_cnt, dn_off, dn_cnt = [], [], [], []
    for lower, offs, cnts in ((True, up_off, up_cnt), (False, dn_off, dn_cnt)):
        for v in order:
            grp = sorted((rank[u], u) for (u, x) in edges if x == v and (rank[u] < rank[v]) == lower)
            offs.append(len(csrc))
            cnts.append(len(grp))
            for _, u in grp:
                slot_of[(u, v)] = len(csrc)
                csrc.append(u)
            if len(grp) % 2:
                csrc.append(-1)
    cw = len(csrc)
    inf_slot, dummy_slot = cw, cw + 1
    orig = set(zip(tails, heads))
    corig = [slot_of[(tails[k], heads[k])] for k in range(len(tails))]
    # customization batches: an edge is final once the two lower edges of each of its triangles are.
    # Two tables: 32 edges per batch (one warp per scenario) and N edges per batch (grouped kernel: N
    # lanes per scenario, several scenarios per group of warps); rows are padded to 32 words either way.
    def table(width):
        tab, batch_t = [], []
        for chunk in _cch_batches(edges, tris, _cch_levels(edges, tris), key=lambda e: slot_of[e], width=width):
            T = len(tris[chunk[0]])
            batch_t.append(T)
            tab += [(slot_of[e] * 8) | (0 if e in orig else 0x8000) for e in chunk]
            tab += [(dummy_slot * 8) | 0x8000] * (32 - len(chunk))
            for t in range(T):
                for e in chunk:
                    if t < len(tris[e]):
                        v = tris[e][t]
                        tab.append((slot_of[(e[0], v)] * 8) | ((slot_of[(v, e[1])] * 8) << 16))
                    else:
                        tab.append((inf_slot * 8) | ((inf_slot * 8) << 16))
                tab += [(inf_slot * 8) | ((inf_slot * 8) << 16)] * (32 - len(chunk))
        return tab, batch_t

    tab, batch_t = table(32)
    gtab, gbatch_t = table(min(N, 32))

    def arr(xs):
        return "{" + ", ".join(str(int(x)) for x in xs) + "}"

    ntri = sum(len(v) for v in tris.values())
    return f"""    // ---- customizable contraction hierarchy (codegen.cch_plan): {len(edges)} filled edges, {ntri} lower
    // triangles, {len(batch_t)} customization batches.  CORD = elimination order (internal node ids); W slot
    // groups of the node eliminated r-th: in-edges from lower-ranked tails at CUP_OFF[r] (CUP_CNT[r] of
    // them), from higher-ranked tails at CDN_OFF[r]; CSRC = tail of each W slot (-1 = padding); CORIG = W
    // slot of the original link in CSC slot k; slot CW holds +inf, slot CW+1 is a write-only dummy.
    // CZ_TAB: per batch 32 edge words (W byte offset | 0x8000 if the edge is a pure shortcut), then
    // CZ_T[b] x 32 triangle words (byte offset of the lower edge (u,v) | (v,x) << 16).
    static constexpr bool STRONG = {"true" if _strongly_connected(N, tails, heads) else "false"};
    static constexpr int CF = {len(edges)}, CW = {cw}, CZ_NB = {len(batch_t)}, CZ_LEN = {len(tab)};
    static constexpr int CORD[{N}] = {arr(order)};
    static constexpr int CUP_OFF[{N}] = {arr(up_off)};
    static constexpr int CUP_CNT[{N}] = {arr(up_cnt)};
    static constexpr int CDN_OFF[{N}] = {arr(dn_off)};
    static constexpr int CDN_CNT[{N}] = {arr(dn_cnt)};
    static constexpr int CSRC[{cw}] = {arr(csrc)};
    static constexpr int CORIG[{len(corig)}] = {arr(corig)};
    static constexpr int CZ_T[{max(len(batch_t), 1)}] = {arr(batch_t or [0])};
    static constexpr unsigned CZ_TAB[{max(len(tab), 1)}] = {arr(tab or [0])};
    // the same customization in batches of N edges (grouped kernel: lane o of a scenario takes word o of a row)
    static constexpr int CZG_NB = {len(gbatch_t)}, CZG_LEN = {len(gtab)};
    static constexpr int CZG_T[{max(len(gbatch_t), 1)}] = {arr(gbatch_t or [0])};
    static constexpr unsigned CZG_TAB[{max(len(gtab), 1)}] = {arr(gtab or [0])};
"""


def emit(name: str, guard: str, N, row, col, t0, origins, source: str) -> str:
    order, root, cost = choose_order(N, row, col, t0, origins)
    pos = np.empty(N, np.int64)
    pos[order] = np.arange(N)
    E = len(row)
    koff, src, eid = [0], [], []
    for i in range(N):
        v = order[i]
        links = sorted((int(row[e]), e) for e in range(E) if col[e] == v)  # ascending ORIGINAL tail id
        for u, e in links:
            src.append(int(pos[u]))
            eid.append(e)
        koff.append(len(src))
    maxd = max(koff[i + 1] - koff[i] for i in range(N))

    def arr(xs):
        return "{" + ", ".join(str(int(x)) for x in xs) + "}"

    return f"""// GENERATED by tools/gen_static_topology.py from {source} - data only, do not edit.
// Topology of one network for the network-specialised kernel (b200_assign_static.cuh).
// Internal node index = position in the Gauss-Seidel sweep order (ascending free-flow distance
// from original node {root}; {cost} sweeps over the 6 calibration patterns).
#pragma once
struct {name} {{
    static constexpr int N = {N}, E = {E}, MAXD = {maxd};
    // original 0-based node id of internal node i
    static constexpr int NODE[{N}] = {arr(order)};
    // CSC over internal nodes: in-links of node i are slots KOFF[i] .. KOFF[i+1]-1, listed by
    // ascending ORIGINAL tail id; SRC = internal tail index, EID = link id (TNTP row order)
    static constexpr int KOFF[{N + 1}] = {arr(koff)};
    static constexpr int SRC[{E}] = {arr(src)};
    static constexpr int EID[{E}] = {arr(eid)};
    // the link list this header was generated from (row -> col, 0-based), matched at graph upload
    static constexpr int ROW[{E}] = {arr(row)};
    static constexpr int COL[{E}] = {arr(col)};
{emit_cch(N, koff, src)}}};
"""
