"""Static network: host arrays in the reference's layout + the device-resident graph handle.

Reference: the arrays ``RepairEnv.__init__`` derives from ``GraphData``
(src/env/repair_env.py:85-97): ``edge_index`` int64 [2,E] 0-based, ``initial_capacities`` /
``t0`` fp32, ``edge_id_map``, ``total_demand``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import numpy as np

from .. import _native


class Network:
    def __init__(self, graph_data, device_index: int = 0, specialize: bool = False):
        """``specialize=True``: compile (once, cached) and use a straight-line assignment kernel
        generated for THIS topology, as Sioux Falls gets out of the box (small networks only,
        ``_native.specialize``)."""
        self.specialize = bool(specialize)
        g = graph_data
        self.graph_data = g
        self.num_nodes = int(g.num_nodes)
        self.edges = g.edges
        self.num_edges = len(g.edges)
        self.edge_index = np.array([[e.u - 1 for e in g.edges], [e.v - 1 for e in g.edges]], dtype=np.int64)
        self.initial_capacities = np.array([e.capacity for e in g.edges], dtype=np.float32)
        self.t0 = np.array([e.t0 for e in g.edges], dtype=np.float32)
        self.max_capacity = float(np.max(self.initial_capacities)) if self.num_edges > 0 else 1.0
        self.max_t0 = float(np.max(self.t0)) if self.num_edges > 0 else 1.0
        self.edge_id_map: Dict[Tuple[int, int], int] = {(e.u - 1, e.v - 1): i for i, e in enumerate(g.edges)}
        self.total_demand = float(np.sum(list(g.od_demand.values()))) if g.od_demand else 0.0
        self.device_index = int(device_index)
        self._handle = None
        self._od = (
            np.ascontiguousarray([o - 1 for (o, _d) in g.od_demand.keys()], dtype=np.int32),
            np.ascontiguousarray([d - 1 for (_o, d) in g.od_demand.keys()], dtype=np.int32),
            np.ascontiguousarray(list(g.od_demand.values()), dtype=np.float64),
        )
        self.origins = np.unique(self._od[0]).astype(np.int32)

    # -- device handle ------------------------------------------------------------------------
    @property
    def handle(self):
        if self._handle is None:
            lib = _native.lib()
            if self.specialize:
                _native.specialize(self.num_nodes, self.edge_index[0], self.edge_index[1], self.t0, self.origins)
            row = np.ascontiguousarray(self.edge_index[0], dtype=np.int32)
            col = np.ascontiguousarray(self.edge_index[1], dtype=np.int32)
            out = C.c_void_p()
            rc = lib.b200_graph_create(
                self.device_index, self.num_nodes, self.num_edges, row.ctypes.data, col.ctypes.data,
                self.t0.ctypes.data, self.initial_capacities.ctypes.data, int(self._od[0].shape[0]),
                self._od[0].ctypes.data, self._od[1].ctypes.data, self._od[2].ctypes.data, C.byref(out))
            _native.check(rc, "b200_graph_create")
            self._handle = out
            info = (C.c_int64 * 8)()
            _native.check(lib.b200_graph_info(self._handle, info), "b200_graph_info")
            self.info = dict(N=info[0], E=info[1], O=info[2], num_od=info[3], max_in_degree=info[4],
                             smem_f32=info[5], smem_f64=info[6], num_sms=info[7])
        return self._handle

    @property
    def num_origins(self) -> int:
        return int(self.origins.shape[0])

    def close(self) -> None:
        if self._handle is not None:
            _native.lib().b200_graph_destroy(self._handle)
            self._handle = None

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:
            pass

    # -- host-side connectivity test used by the damage sampler --------------------------------
    def active_subgraph_strongly_connected(self, damaged_mask: np.ndarray) -> bool:
        """Strong connectivity of the edge-induced subgraph of undamaged links (the test the
        reference runs through networkx, src/env/repair_env.py:178-186): nodes that lost every
        link are not part of the subgraph."""
        act = np.flatnonzero(np.asarray(damaged_mask) == 0)
        if act.size == 0:
            return False
        src, dst = self.edge_index[0][act], self.edge_index[1][act]
        nodes = np.unique(np.concatenate([src, dst]))
        n = self.num_nodes
        for a, b in ((src, dst), (dst, src)):  # forward, then reverse reachability from nodes[0]
            order = np.argsort(a, kind="stable")
            a_s, b_s = a[order], b[order]
            ptr = np.searchsorted(a_s, np.arange(n + 1))
            seen = np.zeros(n, dtype=bool)
            seen[nodes[0]] = True
            stack = [int(nodes[0])]
            while stack:
                u = stack.pop()
                for v in b_s[ptr[u]:ptr[u + 1]]:
                    if not seen[v]:
                        seen[v] = True
                        stack.append(int(v))
            if not seen[nodes].all():
                return False
        return True

    def sample_damage(self, rng: np.random.Generator, damaged_ratio: float, max_retries: int = 50) -> np.ndarray:
        """Damage draw with the reference's retry rule (src/env/repair_env.py:168-190)."""
        count = max(1, int(self.num_edges * damaged_ratio))
        for _ in range(max_retries):
            cand = rng.choice(self.num_edges, size=count, replace=False)
            mask = np.zeros(self.num_edges, dtype=np.float32)
            mask[cand] = 1.0
            if self.active_subgraph_strongly_connected(mask):
                return cand
        return rng.choice(self.num_edges, size=count, replace=False)
