"""ORACLE (test infrastructure) - ctypes front end of oracle/_build/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
``build()`` compiles the C restatement with gcc (oracle/Makefile); building the checker is not
using it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
_lib = None

METHODS = {"msa": 0, "fw": 1, "cfw": 2, "fw_ls": 3}


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in ("assign_oracle.c", "assign_body.inc", "Makefile")]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
    return _lib


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


class COracle:
    """Holds the static network as contiguous arrays and exposes the stage functions."""

    def __init__(self, num_nodes, row, col, t0, cap0, od):
        self.N = int(num_nodes)
        self.row = np.ascontiguousarray(row, dtype=np.int32)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.E = int(self.row.shape[0])
        self.t0 = np.ascontiguousarray(t0, dtype=np.float32)
        self.cap0 = np.ascontiguousarray(cap0, dtype=np.float32)
        self.od_o = np.ascontiguousarray([o for (o, _, _) in od], dtype=np.int32)
        self.od_d = np.ascontiguousarray([d for (_, d, _) in od], dtype=np.int32)
        self.od_q = np.ascontiguousarray([q for (_, _, q) in od], dtype=np.float64)
        self.origins = np.unique(self.od_o).astype(np.int32)
        self.total_demand = float(self.od_q.sum())

    @classmethod
    def from_graph_data(cls, g):
        return cls(g.num_nodes, [e.u - 1 for e in g.edges], [e.v - 1 for e in g.edges], [e.t0 for e in g.edges],
                   [e.capacity for e in g.edges], [(o - 1, d - 1, q) for (o, d), q in g.od_demand.items()])

    # -- stages -------------------------------------------------------------------------------
    def bpr(self, flow, dmg, alpha=0.15, beta=4.0, dtype=np.float32):
        flow = np.ascontiguousarray(flow, dtype=dtype)
        dmg = np.ascontiguousarray(dmg, dtype=np.uint8)
        t = np.empty(self.E, dtype=dtype)
        ct = C.c_float if dtype == np.float32 else C.c_double
        fn = lib().orc_bpr_f32 if dtype == np.float32 else lib().orc_bpr_f64
        fn(C.c_int(self.E), _p(flow, ct), _p(self.cap0, C.c_float), _p(dmg, C.c_ubyte), _p(self.t0, C.c_float),
           C.c_double(alpha), C.c_double(beta), _p(t, ct))
        return t

    def sssp(self, t, origins=None):
        origins = self.origins if origins is None else np.ascontiguousarray(origins, dtype=np.int32)
        w = np.ascontiguousarray(t, dtype=np.float64)
        no = origins.shape[0]
        dist = np.empty((no, self.N), np.float64)
        pred = np.empty((no, self.N), np.int32)
        pedge = np.empty((no, self.N), np.int32)
        lib().orc_sssp(C.c_int(self.N), C.c_int(self.E), _p(self.row, C.c_int), _p(self.col, C.c_int), _p(w, C.c_double),
                       C.c_int(no), _p(origins, C.c_int), _p(dist, C.c_double), _p(pred, C.c_int), _p(pedge, C.c_int))
        return dist, pred, pedge

    def aon(self, pred, pedge, dtype=np.float32):
        pred = np.ascontiguousarray(pred, np.int32)
        pedge = np.ascontiguousarray(pedge, np.int32)
        aux = np.empty(self.E, dtype)
        un = C.c_double(0)
        ct = C.c_float if dtype == np.float32 else C.c_double
        fn = lib().orc_aon_f32 if dtype == np.float32 else lib().orc_aon_f64
        fn(C.c_int(self.N), C.c_int(self.E), _p(self.row, C.c_int), _p(self.col, C.c_int), C.c_int(self.od_o.shape[0]),
           _p(self.od_o, C.c_int), _p(self.od_d, C.c_int), _p(self.od_q, C.c_double), _p(pred, C.c_int),
           _p(pedge, C.c_int), _p(aux, ct), C.byref(un))
        return aux, un.value

    def pairwise_sum(self, a):
        a = np.ascontiguousarray(a)
        if a.dtype == np.float32:
            lib().orc_pairwise_sum_f32.restype = C.c_float
            return lib().orc_pairwise_sum_f32(_p(a, C.c_float), C.c_long(a.shape[0]))
        lib().orc_pairwise_sum_f64.restype = C.c_double
        return lib().orc_pairwise_sum_f64(_p(a.astype(np.float64), C.c_double), C.c_long(a.shape[0]))

    def betweenness(self, dmg):
        dmg = np.ascontiguousarray(dmg, dtype=np.uint8)
        bw = np.empty(self.N, np.float64)
        lib().orc_betweenness(C.c_int(self.N), C.c_int(self.E), _p(self.row, C.c_int), _p(self.col, C.c_int),
                              _p(dmg, C.c_ubyte), _p(bw, C.c_double))
        return bw

    # -- full loop ----------------------------------------------------------------------------
    def assign(self, dmg, flow0=None, iters=20, method="msa", alpha=0.15, beta=4.0, penalty=2e7, ls_iters=24,
               dtype=np.float32, trace=False):
        """Returns dict(flow, tstt, gap, unassigned[, t, pred, pedge, dist, aux, flows, gaps])."""
        dmg = np.ascontiguousarray(dmg, dtype=np.uint8)
        flow = np.zeros(self.E, dtype) if flow0 is None else np.ascontiguousarray(flow0, dtype=dtype).copy()
        ct = C.c_float if dtype == np.float32 else C.c_double
        fn = lib().orc_assign_f32 if dtype == np.float32 else lib().orc_assign_f64
        tstt, gap, un = C.c_double(0), C.c_double(0), C.c_double(0)
        K, no = int(iters), int(self.origins.shape[0])
        tr = {}
        if trace:
            tr = dict(t=np.empty((K, self.E), dtype), pred=np.empty((K, no, self.N), np.int32),
                      pedge=np.empty((K, no, self.N), np.int32), dist=np.empty((K, no, self.N), np.float64),
                      aux=np.empty((K, self.E), dtype), flows=np.empty((K, self.E), dtype), gaps=np.empty(K, np.float64))
        rc = fn(C.c_int(self.N), C.c_int(self.E), _p(self.row, C.c_int), _p(self.col, C.c_int), _p(self.t0, C.c_float),
                _p(self.cap0, C.c_float), C.c_int(self.od_o.shape[0]), _p(self.od_o, C.c_int), _p(self.od_d, C.c_int),
                _p(self.od_q, C.c_double), _p(dmg, C.c_ubyte), _p(flow, ct), C.c_int(K), C.c_int(METHODS[method]),
                C.c_double(alpha), C.c_double(beta), C.c_double(penalty), C.c_int(ls_iters),
                C.byref(tstt), C.byref(gap), C.byref(un),
                _p(tr.get("t"), ct), _p(tr.get("pred"), C.c_int), _p(tr.get("pedge"), C.c_int),
                _p(tr.get("dist"), C.c_double), _p(tr.get("aux"), ct), _p(tr.get("flows"), ct),
                _p(tr.get("gaps"), C.c_double))
        if rc != 0:
            raise ValueError("assignment_iters must be > 0 to update TSTT.")
        out = dict(flow=flow, tstt=tstt.value, gap=gap.value, unassigned=un.value)
        out.update(tr)
        return out
