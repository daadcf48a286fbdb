"""ORACLE (test infrastructure) - the reference itself, imported from the verbatim copy in
oracle/_ref/ (made by oracle/fetch_ref.py; git-ignored, travels to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's CPU / reference legs may import this module.
It adds NOTHING to the reference: it puts the copy on sys.path, imports the reference's own modules
(`src.env.repair_env`, `src.data.tntp_parser`, `src.baselines`, `src.eval`) and offers the few lines
of driver code a caller of the reference would write (construct an env, poke a damage mask the way the
reference's greedy baseline does - src/baselines/__init__.py:43-67 - and call
`compute_flow_assignment()`, src/env/repair_env.py:299).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_ROOT, "src", "env", "repair_env.py"))


def root() -> str:
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/fetch_ref.py` in the build container "
                           "(it copies the unmodified reference; the copy travels to the GPU box)")
    return REF_ROOT


def modules():
    """(repair_env, tntp_parser, baselines) modules of the reference."""
    r = root()
    if r not in sys.path:
        sys.path.insert(0, r)
    import importlib
    return (importlib.import_module("src.env.repair_env"), importlib.import_module("src.data.tntp_parser"),
            importlib.import_module("src.baselines"))


def eval_module():
    modules()
    import importlib
    return importlib.import_module("src.eval")


def sioux_falls_graph():
    _, parser, _ = modules()
    d = os.path.join(root(), "data", "SiouxFalls")
    return parser.load_graph_data(os.path.join(d, "SiouxFalls_net.tntp"), os.path.join(d, "SiouxFalls_trips.tntp"))


def make_env(graph=None, quiet: bool = True, **kw):
    """The reference's RepairEnv on its CPU path (scipy shortest paths, numpy BPR) unless told otherwise."""
    env_mod, _, _ = modules()
    kw.setdefault("sp_backend", "scipy")
    kw.setdefault("use_torch", False)
    kw.setdefault("device", "cpu")
    graph = sioux_falls_graph() if graph is None else graph
    with (contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()):
        return env_mod.RepairEnv(graph, **kw)


def assign_cold(env, damaged_mask: np.ndarray) -> float:
    """One cold-start compute_flow_assignment() of the reference env for a given damage mask (the
    attribute poke of the reference's own greedy baseline, then the reference's own method)."""
    m = np.asarray(damaged_mask) > 0
    env.is_damaged = m.astype(np.float32)
    env.capacities = env.initial_capacities.copy()
    env.capacities[m] = env.capacity_damage
    env.flow = np.zeros(env.num_edges, dtype=np.float32)
    env.compute_flow_assignment()
    return float(env.tstt)


def time_assignments(env, masks: np.ndarray) -> float:
    """Wall seconds of len(masks) cold-start assignments on `env`."""
    t0 = time.perf_counter()
    for m in masks:
        assign_cold(env, m)
    return time.perf_counter() - t0
