"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs the read-only reference checkout):

    python tests/golden/make_golden.py [--only sf_traces,sf_batch,...] [--procs 8]

The reference (``/root/reference``, Python) is imported unmodified; per-iteration traces are taken
by wrapping ``scipy.sparse.csgraph.dijkstra`` and ``RepairEnv._all_or_nothing`` from the outside.
Nothing here is read at test time except the ``.npz`` files it writes.  Host that produced the
committed files: see ``meta.json`` (numpy SIMD dispatch matters at the fp32-ulp level).
"""
from __future__ import annotations

import argparse
import importlib
import io
import json
import os
import platform
import sys
import contextlib
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

import scipy.sparse.csgraph as csgraph  # noqa: E402
from src.data.tntp_parser import load_graph_data as ref_load  # noqa: E402
from src.data.tntp_parser import GraphData as RefGraph, EdgeData as RefEdge  # noqa: E402
from src.env.repair_env import RepairEnv as RefEnv  # noqa: E402

pkg = importlib.import_module("sac-gat-her_transportationrl_b200")
synthetic = importlib.import_module("sac-gat-her_transportationrl_b200.data.synthetic")

SF_NET = os.path.join(REF, "data/SiouxFalls/SiouxFalls_net.tntp")
SF_TRIPS = os.path.join(REF, "data/SiouxFalls/SiouxFalls_trips.tntp")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def to_ref_graph(g) -> RefGraph:
    return RefGraph(num_nodes=g.num_nodes,
                    edges=[RefEdge(u=e.u, v=e.v, capacity=e.capacity, t0=e.t0, length=e.length, b=e.b, power=e.power)
                           for e in g.edges],
                    od_demand=dict(g.od_demand))


class Tracer:
    """Records (t, pred, aux, flow) per iteration of one compute_flow_assignment call."""

    def __init__(self, env):
        self.env = env
        self.t, self.pred, self.aux, self.flow = [], [], [], []
        self._orig_dij = csgraph.dijkstra
        self._orig_aon = env._all_or_nothing
        self._orig_tt = env.compute_travel_time

    def __enter__(self):
        def dij(*a, **k):
            out = self._orig_dij(*a, **k)
            self.pred.append(np.asarray(out[1]).copy())
            return out

        def aon(t):
            self.t.append(np.asarray(t).copy())
            aux, un = self._orig_aon(t)
            self.aux.append(aux.copy())
            return aux, un

        def tt(flow):
            # called once before the loop and once per iteration *after* the update
            self.flow.append(np.asarray(flow).copy())
            return self._orig_tt(flow)

        csgraph.dijkstra = dij
        self.env._all_or_nothing = aon
        self.env.compute_travel_time = tt
        return self

    def __exit__(self, *exc):
        csgraph.dijkstra = self._orig_dij
        self.env._all_or_nothing = self._orig_aon
        self.env.compute_travel_time = self._orig_tt

    def arrays(self, origins=None):
        pred = np.stack(self.pred).astype(np.int16)
        if origins is not None:
            pred = pred[:, origins, :]
        return dict(t=np.stack(self.t), pred=pred, aux=np.stack(self.aux), flow=np.stack(self.flow[1:]))


# ---------------------------------------------------------------------------------------------

def gen_sf_traces():
    g = ref_load(SF_NET, SF_TRIPS)
    out = {}
    for seed in (0, 1, 2):
        for method in ("msa", "fw", "cfw"):
            env = quiet(RefEnv, g, seed=seed, assignment_method=method, assignment_iters=20)
            key = f"s{seed}_{method}"
            dmg_idx = np.flatnonzero(env.is_damaged > 0)
            # replay the reset assignment with tracing (cold start)
            env.flow = np.zeros(env.num_edges, np.float32)
            env.is_reset = True
            with Tracer(env) as tr:
                env.compute_flow_assignment()
            env.is_reset = False
            for k, v in tr.arrays().items():
                out[f"{key}_reset_{k}"] = v
            out[f"{key}_dmg"] = dmg_idx.astype(np.int16)
            out[f"{key}_reset_tstt"] = np.float64(env.tstt)
            assert env.tstt == env.initial_tstt
            a = int(dmg_idx[0])
            with Tracer(env) as tr:
                _, reward, done, info = env.step(a)
            for k, v in tr.arrays().items():
                out[f"{key}_step_{k}"] = v
            out[f"{key}_step_tstt"] = np.float64(info["tstt"])
            out[f"{key}_step_reward"] = np.float64(reward)
            out[f"{key}_step_done"] = np.bool_(done)
    np.savez_compressed(os.path.join(HERE, "sf_traces.npz"), **out)


def gen_sf_kat():
    g = ref_load(SF_NET, SF_TRIPS)
    out = {}
    for method in ("msa", "fw", "cfw"):
        for K in (20, 50, 60, 200):
            env = quiet(RefEnv, g, seed=0, assignment_method=method, assignment_iters=K)
            env.is_damaged[:] = 0.0
            env.capacities = env.initial_capacities.copy()
            env.flow = np.zeros(env.num_edges, np.float32)
            env.compute_flow_assignment()
            out[f"undamaged_{method}_K{K}_tstt"] = np.float64(env.tstt)
            out[f"undamaged_{method}_K{K}_flow"] = env.flow.copy()
    for method in ("msa", "fw"):
        tst = []
        for seed in range(1, 6):
            env = quiet(RefEnv, g, seed=seed, assignment_method=method, assignment_iters=20)
            tst.append(env.tstt)
        out[f"seeds1to5_{method}_K20_tstt"] = np.array(tst)
    # reward modes
    env = quiet(RefEnv, g, seed=0)
    rows = []
    gm = env.goal_mask.copy()
    none = np.zeros_like(gm)
    for mode in ("delta", "log_delta", "neg_tstt", "minimize_tstt", "rel_improve"):
        for prev, curr in ((68636.0, 68384.7), (30.0, 20.8), (0.5, 0.7)):
            for dm in (gm, none):
                for clip in (0.0, 2.0):
                    r = env.compute_reward_with_goal(prev, curr, gm, dm, alpha=1.0, beta=10.0, gamma=0.1, mode=mode, clip=clip)
                    rows.append((mode, prev, curr, float(dm.sum() == 0), clip, float(r)))
    out["reward_modes"] = np.array([r[0] for r in rows])
    out["reward_rows"] = np.array([r[1:] for r in rows], dtype=np.float64)
    out["reward_initial_tstt"] = np.float64(env.initial_tstt)
    np.savez_compressed(os.path.join(HERE, "sf_kat.npz"), **out)


def _batch_one(b):
    g = ref_load(SF_NET, SF_TRIPS)
    env = quiet(RefEnv, g, seed=1000 + b, assignment_method="fw", assignment_iters=20)
    dmg = np.flatnonzero(env.is_damaged > 0).astype(np.int16)
    flow_r, tstt_r = env.flow.copy(), env.tstt
    _, reward, done, info = env.step(int(dmg[0]))
    return dmg, flow_r, tstt_r, env.flow.copy(), info["tstt"], reward


def gen_sf_batch(procs: int, B: int = 4096, keep_flows: int = 512):
    with Pool(procs) as pool:
        res = pool.map(_batch_one, range(B), chunksize=16)
    np.savez_compressed(
        os.path.join(HERE, "sf_batch4096_fw_k20.npz"),
        dmg=np.stack([r[0] for r in res]),
        flow_reset=np.stack([r[1] for r in res[:keep_flows]]),
        tstt_reset=np.array([r[2] for r in res], dtype=np.float64),
        flow_step=np.stack([r[3] for r in res[:keep_flows]]),
        tstt_step=np.array([r[4] for r in res], dtype=np.float64),
        reward_step=np.array([r[5] for r in res], dtype=np.float64),
    )


def gen_sf_features():
    g = ref_load(SF_NET, SF_TRIPS)
    out = {}
    for seed in range(6):
        env = quiet(RefEnv, g, seed=seed, assignment_method="fw", assignment_iters=20)
        st = env.get_state()
        out[f"s{seed}_dmg0"] = env.is_damaged.copy()
        out[f"s{seed}_node0"], out[f"s{seed}_edge0"] = st.node_features, st.edge_features
        acts = np.flatnonzero(env.is_damaged > 0)[:3]
        for a in acts:
            st, _, _, _ = env.step(int(a))
        out[f"s{seed}_acts"] = acts.astype(np.int16)
        out[f"s{seed}_node3"], out[f"s{seed}_edge3"] = st.node_features, st.edge_features
        out[f"s{seed}_mask3"] = st.action_mask
        out[f"s{seed}_logtstt3"] = np.float64(st.log_tstt)
    out["static_betweenness"] = env.betweenness_vec.copy()
    np.savez_compressed(os.path.join(HERE, "sf_features.npz"), **out)


def gen_sf_episode():
    g = ref_load(SF_NET, SF_TRIPS)
    out = {}
    for method, K, seed in (("fw", 50, 43), ("msa", 20, 7)):
        env = quiet(RefEnv, g, seed=seed, assignment_method=method, assignment_iters=K, unassigned_penalty=1e4)
        dmg = np.flatnonzero(env.is_damaged > 0)
        curve, rew, dones = [env.tstt], [], []
        # one invalid action in the middle exercises the -1 branch
        for i, a in enumerate(dmg):
            if i == 3:
                _, r, d, info = env.step(int(dmg[0]))
                curve.append(info["tstt"]); rew.append(r); dones.append(d)
            _, r, d, info = env.step(int(a))
            curve.append(info["tstt"]); rew.append(r); dones.append(d)
        key = f"{method}_K{K}_s{seed}"
        out[f"{key}_dmg"] = dmg.astype(np.int16)
        out[f"{key}_tstt"] = np.array(curve)
        out[f"{key}_reward"] = np.array(rew)
        out[f"{key}_done"] = np.array(dones)
        out[f"{key}_final_flow"] = env.flow.copy()
    np.savez_compressed(os.path.join(HERE, "sf_episode.npz"), **out)


def _synth_case(args):
    kind, net_seed, env_seed, method, K = args
    g = synthetic.anaheim_size(net_seed) if kind == "an" else synthetic.chicago_sketch_size(net_seed)
    rg = to_ref_graph(g)
    env = quiet(RefEnv, rg, seed=env_seed, assignment_method=method, assignment_iters=K)
    origins = sorted({o - 1 for (o, _) in g.od_demand})
    env.flow = np.zeros(env.num_edges, np.float32)
    env.is_reset = True
    with Tracer(env) as tr:
        env.compute_flow_assignment()
    env.is_reset = False
    arr = tr.arrays(origins=origins if kind == "an" else origins[:8])
    return dict(dmg=np.flatnonzero(env.is_damaged > 0).astype(np.int16), tstt=np.float64(env.tstt),
                flow=env.flow.copy(), pred=arr["pred"], t=arr["t"], aux=arr["aux"])


def gen_an(procs):
    cases = [("an", 0, s, m, 5) for s in (0, 1) for m in ("fw", "msa")]
    with Pool(min(procs, len(cases))) as pool:
        res = pool.map(_synth_case, cases)
    out = {}
    for c, r in zip(cases, res):
        for k, v in r.items():
            out[f"s{c[2]}_{c[3]}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "an_traces.npz"), **out)


def gen_cs():
    r = _synth_case(("cs", 0, 0, "msa", 2))
    np.savez_compressed(os.path.join(HERE, "cs_k2.npz"), **r)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="sf_traces,sf_kat,sf_features,sf_episode,sf_batch,an,cs")
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    todo = set(args.only.split(","))
    if "sf_traces" in todo:
        gen_sf_traces()
    if "sf_kat" in todo:
        gen_sf_kat()
    if "sf_features" in todo:
        gen_sf_features()
    if "sf_episode" in todo:
        gen_sf_episode()
    if "sf_batch" in todo:
        gen_sf_batch(args.procs)
    if "an" in todo:
        gen_an(args.procs)
    if "cs" in todo:
        gen_cs()
    import scipy, networkx
    meta = dict(numpy=np.__version__, scipy=scipy.__version__, networkx=networkx.__version__,
                python=platform.python_version(), machine=platform.machine(),
                cpu_features=sorted(f for f in np._core._multiarray_umath.__cpu_features__
                                    if np._core._multiarray_umath.__cpu_features__[f]),
                reference="pop-pop-pOp-dev/SAC-GAT-HER_transportationRL (read-only checkout, unmodified)")
    with open(os.path.join(HERE, "meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1)


if __name__ == "__main__":
    main()
