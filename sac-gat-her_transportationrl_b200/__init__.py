"""B200-native batched traffic-assignment environment.

Drop-in for the env-step hot path of pop-pop-pOp-dev/SAC-GAT-HER_transportationRL
(``src/env/repair_env.py``): ``RepairEnv`` / ``EnvState`` keep the reference API, while
``BatchedRepairEnv`` advances thousands of independent damage/repair scenarios per launch through
hand-written sm_100a CUDA kernels behind the C ABI in ``include/b200_assign.h``.
"""
from . import _native
from .data.tntp_parser import EdgeData, GraphData, load_graph_data, parse_net_tntp, parse_trips_tntp
from .data import synthetic
from .env.network import Network
from .env.batched_env import AssignmentEngine, BatchedEnvState, BatchedRepairEnv
from .env.repair_env import EnvState, RepairEnv
from .models.gat import GATConvB200, GATEncoderB200, gat_edge_softmax_aggregate
from . import baselines
from . import rl

import os as _os

NETWORK_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "data", "networks")


def load_sioux_falls() -> GraphData:
    """The Sioux Falls network shipped with this package (24 nodes, 76 links, 528 OD pairs)."""
    return load_graph_data(_os.path.join(NETWORK_DIR, "SiouxFalls_net.tntp"),
                           _os.path.join(NETWORK_DIR, "SiouxFalls_trips.tntp"))


def build_native(force: bool = False, verbose: bool = False) -> str:
    return _native.build(force=force, verbose=verbose)


__all__ = [
    "AssignmentEngine", "BatchedEnvState", "GATConvB200", "GATEncoderB200", "gat_edge_softmax_aggregate", "BatchedRepairEnv", "EdgeData", "EnvState", "GraphData", "Network",
    "RepairEnv", "build_native", "load_graph_data", "load_sioux_falls", "parse_net_tntp", "parse_trips_tntp",
    "synthetic", "baselines", "rl",
]
