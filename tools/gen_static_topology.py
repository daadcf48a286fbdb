#!/usr/bin/env python
"""Emit a constexpr topology header for the network-specialised assignment kernel
(csrc/b200_assign_static.cuh).  DATA ONLY: node order, CSC in-link lists and link ids of one
network; the kernel template that consumes it is hand-written.

    python tools/gen_static_topology.py            # Sioux Falls -> csrc/b200_topo_siouxfalls.h

Internal node numbering = position in the Gauss-Seidel sweep order.  The order is the ascending
free-flow distance from the root that needs the fewest alternating sweeps (all origins at once,
what the kernel does) on the free-flow costs and five deterministic damage/jitter patterns.
In-links of a node are listed by ascending ORIGINAL tail node id (the scipy tie rule picks the
largest tail among equal candidates).
"""
from __future__ import annotations

import os
import sys


REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


from importlib import import_module

emit = import_module("sac-gat-her_transportationrl_b200.codegen").emit


def main() -> None:
    import b200env

    g = b200env.load_sioux_falls()
    net = b200env.Network(g)
    text = emit("TopoSiouxFalls", "B200_TOPO_SIOUXFALLS_H", net.num_nodes, net.edge_index[0], net.edge_index[1],
                net.t0, [int(o) for o in net.origins], "data/networks/SiouxFalls_net.tntp")
    out = os.path.join(REPO, "sac-gat-her_transportationrl_b200", "csrc", "b200_topo_siouxfalls.h")
    with open(out, "w") as fh:
        fh.write(text)
    print("wrote", out)


if __name__ == "__main__":
    main()
