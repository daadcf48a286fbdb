"""CPU tests of the host logic: TNTP reader/writer, synthetic generators, damage sampler,
C-ABI library loading (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import ref_port as rp


def test_sioux_falls_network_shape(sf_graph):
    assert sf_graph.num_nodes == 24 and len(sf_graph.edges) == 76 and len(sf_graph.od_demand) == 528
    assert sum(sf_graph.od_demand.values()) == 360600.0
    e0 = sf_graph.edges[0]
    assert (e0.u, e0.v, e0.t0, e0.b, e0.power) == (1, 2, 6.0, 0.15, 4.0) and abs(e0.capacity - 25900.20064) < 1e-9
    assert all(q > 0 for q in sf_graph.od_demand.values())


def test_tntp_round_trip_and_parser_quirks(b200env, sf_graph, tmp_path):
    tp = b200env.package.data.tntp_parser
    net, trips = str(tmp_path / "n.tntp"), str(tmp_path / "t.tntp")
    tp.write_net_tntp(net, sf_graph, num_zones=24)
    tp.write_trips_tntp(trips, sf_graph)
    g2 = tp.load_graph_data(net, trips)
    assert [vars(e) for e in g2.edges] == [vars(e) for e in sf_graph.edges]
    assert list(g2.od_demand.items()) == list(sf_graph.od_demand.items())
    # quirks kept from the reference reader: FIRST THRU NODE ignored, short rows skipped,
    # zero / negative demand dropped, '~' comment lines ignored
    (tmp_path / "q.tntp").write_text(
        "<NUMBER OF NODES> 3\n<FIRST THRU NODE> 2\n<ORIGINAL HEADER>~ Init node Term node\n<END OF METADATA>\n"
        "~ comment ;\n 1 2 100.5 1 2 0.15 4 0 0 1 ;\n 2 3 7 ;\n 2 3 50 1 3 0.2 ;\n")
    n, edges = tp.parse_net_tntp(str(tmp_path / "q.tntp"))
    assert n == 3 and [(e.u, e.v, e.t0, e.b, e.power) for e in edges] == [(1, 2, 2.0, 0.15, 4.0), (2, 3, 3.0, 0.2, 4.0)]
    (tmp_path / "d.tntp").write_text("<TOTAL OD FLOW> 1\n\nOrigin 1\n 1 : 0.0; 2 : 5.0; 3 : -1;\nOrigin  2 \n 1 : 2.5;\n")
    assert tp.parse_trips_tntp(str(tmp_path / "d.tntp")) == {(1, 2): 5.0, (2, 1): 2.5}


def test_synthetic_networks_match_named_sizes(b200env):
    an = b200env.synthetic.anaheim_size(0)
    assert (an.num_nodes, len(an.edges), len(an.od_demand)) == (416, 914, 1406)
    assert len({o for (o, _) in an.od_demand}) == 38
    pairs = [(e.u, e.v) for e in an.edges]
    assert pairs == sorted(pairs) and len(set(pairs)) == len(pairs) and all(u != v for u, v in pairs)
    assert all((v, u) in set(pairs) for u, v in pairs)
    assert all(float(e.t0).is_integer() and 1 <= e.t0 <= 10 for e in an.edges)
    assert all(float(q).is_integer() for q in an.od_demand.values())
    again = b200env.synthetic.anaheim_size(0)
    assert [vars(e) for e in again.edges] == [vars(e) for e in an.edges] and again.od_demand == an.od_demand
    net = b200env.Network(an)
    assert net.active_subgraph_strongly_connected(np.zeros(914))


def test_damage_sampler_matches_port_and_reference_goldens(b200env, sf_graph, golden):
    net = b200env.Network(sf_graph)
    pnet = rp.PortNetwork.from_graph_data(sf_graph)
    G = golden("sf_batch4096_fw_k20.npz")
    for b in range(0, 200, 7):
        idx = net.sample_damage(np.random.default_rng(1000 + b), 0.3)
        assert np.array_equal(np.sort(idx), G["dmg"][b])  # what the reference env drew with this seed
        idx2 = rp.sample_damage(pnet, np.random.default_rng(1000 + b), 0.3)
        assert np.array_equal(idx, idx2)
    rng = np.random.default_rng(3)
    for _ in range(200):
        m = (rng.random(76) < 0.45).astype(np.float32)
        assert net.active_subgraph_strongly_connected(m) == rp.active_subgraph_strongly_connected(pnet, m)
    assert not net.active_subgraph_strongly_connected(np.ones(76))


def test_c_abi_library_loads_and_exports_header_symbols(b200env, native_lib):
    header = open(os.path.join(os.path.dirname(__file__), "..", "include", "b200_assign.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(b200env._native.EXPORTS)
    for name in declared:
        assert getattr(native_lib, name) is not None
    assert native_lib.b200_version() >= 100
    # argument validation happens before any CUDA call
    out = ctypes.c_void_p()
    rc = native_lib.b200_graph_create(0, 0, 0, None, None, None, None, 0, None, None, None, ctypes.byref(out))
    assert rc == -1 and b"N > 0" in native_lib.b200_last_error()


def test_no_cpu_fallback(b200env, sf_graph):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(b200env._native.NativeLibraryError):
        b200env.AssignmentEngine(b200env.Network(sf_graph), "cuda")
    with pytest.raises(b200env._native.NativeLibraryError):
        b200env.RepairEnv(sf_graph, device="cpu")


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may import or load it."""
    root = os.path.join(os.path.dirname(__file__), "..", "sac-gat-her_transportationrl_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
                assert "liboracle" not in text and "ref_port" not in text, f


def test_topology_codegen_is_deterministic_and_consistent(b200env, sf_graph):
    """The generated topology header is data only; its CSC must describe exactly the uploaded link
    list, in-links by ascending original tail id, nodes renumbered by sweep position."""
    import importlib
    import re
    cg = importlib.import_module("sac-gat-her_transportationrl_b200.codegen")
    net = b200env.Network(sf_graph)
    row, col = net.edge_index
    text = cg.emit("TopoX", "X", net.num_nodes, row, col, net.t0, [int(o) for o in net.origins], "test")
    assert text == cg.emit("TopoX", "X", net.num_nodes, row, col, net.t0, [int(o) for o in net.origins], "test")
    arr = {m.group(1): [int(x) for x in m.group(2).split(",")]
           for m in re.finditer(r"constexpr int (\w+)\[\d+\] = \{([^}]*)\}", text)}
    node, koff, src, eid = arr["NODE"], arr["KOFF"], arr["SRC"], arr["EID"]
    assert sorted(node) == list(range(24)) and sorted(eid) == list(range(76)) and koff[-1] == 76
    for i in range(24):
        links = list(range(koff[i], koff[i + 1]))
        assert all(col[eid[k]] == node[i] for k in links)           # slot k enters internal node i
        assert all(node[src[k]] == row[eid[k]] for k in links)      # from the right tail
        tails = [int(row[eid[k]]) for k in links]
        assert tails == sorted(tails)                                # ascending ORIGINAL tail id (tie rule)
    # the committed built-in header is what the generator produces today
    committed = open(os.path.join(os.path.dirname(__file__), "..", "sac-gat-her_transportationrl_b200", "csrc",
                                  "b200_topo_siouxfalls.h")).read()
    assert arr["SRC"] == [int(x) for x in re.search(r"SRC\[76\] = \{([^}]*)\}", committed).group(1).split(",")]
