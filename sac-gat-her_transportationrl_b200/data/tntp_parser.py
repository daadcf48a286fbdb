"""TNTP network / trip-table reader and writer.

Boundary input of the assignment path.  The two dataclasses keep the
field names of the reference's data model (reference: src/data/tntp_parser.py:10-25) because
``RepairEnv(graph_data, ...)`` consumes them by attribute; the reader keeps the reference's
observable semantics (reference: src/data/tntp_parser.py:33-99):

* the node count comes from the ``NUMBER OF NODES`` metadata line, ``FIRST THRU NODE`` is ignored;
* link rows start after the first non-comment line that mentions ``init node`` / ``init_node``;
  rows with fewer than six fields are skipped; ``;`` is a separator, not data;
* per-link ``b`` / ``power`` are parsed but the env uses the global ``bpr_alpha`` / ``bpr_beta``;
* OD entries with demand <= 0 are dropped; node ids stay 1-based.

The writer is new (the reference has none); it is used to ship networks inside this repository
and to hand synthetic networks to any TNTP consumer.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple


@dataclass
class EdgeData:
    u: int
    v: int
    capacity: float
    t0: float
    length: float
    b: float
    power: float


@dataclass
class GraphData:
    num_nodes: int
    edges: List[EdgeData]
    od_demand: Dict[Tuple[int, int], float]




def _significant_lines(path: str):
    with open(path, "r", encoding="utf-8", errors="ignore") as fh:
        for raw in fh:
            line = raw.strip()
            if line and not line.startswith("~"):
                yield line


def parse_net_tntp(path: str) -> Tuple[int, List[EdgeData]]:
    num_nodes = 0
    edges: List[EdgeData] = []
    in_table = False
    for line in _significant_lines(path):
        low = line.lower()
        if "number of nodes" in low:
            tail = line.split(">")[-1] if ">" in line else line.split()[-1]
            num_nodes = int(tail.strip())
        if "init_node" in low or "init node" in low:
            in_table = True
            continue
        if not in_table:
            continue
        fields = line.replace(";", " ").split()
        if len(fields) < 6:
            continue
        edges.append(
            EdgeData(
                u=int(fields[0]),
                v=int(fields[1]),
                capacity=float(fields[2]),
                length=float(fields[3]),
                t0=float(fields[4]),
                b=float(fields[5]),
                power=float(fields[6]) if len(fields) > 6 else 4.0,
            )
        )
    return num_nodes, edges


def parse_trips_tntp(path: str) -> Dict[Tuple[int, int], float]:
    demand: Dict[Tuple[int, int], float] = {}
    origin = None
    for line in _significant_lines(path):
        if line.lower().startswith("origin"):
            origin = int(line.split()[1])
            continue
        if origin is None:
            continue
        for chunk in line.split(";"):
            if ":" not in chunk:
                continue
            dest_s, val_s = chunk.split(":")
            val = float(val_s.strip())
            if val > 0:
                demand[(origin, int(dest_s.strip()))] = val
    return demand


def load_graph_data(net_path: str, trips_path: str) -> GraphData:
    num_nodes, edges = parse_net_tntp(net_path)
    return GraphData(num_nodes=num_nodes, edges=edges, od_demand=parse_trips_tntp(trips_path))


# --------------------------------------------------------------------------------------------
# writer (new)
# --------------------------------------------------------------------------------------------

def _fmt(x: float) -> str:
    return repr(float(x)) if float(x) != int(x) else str(int(x))


def write_net_tntp(path: str, graph: GraphData, num_zones: int | None = None) -> None:
    zones = num_zones if num_zones is not None else len({o for (o, _) in graph.od_demand})
    with open(path, "w", encoding="utf-8") as fh:
        fh.write(f"<NUMBER OF ZONES> {zones}\n<NUMBER OF NODES> {graph.num_nodes}\n")
        fh.write(f"<FIRST THRU NODE> 1\n<NUMBER OF LINKS> {len(graph.edges)}\n")
        # TNTP readers (the reference's included) find the table through a non-comment line
        # that mentions "init node"; real TNTP files carry it in this metadata tag.
        fh.write("<ORIGINAL HEADER>~ Init node Term node Capacity Length Free Flow Time B Power Speed Toll Type ;\n")
        fh.write("<END OF METADATA>\n\n")
        fh.write("~ init_node term_node capacity length free_flow_time b power speed toll link_type ;\n")
        for e in graph.edges:
            fh.write(
                f"{e.u} {e.v} {repr(float(e.capacity))} {_fmt(e.length)} {_fmt(e.t0)} "
                f"{_fmt(e.b)} {_fmt(e.power)} 0 0 1 ;\n"
            )


def write_trips_tntp(path: str, graph: GraphData) -> None:
    by_origin: Dict[int, List[Tuple[int, float]]] = {}
    for (o, d), q in graph.od_demand.items():
        by_origin.setdefault(o, []).append((d, q))
    total = float(sum(graph.od_demand.values()))
    with open(path, "w", encoding="utf-8") as fh:
        fh.write(f"<NUMBER OF ZONES> {len(by_origin)}\n<TOTAL OD FLOW> {total}\n<END OF METADATA>\n")
        for o, row in by_origin.items():
            fh.write(f"\nOrigin {o}\n")
            for i in range(0, len(row), 5):
                fh.write(" ".join(f"{d} : {repr(float(q))};" for d, q in row[i : i + 5]) + "\n")
