// Network-specialised kernel: translation unit (kernel template in b200_assign_static.cuh,
// topology data in b200_topo_*.h).
#include "b200_assign_internal.cuh"

using namespace b200i;

namespace {
#include "b200_assign_static.cuh"
}  // namespace

int b200_launch_static(const b200_graph *g, const Params &p, int prec, cudaStream_t st)
{
    return prec == B200_F32 ? launch_static<float>(g, p, st) : launch_static<double>(g, p, st);
}

// Network-specialised kernel: eligibility (link list equals the compiled-in topology) + tables.
int b200_static_prepare(b200_graph *g, const int32_t *row, const int32_t *col, const float *t0, const float *cap0,
                        int num_od, const int32_t *od_o, const int32_t *od_d, const double *od_q,
                        const int32_t *origin_slot)
{
    g->static_ok = 0;
    if (!static_topology_matches<TopoSiouxFalls>(g->N, g->E, row, col)) return 0;
    bool intq = g->total_demand < 16777216.0;  // 2^24: fp32 link sums stay exact
    for (int k = 0; k < num_od && intq; ++k) intq = od_q[k] >= 0.0 && od_q[k] == floor(od_q[k]);
    g->st_intq = intq ? 1 : 0;
    g->static_ok = static_prepare<TopoSiouxFalls>(g, t0, cap0, num_od, od_o, od_d, od_q, origin_slot, intq);
    return g->static_ok;
}

