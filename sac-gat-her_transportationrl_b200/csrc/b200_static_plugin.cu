// On-demand instance of the network-specialised kernel for ONE topology header
// (tools/gen_static_topology.py / codegen.emit).  Compiled by _native.specialize with
//   -DB200_TOPO_HEADER=\"<file>\" -DB200_TOPO_NAME=<struct>
// into its own shared object and registered with b200_register_static_plugin.
#include "b200_assign_internal.cuh"
#include B200_TOPO_HEADER

using namespace b200i;

namespace {
#include "b200_assign_static.cuh"
}  // namespace

// the plugin does not link the main library: errors are reported through the return code only
void b200_set_error(const char *, ...) {}

extern "C" int b200_plugin_abi(void) { return B200_PLUGIN_ABI; }

extern "C" int b200_plugin_prepare(b200_graph *g, const int32_t *row, const int32_t *col, const float *t0,
                                   const float *cap0, int num_od, const int32_t *od_o, const int32_t *od_d,
                                   const double *od_q, const int32_t *origin_slot)
{
    return static_try_prepare<B200_TOPO_NAME>(g, row, col, t0, cap0, num_od, od_o, od_d, od_q, origin_slot);
}

extern "C" int b200_plugin_launch(const b200_graph *g, const Params *p, int prec, cudaStream_t st)
{
    return launch_static<B200_TOPO_NAME>(g, *p, prec, st);
}
