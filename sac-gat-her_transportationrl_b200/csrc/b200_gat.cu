// placeholder replaced below
#include "b200_common.cuh"
extern "C" int b200_gat_edge_softmax_aggregate_fwd(const b200_graph *, int, int, int, const float *, const float *,
                                                   const float *, const float *, float, float *, float *, void *)
{ b200_set_error("gat kernel not built yet"); return B200_EUNSUPPORTED; }
extern "C" int b200_gat_edge_softmax_aggregate_bwd(const b200_graph *, int, int, int, const float *, const float *,
                                                   const float *, const float *, const float *, const float *, float,
                                                   float *, float *, float *, float *, void *)
{ b200_set_error("gat kernel not built yet"); return B200_EUNSUPPORTED; }
