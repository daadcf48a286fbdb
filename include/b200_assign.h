/*
 * b200_assign.h - C ABI of the B200-native batched traffic-assignment path.
 *
 * The reference (pop-pop-pOp-dev/SAC-GAT-HER_transportationRL) has NO FFI for this path: the
 * boundary is the Python class RepairEnv (reference: src/env/repair_env.py:22-50, imported at
 * src/train.py:23, src/eval.py:15, src/baselines/__init__.py:8).  This header is the thin native
 * layer underneath our drop-in RepairEnv / BatchedRepairEnv; every entry point names the
 * reference method whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - every function returns 0 on success, a negative B200_E* code or a positive cudaError_t;
 *    b200_last_error() returns a static string for the calling thread's last failure;
 *  - all array arguments of the compute calls are DEVICE pointers owned by the caller (torch
 *    tensors in our host code); the library allocates nothing per call (one exception: the general
 *    fused kernel keeps per-CTA workspaces - previous trees, hierarchy weights - inside the handle, grown on
 *    first use); calls are asynchronous on the caller's stream (pass the cudaStream_t as a void*; NULL =
 *    legacy default stream).  A HANDLE SERVES ONE STREAM / THREAD AT A TIME: those workspaces and the
 *    "last launch" fields are per handle, so concurrent b200_assign calls on one handle would share them.
 *    Use one handle (one b200_graph_create) per concurrent stream;
 *  - batch layout: [B][E] / [B][O][N] row-major, one scenario per row ("scenario-major");
 *  - prec: B200_F32 keeps the reference's dtypes (fp32 flow/cost/TSTT sum, fp64 distances),
 *    B200_F64 keeps all link state in double.  `real` below means float or double by prec;
 *  - dmg is uint8 [B][E], 1 = damaged (reference attribute is_damaged, :169,193); the capacity of
 *    an undamaged link is the network's initial capacity; a damaged link costs exactly 1e6
 *    (reference :674-676), so capacity_damage never enters the assignment;
 *  - O = number of origins with demand (ascending node order), N nodes, E links.
 */
#ifndef B200_ASSIGN_H
#define B200_ASSIGN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_graph b200_graph; /* opaque: static network resident in HBM */

enum { B200_F32 = 0, B200_F64 = 1 };
/* assignment_method (reference :55, :318-335); FW_LS = fw with a bisection line search (extension;
 * the first step is 1 like fw's 2/(it+2), later steps solve sum_e (aux_e - f_e) t_e(f + lam (aux - f)) = 0) */
enum { B200_MSA = 0, B200_FW = 1, B200_CFW = 2, B200_FW_LS = 3 };
enum { B200_EINVAL = -1, B200_ENOMEM = -2, B200_EUNSUPPORTED = -3 };
/* bits of the per-scenario status word */
enum { B200_ST_NAN_FLOW = 1, B200_ST_UNASSIGNED = 2, B200_ST_SSSP_CAP = 4 };
#define B200_NO_PRED (-9999) /* scipy's "no predecessor" sentinel (reference :712) */

const char *b200_last_error(void);
int b200_version(void);

/* Static network upload.  Replaces the arrays RepairEnv.__init__ builds (reference :85-97):
 * row/col = edge_index (0-based), t0, cap0 = initial_capacities, OD triples in trip-table order
 * (od_o/od_d 0-based).  All pointers are HOST pointers; parallel links and self loops are
 * rejected (csr_matrix would sum them, reference :488).  device = CUDA ordinal. */
int b200_graph_create(int device, int num_nodes, int num_links, const int32_t *row, const int32_t *col,
                      const float *t0, const float *cap0, int num_od, const int32_t *od_o, const int32_t *od_d,
                      const double *od_q, b200_graph **out);
int b200_graph_destroy(b200_graph *g);
/* info[0..7] = N, E, O, num_od, max_in_degree, fused-kernel dynamic smem bytes (f32), same (f64),
 * SM count */
int b200_graph_info(const b200_graph *g, int64_t info[8]);
/* origins[O] (host) = 0-based origin nodes in the order the [B][O][N] arrays use */
int b200_graph_origins(const b200_graph *g, int32_t *origins_host);

/* S1 - compute_travel_time, numpy branch (reference :667-677).  flow,t: real [B][E]. */
int b200_bpr(const b200_graph *g, int B, int prec, const void *flow, const uint8_t *dmg, void *t, double alpha,
             double beta, void *stream);

/* S2 - scipy.sparse.csgraph.dijkstra(..., return_predecessors=True) as called at reference
 * :485-489, restricted to origins with demand.  t: real [B][E] -> dist f64 [B][O][N],
 * pred int32 [B][O][N] (B200_NO_PRED = none), pedge int32 [B][O][N] (link id of (pred[v],v), -1 = none). */
int b200_sssp(const b200_graph *g, int B, int prec, const void *t, double *dist, int32_t *pred, int32_t *pedge,
              void *stream);

/* S3 - all-or-nothing loading + predecessor walk (reference :490-503, :707-722).
 * pedge int32 [B][O][N] -> aux real [B][E], unassigned f64 [B]. */
int b200_aon(const b200_graph *g, int B, int prec, const int32_t *pedge, void *aux, double *unassigned,
             void *stream);

/* S4 - one loop body of compute_flow_assignment after the AON call (reference :317-343) plus
 * compute_tstt (reference :724-735).  In: flow (current), aux, t = costs of `flow`, iteration
 * index `it`.  Out: flow updated in place, t := costs of the new flow, gap[B] = relative gap of
 * the OLD flow (sum f t - sum aux t)/sum f t, tstt[B] = TSTT of the NEW flow given
 * unassigned[B].  dprev: real [B][E] scratch, only for B200_CFW (direction of the previous
 * iteration; ignored when it == 0).  status[B] is OR-ed with B200_ST_* bits. */
int b200_update_reduce(const b200_graph *g, int B, int prec, int method, int it, void *flow, const void *aux,
                       void *t, const uint8_t *dmg, void *dprev, const double *unassigned, double *tstt,
                       double *gap, int32_t *status, double alpha, double beta, double penalty, int ls_iters,
                       void *stream);

/* Fused path - compute_flow_assignment() for B scenarios (reference :299-345): K iterations of
 * S1..S4 with all per-scenario state on chip.  flow: real [B][E], read as warm start (pass zeros
 * for a cold start, reference :302-303) and overwritten with the result.  Outputs tstt/gap/
 * unassigned f64 [B], status int32 [B] (overwritten).  dprev: real [B][E] scratch for CFW or NULL.
 * active: uint8 [B] or NULL; rows with active[b] == 0 are skipped and keep every output (step()
 * leaves a scenario untouched when the action hits an undamaged link, reference :210-212).
 * K <= 0 returns B200_EINVAL (reference raises ValueError, :300-301). */
int b200_assign(const b200_graph *g, int B, int prec, int K, int method, const uint8_t *dmg, void *flow,
                void *dprev, double *tstt, double *gap, double *unassigned, int32_t *status,
                const uint8_t *active, double alpha, double beta, double penalty, int ls_iters, void *stream);

/* How many kernels the last b200_assign on this handle launched (for bench.py's gpu_launches). */
int b200_assign_launches(const b200_graph *g);
/* Launch configuration of the last b200_assign on this handle: cfg[0] = kernels launched, cfg[1] =
 * thread-block cluster size (1 = one CTA per scenario), cfg[2] = 1 when iterations after the first
 * warm-started their shortest paths from the previous trees, cfg[3] = warps per CTA, cfg[4] = 1 when
 * shortest paths came from the contraction-hierarchy sweeps (exactly summable fp32 costs) rather
 * than from label correcting, 2 when additionally the grouped variant of the network-specialised
 * kernel ran (several scenarios per group of warps, no idle origin lanes), cfg[5..7] reserved (0).  Parity tests use it to prove that the path
 * they mean to cover (cluster mode, warm start, hierarchy) is the one that ran. */
int b200_assign_last_config(const b200_graph *g, int32_t cfg[8]);
/* Host-only self-check of the contraction-hierarchy planner behind the general fused kernel
 * (csrc/b200_cch.cu): plans the hierarchy of the link list, then replays on the CPU exactly the
 * schedule the kernel executes (customization rows, up phase, down-sweep batches) for one cost
 * vector (fp64 [E], link order) and one origin; dist_out fp64 [N] (+inf = unreachable) must equal
 * scipy's Dijkstra distances bit for bit when the costs sum exactly (reference :485-489).
 * stats: optional int64[8] = filled edges, lower triangles, down-region entries, up-region entries,
 * distance slots, down batches, customization rows, 0.  Needs no GPU. */
int b200_cch_host_sssp(int num_nodes, int num_links, const int32_t *row, const int32_t *col, const double *cost,
                       int origin, double *dist_out, int64_t *stats);
/* Which fused kernel b200_assign dispatches to for this network ("assign_static_kernel",
 * "assign_small_kernel" or "assign_fused_kernel"; static string). */
const char *b200_assign_kernel_name(const b200_graph *g);
/* Register a network-specialised kernel compiled on demand for one small topology (a shared object
 * built from csrc/b200_static_plugin.cu; see _native.specialize).  Networks uploaded AFTERWARDS whose
 * link list equals the plugin's topology take its straight-line kernel, like Sioux Falls takes the
 * built-in one.  Not in the reference (it has a single generic Python path). */
int b200_register_static_plugin(const char *path);

/* S5 - get_state features (reference :751-819).  Inputs: flow real [B][E], dmg/goal uint8 [B][E],
 * tstt f64 [B].  Outputs fp32: node_feat [B][N][4], edge_feat [B][E][6]; capacity_damage is the
 * capacity the features show for damaged links (reference :196, :798).  betweenness_raw: optional
 * fp32 [B][N] (NULL to skip) = the normalised betweenness before the division by its maximum
 * (what the reference keeps in betweenness_vec for the undamaged network, :163-165). */
int b200_state_features(const b200_graph *g, int B, int prec, const void *flow, const uint8_t *dmg,
                        const uint8_t *goal, const double *tstt, double capacity_damage, float *node_feat,
                        float *edge_feat, float *betweenness_raw, void *stream);

/* step() glue (reference :207-237, :244-294): apply actions, mark which scenarios need an
 * assignment.  action int32 [B]; dmg updated in place; valid uint8 [B] = 1 where the action hit a
 * damaged link (reference :210-212 returns reward -1 otherwise).  Returns B200_EINVAL via the
 * status word for out-of-range actions (reference raises ValueError, :208-209). */
int b200_apply_actions(const b200_graph *g, int B, const int32_t *action, uint8_t *dmg, uint8_t *valid,
                       int32_t *status, void *stream);
/* reset() damage sampling on the device (reference :167-190): for every row with select[b] != 0
 * (select NULL = all rows) draw `count` distinct damaged links, retrying up to max_tries times until the
 * undamaged links form a strongly connected subgraph on the nodes they touch, then one unconstrained
 * draw (the reference's rule, max_tries = 50).  dmg uint8 [B][E] rows are overwritten; counters uint64 [B]
 * (device, in/out) is the per-scenario position in a counter-based random stream keyed by `seed` - NOT
 * numpy's Generator stream: masks that must equal the reference's draw for a seed are sampled on the host
 * (env/network.py) and passed in.  tries int32 [B] (optional): tries used, max_tries + 1 = unconstrained. */
int b200_sample_damage(const b200_graph *g, int B, int count, int max_tries, uint64_t seed, uint64_t *counters,
                       const uint8_t *select, uint8_t *dmg, int32_t *tries, void *stream);
/* reward_mode: 0 delta, 1 log_delta, 2 neg_tstt, 3 minimize_tstt, 4 rel_improve (reference :256-291) */
int b200_reward_done(const b200_graph *g, int B, int reward_mode, const double *prev_tstt, const double *tstt,
                     const double *initial_tstt, const uint8_t *dmg, const uint8_t *goal, const uint8_t *valid,
                     double r_alpha, double r_beta, double r_gamma, double r_clip, double *reward, uint8_t *done,
                     void *stream);

/* G1 - GAT edge-softmax + neighbour aggregation over incoming links and the self loop
 * (torch_geometric GATConv message/aggregate as used at reference src/models/gat_encoder.py:23-27;
 * third-party, unpinned).  Fixed topology shared by all B graphs.
 *  xw       fp32 [B][N][H][C]  projected node features (lin(x))
 *  a_src/a_dst fp32 [B][N][H]  per-node attention terms
 *  a_edge   fp32 [B][E+N][H]   per-link attention terms (rows E..E+N-1 = self loops)
 *  out      fp32 [B][N][H][C]  sum_e alpha_e * xw[src(e)]   (bias/concat/mean stay in torch)
 *  alpha    fp32 [B][E+N][H]   attention coefficients (saved for backward / return_attention) */
int b200_gat_edge_softmax_aggregate_fwd(const b200_graph *g, int B, int H, int C, const float *xw,
                                        const float *a_src, const float *a_dst, const float *a_edge,
                                        float negative_slope, float *out, float *alpha, void *stream);
int b200_gat_edge_softmax_aggregate_bwd(const b200_graph *g, int B, int H, int C, const float *xw,
                                        const float *a_src, const float *a_dst, const float *a_edge,
                                        const float *alpha, const float *grad_out, float negative_slope,
                                        float *grad_xw, float *grad_a_src, float *grad_a_dst, float *grad_a_edge,
                                        void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_ASSIGN_H */
