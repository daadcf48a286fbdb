// Internals shared by the assignment translation units (b200_assign.cu: stage kernels, general
// fused kernel, C ABI; b200_assign_small.cu: generic lane-per-origin kernel; b200_assign_static.cu:
// network-specialised kernel).  Everything here is inline / template code.
#pragma once
#include "b200_common.cuh"

#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include <vector>

namespace b200i {


constexpr int kMaxWarps = 16;
constexpr int kMinWarps = 2;
constexpr size_t kSmemLimit = 227 * 1024;

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// Per-warp scratch of the general kernel (one origin at a time):
//   dist f64 [N] (aliased by the subtree sums once the tree is chosen) | pk u16 [N] tree in-link slot
//   | q0, q1 u16 [N] frontier queues | cnt u16 [N] pending children (updated with 32-bit atomics on
//   the containing word) | mask u32 [ceil(N/32)] "already queued" bits | tail counter
struct WarpOffsets {
    size_t pk, q0, q1, cnt, mask, tail, total;
};
// alias_cnt: the subtree sums are 4-byte (int32 or float), so they fill only the lower half of the
// dead dist[] array and the pending-children counters go into the upper half
// dist_slots: the contraction-hierarchy sweeps keep N labels, the +inf slot and the partial slots of split
// destinations in the distance array (0 = N, the label-correcting kernel)
// lean: 0 = all scratch arrays, 1 = hierarchy sweeps + packed subtree words (one queue), 2 = ... with that queue inside dist[]
__host__ __device__ inline WarpOffsets warp_offsets(int N, bool alias_cnt, int dist_slots = 0, int lean = 0)
{
    WarpOffsets W;
    const size_t dist_bytes = align16((size_t)(dist_slots > N ? dist_slots : N) * 8),
                 cnt_bytes = align16((size_t)((N + 1) / 2) * 4);
    size_t o = dist_bytes;
    W.pk = o;   o += align16((size_t)N * 2);
    if (lean) {
        // hierarchy sweeps + packed subtree words (warp_aon_chain): one queue (the leaves) is all the origin needs.
        // lean == 2: it lives in the distance array too - the labels are dead once the predecessors are chosen, the
        // 32-bit subtree words take bytes [0, 4N) of it and the 16-bit leaf list the bytes after them (4N + 2N <= 8N).
        // Taken when it buys warps (Chicago-Sketch-size 14 -> 16, +8 %); measured 1.5-3 % slower at equal warp count.
        if (lean == 2) W.q0 = align16((size_t)N * 4);
        else { W.q0 = o; o += align16((size_t)N * 2); }
        W.q1 = W.q0; W.cnt = W.q0; W.mask = W.q0; W.tail = W.q0;
        W.total = o;
        return W;
    }
    W.q0 = o;   o += align16((size_t)N * 2);
    W.q1 = o;   o += align16((size_t)N * 2);
    if (alias_cnt && align16((size_t)N * 4) + cnt_bytes <= dist_bytes) W.cnt = align16((size_t)N * 4);
    else { W.cnt = o; o += cnt_bytes; }
    W.mask = o; o += align16((size_t)((N + 31) / 32) * 4);
    W.tail = o; o += 16;
    W.total = o;
    return W;
}

struct SmemLayout {
    size_t off_bar, off_topo, off_w, off_aux, off_flow, off_red, off_cchw, off_warp, warp_stride, total;
};

// cch_w_bytes: contraction-hierarchy weights kept in shared memory (0 = none / label-correcting kernel)
__host__ __device__ inline SmemLayout make_layout(int N, int E, int topo_bytes, int warps, int real_bytes, bool alias_cnt,
                                                  int dist_slots = 0, size_t cch_w_bytes = 0, int lean = 0)
{
    SmemLayout L;
    size_t o = 0;
    L.off_bar = o;  o += 16;
    L.off_topo = o; o += align16((size_t)topo_bytes);
    L.off_w = o;    o += align16((size_t)E * 8);
    L.off_aux = o;  o += align16((size_t)E * real_bytes);
    L.off_flow = o; o += align16((size_t)E * real_bytes);
    L.off_red = o;  o += align16((size_t)(kMaxWarps * 4 + 8) * 8);
    L.off_cchw = o; o += align16(cch_w_bytes);
    L.off_warp = o;
    L.warp_stride = warp_offsets(N, alias_cnt, dist_slots, lean).total;
    L.total = o + L.warp_stride * warps;
    return L;
}

// shared-memory layout of the small-network kernel, computed on the host (b200_assign_small.cuh)
struct SmallOffsets {
    int blob_bytes, off_rec2, off_q, off_dist, off_sacc, off_cnt, off_stk, off_pu, off_pe, off_w, off_aux, off_flow,
        warp_stride;
};

// contraction-hierarchy tables of the general fused kernel (b200_cch.cu), device pointers
struct CchParams {
    const uint32_t *dn_rows, *up_rows, *sel_rows, *up_off, *cz_init, *cz_desc, *cz_tab, *cz_lvl;
    double *ws;            // per-CTA weight workspace in HBM (the part of W outside shared memory)
    int F, wd, wu, slots, dn_batches, D, SD, cz_rows, cz_levels;
    int topo_bytes;        // staged part of the 16-bit topology blob: in_ptr | in_src | in_eid
    int wmode;             // 0: W in HBM/L2, 1: down region in shared memory, 2: all of W in shared memory
    int leaf_alias;        // 1: the leaf list of the tree peel lives inside dist[] (warp_offsets, lean == 2)
    int ws_stride;         // doubles per CTA in ws
    unsigned exact_hi;     // high word of the exactness bound (labels >= it trigger the Bellman-Ford fallback)
};

struct Params {
    SmallOffsets so;
    CchParams cch;
    // static network
    int N, E, O, topo_ints, off_in_src, off_in_eid;
    // general kernels: 16-bit topology blob, offsets in uint16 elements
    const uint16_t *topo16;
    int topo16_bytes, o16_in_src, o16_in_eid, o16_out_ptr, o16_out_dst;
    int alias32, alias64;  // general kernel: counters aliased into dist[] (per precision; see warp_offsets)
    int warm;  // general kernel: warm-start shortest paths from the previous iteration's trees
    int intq;  // general kernel: all demands are integers (total < 2^31) -> exact int32 link loads
    int kahn_pack;  // general kernel: integer demands < 2^28 per origin and out-degree <= 15 -> packed subtree words
    const int32_t *topo, *row, *origins, *od_ptr, *od_dst;
    const float *t0_csc, *cap0_csc, *t0, *cap0;
    const void *od_q;  // REAL
    // small-network kernel only: TMA blob (node records + dense demand), per-slot link id / cost index
    const unsigned char *small_blob;
    const uint8_t *small_eid, *small_wpos;
    // network-specialised kernel only: TMA blob (dense demand + per-slot tables), origin positions
    const unsigned char *st_blob;
    const int32_t *st_origpos;
    double total_demand;
    // batch
    int B, K, method, it0, ls_iters;
    const uint8_t *dmg, *active;
    void *flow, *dprev, *t, *aux;          // real [B][E]
    double *tstt, *gap, *unassigned;       // [B]
    const double *unassigned_in;
    int32_t *status;
    double *dist_out;                      // [B][O][N]
    int32_t *pred_out, *pedge_out;
    uint16_t *tree_ws;                     // general fused kernel: [gridDim.x][O][N] previous trees
    const int32_t *pedge_in;
    double alpha, beta, penalty;
};

template <typename REAL> __device__ __forceinline__ REAL rmul(REAL a, REAL b);
template <> __device__ __forceinline__ float rmul<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double rmul<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename REAL> __device__ __forceinline__ REAL radd(REAL a, REAL b);
template <> __device__ __forceinline__ float radd<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double radd<double>(double a, double b) { return __dadd_rn(a, b); }

template <typename REAL>
__device__ __forceinline__ REAL link_cost(const Params &p, int k, REAL f, bool damaged)
{
    if (damaged) return (REAL)1e6;
    return Bpr<REAL>::eval(f, (REAL)__ldg(p.cap0_csc + k), (REAL)__ldg(p.t0_csc + k), (REAL)p.alpha, (REAL)p.beta,
                           p.beta == 4.0);
}

// numpy's pairwise summation order (np.sum of a contiguous fp array, reference :727), evaluated
// cooperatively by one warp: lanes j, j+8, .. replicate the 8 running sums of a <=128 block.
template <typename REAL>
__device__ REAL warp_pairwise(const REAL *a, int n)
{
    const int lane = threadIdx.x & 31;
    if (n < 8) {
        REAL r = 0;
        for (int i = 0; i < n; ++i) r = radd<REAL>(r, a[i]);
        return r;
    }
    if (n <= 128) {
        const int j = lane & 7, lim = n - (n % 8);
        REAL r = a[j];
        for (int i = 8; i < lim; i += 8) r = radd<REAL>(r, a[i + j]);
        r = radd<REAL>(r, __shfl_xor_sync(0xffffffffu, r, 1));
        r = radd<REAL>(r, __shfl_xor_sync(0xffffffffu, r, 2));
        r = radd<REAL>(r, __shfl_xor_sync(0xffffffffu, r, 4));
        for (int i = lim; i < n; ++i) r = radd<REAL>(r, a[i]);
        return r;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    const REAL lo = warp_pairwise<REAL>(a, n2);
    const REAL hi = warp_pairwise<REAL>(a + n2, n - n2);
    return radd<REAL>(lo, hi);
}


}  // namespace b200i

// On-demand static-kernel plugins (b200_static_plugin.cu).  The ABI stamp changes whenever Params or
// b200_graph change layout, so a stale cached plugin is rejected instead of misreading the structs.
#define B200_PLUGIN_ABI (B200_VERSION * 1000 + (int)(sizeof(b200i::Params) % 997))
typedef int (*b200_plugin_prepare_fn)(b200_graph *, const int32_t *, const int32_t *, const float *, const float *, int,
                                      const int32_t *, const int32_t *, const double *, const int32_t *);
typedef int (*b200_plugin_launch_fn)(const b200_graph *, const b200i::Params *, int, cudaStream_t);

// launchers implemented in the other translation units
int b200_launch_small(const b200_graph *g, const b200i::Params &p, int prec, cudaStream_t st);
int b200_launch_static(const b200_graph *g, const b200i::Params &p, int prec, cudaStream_t st);
