// Batched traffic assignment on B200 (sm_100a): stage kernels S1..S4 and the fused persistent
// kernel.  One CTA owns one damage scenario at a time; its warps own origins.
//
// Reference path being replaced (per scenario, on one CPU thread):
//   compute_flow_assignment   src/env/repair_env.py:299-345
//   compute_travel_time       src/env/repair_env.py:667-677   (S1)
//   _all_or_nothing           src/env/repair_env.py:481-503   (S2 scipy dijkstra + S3 loading)
//   _path_edges_from_pred.    src/env/repair_env.py:707-722   (S3)
//   compute_tstt              src/env/repair_env.py:724-735   (S4)
//
// Shared-memory layout of a CTA (all offsets 16-byte aligned):
//   mbarrier | topology blob (in_ptr, in_src, in_eid; staged once per CTA by a TMA bulk copy)
//   | w[E] double link costs in CSC-slot order | aux[E] real | flow[E] real | red[] scratch
//   | per warp: dist[N] double, pk[N] int32 (CSC slot of the tree in-link, -1 = none)
//
// Numerics: distances are fp64 sums of the (fp32- or fp64-valued) link costs, relaxed to the
// unique fixed point dist[v] = min_u fl(dist[u] + w(u,v)) - the same values Dijkstra produces for
// positive costs, so predecessor selection "tight in-link with the smallest dist[u], then the
// largest u" reproduces scipy's trees bit for bit.
#include "b200_assign_internal.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

using namespace b200i;

namespace {

// Everything a CTA needs, carved out of dynamic shared memory.
template <typename REAL>
struct Cta {
    uint64_t *bar;
    const uint16_t *in_ptr, *in_src, *in_eid, *out_ptr, *out_dst;   // 16-bit topology blob
    uint16_t *topo;
    double *w;
    REAL *aux, *flow;
    double *red;
    unsigned char *warp_base;
    size_t warp_stride;
    int N, E;
    bool intq;   // link loads are accumulated as int32 in the aux array (exact, order-independent)
    bool alias_cnt;   // pending-children counters live in the upper half of dist[] (see warp_offsets)
    // contraction-hierarchy weights (b200_cch.cu): down region / up region, each either in shared memory
    // or in this CTA's slice of the HBM workspace; position pos of the logical array is wd[pos] below
    // cch_wd and wu[pos - cch_wd] from there on (the +inf and dummy slots follow the up region)
    double *wd, *wu;
    int cch_wd;
    int lean;    // per-warp scratch without the second queue / counters (warp_offsets)
    int topo_bytes;   // staged part of the topology blob

    // link load of link e as REAL, whatever the accumulation type
    __device__ __forceinline__ REAL auxv(int e) const { return intq ? (REAL)((const int *)aux)[e] : aux[e]; }

    __device__ void carve(unsigned char *smem, const Params &p, int warps, bool cch = false)
    {
        alias_cnt = (sizeof(REAL) == 4 ? p.alias32 : p.alias64) != 0;
        const int wmode = cch ? p.cch.wmode : 0;
        const size_t w_smem = !cch ? 0 : (wmode == 2 ? (size_t)(p.cch.wd + p.cch.wu + 2) * 8 : wmode == 1 ? (size_t)p.cch.wd * 8 : 0);
        topo_bytes = cch ? p.cch.topo_bytes : p.topo16_bytes;
        SmemLayout L = make_layout(p.N, p.E, topo_bytes, warps, (int)sizeof(REAL), alias_cnt, cch ? p.cch.slots : 0,
                                   w_smem, (cch && p.kahn_pack) ? 1 + p.cch.leaf_alias : 0);
        lean = (cch && p.kahn_pack) ? 1 + p.cch.leaf_alias : 0;
        if (cch) {
            double *ws_cta = p.cch.ws + (size_t)blockIdx.x * p.cch.ws_stride;
            double *sm = (double *)(smem + L.off_cchw);
            cch_wd = p.cch.wd;
            wd = wmode >= 1 ? sm : ws_cta;
            wu = wmode == 2 ? sm + p.cch.wd : (wmode == 1 ? ws_cta : ws_cta + p.cch.wd);
        }
        bar = (uint64_t *)(smem + L.off_bar);
        topo = (uint16_t *)(smem + L.off_topo);
        in_ptr = topo;
        in_src = topo + p.o16_in_src;
        in_eid = topo + p.o16_in_eid;
        out_ptr = topo + p.o16_out_ptr;
        out_dst = topo + p.o16_out_dst;
        intq = p.intq != 0;
        w = (double *)(smem + L.off_w);
        aux = (REAL *)(smem + L.off_aux);
        flow = (REAL *)(smem + L.off_flow);
        red = (double *)(smem + L.off_red);
        warp_base = smem + L.off_warp;
        warp_stride = L.warp_stride;
        N = p.N;
        E = p.E;
    }
    __device__ double *dist(int warp) const { return (double *)(warp_base + warp_stride * warp); }
    __device__ __forceinline__ double *wpos(int pos) const { return pos < cch_wd ? wd + pos : wu + (pos - cch_wd); }

    // Stage the static topology into shared memory with one TMA bulk copy (UBLKCP).
    __device__ void stage_topology(const Params &p)
    {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)topo_bytes;
            mbar_expect_tx(bar, bytes);
            tma_bulk_g2s(topo, p.topo16, bytes, bar);
        }
        mbar_wait(bar, 0);
    }

    // deterministic block-wide fp64 sum; every thread gets the result
    __device__ double block_sum(double v)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
        v = warp_sum_f64(v);
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double s = 0.0;
        for (int i = 0; i < W; ++i) s += red[i];
        return s;
    }
};

// ---------------------------------------------------------------------------------------------
// S1: link costs of the CTA's flow vector into w[] (CSC-slot order, as double)
// ---------------------------------------------------------------------------------------------
template <typename REAL>
__device__ void cta_costs_from_flow(Cta<REAL> &c, const Params &p, const uint8_t *dmg_b)
{
    for (int k = threadIdx.x; k < c.E; k += blockDim.x) {
        const int e = c.in_eid[k];
        c.w[k] = (double)link_cost<REAL>(p, k, c.flow[e], dmg_b[e] != 0);
    }
}

// ---------------------------------------------------------------------------------------------
// S2: one warp, one origin - frontier-driven label correcting + predecessor selection
// ---------------------------------------------------------------------------------------------
constexpr int kNoPk = 0xFFFF;

struct WarpScratch {
    double *dist;
    uint16_t *pk, *q0, *q1;
    uint32_t *cntw, *mask;
    int *tail;
};
template <typename REAL>
__device__ __forceinline__ WarpScratch warp_scratch(const Cta<REAL> &c, int warp, int dist_slots = 0)
{
    const WarpOffsets W = warp_offsets(c.N, c.alias_cnt, dist_slots, c.lean);
    unsigned char *base = c.warp_base + c.warp_stride * warp;
    WarpScratch s;
    s.dist = (double *)base;
    s.pk = (uint16_t *)(base + W.pk);
    s.q0 = (uint16_t *)(base + W.q0);
    s.q1 = (uint16_t *)(base + W.q1);
    s.cntw = (uint32_t *)(base + W.cnt);
    s.mask = (uint32_t *)(base + W.mask);
    s.tail = (int *)(base + W.tail);
    return s;
}

// Distances are relaxed to the least fixed point of dist[v] = min_u fl(dist[u] + w(u,v)) (see the
// file header).  Work is driven by a frontier: a node is re-examined (pull over its in-links) only
// after one of its in-neighbours improved; the "queued" bit of a node is cleared before its
// in-neighbours are read, so an improvement that lands after the read re-queues the node for the
// next round.  Rounds <= longest shortest path in hops.  One warp owns one origin.
//
// Warm start (tree != nullptr: this origin's tree of the previous iteration, in-link slot per node):
// the old tree is re-evaluated top-down under the new costs - every label is then the length of a
// real path, i.e. a valid upper bound - and ALL nodes are queued once, so the first round is a
// full verification pass and later rounds only follow actual changes.  The fixed point, and with
// it the result, is the same as from a cold start.
// Queue tails live in a warp-uniform register: new entries are placed by ballot + popcount, the
// only shared-memory atomics are the "already queued" bits.  Link loops are unrolled by kLinkUnroll
// so the dependent shared-memory loads of one step are issued together.
template <typename REAL>
__device__ void warp_select_pred(const Cta<REAL> &c, int origin, const WarpScratch &s);

template <typename REAL, int kLinkUnroll = 4>
__device__ int warp_sssp(const Cta<REAL> &c, int origin, const WarpScratch &s, const uint16_t *tree = nullptr)
{
    const int lane = threadIdx.x & 31, N = c.N;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;   // lower lanes
    double *dist = s.dist;
    const uint16_t *__restrict__ in_ptr = c.in_ptr, *__restrict__ in_src = c.in_src, *__restrict__ out_ptr = c.out_ptr,
                   *__restrict__ out_dst = c.out_dst;
    const double *__restrict__ w = c.w;
    for (int v = lane; v < N; v += 32) dist[v] = (v == origin) ? 0.0 : B200_INF_F64;
    uint16_t *cur = s.q0, *nxt = s.q1;
    int count = 0;
    if (tree) {   // warp-uniform (the iteration index decides)
        for (int v = lane; v < N; v += 32) s.pk[v] = tree[v];
        if (lane == 0) cur[0] = (uint16_t)origin;
        count = 1;
        __syncwarp();
        while (count > 0) {   // breadth-first over the old tree's children
            int ntail = 0;
            for (int base = 0; base < count; base += 32) {
                const bool valid = base + lane < count;
                const int u = valid ? (int)cur[base + lane] : 0;
                const double du = valid ? dist[u] : 0.0;
                const int j0 = valid ? out_ptr[u] : 0, j1 = valid ? out_ptr[u + 1] : 0;
                for (int jb = 0; __any_sync(FULL, j0 + jb < j1); jb += kLinkUnroll) {
                    int xs[kLinkUnroll], ks[kLinkUnroll];
#pragma unroll
                    for (int i = 0; i < kLinkUnroll; ++i) xs[i] = j0 + jb + i < j1 ? out_dst[j0 + jb + i] : -1;
#pragma unroll
                    for (int i = 0; i < kLinkUnroll; ++i) ks[i] = xs[i] >= 0 ? (int)s.pk[xs[i]] : kNoPk;
#pragma unroll
                    for (int i = 0; i < kLinkUnroll; ++i) {
                        const bool child = ks[i] != kNoPk && in_src[ks[i]] == u;
                        if (child) dist[xs[i]] = du + w[ks[i]];
                        const unsigned m = __ballot_sync(FULL, child);
                        if (child) nxt[ntail + __popc(m & lt)] = (uint16_t)xs[i];
                        ntail += __popc(m);
                    }
                }
            }
            __syncwarp();
            count = ntail;
            uint16_t *t = cur; cur = nxt; nxt = t;
        }
        // queue every node but the origin
        for (int v = lane; v < N; v += 32)
            if (v != origin) cur[v - (v > origin ? 1 : 0)] = (uint16_t)v;
        for (int i = lane; i < (N + 31) / 32; i += 32) s.mask[i] = 0xffffffffu;
        count = N - 1;
        __syncwarp();
        if (lane == 0) atomicAnd(&s.mask[origin >> 5], ~(1u << (origin & 31)));
    } else {
        for (int i = lane; i < (N + 31) / 32; i += 32) s.mask[i] = 0u;
        __syncwarp();
        const int o0 = out_ptr[origin];
        count = out_ptr[origin + 1] - o0;
        for (int j = lane; j < count; j += 32) {
            const int x = out_dst[o0 + j];
            cur[j] = (uint16_t)x;
            atomicOr(&s.mask[x >> 5], 1u << (x & 31));
        }
    }
    __syncwarp();
    int rounds = 0, capped = 0;
    while (count > 0) {
        int ntail = 0;
        for (int base = 0; base < count; base += 32) {
            const bool valid = base + lane < count;
            const int v = valid ? (int)cur[base + lane] : 0;
            if (valid) atomicAnd(&s.mask[v >> 5], ~(1u << (v & 31)));
            __syncwarp();
            bool improved = false;
            if (valid) {
                const double dv = dist[v];
                double best = dv;
                const int k0 = in_ptr[v], k1 = in_ptr[v + 1];
                for (int kb = k0; kb < k1; kb += kLinkUnroll) {
                    int us[kLinkUnroll];
                    double ws[kLinkUnroll];
#pragma unroll
                    for (int i = 0; i < kLinkUnroll; ++i) {
                        const bool ok = kb + i < k1;
                        us[i] = ok ? in_src[kb + i] : v;
                        ws[i] = ok ? w[kb + i] : B200_INF_F64;
                    }
#pragma unroll
                    for (int i = 0; i < kLinkUnroll; ++i) {
                        const double cand = dist[us[i]] + ws[i];
                        best = cand < best ? cand : best;
                    }
                }
                improved = best < dv;
                if (improved) dist[v] = best;
            }
            const int j0 = improved ? out_ptr[v] : 0, j1 = improved ? out_ptr[v + 1] : 0;
            for (int jb = 0; __any_sync(FULL, j0 + jb < j1); jb += kLinkUnroll) {
                int xs[kLinkUnroll];
                bool fresh[kLinkUnroll];
#pragma unroll
                for (int i = 0; i < kLinkUnroll; ++i) xs[i] = j0 + jb + i < j1 ? out_dst[j0 + jb + i] : -1;
#pragma unroll
                for (int i = 0; i < kLinkUnroll; ++i) {
                    fresh[i] = false;
                    if (xs[i] >= 0 && xs[i] != origin) {
                        const uint32_t bit = 1u << (xs[i] & 31);
                        fresh[i] = !(atomicOr(&s.mask[xs[i] >> 5], bit) & bit);
                    }
                }
#pragma unroll
                for (int i = 0; i < kLinkUnroll; ++i) {
                    const unsigned m = __ballot_sync(FULL, fresh[i]);
                    if (fresh[i]) nxt[ntail + __popc(m & lt)] = (uint16_t)xs[i];
                    ntail += __popc(m);
                }
            }
            __syncwarp();
        }
        count = ntail;
        uint16_t *t = cur; cur = nxt; nxt = t;
        if (++rounds > N + 1) { capped = count > 0; break; }
    }
    warp_select_pred<REAL>(c, origin, s);
    return capped;
}

// predecessor: tight in-link with the smallest dist[u], then the largest u (scipy's rule, SURVEY.md
// Appendix A.1; in-links of a node are stored by ascending tail id)
template <typename REAL>
__device__ void warp_select_pred(const Cta<REAL> &c, int origin, const WarpScratch &s)
{
    const int lane = threadIdx.x & 31, N = c.N;
    const double *dist = s.dist;
    const uint16_t *__restrict__ in_ptr = c.in_ptr, *__restrict__ in_src = c.in_src;
    const double *__restrict__ w = c.w;
    for (int v = lane; v < N; v += 32) {
        const double dv = dist[v];
        int bk = kNoPk;
        double bdu = B200_INF_F64;
        if (v != origin && dv < B200_INF_F64) {
            // in-links come by ascending tail id, so "<=" on dist[u] leaves the largest u among equals
            const int k1 = in_ptr[v + 1];
            int k = in_ptr[v];
            for (; k + 2 <= k1; k += 2) {
                const double du0 = dist[in_src[k]], du1 = dist[in_src[k + 1]];
                const double c0 = du0 + w[k], c1 = du1 + w[k + 1];
                if (c0 == dv && du0 <= bdu) { bk = k; bdu = du0; }
                if (c1 == dv && du1 <= bdu) { bk = k + 1; bdu = du1; }
            }
            if (k < k1) {
                const double du0 = dist[in_src[k]];
                if (du0 + w[k] == dv && du0 <= bdu) bk = k;
            }
        }
        s.pk[v] = (uint16_t)bk;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// S2 through the customizable contraction hierarchy (tables: b200_cch.cu)
// ---------------------------------------------------------------------------------------------
// Customization, once per (scenario, iteration), all warps of the CTA: every filled edge starts from
// its own link cost (+inf for a pure shortcut), then the rows of each level - 32 edges with their
// lower triangles, most triangles first - are dealt to the warps round-robin; a CTA barrier separates
// the levels (the two lower edges of a triangle are final one level earlier).  A warp goes straight to
// its rows (per-level row ranges, per-row table offsets) and fetches four triangle words ahead of the
// shared-memory reads they address: the table (Chicago-Sketch-size: 200 KB) lives in L2.
template <typename REAL, int G>   // G: triangle words in flight per lane (8 when registers allow, else 4)
__device__ void cta_customize(const Cta<REAL> &c, const Params &p)
{
    const CchParams &h = p.cch;
    for (int e = threadIdx.x; e < h.F; e += blockDim.x) {
        const uint32_t wi = __ldg(h.cz_init + e);
        const int slot = (int)(wi >> 16);
        *c.wpos((int)(wi & 0xffffu)) = slot == 0xffff ? B200_INF_F64 : c.w[slot];
    }
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto relax = [&](double acc, uint32_t tri) {
        const double cand = *c.wpos((int)(tri & 0xffffu)) + *c.wpos((int)(tri >> 16));
        return cand < acc ? cand : acc;
    };
    const uint32_t inf_pair = (uint32_t)(h.wd + h.wu) * 0x10001u;   // both halves point at the +inf slot
    for (int L = 0; L < h.cz_levels; ++L) {
        __syncthreads();
        const int r0 = (int)__ldg(h.cz_lvl + L), r1 = (int)__ldg(h.cz_lvl + L + 1);
        for (int r = r0 + warp; r < r1; r += W) {
            const uint32_t desc = __ldg(h.cz_desc + r);
            const int T = (int)(desc & 0xfffu);
            const uint32_t *row = h.cz_tab + (size_t)(desc >> 12) * 32 + lane;
            double *dst = c.wpos((int)__ldg(row));
            row += 32;
            double acc = *dst;
            // groups of G triangle words, the next group in flight while this one is relaxed (a row of the dense top
            // of the hierarchy has dozens of triangles per edge and is the critical path of its level)
            uint32_t cur[G], nxt[G];
#pragma unroll
            for (int i = 0; i < G; ++i) cur[i] = i < T ? __ldg(row + 32 * i) : inf_pair;
            for (int t = 0; t < T; t += G) {
                row += 32 * G;
#pragma unroll
                for (int i = 0; i < G; ++i) nxt[i] = t + G + i < T ? __ldg(row + 32 * i) : inf_pair;
#pragma unroll
                for (int i = 0; i < G; ++i) acc = relax(acc, cur[i]);
#pragma unroll
                for (int i = 0; i < G; ++i) cur[i] = nxt[i];
            }
            *dst = acc;
        }
    }
    __syncthreads();
}

// Bellman-Ford to the least fixed point with nothing but dist[]: the fallback for an origin whose
// labels leave the exactly-summable range (never on the shipped networks; B200_CCH_LIMIT forces it
// in the tests).
template <typename REAL>
__device__ int warp_sssp_bf(const Cta<REAL> &c, const Params &p, int origin, const WarpScratch &s)
{
    const int lane = threadIdx.x & 31, N = c.N;
    double *dist = s.dist;
    const int32_t *in_ptr = p.topo, *in_src = p.topo + p.off_in_src;   // the 32-bit blob in HBM: this path is cold
    for (int v = lane; v < N; v += 32) dist[v] = (v == origin) ? 0.0 : B200_INF_F64;
    __syncwarp();
    for (int round = 0; round <= N; ++round) {
        bool changed = false;
        for (int v = lane; v < N; v += 32) {
            const double dv = dist[v];
            double best = dv;
            for (int k = __ldg(in_ptr + v); k < __ldg(in_ptr + v + 1); ++k) {
                const double cand = dist[__ldg(in_src + k)] + c.w[k];
                best = cand < best ? cand : best;
            }
            if (best < dv) { dist[v] = best; changed = true; }
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, changed)) return 0;
    }
    return 1;
}

#ifdef B200_PHASE_TIMING
// tuning builds only (B200_NVCC_EXTRA=-DB200_PHASE_TIMING): SM clocks CTA 0 spends per phase, summed over iterations
__device__ unsigned long long g_phase_clocks[8], g_sub_clocks[8];
__device__ int g_dbg_flags;   // bit 0: skip the link-load atomics of the chain walk (timing experiment, wrong results)
#define B200_PHASE_MARK(i)                                                         \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                                     \
        const long long now_ = clock64();                                          \
        g_phase_clocks[i] += (unsigned long long)(now_ - mark_);                   \
        mark_ = now_;                                                              \
    }
// inside one origin solve: warp 0 of CTA 0
#define B200_SUB_BEGIN long long sub_ = clock64();
#define B200_SUB_MARK(i)                                                           \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                                     \
        const long long now_ = clock64();                                          \
        g_sub_clocks[i] += (unsigned long long)(now_ - sub_);                      \
        sub_ = now_;                                                               \
    }
#define B200_DBG_SKIP_AUX (g_dbg_flags & 1)
#define B200_DBG_COUNT(i, n) if (blockIdx.x == 0 && threadIdx.x == 0) g_sub_clocks[i] += (n);
#else
#define B200_PHASE_MARK(i)
#define B200_SUB_BEGIN
#define B200_SUB_MARK(i)
#define B200_DBG_SKIP_AUX 0
#define B200_DBG_COUNT(i, n)
#endif


// Down stream of the hierarchy (b200_cch.cu): batches of a header row + D edge rows, all of one shape, so
// the loop body is branch-free; the words of the NEXT batch are fetched into registers before the current
// one is relaxed, which hides the latency of whatever level of the memory hierarchy holds the table.
template <int D>
__device__ __forceinline__ void cch_down_sweep(const uint32_t *__restrict__ rows, int batches, double *dist,
                                               const double *__restrict__ wd)
{
    if (batches <= 0) return;
    const uint32_t *rp = rows + (threadIdx.x & 31);
    uint32_t cur[D + 1], nxt[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i) cur[i] = __ldg(rp + i * 32);
    for (int b = 0; b < batches; ++b) {
        rp += (D + 1) * 32;
        if (b + 1 < batches) {
#pragma unroll
            for (int i = 0; i <= D; ++i) nxt[i] = __ldg(rp + i * 32);
        }
        __syncwarp();   // the labels this batch pulls from (earlier levels / stages) are complete
        double *dst = dist + (cur[0] & 0xffffu);
        double best = (cur[0] & 0x10000u) ? B200_INF_F64 : *dst;   // partial slots are reused from level to level
        double cand[D];
#pragma unroll
        for (int i = 0; i < D; ++i) cand[i] = dist[cur[i + 1] & 0xffffu] + wd[cur[i + 1] >> 16];
#pragma unroll
        for (int i = 0; i < D; ++i) best = cand[i] < best ? cand[i] : best;
        // a destination with more than D in-edges owns 2^span adjacent lanes of the batch: minimum over the group
        const unsigned span = (cur[0] >> 17) & 7u;
        if (__any_sync(0xffffffffu, span != 0u)) {   // warp-uniform
#pragma unroll
            for (int st = 0; st < 5; ++st) {
                const double other = __shfl_xor_sync(0xffffffffu, best, 1 << st);
                if (st < (int)span) best = other < best ? other : best;
            }
        }
        if (((threadIdx.x & 31) & ((1u << span) - 1u)) == 0u) *dst = best;
#pragma unroll
        for (int i = 0; i <= D; ++i) cur[i] = nxt[i];
    }
}

// One warp, one origin: up stream of the origin (the nodes it reaches through edges to higher ranks push
// along those edges, in rising rank), then the down stream shared by all origins; labels are exact sums,
// checked below.
template <typename REAL>
__device__ int warp_sssp_cch(const Cta<REAL> &c, const Params &p, int slot, int origin, const WarpScratch &s)
{
    const CchParams &h = p.cch;
    const int lane = threadIdx.x & 31, N = c.N;
    double *dist = s.dist;
    B200_SUB_BEGIN
    for (int v = lane; v < h.slots; v += 32) dist[v] = (v == origin) ? 0.0 : B200_INF_F64;
    {
        const double *__restrict__ wu = c.wu;
        const int q0 = (int)__ldg(h.up_off + slot), q1 = (int)__ldg(h.up_off + slot + 1);
        const uint32_t *rp = h.up_rows + (size_t)q0 * 64 + lane;
        uint32_t a = 0, e = 0;
        if (q0 < q1) { a = __ldg(rp); e = __ldg(rp + 32); }
        for (int q = q0; q < q1; ++q) {
            rp += 64;
            uint32_t an = 0, en = 0;
            if (q + 1 < q1) { an = __ldg(rp); en = __ldg(rp + 32); }
            __syncwarp();   // every lower node has pushed: the label of this one is final
            const double cand = dist[a & 0xffffu] + wu[e >> 16];
            double *dx = dist + (e & 0xffffu);
            if (cand < *dx) *dx = cand;
            a = an; e = en;
        }
    }
    __syncwarp();
    B200_SUB_MARK(0)   // label init + up phase
    switch (h.D) {
        case 2: cch_down_sweep<2>(h.dn_rows, h.dn_batches, dist, c.wd); break;
        case 3: cch_down_sweep<3>(h.dn_rows, h.dn_batches, dist, c.wd); break;
        case 4: cch_down_sweep<4>(h.dn_rows, h.dn_batches, dist, c.wd); break;
        case 5: cch_down_sweep<5>(h.dn_rows, h.dn_batches, dist, c.wd); break;
        case 6: cch_down_sweep<6>(h.dn_rows, h.dn_batches, dist, c.wd); break;
        default: cch_down_sweep<8>(h.dn_rows, h.dn_batches, dist, c.wd); break;
    }
    __syncwarp();
    B200_SUB_MARK(1)   // down sweep
    unsigned mx = 0;
    for (int v = lane; v < N; v += 32) {
        const unsigned hi = (unsigned)__double2hiint(dist[v]);
        mx = hi < 0x7ff00000u && hi > mx ? hi : mx;
    }
    int capped = 0;
    if (__reduce_max_sync(0xffffffffu, mx) >= h.exact_hi) capped = warp_sssp_bf<REAL>(c, p, origin, s);
    // predecessor: tight in-link with the smallest dist[u], then the largest u - from the select stream (rows of
    // "tail | CSC slot" per block of 32 nodes, ascending tail, so "<=" leaves the largest u among equals)
    const double *__restrict__ w = c.w;
    const uint32_t *sp = h.sel_rows + lane;
    // (tried: fetching the next block's rows ahead like the down stream does - no gain on Chicago-Sketch-size and
    // spills under the 64-register bound of the two-CTA instantiation)
    for (int v = lane; v < ((N + 31) & ~31); v += 32, sp += h.SD * 32) {
        const double dv = v < N ? dist[v] : B200_INF_F64;
        int bk = kNoPk;
        double bdu = B200_INF_F64;
#pragma unroll 4
        for (int j = 0; j < h.SD; ++j) {
            const uint32_t e = __ldg(sp + j * 32);
            const double du = dist[e & 0xffffu];
            const bool take = du + w[e >> 16] == dv && du <= bdu;
            bk = take ? (int)(e >> 16) : bk;
            bdu = take ? du : bdu;
        }
        if (v < N) s.pk[v] = (uint16_t)((v != origin && dv < B200_INF_F64) ? bk : kNoPk);
    }
    __syncwarp();
    B200_SUB_MARK(2)   // exactness check + predecessor selection
    return capped;
}

// ---------------------------------------------------------------------------------------------
// S3: one warp, one origin - subtree demand sums, leaves first (Kahn), link loads into c.aux
// ---------------------------------------------------------------------------------------------
// Every tree link is handled once: a node whose children are all done adds its subtree sum to
// its in-link and to its parent.  The frontier of "ready" nodes is a queue; pending-children
// counters are 16-bit, two per 32-bit word, updated with 32-bit shared-memory atomics.
// INTQ: sums are int32 (exact, order-independent); otherwise REAL atomics (order-dependent in the
// last bits, like any parallel sum).  Returns this lane's share of the unassigned demand.
template <typename REAL, bool INTQ>
__device__ double warp_aon_peel(Cta<REAL> &c, const Params &p, int slot, int origin, const WarpScratch &s)
{
    using QT = typename std::conditional<INTQ, int, REAL>::type;
    const int lane = threadIdx.x & 31, N = c.N;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    const uint16_t *__restrict__ in_src = c.in_src, *__restrict__ in_eid = c.in_eid;
    QT *sacc = (QT *)s.dist;   // dist is dead once pk is chosen
    QT *auxq = (QT *)c.aux;
    const REAL *q = (const REAL *)p.od_q;
    for (int v = lane; v < N; v += 32) sacc[v] = 0;
    for (int i = lane; i < (N + 1) / 2; i += 32) s.cntw[i] = 0u;
    __syncwarp();
    const int jd1 = __ldg(p.od_ptr + slot + 1);
    for (int j = __ldg(p.od_ptr + slot) + lane; j < jd1; j += 32) sacc[__ldg(p.od_dst + j)] = (QT)q[j];
    __syncwarp();
    double un = 0.0;
    for (int v = lane; v < N; v += 32) {
        const int k = s.pk[v];
        if (k != kNoPk) {
            const int u = in_src[k];
            atomicAdd(&s.cntw[u >> 1], (u & 1) ? 0x10000u : 1u);
        } else {
            un += (double)sacc[v];   // unreachable destination, or demand to the origin itself (reference :496)
        }
    }
    __syncwarp();
    uint16_t *cur = s.q0, *nxt = s.q1;
    int count = 0;
    for (int v0 = 0; v0 < N; v0 += 32) {
        const int v = v0 + lane;
        bool leaf = false;
        if (v < N && s.pk[v] != kNoPk) leaf = ((s.cntw[v >> 1] >> ((v & 1) * 16)) & 0xffffu) == 0u;
        const unsigned m = __ballot_sync(FULL, leaf);
        if (leaf) cur[count + __popc(m & lt)] = (uint16_t)v;
        count += __popc(m);
    }
    __syncwarp();
    while (count > 0) {
        int ntail = 0;
        for (int base = 0; base < count; base += 32) {
            bool ready = false;
            int u = 0;
            if (base + lane < count) {
                const int v = cur[base + lane], k = s.pk[v];
                u = in_src[k];
                const QT sv = sacc[v];
                if (sv != (QT)0) {
                    atomicAdd(&auxq[in_eid[k]], sv);
                    atomicAdd(&sacc[u], sv);
                }
                const uint32_t old = atomicSub(&s.cntw[u >> 1], (u & 1) ? 0x10000u : 1u);
                ready = ((old >> ((u & 1) * 16)) & 0xffffu) == 1u && u != origin;   // that was the last child
            }
            const unsigned m = __ballot_sync(FULL, ready);
            if (ready) nxt[ntail + __popc(m & lt)] = (uint16_t)u;
            ntail += __popc(m);
        }
        __syncwarp();
        count = ntail;
        uint16_t *t = cur; cur = nxt; nxt = t;
    }
    return un;
}

// Integer demands whose per-origin total stays below 2^28 on a network with out-degree <= 15 (every
// shipped / generated network): ONE 32-bit word per node = children still pending << 28 | demand sum of
// the subtree so far, so delivering a subtree to its parent is a single shared-memory atomic, and the
// lane that delivers the LAST child keeps walking up with the parent - no ready queues, no ballots.
// Lanes start from the leaves (lane l takes leaves l, l + 32, ..) and pick the next one when their
// chain ends at a node that still waits for other children.
template <typename REAL>
__device__ double warp_aon_chain(Cta<REAL> &c, const Params &p, int slot, int origin, const WarpScratch &s)
{
    constexpr unsigned kChild = 1u << 28, kSum = kChild - 1u;
    const int lane = threadIdx.x & 31, N = c.N;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    const uint16_t *__restrict__ in_src = c.in_src, *__restrict__ in_eid = c.in_eid;
    unsigned *word = (unsigned *)s.dist;   // dist is dead once pk is chosen
    int *auxq = (int *)c.aux;
    const REAL *q = (const REAL *)p.od_q;
    B200_SUB_BEGIN
    for (int v = lane; v < N; v += 32) word[v] = 0u;
    __syncwarp();
    const int jd1 = __ldg(p.od_ptr + slot + 1);
    for (int j = __ldg(p.od_ptr + slot) + lane; j < jd1; j += 32) word[__ldg(p.od_dst + j)] = (unsigned)(int)q[j];
    __syncwarp();
    double un = 0.0;
    for (int v = lane; v < N; v += 32) {
        const int k = s.pk[v];
        if (k != kNoPk) atomicAdd(&word[in_src[k]], kChild);
        else un += (double)(word[v] & kSum);   // unreachable destination, or demand to the origin itself (reference :496)
    }
    __syncwarp();
    uint16_t *leaves = s.q0;
    int count = 0;
    for (int v0 = 0; v0 < N; v0 += 32) {
        const int v = v0 + lane;
        const bool leaf = v < N && s.pk[v] != kNoPk && word[v] < kChild;
        const unsigned m = __ballot_sync(FULL, leaf);
        if (leaf) leaves[count + __popc(m & lt)] = (uint16_t)v;
        count += __popc(m);
    }
    __syncwarp();
    B200_SUB_MARK(3)   // demand scatter, children counts, leaf list
    // Tried and dropped: the tree parent packed into the word as well (the delivering atomic then returns where to
    // go next: one dependent shared-memory operation per step instead of three) - no gain, the walk is bound by its
    // ~70 serial steps per origin (tree height) at ~6 active lanes, not by the latency of a step.
    int next = lane, v = -1;
    unsigned sv = 0u;
    for (;;) {
        if (v < 0 && next < count) {
            v = leaves[next];
            sv = word[v];   // a leaf: no children, the word is its own demand
            next += 32;
        }
        if (!__any_sync(FULL, v >= 0)) break;
#ifdef B200_PHASE_TIMING
        {
            const int act_ = __popc(__ballot_sync(FULL, v >= 0));   // every lane takes part in the ballot
            B200_DBG_COUNT(5, 1)
            B200_DBG_COUNT(6, act_)
        }
#endif
        if (v >= 0) {
            const int k = s.pk[v], u = in_src[k];
            if (sv != 0u && !B200_DBG_SKIP_AUX) atomicAdd(&auxq[in_eid[k]], (int)sv);
            const unsigned old = atomicAdd(&word[u], sv - kChild);
            // the last child to arrive carries the parent's completed subtree sum upwards
            const bool last = (old >> 28) == 1u && u != origin;
            sv = (old + sv) & kSum;
            v = last ? u : -1;
        }
    }
    B200_SUB_MARK(4)   // chain walk
    return un;
}

template <typename REAL>
__device__ __forceinline__ double warp_aon(Cta<REAL> &c, const Params &p, int slot, int origin, const WarpScratch &s)
{
    if (p.kahn_pack) return warp_aon_chain<REAL>(c, p, slot, origin, s);
    return c.intq ? warp_aon_peel<REAL, true>(c, p, slot, origin, s) : warp_aon_peel<REAL, false>(c, p, slot, origin, s);
}

// ---------------------------------------------------------------------------------------------
// S4: step size, flow update, new costs (fused S1), optional gap; block-wide
// ---------------------------------------------------------------------------------------------
template <typename REAL>
__device__ double ls_derivative(Cta<REAL> &c, const Params &p, const uint8_t *dmg_b, double lam)
{
    double g = 0.0;
    for (int k = threadIdx.x; k < c.E; k += blockDim.x) {
        const int e = c.in_eid[k];
        const double f = (double)c.flow[e], d = (double)c.auxv(e) - f;
        double te;
        if (dmg_b[e]) te = 1e6;
        else te = Bpr<double>::eval(f + lam * d, (double)__ldg(p.cap0_csc + k), (double)__ldg(p.t0_csc + k), p.alpha,
                                    p.beta, p.beta == 4.0);
        g += d * te;
    }
    return c.block_sum(g);
}

template <typename REAL>
__device__ void cta_update(Cta<REAL> &c, const Params &p, const uint8_t *dmg_b, int it, REAL *dprev_b, bool want_gap,
                           double *gap_out, int *nan_seen)
{
    if (want_gap) {
        double sft = 0.0, sat = 0.0;
        for (int k = threadIdx.x; k < c.E; k += blockDim.x) {
            const int e = c.in_eid[k];
            sft += (double)c.flow[e] * c.w[k];
            sat += (double)c.auxv(e) * c.w[k];
        }
        sft = c.block_sum(sft);
        sat = c.block_sum(sat);
        if (threadIdx.x == 0) *gap_out = sft > 0.0 ? (sft - sat) / sft : B200_INF_F64;
    }
    int saw_nan = 0;
    if (p.method == B200_CFW) {
        REAL bb = 0;
        if (it > 0) {
            double num = 0.0, den = 0.0;
            for (int e = threadIdx.x; e < c.E; e += blockDim.x) {
                const REAL dir = c.auxv(e) - c.flow[e], dp = dprev_b[e];
                num += (double)dir * (double)(dir - dp);
                den += (double)dp * (double)dp;
            }
            num = c.block_sum(num);
            den = c.block_sum(den);
            double b = num / (den + 1e-12);
            if (!(b > 0.0)) b = 0.0;
            bb = (REAL)b;
        }
        const REAL lam = (REAL)(2.0 / (it + 2.0));
        for (int k = threadIdx.x; k < c.E; k += blockDim.x) {
            const int e = c.in_eid[k];
            REAL dir = c.auxv(e) - c.flow[e];
            if (it > 0) dir = radd<REAL>(dir, rmul<REAL>(bb, dprev_b[e]));
            REAL f = radd<REAL>(c.flow[e], rmul<REAL>(lam, dir));
            f = f > (REAL)0 ? f : (REAL)0;  // np.maximum(., 0); a NaN ends as 0 after nan_to_num
            dprev_b[e] = dir;
            c.flow[e] = f;
            c.w[k] = (double)link_cost<REAL>(p, k, f, dmg_b[e] != 0);
        }
    } else {
        double lam64;
        if (p.method == B200_MSA) lam64 = 1.0 / (it + 1.0);
        else if (p.method == B200_FW) lam64 = 2.0 / (it + 2.0);
        else {
            // bisection on g(lam) = sum_e (aux_e - f_e) t_e(f + lam (aux - f)), fixed probe count
            // first step = the all-or-nothing load itself (as fw's 2/(it+2)): flow = 0 is not feasible
            if (it == 0 || ls_derivative<REAL>(c, p, dmg_b, 1.0) <= 0.0) lam64 = 1.0;
            else {
                double lo = 0.0, hi = 1.0;
                for (int i = 0; i < p.ls_iters; ++i) {
                    const double mid = 0.5 * (lo + hi);
                    if (ls_derivative<REAL>(c, p, dmg_b, mid) > 0.0) hi = mid; else lo = mid;
                }
                lam64 = 0.5 * (lo + hi);
            }
        }
        const REAL lam = (REAL)lam64, oml = (REAL)(1.0 - lam64);
        for (int k = threadIdx.x; k < c.E; k += blockDim.x) {
            const int e = c.in_eid[k];
            REAL f = radd<REAL>(rmul<REAL>(oml, c.flow[e]), rmul<REAL>(lam, c.auxv(e)));
            if (f != f) { f = 0; saw_nan = 1; }
            c.flow[e] = f;
            c.w[k] = (double)link_cost<REAL>(p, k, f, dmg_b[e] != 0);
        }
    }
    if (saw_nan) *nan_seen = 1;
}

// TSTT of (flow, w) - products in link order into `scratch`, numpy-ordered sum by warp 0.
template <typename REAL>
__device__ void cta_tstt(Cta<REAL> &c, const Params &p, REAL *scratch, double unassigned, double *tstt_out)
{
    for (int k = threadIdx.x; k < c.E; k += blockDim.x) {
        const int e = c.in_eid[k];
        scratch[e] = rmul<REAL>(c.flow[e], (REAL)c.w[k]);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const REAL base = radd<REAL>((REAL)0, warp_pairwise<REAL>(scratch, c.E));
        if (threadIdx.x == 0) {
            const double total = p.total_demand > 1.0 ? p.total_demand : 1.0;
            double v = (double)base / total;
            if (unassigned > 0.0) v += p.penalty * (unassigned / total);
            *tstt_out = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// S1 is the one stage that really streams from HBM (9 bytes per link: flow + mask in, cost out).
// Grid-stride sweep, VEC links per thread per step (VEC = 4: 16-byte flow loads / cost stores and
// 4-byte mask loads when rows are 16-byte multiples), link index within the row kept by an
// incremental wrap instead of a 64-bit modulo per element; t0 / cap0 come from L1 (__ldg).
template <typename REAL, int VEC>
__global__ void __launch_bounds__(256) bpr_kernel(Params p)
{
    const long long total = (long long)p.B * p.E / VEC;           // vectors
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int E = p.E;
    int e = (int)((i * VEC) % E);
    const int step = (int)((stride * VEC) % E);
    const bool beta4 = p.beta == 4.0;
    const REAL alpha = (REAL)p.alpha, beta = (REAL)p.beta;
    if constexpr (VEC == 4 && sizeof(REAL) == 4) {
        auto eval4 = [&](const float4 &f, const uchar4 &d, int ee) {
            const float4 c = __ldg((const float4 *)(p.cap0 + ee)), z = __ldg((const float4 *)(p.t0 + ee));
            float4 t;
            t.x = d.x ? 1e6f : Bpr<float>::eval(f.x, c.x, z.x, alpha, beta, beta4);
            t.y = d.y ? 1e6f : Bpr<float>::eval(f.y, c.y, z.y, alpha, beta, beta4);
            t.z = d.z ? 1e6f : Bpr<float>::eval(f.z, c.z, z.z, alpha, beta, beta4);
            t.w = d.w ? 1e6f : Bpr<float>::eval(f.w, c.w, z.w, alpha, beta, beta4);
            return t;
        };
        // two vectors per trip: both loads are in flight before the first result is needed
        for (; i + stride < total; i += 2 * stride) {
            const float4 f0 = ((const float4 *)p.flow)[i], f1 = ((const float4 *)p.flow)[i + stride];
            const uchar4 d0 = ((const uchar4 *)p.dmg)[i], d1 = ((const uchar4 *)p.dmg)[i + stride];
            int e1 = e + step;
            if (e1 >= E) e1 -= E;
            ((float4 *)p.t)[i] = eval4(f0, d0, e);
            ((float4 *)p.t)[i + stride] = eval4(f1, d1, e1);
            e = e1 + step;
            if (e >= E) e -= E;
        }
        if (i < total) ((float4 *)p.t)[i] = eval4(((const float4 *)p.flow)[i], ((const uchar4 *)p.dmg)[i], e);
    } else {
        for (; i < total; i += stride) {
            ((REAL *)p.t)[i] = p.dmg[i] ? (REAL)1e6
                                        : Bpr<REAL>::eval(((const REAL *)p.flow)[i], (REAL)__ldg(p.cap0 + e),
                                                          (REAL)__ldg(p.t0 + e), alpha, beta, beta4);
            e += step;
            if (e >= E) e -= E;
        }
    }
}

template <typename REAL>
__global__ void __launch_bounds__(kMaxWarps * 32) sssp_kernel(Params p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    Cta<REAL> c;
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    c.carve(smem, p, W);
    c.stage_topology(p);
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const REAL *t_b = (const REAL *)p.t + (size_t)b * p.E;
        __syncthreads();
        for (int k = threadIdx.x; k < c.E; k += blockDim.x) c.w[k] = (double)t_b[c.in_eid[k]];
        __syncthreads();
        int capped = 0;
        for (int s = warp; s < p.O; s += W) {
            const int o = __ldg(p.origins + s);
            const WarpScratch ws = warp_scratch(c, warp);
            capped |= warp_sssp<REAL>(c, o, ws);
            const size_t base = ((size_t)b * p.O + s) * p.N;
            for (int v = lane; v < p.N; v += 32) {
                const int k = ws.pk[v];
                p.dist_out[base + v] = ws.dist[v];
                p.pred_out[base + v] = k == kNoPk ? B200_NO_PRED : c.in_src[k];
                p.pedge_out[base + v] = k == kNoPk ? -1 : c.in_eid[k];
            }
            __syncwarp();
        }
        if (capped && lane == 0 && p.status) atomicOr(p.status + b, B200_ST_SSSP_CAP);
    }
}

template <typename REAL>
__global__ void __launch_bounds__(kMaxWarps * 32) aon_kernel(Params p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    Cta<REAL> c;
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    c.carve(smem, p, W);
    c.stage_topology(p);
    // CSC slot of each link id, to turn pedge back into slots (scratch: reuse w[] as int32)
    int32_t *slot_of = (int32_t *)c.w;
    for (int k = threadIdx.x; k < c.E; k += blockDim.x) slot_of[c.in_eid[k]] = k;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        for (int e = threadIdx.x; e < c.E; e += blockDim.x) c.aux[e] = 0;
        if (threadIdx.x == 0) c.red[kMaxWarps * 2] = 0.0;
        __syncthreads();
        double un = 0.0;
        for (int s = warp; s < p.O; s += W) {
            const int o = __ldg(p.origins + s);
            const WarpScratch ws = warp_scratch(c, warp);
            const int32_t *pe = p.pedge_in + ((size_t)b * p.O + s) * p.N;
            for (int v = lane; v < p.N; v += 32) { const int e = pe[v]; ws.pk[v] = (uint16_t)(e < 0 ? kNoPk : slot_of[e]); }
            __syncwarp();
            un += warp_aon<REAL>(c, p, s, o, ws);
            __syncwarp();
        }
        un = warp_sum_f64(un);
        if (lane == 0 && un != 0.0) atomicAdd(&c.red[kMaxWarps * 2], un);
        __syncthreads();
        REAL *aux_b = (REAL *)p.aux + (size_t)b * p.E;
        for (int e = threadIdx.x; e < c.E; e += blockDim.x) aux_b[e] = c.auxv(e);
        if (threadIdx.x == 0) p.unassigned[b] = c.red[kMaxWarps * 2];
    }
}

template <typename REAL>
__global__ void __launch_bounds__(kMaxWarps * 32) update_kernel(Params p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    Cta<REAL> c;
    const int W = blockDim.x >> 5;
    c.carve(smem, p, W);
    c.intq = false;   // aux comes from HBM as REAL here
    c.stage_topology(p);
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const size_t off = (size_t)b * p.E;
        REAL *flow_b = (REAL *)p.flow + off, *t_b = (REAL *)p.t + off;
        const REAL *aux_b = (const REAL *)p.aux + off;
        const uint8_t *dmg_b = p.dmg + off;
        REAL *dprev_b = p.dprev ? (REAL *)p.dprev + off : nullptr;
        __syncthreads();
        for (int k = threadIdx.x; k < c.E; k += blockDim.x) {
            const int e = c.in_eid[k];
            c.w[k] = (double)t_b[e];
            c.flow[e] = flow_b[e];
            c.aux[e] = aux_b[e];
        }
        __syncthreads();
        int nan_seen = 0;
        cta_update<REAL>(c, p, dmg_b, p.it0, dprev_b, true, p.gap + b, &nan_seen);
        __syncthreads();
        for (int k = threadIdx.x; k < c.E; k += blockDim.x) {
            const int e = c.in_eid[k];
            flow_b[e] = c.flow[e];
            t_b[e] = (REAL)c.w[k];
        }
        __syncthreads();
        cta_tstt<REAL>(c, p, c.aux, p.unassigned_in ? p.unassigned_in[b] : 0.0, p.tstt + b);
        if (p.status) {
            int st = nan_seen ? B200_ST_NAN_FLOW : 0;
            if (threadIdx.x == 0 && p.unassigned_in && p.unassigned_in[b] > 0.0) st |= B200_ST_UNASSIGNED;
            if (st) atomicOr(p.status + b, st);
        }
    }
}

// Thread-block cluster mode (few scenarios, many origins: e.g. 32 Chicago-Sketch-size scenarios
// per GPU): the C CTAs of a cluster share ONE scenario and split its origins.  Each CTA loads its
// origins' trees into its own aux[]; after a cluster barrier every CTA pulls the other CTAs'
// partial link loads through distributed shared memory, in rank order, so all C copies of the
// total are bit-identical; each CTA then runs S4 redundantly on its own copy of flow / w (cheap:
// O(E)), which avoids any further exchange.  Two cluster barriers per iteration.
template <typename REAL, bool INTQ>
__device__ void cluster_combine(Cta<REAL> &c, cg::cluster_group &cluster, double *un_s, int *flag_s)
{
    using QT = typename std::conditional<INTQ, int, REAL>::type;
    constexpr int CH = 8;
    const unsigned C = cluster.num_blocks();
    QT *mine = (QT *)c.aux;
    double un = 0.0;
    int flags = 0;
    if (threadIdx.x == 0) {
        for (unsigned r = 0; r < C; ++r) {
            un += *cluster.map_shared_rank(un_s, r);
            const int *f = cluster.map_shared_rank(flag_s, r);
            flags |= (f[0] ? 1 : 0) | (f[1] ? 2 : 0);
        }
    }
    for (int base = 0; base < c.E; base += CH * (int)blockDim.x) {
        QT acc[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int e = base + j * (int)blockDim.x + (int)threadIdx.x;
            acc[j] = 0;
            if (e < c.E)
                for (unsigned r = 0; r < C; ++r) acc[j] += ((const QT *)cluster.map_shared_rank(c.aux, r))[e];
        }
        cluster.sync();   // everybody has read this range (and, first pass, the scalars)
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int e = base + j * (int)blockDim.x + (int)threadIdx.x;
            if (e < c.E) mine[e] = acc[j];
        }
    }
    if (threadIdx.x == 0) {
        *un_s = un;
        flag_s[0] = flags & 1;
        flag_s[1] = (flags >> 1) & 1;
    }
    __syncthreads();
}

// The fused path: K iterations of S1..S4 with every per-scenario array in shared memory.
// MINB = 2: at most 64 registers per thread, for networks small enough that several CTAs share an
// SM (occupancy is what this latency-bound kernel lives on); MINB = 1: up to 128 registers when
// shared memory allows only one CTA per SM anyway (Chicago-Sketch-size)
template <typename REAL, int MINB, bool CCH>
__global__ void __launch_bounds__(kMaxWarps * 32, MINB) assign_fused_kernel(Params p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    Cta<REAL> c;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    c.carve(smem, p, W, CCH);
    c.stage_topology(p);
    if (CCH) {
        // zero weights of the ELL padding / partial-combine entries (never touched again), the +inf slot
        for (int i = threadIdx.x; i < p.cch.wd; i += blockDim.x) c.wd[i] = 0.0;
        if (threadIdx.x == 0) *c.wpos(p.cch.wd + p.cch.wu) = B200_INF_F64;
    }
    double *un_s = &c.red[kMaxWarps * 2];
    int *flag_s = (int *)&c.red[kMaxWarps * 2 + 1];
    const int cluster_id = blockIdx.x / C, num_clusters = gridDim.x / C;
    for (int b = cluster_id; b < p.B; b += num_clusters) {
        if (p.active && !p.active[b]) continue;  // cluster-uniform
        const size_t off = (size_t)b * p.E;
        REAL *flow_b = (REAL *)p.flow + off;
        const uint8_t *dmg_b = p.dmg + off;
        REAL *dprev_b = p.dprev ? (REAL *)p.dprev + off : nullptr;
        __syncthreads();
        for (int e = threadIdx.x; e < c.E; e += blockDim.x) c.flow[e] = flow_b[e];
        if (threadIdx.x == 0) { flag_s[0] = 0; flag_s[1] = 0; }
        __syncthreads();
        cta_costs_from_flow<REAL>(c, p, dmg_b);
#ifdef B200_PHASE_TIMING
        long long mark_ = clock64();
#endif
        for (int it = 0; it < p.K; ++it) {
            for (int e = threadIdx.x; e < c.E; e += blockDim.x) c.aux[e] = 0;
            if (threadIdx.x == 0) *un_s = 0.0;
            __syncthreads();   // (neighbours finished pulling last iteration's loads before its last cluster barrier)
            B200_PHASE_MARK(0)
            if (CCH) cta_customize<REAL, (MINB == 1 ? 8 : 4)>(c, p);
            B200_PHASE_MARK(1)
            double un = 0.0;
            int capped = 0;
            for (int s = rank * W + warp; s < p.O; s += C * W) {
                const int o = __ldg(p.origins + s);
                const WarpScratch ws = warp_scratch(c, warp, CCH ? p.cch.slots : 0);
                if (CCH) {
                    capped |= warp_sssp_cch<REAL>(c, p, s, o, ws);
                } else {
                    uint16_t *tree = p.tree_ws + ((size_t)cluster_id * p.O + s) * p.N;
                    capped |= warp_sssp<REAL>(c, o, ws, (it > 0 && p.warm) ? tree : nullptr);
                    if (p.warm)
                        for (int v = lane; v < p.N; v += 32) tree[v] = ws.pk[v];   // next iteration's warm start
                }
                un += warp_aon<REAL>(c, p, s, o, ws);
                __syncwarp();
            }
            un = warp_sum_f64(un);
            if (lane == 0 && un != 0.0) atomicAdd(un_s, un);
            if (__any_sync(0xffffffffu, capped) && lane == 0) flag_s[1] = 1;
            B200_PHASE_MARK(2)   // warp 0's own origins
            __syncthreads();
            B200_PHASE_MARK(3)   // waiting for the CTA's other warps
            if (C > 1) {
                cluster.sync();   // every CTA's partial loads are complete
                B200_PHASE_MARK(4)
                if (c.intq) cluster_combine<REAL, true>(c, cluster, un_s, flag_s);
                else cluster_combine<REAL, false>(c, cluster, un_s, flag_s);
                B200_PHASE_MARK(5)
            }
            int nan_seen = 0;
            cta_update<REAL>(c, p, dmg_b, it, dprev_b, it == p.K - 1, p.gap + b, &nan_seen);
            if (nan_seen) flag_s[0] = 1;
            __syncthreads();
            B200_PHASE_MARK(6)
        }
        const double un_final = *un_s;
        cta_tstt<REAL>(c, p, c.aux, un_final, p.tstt + b);
        if (rank == 0) {
            for (int e = threadIdx.x; e < c.E; e += blockDim.x) flow_b[e] = c.flow[e];
            if (threadIdx.x == 0) {
                p.unassigned[b] = un_final;
                p.status[b] = (flag_s[0] ? B200_ST_NAN_FLOW : 0) | (un_final > 0.0 ? B200_ST_UNASSIGNED : 0) |
                              (flag_s[1] ? B200_ST_SSSP_CAP : 0);
            }
        }
    }
    if (C > 1) cluster.sync();   // no CTA exits while a neighbour may still read its shared memory
}


// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// subtree sums are 4-byte (int32 / float) -> the counters fit into the dead half of dist[]
static int gen_alias(const b200_graph *g, int real_bytes)
{
    return (g->gen_intq || real_bytes == 4) && getenv("B200_NO_ALIAS") == nullptr;
}

Params base_params(const b200_graph *g, int B, int prec, double alpha, double beta)
{
    Params p;
    memset(&p, 0, sizeof(p));
    p.N = g->N; p.E = g->E; p.O = g->O; p.topo_ints = g->topo_ints;
    p.off_in_src = g->off_in_src; p.off_in_eid = g->off_in_eid;
    p.topo16 = g->topo16; p.topo16_bytes = g->topo16_bytes; p.o16_in_src = g->o16_in_src; p.o16_in_eid = g->o16_in_eid;
    p.o16_out_ptr = g->o16_out_ptr; p.o16_out_dst = g->o16_out_dst;
    p.intq = g->gen_intq; p.warm = g->warm_start; p.kahn_pack = g->kahn_pack;
    p.alias32 = gen_alias(g, 4); p.alias64 = gen_alias(g, 8);
    p.topo = g->topo; p.row = g->row; p.origins = g->origins; p.od_ptr = g->od_ptr; p.od_dst = g->od_dst;
    p.t0_csc = g->t0_csc; p.cap0_csc = g->cap0_csc; p.t0 = g->t0; p.cap0 = g->cap0;
    p.od_q = prec == B200_F32 ? (const void *)g->od_q32 : (const void *)g->od_q64;
    p.small_blob = g->small_intq ? g->small_blob_i32 : (prec == B200_F32 ? g->small_blob32 : g->small_blob64);
    p.small_eid = g->small_eid; p.small_wpos = g->small_wpos;
    p.st_origpos = g->st_origpos;
    p.total_demand = g->total_demand;
    if (g->cch_ok) {
        p.cch.dn_rows = g->cch_dn_rows; p.cch.up_rows = g->cch_up_rows; p.cch.sel_rows = g->cch_sel_rows;
        p.cch.up_off = g->cch_up_off; p.cch.cz_init = g->cch_cz_init;
        p.cch.cz_desc = g->cch_cz_desc; p.cch.cz_tab = g->cch_cz_tab; p.cch.cz_lvl = g->cch_cz_lvl;
        p.cch.F = g->cch_F; p.cch.wd = g->cch_wd; p.cch.wu = g->cch_wu; p.cch.slots = g->cch_slots;
        p.cch.dn_batches = g->cch_dn_batches; p.cch.D = g->cch_D; p.cch.SD = g->cch_SD;
        p.cch.cz_rows = g->cch_cz_rows; p.cch.cz_levels = g->cch_cz_levels; p.cch.wmode = g->cch_wmode; p.cch.leaf_alias = g->cch_leaf_alias;
        p.cch.topo_bytes = g->o16_out_ptr * 2;   // in_ptr | in_src | in_eid: the out-link arrays are not needed
        union { double d; unsigned long long u; } lim;
        lim.d = g->cch_exact_limit;
        p.cch.exact_hi = (unsigned)(lim.u >> 32);
    }
    p.B = B; p.alpha = alpha; p.beta = beta;
    return p;
}

template <typename K>
int launch_cta_kernel(K kernel, const b200_graph *g, int prec, const Params &p_in, cudaStream_t st, bool fused = false,
                      bool cch = false)
{
    Params p = p_in;
    const int warps = cch ? g->warps_cch : (prec == B200_F32 ? g->warps_f32 : g->warps_f64);
    const size_t smem = cch ? g->smem_cch : (prec == B200_F32 ? g->smem_f32 : g->smem_f64);
    B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem));
    if (per_sm < 1) { b200_set_error("kernel does not fit on an SM (smem %zu)", smem); return B200_EUNSUPPORTED; }
    // Small networks: one CTA per scenario, the hardware block scheduler balances the tail and
    // re-staging a sub-8-KB topology is free.  Large networks: persistent CTAs, topology staged once.
    const int capacity = per_sm * g->num_sms;
    int grid = (g->topo16_bytes <= 5120) ? p.B : std::min(p.B, capacity);
    int cluster = 1;
    if (fused) {
        // Fewer scenarios than resident CTAs: split every scenario's origins over a thread-block
        // cluster (power of two <= 8, no more CTAs than batches of `warps` origins).  cfw keeps one
        // CTA per scenario (its direction scratch lives in the caller's [B][E] array).
        if (g->cluster_limit != 1 && p.method != B200_CFW) {
            const int want = std::min({8, g->cluster_limit > 0 ? g->cluster_limit : 8, capacity / std::max(p.B, 1),
                                       (g->O + warps - 1) / warps});
            while (cluster * 2 <= want) cluster *= 2;
        }
        if (cluster > 1) {
            cudaLaunchConfig_t probe = {};
            probe.gridDim = dim3(cluster);
            probe.blockDim = dim3(warps * 32);
            probe.dynamicSmemBytes = smem;
            cudaLaunchAttribute pa[1];
            pa[0].id = cudaLaunchAttributeClusterDimension;
            pa[0].val.clusterDim.x = cluster; pa[0].val.clusterDim.y = 1; pa[0].val.clusterDim.z = 1;
            probe.attrs = pa; probe.numAttrs = 1;
            int max_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &probe) != cudaSuccess || max_clusters < 1) {
                cudaGetLastError();
                cluster = 1;
            } else {
                grid = cluster * std::min(p.B, max_clusters);
            }
        }
        if (cluster == 1) grid = (g->topo16_bytes <= 5120) ? p.B : std::min(p.B, capacity);
        // previous-iteration trees of the resident scenarios (the only allocation the library ever
        // makes after graph upload; it grows monotonically and is reused by later calls)
        const size_t need = cch ? 0 : (size_t)(grid / cluster) * g->O * g->N * sizeof(uint16_t);
        b200_graph *gm = const_cast<b200_graph *>(g);
        if (cch) {
            // the part of the hierarchy weights that lives outside shared memory: one slice per CTA
            const int wtot = g->cch_wd + g->cch_wu + 2;
            p.cch.ws_stride = g->cch_wmode == 2 ? 0 : (g->cch_wmode == 1 ? g->cch_wu + 2 : wtot);
            const size_t wneed = (size_t)grid * p.cch.ws_stride * sizeof(double);
            if (wneed > gm->cch_ws_bytes) {
                if (gm->cch_ws) { B200_CUDA(cudaStreamSynchronize(st)); B200_CUDA(cudaFree(gm->cch_ws)); gm->cch_ws = nullptr; gm->cch_ws_bytes = 0; }
                B200_CUDA(cudaMalloc((void **)&gm->cch_ws, wneed));
                gm->cch_ws_bytes = wneed;
            }
            p.cch.ws = gm->cch_ws;
        }
        if (need > gm->tree_ws_bytes) {
            if (gm->tree_ws) { B200_CUDA(cudaStreamSynchronize(st)); B200_CUDA(cudaFree(gm->tree_ws)); gm->tree_ws = nullptr; gm->tree_ws_bytes = 0; }
            B200_CUDA(cudaMalloc((void **)&gm->tree_ws, need));
            gm->tree_ws_bytes = need;
        }
        p.tree_ws = gm->tree_ws;
        gm->last_cluster = cluster;
        gm->last_warps = warps;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(warps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cluster > 1 ? 1 : 0;
    B200_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
    B200_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

static bool gen_alias_host(const b200_graph *g, int real_bytes)
{
    return (g->gen_intq || real_bytes == 4) && getenv("B200_NO_ALIAS") == nullptr;
}

// Warps per CTA: enough to spread the origins evenly, bounded by shared memory.
int b200_pick_warps(const b200_graph *g, int prec, size_t *smem_out)
{
    const int rb = prec == B200_F32 ? 4 : 8;
    int best = -1;
    if (g->N > 65534 || g->E > 65534) return -1;   // tree slots and queues are 16-bit
    // The kernel is latency-bound (dependent shared-memory loads, few active lanes per round), so what
    // counts is how many origins an SM works on at once: resident warps = warps per CTA x CTAs per
    // SM (shared memory: 227 KB per SM, 1 KB reserved per CTA), weighted by how evenly the origins
    // divide among the warps.  Ties go to fewer warps per CTA (more CTAs = more scenarios in flight).
    double best_score = -1.0;
    const int forced = getenv("B200_WARPS") ? atoi(getenv("B200_WARPS")) : 0;   // tuning experiments
    for (int w = kMinWarps; w <= kMaxWarps; ++w) {
        SmemLayout L = make_layout(g->N, g->E, g->topo16_bytes, w, rb, gen_alias_host(g, rb));
        if (L.total > kSmemLimit) break;
        if (forced && w != forced) continue;
        const int O = std::max(g->O, 1), batch = w;
        const double eff = (double)O / (double)(((O + batch - 1) / batch) * batch);
        const int ctas = std::min<int>((int)((228 * 1024) / (L.total + 1024)), std::min(32, 64 / w));
        const double score = eff * batch * std::max(ctas, 1);
        if (score > best_score + 1e-9) { best_score = score; best = w; }
    }
    if (best > 0 && smem_out) *smem_out = make_layout(g->N, g->E, g->topo16_bytes, best, rb, gen_alias_host(g, rb)).total;
    return best;
}

// Contraction-hierarchy variant of the fused kernel (fp32 link state): where the weights live and how
// many warps a CTA gets.  Small hierarchies keep all of W in shared memory; otherwise the down region
// (read in full by every origin) stays there if it leaves room for enough warps, and the rest - or all
// of it - is served from the CTA's slice of the HBM workspace through L1/L2.  B200_CCH_W=0/1/2 forces
// a placement (measurements).
int b200_pick_warps_cch(b200_graph *g)
{
    if (!g->cch_ok) return -1;
    const int forced_w = getenv("B200_WARPS") ? atoi(getenv("B200_WARPS")) : 0;
    const int forced_m = getenv("B200_CCH_W") ? atoi(getenv("B200_CCH_W")) : -1;
    double best_score = -1.0;
    int best_w = -1, best_m = 0, best_la = 0;
    size_t best_smem = 0;
    for (int mode = 2; mode >= 0; --mode) {
        if (forced_m >= 0 && mode != forced_m) continue;
        const size_t wb = mode == 2 ? (size_t)(g->cch_wd + g->cch_wu + 2) * 8 : mode == 1 ? (size_t)g->cch_wd * 8 : 0;
        for (int la = 0; la <= (g->kahn_pack ? 1 : 0); ++la)   // leaf list inside dist[]: only if it scores higher
        for (int w = kMinWarps; w <= kMaxWarps; ++w) {
            SmemLayout L = make_layout(g->N, g->E, g->o16_out_ptr * 2, w, 4, true, g->cch_slots, wb, g->kahn_pack ? 1 + la : 0);
            if (L.total > kSmemLimit) break;
            if (forced_w && w != forced_w) continue;
            const int O = std::max(g->O, 1);
            const double eff = (double)O / (double)(((O + w - 1) / w) * w);
            const int ctas = std::min<int>((int)((228 * 1024) / (L.total + 1024)), std::min(32, 64 / w));
            // resident warps, with a premium for weights in shared memory: every origin gathers from the whole
            // down region, and the up phase is a chain of dependent reads (L2 latency per ancestor otherwise)
            const double score = eff * w * std::max(ctas, 1) * (mode == 2 ? 1.3 : mode == 1 ? 1.15 : 1.0) * (la ? 0.97 : 1.0);
            if (score > best_score + 1e-9) { best_score = score; best_w = w; best_m = mode; best_smem = L.total; best_la = la; }
        }
    }
    if (best_w < 0) return -1;
    g->warps_cch = best_w;
    g->cch_leaf_alias = best_la;
    g->cch_wmode = best_m;
    g->smem_cch = best_smem;
    return best_w;
}

#ifdef B200_PHASE_TIMING
extern "C" int b200_debug_flags(int flags)
{
    B200_CUDA(cudaMemcpyToSymbol(g_dbg_flags, &flags, sizeof(int)));
    return 0;
}

extern "C" int b200_debug_phase_clocks(unsigned long long out[16], int reset)
{
    B200_CUDA(cudaDeviceSynchronize());
    B200_CUDA(cudaMemcpyFromSymbol(out, g_phase_clocks, sizeof(unsigned long long) * 8));
    B200_CUDA(cudaMemcpyFromSymbol(out + 8, g_sub_clocks, sizeof(unsigned long long) * 8));
    if (reset) {
        unsigned long long z[8] = {0};
        B200_CUDA(cudaMemcpyToSymbol(g_phase_clocks, z, sizeof(z)));
        B200_CUDA(cudaMemcpyToSymbol(g_sub_clocks, z, sizeof(z)));
    }
    return 0;
}
#endif

#define B200_PREC_OK(prec) B200_REQUIRE((prec) == B200_F32 || (prec) == B200_F64, "prec must be B200_F32 or B200_F64")

extern "C" int b200_bpr(const b200_graph *g, int B, int prec, const void *flow, const uint8_t *dmg, void *t,
                        double alpha, double beta, void *stream)
{
    B200_REQUIRE(g && flow && dmg && t && B >= 0, "bad argument");
    B200_PREC_OK(prec);
    if (B == 0) return 0;
    Params p = base_params(g, B, prec, alpha, beta);
    p.flow = (void *)flow; p.dmg = dmg; p.t = t;
    const long long total = (long long)B * g->E;
    const int threads = 256;
    // vector path: rows are multiples of 16 bytes and every base pointer is 16-byte aligned
    const bool vec = prec == B200_F32 && g->E % 4 == 0 &&
                     (((uintptr_t)flow | (uintptr_t)t | (uintptr_t)g->cap0 | (uintptr_t)g->t0) % 16 == 0) &&
                     (uintptr_t)dmg % 4 == 0;
    const long long units = vec ? total / 4 : total;
    const int grid = (int)std::min<long long>((units + threads - 1) / threads, (long long)g->num_sms * 8);
    if (vec) bpr_kernel<float, 4><<<grid, threads, 0, (cudaStream_t)stream>>>(p);
    else if (prec == B200_F32) bpr_kernel<float, 1><<<grid, threads, 0, (cudaStream_t)stream>>>(p);
    else bpr_kernel<double, 1><<<grid, threads, 0, (cudaStream_t)stream>>>(p);
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200_sssp(const b200_graph *g, int B, int prec, const void *t, double *dist, int32_t *pred,
                         int32_t *pedge, void *stream)
{
    B200_REQUIRE(g && t && dist && pred && pedge && B >= 0, "bad argument");
    B200_PREC_OK(prec);
    if (B == 0 || g->O == 0) return 0;
    Params p = base_params(g, B, prec, 0.15, 4.0);
    p.t = (void *)t; p.dist_out = dist; p.pred_out = pred; p.pedge_out = pedge;
    return prec == B200_F32 ? launch_cta_kernel(sssp_kernel<float>, g, prec, p, (cudaStream_t)stream)
                            : launch_cta_kernel(sssp_kernel<double>, g, prec, p, (cudaStream_t)stream);
}

extern "C" int b200_aon(const b200_graph *g, int B, int prec, const int32_t *pedge, void *aux, double *unassigned,
                        void *stream)
{
    B200_REQUIRE(g && pedge && aux && unassigned && B >= 0, "bad argument");
    B200_PREC_OK(prec);
    if (B == 0) return 0;
    Params p = base_params(g, B, prec, 0.15, 4.0);
    p.pedge_in = pedge; p.aux = aux; p.unassigned = unassigned;
    return prec == B200_F32 ? launch_cta_kernel(aon_kernel<float>, g, prec, p, (cudaStream_t)stream)
                            : launch_cta_kernel(aon_kernel<double>, g, prec, p, (cudaStream_t)stream);
}

extern "C" int b200_update_reduce(const b200_graph *g, int B, int prec, int method, int it, void *flow,
                                  const void *aux, void *t, const uint8_t *dmg, void *dprev,
                                  const double *unassigned, double *tstt, double *gap, int32_t *status, double alpha,
                                  double beta, double penalty, int ls_iters, void *stream)
{
    B200_REQUIRE(g && flow && aux && t && dmg && tstt && gap && B >= 0 && it >= 0, "bad argument");
    B200_PREC_OK(prec);
    B200_REQUIRE(method >= B200_MSA && method <= B200_FW_LS, "unknown assignment method");
    B200_REQUIRE(method != B200_CFW || dprev, "cfw needs the dprev scratch array");
    if (B == 0) return 0;
    Params p = base_params(g, B, prec, alpha, beta);
    p.method = method; p.it0 = it; p.flow = flow; p.aux = (void *)aux; p.t = t; p.dmg = dmg; p.dprev = dprev;
    p.unassigned_in = unassigned; p.tstt = tstt; p.gap = gap; p.status = status; p.penalty = penalty;
    p.ls_iters = ls_iters;
    return prec == B200_F32 ? launch_cta_kernel(update_kernel<float>, g, prec, p, (cudaStream_t)stream)
                            : launch_cta_kernel(update_kernel<double>, g, prec, p, (cudaStream_t)stream);
}

extern "C" int b200_assign(const b200_graph *g, int B, int prec, int K, int method, const uint8_t *dmg, void *flow,
                           void *dprev, double *tstt, double *gap, double *unassigned, int32_t *status,
                           const uint8_t *active, double alpha, double beta, double penalty, int ls_iters,
                           void *stream)
{
    B200_REQUIRE(g && B >= 0, "bad argument");
    B200_PREC_OK(prec);
    B200_REQUIRE(K > 0, "assignment_iters must be > 0 to update TSTT.");
    B200_REQUIRE(method >= B200_MSA && method <= B200_FW_LS, "unknown assignment method");
    const_cast<b200_graph *>(g)->last_launches = 0;
    const_cast<b200_graph *>(g)->last_cch = 0;
    if (B == 0) return 0;  // empty batch: nothing to do, pointers may be null
    B200_REQUIRE(dmg && flow && tstt && gap && unassigned && status, "null array argument");
    B200_REQUIRE(method != B200_CFW || dprev, "cfw needs the dprev scratch array");
    Params p = base_params(g, B, prec, alpha, beta);
    p.K = K; p.method = method; p.dmg = dmg; p.active = active; p.flow = flow; p.dprev = dprev; p.tstt = tstt; p.gap = gap;
    p.unassigned = unassigned; p.status = status; p.penalty = penalty; p.ls_iters = ls_iters;
    int rc;
    if (g->static_ok && !g->force_general && !g->force_small) {
        // compiled-in topology (Sioux Falls): straight-line lane-per-origin kernel
        rc = b200_launch_static(g, p, prec, (cudaStream_t)stream);
    } else if (g->small_ok && !g->force_general) {
        // Sioux-Falls class: one warp per scenario, lane = origin
        rc = b200_launch_small(g, p, prec, (cudaStream_t)stream);
    } else {
        const size_t smem = prec == B200_F32 ? g->smem_f32 : g->smem_f64;
        const bool one_cta = 2 * (smem + 1024) > 228 * 1024;   // shared memory admits a single CTA per SM
        cudaStream_t st = (cudaStream_t)stream;
        // fp32 link costs (>= t0 when alpha, beta >= 0) sum exactly in fp64 below the bound the kernel
        // checks per origin: shortest paths from the contraction hierarchy instead of label correcting
        const bool cch = prec == B200_F32 && g->cch_ok && !g->no_cch && g->warps_cch > 0 && alpha >= 0.0 && beta >= 0.0;
        if (cch) {
            const bool one_cta_c = 2 * (g->smem_cch + 1024) > 228 * 1024;
            const_cast<b200_graph *>(g)->last_cch = 1;
            rc = one_cta_c ? launch_cta_kernel(assign_fused_kernel<float, 1, true>, g, prec, p, st, true, true)
                           : launch_cta_kernel(assign_fused_kernel<float, 2, true>, g, prec, p, st, true, true);
        } else {
#define B200_LAUNCH_FUSED(REAL)                                                                 \
    (one_cta ? launch_cta_kernel(assign_fused_kernel<REAL, 1, false>, g, prec, p, st, true)      \
             : launch_cta_kernel(assign_fused_kernel<REAL, 2, false>, g, prec, p, st, true))
            rc = prec == B200_F32 ? B200_LAUNCH_FUSED(float) : B200_LAUNCH_FUSED(double);
        }
#undef B200_LAUNCH_FUSED
    }
    if (rc == 0) const_cast<b200_graph *>(g)->last_launches = 1;
    return rc;
}
