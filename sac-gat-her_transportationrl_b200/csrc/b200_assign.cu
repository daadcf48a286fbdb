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
#include "b200_common.cuh"

#include <algorithm>
#include <math.h>
#include <string.h>
#include <type_traits>
#include <vector>

namespace {

constexpr int kMaxWarps = 16;
constexpr int kMinWarps = 2;
constexpr size_t kSmemLimit = 227 * 1024;

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

struct SmemLayout {
    size_t off_bar, off_topo, off_w, off_aux, off_flow, off_red, off_warp, warp_stride, total;
};

__host__ __device__ inline SmemLayout make_layout(int N, int E, int topo_ints, int warps, int real_bytes)
{
    SmemLayout L;
    size_t o = 0;
    L.off_bar = o;  o += 16;
    L.off_topo = o; o += align16((size_t)topo_ints * 4);
    L.off_w = o;    o += align16((size_t)E * 8);
    L.off_aux = o;  o += align16((size_t)E * real_bytes);
    L.off_flow = o; o += align16((size_t)E * real_bytes);
    L.off_red = o;  o += align16((size_t)(kMaxWarps * 4 + 8) * 8);
    L.off_warp = o;
    L.warp_stride = align16((size_t)N * 8) + align16((size_t)N * 4);
    L.total = o + L.warp_stride * warps;
    return L;
}

// shared-memory layout of the small-network kernel, computed on the host (b200_assign_small.cuh)
struct SmallOffsets {
    int blob_bytes, off_rec2, off_q, off_dist, off_sacc, off_cnt, off_stk, off_pu, off_pe, off_w, off_aux, off_flow,
        warp_stride;
};

struct Params {
    SmallOffsets so;
    // static network
    int N, E, O, topo_ints, off_in_src, off_in_eid;
    const int32_t *topo, *row, *origins, *od_ptr, *od_dst;
    const float *t0_csc, *cap0_csc, *t0, *cap0;
    const void *od_q;  // REAL
    // small-network kernel only: TMA blob (node records + dense demand), per-slot link id / cost index
    const unsigned char *small_blob;
    const uint8_t *small_eid, *small_wpos;
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
    const int32_t *pedge_in;
    double alpha, beta, penalty;
};

template <typename REAL> __device__ __forceinline__ REAL rmul(REAL a, REAL b);
template <> __device__ __forceinline__ float rmul<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double rmul<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename REAL> __device__ __forceinline__ REAL radd(REAL a, REAL b);
template <> __device__ __forceinline__ float radd<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double radd<double>(double a, double b) { return __dadd_rn(a, b); }

// Everything a CTA needs, carved out of dynamic shared memory.
template <typename REAL>
struct Cta {
    uint64_t *bar;
    const int32_t *in_ptr, *in_src, *in_eid;
    int32_t *topo;
    double *w;
    REAL *aux, *flow;
    double *red;
    unsigned char *warp_base;
    size_t warp_stride;
    int N, E;

    __device__ void carve(unsigned char *smem, const Params &p, int warps)
    {
        SmemLayout L = make_layout(p.N, p.E, p.topo_ints, warps, (int)sizeof(REAL));
        bar = (uint64_t *)(smem + L.off_bar);
        topo = (int32_t *)(smem + L.off_topo);
        in_ptr = topo;
        in_src = topo + p.off_in_src;
        in_eid = topo + p.off_in_eid;
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
    __device__ int32_t *pk(int warp) const { return (int32_t *)(warp_base + warp_stride * warp + align16((size_t)N * 8)); }

    // Stage the static topology into shared memory with one TMA bulk copy (UBLKCP).
    __device__ void stage_topology(const Params &p)
    {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)p.topo_ints * 4u;
            mbar_expect_tx(bar, bytes);
            tma_bulk_g2s(topo, p.topo, bytes, bar);
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
__device__ __forceinline__ REAL link_cost(const Params &p, int k, REAL f, bool damaged)
{
    if (damaged) return (REAL)1e6;
    return Bpr<REAL>::eval(f, (REAL)__ldg(p.cap0_csc + k), (REAL)__ldg(p.t0_csc + k), (REAL)p.alpha, (REAL)p.beta,
                           p.beta == 4.0);
}

template <typename REAL>
__device__ void cta_costs_from_flow(Cta<REAL> &c, const Params &p, const uint8_t *dmg_b)
{
    for (int k = threadIdx.x; k < c.E; k += blockDim.x) {
        const int e = c.in_eid[k];
        c.w[k] = (double)link_cost<REAL>(p, k, c.flow[e], dmg_b[e] != 0);
    }
}

// ---------------------------------------------------------------------------------------------
// S2: one warp, one origin - label-correcting shortest paths + predecessor selection
// ---------------------------------------------------------------------------------------------
template <typename REAL>
__device__ int warp_sssp(const Cta<REAL> &c, int origin, double *dist, int32_t *pk)
{
    const int lane = threadIdx.x & 31, N = c.N;
    for (int v = lane; v < N; v += 32) dist[v] = (v == origin) ? 0.0 : B200_INF_F64;
    __syncwarp();
    int rounds = 0, capped = 0;
    for (;;) {
        bool changed = false;
        for (int v0 = 0; v0 < N; v0 += 32) {
            const int v = v0 + lane;
            double best = 0.0, dv = 0.0;
            if (v < N) {
                dv = dist[v];
                best = dv;
                const int k1 = c.in_ptr[v + 1];
                for (int k = c.in_ptr[v]; k < k1; ++k) {
                    const double cand = dist[c.in_src[k]] + c.w[k];
                    best = cand < best ? cand : best;
                }
            }
            __syncwarp();  // all reads of this chunk are done before anyone overwrites
            if (v < N && best < dv) {
                dist[v] = best;
                changed = true;
            }
            __syncwarp();
        }
        if (!__any_sync(0xffffffffu, changed)) break;
        if (++rounds > N + 1) { capped = 1; break; }
    }
    // predecessor: tight in-link with the smallest dist[u], then the largest u
    for (int v = lane; v < N; v += 32) {
        const double dv = dist[v];
        int bk = -1, bu = -1;
        double bdu = B200_INF_F64;
        if (v != origin && dv < B200_INF_F64) {
            const int k1 = c.in_ptr[v + 1];
            for (int k = c.in_ptr[v]; k < k1; ++k) {
                const int u = c.in_src[k];
                const double du = dist[u];
                if (du + c.w[k] == dv && (bk < 0 || du < bdu || (du == bdu && u > bu))) {
                    bk = k; bu = u; bdu = du;
                }
            }
        }
        pk[v] = bk;
    }
    __syncwarp();
    return capped;
}

// ---------------------------------------------------------------------------------------------
// S3: one warp, one origin - walk every destination back to the origin, scatter-add demand
// ---------------------------------------------------------------------------------------------
template <typename REAL>
__device__ double warp_aon_walk(Cta<REAL> &c, const Params &p, int slot, int origin, const int32_t *pk)
{
    const int lane = threadIdx.x & 31;
    const REAL *q = (const REAL *)p.od_q;
    double un = 0.0;
    const int j1 = __ldg(p.od_ptr + slot + 1);
    for (int j = __ldg(p.od_ptr + slot) + lane; j < j1; j += 32) {
        const int d = __ldg(p.od_dst + j);
        const REAL qj = q[j];
        if (d == origin || pk[d] < 0) { un += (double)qj; continue; }
        int cur = d, hops = 0;
        while (cur != origin && hops++ <= c.N) {
            const int k = pk[cur];
            atomicAdd(&c.aux[c.in_eid[k]], qj);
            cur = c.in_src[k];
        }
    }
    return un;
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

// ---------------------------------------------------------------------------------------------
// S4: step size, flow update, new costs (fused S1), optional gap; block-wide
// ---------------------------------------------------------------------------------------------
template <typename REAL>
__device__ double ls_derivative(Cta<REAL> &c, const Params &p, const uint8_t *dmg_b, double lam)
{
    double g = 0.0;
    for (int k = threadIdx.x; k < c.E; k += blockDim.x) {
        const int e = c.in_eid[k];
        const double f = (double)c.flow[e], d = (double)c.aux[e] - f;
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
            sat += (double)c.aux[e] * c.w[k];
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
                const REAL dir = c.aux[e] - c.flow[e], dp = dprev_b[e];
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
            REAL dir = c.aux[e] - c.flow[e];
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
            if (ls_derivative<REAL>(c, p, dmg_b, 1.0) <= 0.0) lam64 = 1.0;
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
            REAL f = radd<REAL>(rmul<REAL>(oml, c.flow[e]), rmul<REAL>(lam, c.aux[e]));
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
template <typename REAL>
__global__ void bpr_kernel(Params p)
{
    const REAL *flow = (const REAL *)p.flow;
    REAL *t = (REAL *)p.t;
    const long long total = (long long)p.B * p.E;
    const bool beta4 = p.beta == 4.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i % p.E);
        t[i] = p.dmg[i] ? (REAL)1e6
                        : Bpr<REAL>::eval(flow[i], (REAL)__ldg(p.cap0 + e), (REAL)__ldg(p.t0 + e), (REAL)p.alpha,
                                          (REAL)p.beta, beta4);
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
            double *dist = c.dist(warp);
            int32_t *pk = c.pk(warp);
            capped |= warp_sssp<REAL>(c, o, dist, pk);
            const size_t base = ((size_t)b * p.O + s) * p.N;
            for (int v = lane; v < p.N; v += 32) {
                const int k = pk[v];
                p.dist_out[base + v] = dist[v];
                p.pred_out[base + v] = k < 0 ? B200_NO_PRED : c.in_src[k];
                p.pedge_out[base + v] = k < 0 ? -1 : c.in_eid[k];
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
            int32_t *pk = c.pk(warp);
            const int32_t *pe = p.pedge_in + ((size_t)b * p.O + s) * p.N;
            for (int v = lane; v < p.N; v += 32) { const int e = pe[v]; pk[v] = e < 0 ? -1 : slot_of[e]; }
            __syncwarp();
            un += warp_aon_walk<REAL>(c, p, s, o, pk);
            __syncwarp();
        }
        un = warp_sum_f64(un);
        if (lane == 0 && un != 0.0) atomicAdd(&c.red[kMaxWarps * 2], un);
        __syncthreads();
        REAL *aux_b = (REAL *)p.aux + (size_t)b * p.E;
        for (int e = threadIdx.x; e < c.E; e += blockDim.x) aux_b[e] = c.aux[e];
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

// The fused path: K iterations of S1..S4 with every per-scenario array in shared memory.
template <typename REAL>
__global__ void __launch_bounds__(kMaxWarps * 32) assign_fused_kernel(Params p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    Cta<REAL> c;
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    c.carve(smem, p, W);
    c.stage_topology(p);
    double *un_s = &c.red[kMaxWarps * 2];
    int *flag_s = (int *)&c.red[kMaxWarps * 2 + 1];
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        if (p.active && !p.active[b]) continue;  // CTA-uniform
        const size_t off = (size_t)b * p.E;
        REAL *flow_b = (REAL *)p.flow + off;
        const uint8_t *dmg_b = p.dmg + off;
        REAL *dprev_b = p.dprev ? (REAL *)p.dprev + off : nullptr;
        __syncthreads();
        for (int e = threadIdx.x; e < c.E; e += blockDim.x) c.flow[e] = flow_b[e];
        if (threadIdx.x == 0) { flag_s[0] = 0; flag_s[1] = 0; }
        __syncthreads();
        cta_costs_from_flow<REAL>(c, p, dmg_b);
        for (int it = 0; it < p.K; ++it) {
            for (int e = threadIdx.x; e < c.E; e += blockDim.x) c.aux[e] = 0;
            if (threadIdx.x == 0) *un_s = 0.0;
            __syncthreads();
            double un = 0.0;
            int capped = 0;
            for (int s = warp; s < p.O; s += W) {
                const int o = __ldg(p.origins + s);
                capped |= warp_sssp<REAL>(c, o, c.dist(warp), c.pk(warp));
                un += warp_aon_walk<REAL>(c, p, s, o, c.pk(warp));
                __syncwarp();
            }
            un = warp_sum_f64(un);
            if (lane == 0 && un != 0.0) atomicAdd(un_s, un);
            if (capped && lane == 0) flag_s[1] = 1;
            __syncthreads();
            int nan_seen = 0;
            cta_update<REAL>(c, p, dmg_b, it, dprev_b, it == p.K - 1, p.gap + b, &nan_seen);
            if (nan_seen) flag_s[0] = 1;
            __syncthreads();
        }
        const double un_final = *un_s;
        cta_tstt<REAL>(c, p, c.aux, un_final, p.tstt + b);
        for (int e = threadIdx.x; e < c.E; e += blockDim.x) flow_b[e] = c.flow[e];
        if (threadIdx.x == 0) {
            p.unassigned[b] = un_final;
            p.status[b] = (flag_s[0] ? B200_ST_NAN_FLOW : 0) | (un_final > 0.0 ? B200_ST_UNASSIGNED : 0) |
                          (flag_s[1] ? B200_ST_SSSP_CAP : 0);
        }
    }
}

#include "b200_assign_small.cuh"

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
Params base_params(const b200_graph *g, int B, int prec, double alpha, double beta)
{
    Params p;
    memset(&p, 0, sizeof(p));
    p.N = g->N; p.E = g->E; p.O = g->O; p.topo_ints = g->topo_ints;
    p.off_in_src = g->off_in_src; p.off_in_eid = g->off_in_eid;
    p.topo = g->topo; p.row = g->row; p.origins = g->origins; p.od_ptr = g->od_ptr; p.od_dst = g->od_dst;
    p.t0_csc = g->t0_csc; p.cap0_csc = g->cap0_csc; p.t0 = g->t0; p.cap0 = g->cap0;
    p.od_q = prec == B200_F32 ? (const void *)g->od_q32 : (const void *)g->od_q64;
    p.small_blob = g->small_intq ? g->small_blob_i32 : (prec == B200_F32 ? g->small_blob32 : g->small_blob64);
    p.small_eid = g->small_eid; p.small_wpos = g->small_wpos;
    p.total_demand = g->total_demand;
    p.B = B; p.alpha = alpha; p.beta = beta;
    return p;
}

template <typename K>
int launch_cta_kernel(K kernel, const b200_graph *g, int prec, const Params &p, cudaStream_t st)
{
    const int warps = prec == B200_F32 ? g->warps_f32 : g->warps_f64;
    const size_t smem = prec == B200_F32 ? g->smem_f32 : g->smem_f64;
    B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem));
    if (per_sm < 1) { b200_set_error("kernel does not fit on an SM (smem %zu)", smem); return B200_EUNSUPPORTED; }
    // Small networks: one CTA per scenario, the hardware block scheduler balances the tail and
    // re-staging a sub-8-KB topology is free.  Large networks: persistent CTAs, topology staged once.
    const int grid = ((size_t)g->topo_ints * 4 <= 8192) ? p.B : std::min(p.B, per_sm * g->num_sms);
    kernel<<<grid, warps * 32, smem, st>>>(p);
    B200_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

// Sweep order for the small-network kernel: Gauss-Seidel converges fastest when the node order
// follows the direction paths travel.  Candidates: nodes sorted by free-flow distance from a root,
// for every root; the winner needs the fewest alternating sweeps on the free-flow costs with a
// few deterministic damage patterns (all origins at once, exactly what the kernel does).
static std::vector<int> small_sweep_order(int N, int E, const int32_t *in_ptr, const int32_t *in_src,
                                          const int32_t *in_eid, const float *t0, const int32_t *origins, int O)
{
    auto count_sweeps = [&](const std::vector<int> &order, const std::vector<double> &w) {
        std::vector<double> d((size_t)O * N, INFINITY);
        for (int s = 0; s < O; ++s) d[(size_t)s * N + origins[s]] = 0.0;
        int sweeps = 0;
        for (;; ++sweeps) {
            bool changed = false;
            for (int n = 0; n < N; ++n) {
                const int v = order[(sweeps & 1) ? N - 1 - n : n];
                for (int k = in_ptr[v]; k < in_ptr[v + 1]; ++k)
                    for (int s = 0; s < O; ++s) {
                        const double c = d[(size_t)s * N + in_src[k]] + w[k];
                        if (c < d[(size_t)s * N + v]) { d[(size_t)s * N + v] = c; changed = true; }
                    }
            }
            if (!changed || sweeps > 4 * N) break;
        }
        return sweeps + 1;
    };
    std::vector<std::vector<double>> patterns;
    unsigned long long lcg = 0x9E3779B97F4A7C15ull;
    for (int pat = 0; pat < 6; ++pat) {
        std::vector<double> w(E);
        for (int k = 0; k < E; ++k) {
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            const bool dmg = pat > 0 && ((lcg >> 33) % 10) < 3;
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            const double jitter = 1.0 + 0.5 * (double)((lcg >> 33) % 1000) / 1000.0;
            w[k] = dmg ? 1e6 : (double)t0[in_eid[k]] * (pat > 0 ? jitter : 1.0);
        }
        patterns.push_back(w);
    }
    std::vector<int> best_order(N);
    for (int v = 0; v < N; ++v) best_order[v] = v;
    long best_cost = 0;
    for (auto &w : patterns) best_cost += count_sweeps(best_order, w);
    for (int root = 0; root < N; ++root) {
        // undirected free-flow distances from root (O(N^2) Dijkstra), order = ascending distance
        std::vector<double> dist(N, INFINITY);
        std::vector<char> done(N, 0);
        dist[root] = 0.0;
        for (int it = 0; it < N; ++it) {
            int u = -1;
            for (int v = 0; v < N; ++v) if (!done[v] && (u < 0 || dist[v] < dist[u])) u = v;
            if (u < 0) break;
            done[u] = 1;
            for (int v = 0; v < N; ++v)
                for (int k = in_ptr[v]; k < in_ptr[v + 1]; ++k) {
                    const int a = in_src[k];
                    const double wk = (double)t0[in_eid[k]];
                    if (a == u && dist[u] + wk < dist[v]) dist[v] = dist[u] + wk;
                    if (v == u && dist[u] + wk < dist[a]) dist[a] = dist[u] + wk;
                }
        }
        std::vector<int> order(N);
        for (int v = 0; v < N; ++v) order[v] = v;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return dist[a] < dist[b]; });
        long cost = 0;
        for (auto &w : patterns) cost += count_sweeps(order, w);
        if (cost < best_cost) { best_cost = cost; best_order = order; }
    }
    return best_order;
}

// Small-network eligibility + static tables (called by b200_graph_create with the host CSC and
// the OD list).  Builds TMA-able blobs: node records in sweep order | dense demand Q[node][slot]
// as float, double and (when every demand is an integer) int32.
int b200_small_prepare(b200_graph *g, const int32_t *in_ptr, const int32_t *in_src, const int32_t *in_eid,
                       const float *t0, int num_od, const int32_t *od_o, const int32_t *od_d, const double *od_q,
                       const int32_t *origin_slot)
{
    g->small_ok = 0;
    const int N = g->N, E = g->E, O = g->O;
    if (N > kSmallMaxN || E > kSmallMaxE || O > kSmallMaxO || O < 1 || g->max_in_deg > kSmallMaxD) return 0;
    const int D = g->max_in_deg < 3 ? 3 : g->max_in_deg, DP = (D + 1) & ~1;
    if (N * DP > 255) return 0;  // padded cost index is a byte
    if (make_small_layout(N, E, O, D, 8, 8).total > kSmemLimit) return 0;
    bool intq = g->total_demand < 16777216.0;  // 2^24: fp32 link sums stay exact
    for (int k = 0; k < num_od && intq; ++k) intq = od_q[k] >= 0.0 && od_q[k] == floor(od_q[k]);
    const std::vector<int> order = small_sweep_order(N, E, in_ptr, in_src, in_eid, t0, g->origins_host, O);
    std::vector<uint8_t> eid(E), wpos(E);
    for (int variant = 0; variant < 3; ++variant) {   // 0 float Q, 1 double Q, 2 int32 Q
        if (variant == 2 && !intq) break;
        const int qb = variant == 1 ? 8 : 4;
        const SmallLayout L = make_small_layout(N, E, O, D, qb, qb);
        std::vector<unsigned char> blob(L.blob_bytes, 0);
        uint16_t *rec1 = (uint16_t *)blob.data();
        uint8_t *rec2 = blob.data() + L.off_rec2;
        memset(rec2, 255, (size_t)N * 16);
        for (int i = 0; i < N; ++i) {
            const int v = order[i], deg = in_ptr[v + 1] - in_ptr[v];
            rec1[i * 8 + 0] = (uint16_t)(v * O * 8) | (uint16_t)(deg > 3 ? 0x8000 : 0);
            for (int j = 0; j < 7; ++j) rec1[i * 8 + 1 + j] = (uint16_t)(v * O * 8);  // padding: the node itself
            rec2[i * 16 + 7] = (uint8_t)v;
            for (int k = in_ptr[v], j = 0; k < in_ptr[v + 1]; ++k, ++j) {
                rec1[i * 8 + 1 + j] = (uint16_t)(in_src[k] * O * 8);
                rec2[i * 16 + j] = (uint8_t)in_src[k];
                rec2[i * 16 + 8 + j] = (uint8_t)in_eid[k];
                eid[k] = (uint8_t)in_eid[k];
                wpos[k] = (uint8_t)(i * DP + j);
            }
        }
        for (int k = 0; k < num_od; ++k) {
            const size_t idx = (size_t)od_d[k] * O + origin_slot[od_o[k]];
            if (variant == 0) ((float *)(blob.data() + L.off_q))[idx] += (float)od_q[k];
            else if (variant == 1) ((double *)(blob.data() + L.off_q))[idx] += od_q[k];
            else ((int32_t *)(blob.data() + L.off_q))[idx] += (int32_t)od_q[k];
        }
        unsigned char **dst = variant == 0 ? &g->small_blob32 : variant == 1 ? &g->small_blob64 : &g->small_blob_i32;
        if (cudaMalloc((void **)dst, blob.size()) != cudaSuccess) return 0;
        if (cudaMemcpy(*dst, blob.data(), blob.size(), cudaMemcpyHostToDevice) != cudaSuccess) return 0;
    }
    if (cudaMalloc((void **)&g->small_eid, E) != cudaSuccess || cudaMalloc((void **)&g->small_wpos, E) != cudaSuccess) return 0;
    cudaMemcpy(g->small_eid, eid.data(), E, cudaMemcpyHostToDevice);
    cudaMemcpy(g->small_wpos, wpos.data(), E, cudaMemcpyHostToDevice);
    g->small_D = D;
    g->small_intq = intq ? 1 : 0;
    g->small_ok = 1;
    return 1;
}

// Warps per CTA: enough to spread the origins evenly, bounded by shared memory.
int b200_pick_warps(const b200_graph *g, int prec, size_t *smem_out)
{
    const int rb = prec == B200_F32 ? 4 : 8;
    int best = -1;
    double best_score = -1.0;
    for (int w = kMinWarps; w <= kMaxWarps; ++w) {
        SmemLayout L = make_layout(g->N, g->E, g->topo_ints, w, rb);
        if (L.total > kSmemLimit) break;
        const int O = std::max(g->O, 1);
        const double eff = (double)O / (double)(((O + w - 1) / w) * w);
        const double score = eff - 0.004 * (w > 8 ? w - 8 : 8 - w);  // prefer ~8 warps on ties
        if (score > best_score) { best_score = score; best = w; }
    }
    if (best > 0 && smem_out) *smem_out = make_layout(g->N, g->E, g->topo_ints, best, rb).total;
    return best;
}

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
    const int grid = (int)std::min<long long>((total + threads - 1) / threads, (long long)g->num_sms * 8);
    if (prec == B200_F32) bpr_kernel<float><<<grid, threads, 0, (cudaStream_t)stream>>>(p);
    else bpr_kernel<double><<<grid, threads, 0, (cudaStream_t)stream>>>(p);
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
    if (B == 0) return 0;  // empty batch: nothing to do, pointers may be null
    B200_REQUIRE(dmg && flow && tstt && gap && unassigned && status, "null array argument");
    B200_REQUIRE(method != B200_CFW || dprev, "cfw needs the dprev scratch array");
    Params p = base_params(g, B, prec, alpha, beta);
    p.K = K; p.method = method; p.dmg = dmg; p.active = active; p.flow = flow; p.dprev = dprev; p.tstt = tstt; p.gap = gap;
    p.unassigned = unassigned; p.status = status; p.penalty = penalty; p.ls_iters = ls_iters;
    int rc;
    if (g->small_ok && !g->force_general) {
        // Sioux-Falls class: one warp per scenario, lane = origin
        rc = prec == B200_F32 ? launch_small<float>(g, p, (cudaStream_t)stream)
                              : launch_small<double>(g, p, (cudaStream_t)stream);
    } else {
        rc = prec == B200_F32 ? launch_cta_kernel(assign_fused_kernel<float>, g, prec, p, (cudaStream_t)stream)
                              : launch_cta_kernel(assign_fused_kernel<double>, g, prec, p, (cudaStream_t)stream);
    }
    if (rc == 0) const_cast<b200_graph *>(g)->last_launches = 1;
    return rc;
}
