// Small-network fused kernel (Sioux-Falls class: N <= 64, E <= 255, O <= 32, in-degree <= 7).
// Included by b200_assign.cu inside its anonymous namespace (uses Params, link_cost, rmul/radd,
// warp_pairwise, align16).
//
// ONE WARP owns one scenario for all K iterations; LANE = ORIGIN.  Per-warp shared memory holds
// column-per-lane tables [node][lane] (a lane only ever touches its own column, bank = lane, so
// the tables are conflict-free and need no intra-warp synchronisation):
//   dist  double   label-correcting distances, relaxed by Gauss-Seidel sweeps over the nodes in a
//                  host-chosen order (alternating direction) until a full sweep changes nothing:
//                  the same least fixed point as Dijkstra (see b200_assign.cu header).  A node
//                  has 3 or D in-link slots (padding slots point at the node itself with cost
//                  +inf), so the relaxation is branch-free and unrolled;
//   pu,pe uint8    tree predecessor node / link per (node, origin), scipy tie rule;
//   sacc, cnt, stk (aliased onto dist, which is dead after the selection): subtree demand sums
//                  in Kahn order - all-or-nothing loading handles every tree link once instead
//                  of once per OD pair and path link.
// Link loads: when all demands are integers (< 2^24 in total, so the reference's fp32 sums are
// exact too) subtree sums are int32 and the per-link total over the 32 origin lanes is ONE
// redux.sync per (node, in-link slot) - no atomics at all.  Otherwise the subtree sums are
// scatter-added with shared-memory float atomics (a CAS loop on this hardware).
// Lanes meet only in the per-scenario link arrays (wpad / aux / flow), separated by __syncwarp.
// The static tables (node records + dense demand Q[node][lane]) are one device blob, staged into
// shared memory once per CTA by a TMA bulk copy.
#pragma once

#ifndef B200_SMALL_MINB
#define B200_SMALL_MINB 4
#endif
constexpr int kSmallMaxN = 64, kSmallMaxE = 255, kSmallMaxO = 32, kSmallWarps = 4, kSmallMaxD = 7;

// blob: uint4 rec1[N] | uint4 rec2[N] | Q[N][O] (float, double or int32), records in SWEEP ORDER
//   rec1 halfwords: [0] = v*O*8 | class (bit 15: 1 = node needs all D slots, 0 = 3 slots are enough)
//                   [1+j] = u_j*O*8, j < 7 (padding slots: v*O*8)
//   rec2 bytes:     [j] = u_j (255 = padding), j < 7; [7] = v;  [8+j] = link id of slot j (255 = padding)
struct SmallLayout;
struct SmallLayout {
    size_t blob_bytes, off_rec2, off_q;                 // CTA-shared part (after the 16-byte mbarrier)
    size_t off_dist, off_pu, off_pe, off_w, off_aux, off_flow, warp_stride, total;
    size_t off_sacc, off_cnt, off_stk;                  // inside the dist region
};

__host__ __device__ inline SmallLayout make_small_layout(int N, int E, int O, int D, int real_bytes, int q_bytes)
{
    SmallLayout L;
    const int DP = (D + 1) & ~1;
    L.off_rec2 = (size_t)N * 16;
    L.off_q = (size_t)N * 32;
    L.blob_bytes = L.off_q + align16((size_t)N * O * q_bytes);
    size_t o = 0;
    const size_t col = (size_t)N * O;
    L.off_dist = o;
    L.off_sacc = o;                                     // aliases dist
    L.off_cnt = o + align16(col * q_bytes);
    L.off_stk = L.off_cnt + align16(col);
    {
        const size_t need = align16(col * q_bytes) + 2 * align16(col);
        const size_t dist_bytes = align16(col * 8);
        o += need > dist_bytes ? need : dist_bytes;
    }
    L.off_pu = o;   o += align16(col);
    L.off_pe = o;   o += align16(col);
    L.off_w = o;    o += align16((size_t)N * DP * 8);
    L.off_aux = o;  o += align16((size_t)E * real_bytes);
    L.off_flow = o; o += align16((size_t)E * real_bytes);
    L.warp_stride = o;
    L.total = 16 + L.blob_bytes + o * kSmallWarps;
    return L;
}

__device__ __forceinline__ int rec_half(const uint4 &r, int i)  // i-th halfword, i compile-time
{
    const unsigned wd = i < 2 ? r.x : i < 4 ? r.y : i < 6 ? r.z : r.w;
    return (i & 1) ? (int)(wd >> 16) : (int)(wd & 0xffffu);
}
__device__ __forceinline__ int rec_byte(const uint4 &r, int i)  // i-th byte, i compile-time
{
    const unsigned wd = i < 4 ? r.x : i < 8 ? r.y : i < 12 ? r.z : r.w;
    return (int)((wd >> (8 * (i & 3))) & 0xffu);
}

// One Gauss-Seidel relaxation of the node described by r.  Slots 0..2 are always evaluated, slots
// 3..D-1 only for nodes of the wide class (predicated loads: no branch, no LSU wavefront when off).
template <int D>
__device__ __forceinline__ bool relax_node(const uint4 &r, unsigned char *dcol, const double *wv)
{
    const bool wide = (r.x & 0x8000u) != 0;
    double *pv = (double *)(dcol + (r.x & 0x3fffu));
    const double dv = *pv;
    double c[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        c[j] = B200_INF_F64;
        if (j < 3 || wide) c[j] = *(const double *)(dcol + rec_half(r, 1 + j));
    }
#pragma unroll
    for (int j = 0; j < (D + 1) / 2; ++j) {
        double2 ww = make_double2(0.0, 0.0);
        if (j < 2 || wide) ww = ((const double2 *)wv)[j];
        c[2 * j] += ww.x;
        if (2 * j + 1 < D) c[2 * j + 1] += ww.y;
    }
    double best = c[0];
#pragma unroll
    for (int j = 1; j < D; ++j) best = c[j] < best ? c[j] : best;  // no NaNs here: 3 SASS ops, fmin() is 7
    const bool better = best < dv;
    if (better) *pv = best;
    return better;
}

template <int D>
__device__ __forceinline__ void select_pred(const uint4 &r, const uint4 &r2, const unsigned char *dcol, const double *wv,
                                            int &bu, int &be)
{
    const bool wide = (r.x & 0x8000u) != 0;
    const double dv = *(const double *)(dcol + (r.x & 0x3fffu));
    double bdu = B200_INF_F64;
    const bool reach = dv < B200_INF_F64;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        if (j < 3 || wide) {
            const double du = *(const double *)(dcol + rec_half(r, 1 + j));
            const double2 ww2 = ((const double2 *)wv)[j >> 1];
            const double ww = (j & 1) ? ww2.y : ww2.x;
            // slots are stored by ascending tail node, so "<=" keeps the largest u on ties
            if (du + ww == dv && du <= bdu && reach) { bu = rec_byte(r2, j); be = rec_byte(r2, 8 + j); bdu = du; }
        }
    }
}

// REAL: link-state type; D: in-link slots of the widest node; INTQ: integer demand path
template <typename REAL, int D, bool INTQ>
__global__ void __launch_bounds__(kSmallWarps * 32, B200_SMALL_MINB) assign_small_kernel(const Params p)
{
    constexpr int DP = (D + 1) & ~1;
    using QT = typename std::conditional<INTQ, int, REAL>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.N, E = p.E, O = p.O;
    const SmallOffsets &L = p.so;      // layout computed on the host, read from the constant bank
    uint64_t *bar = (uint64_t *)smem;
    unsigned char *blob = smem + 16;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)L.blob_bytes);
        tma_bulk_g2s(blob, p.small_blob, (uint32_t)L.blob_bytes, bar);
    }
    mbar_wait(bar, 0);
    const uint4 *rec1 = (const uint4 *)blob;
    const uint4 *rec2 = (const uint4 *)(blob + L.off_rec2);

    const int b = blockIdx.x * kSmallWarps + warp;
    if (b >= p.B) return;
    if (p.active && !p.active[b]) return;

    unsigned char *wb = smem + 16 + L.blob_bytes + L.warp_stride * warp;
    const bool on = lane < O;           // lanes >= O shadow lane O-1 (same addresses, same values)
    const int ol = on ? lane : O - 1;
    unsigned char *dcol = wb + L.off_dist + (size_t)ol * 8;  // this lane's column of dist
    QT *sacc = (QT *)(wb + L.off_sacc) + ol;
    uint8_t *cnt = wb + L.off_cnt + ol, *stk = wb + L.off_stk + ol;
    uint8_t *pu = wb + L.off_pu + ol, *pe = wb + L.off_pe + ol;
    double *wpad = (double *)(wb + L.off_w);
    REAL *aux = (REAL *)(wb + L.off_aux), *flow = (REAL *)(wb + L.off_flow);
    const QT *Qc = (const QT *)(blob + L.off_q) + ol;

    const size_t off = (size_t)b * E;
    REAL *flow_b = (REAL *)p.flow + off;
    const uint8_t *dmg_b = p.dmg + off;
    REAL *dprev_b = p.dprev ? (REAL *)p.dprev + off : nullptr;
    const int orig = __ldg(p.origins + ol);
    const int rowb = O * 8;             // bytes per dist row
    const unsigned FULL = 0xffffffffu;

    // per-lane link slots k = lane, lane+32, ..: link id, padded cost index, damage bit in registers
    int eidr[8], wposr[8];
    unsigned dmg_bits = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = lane + 32 * j;
        eidr[j] = k < E ? __ldg(p.small_eid + k) : 0;
        wposr[j] = k < E ? __ldg(p.small_wpos + k) : 0;
        if (k < E && dmg_b[eidr[j]]) dmg_bits |= 1u << j;
    }
    auto for_slots = [&](auto fn) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (32 * j >= E) break;
            const int k = lane + 32 * j;
            if (k < E) fn(k, eidr[j], wposr[j], (dmg_bits >> j) & 1u);
        }
    };
    for (int i = lane; i < N * DP; i += 32) wpad[i] = B200_INF_F64;  // padding slots stay +inf
    for (int e = lane; e < E; e += 32) flow[e] = flow_b[e];
    __syncwarp();
    for_slots([&](int k, int e, int wp, unsigned dm) { wpad[wp] = (double)link_cost<REAL>(p, k, flow[e], dm != 0); });
    __syncwarp();

    double un_total = 0.0, gap = B200_INF_F64;
    int capped = 0, nan_seen = 0;
    for (int it = 0; it < p.K; ++it) {
        if (!INTQ)
            for (int e = lane; e < E; e += 32) aux[e] = 0;
        // ---- S2: all origins at once, Gauss-Seidel sweeps in alternating order ------------------
        for (int v = 0; v < N; ++v) *(double *)(dcol + v * rowb) = (v == orig) ? 0.0 : B200_INF_F64;
        for (int sweep = 0;; ++sweep) {
            bool changed = false;
            const int dir = (sweep & 1) ? -1 : 1;
            const uint4 *rp = rec1 + ((sweep & 1) ? N - 1 : 0);
            const double *wv = wpad + ((sweep & 1) ? (N - 1) * DP : 0);
            for (int n = N; n > 0; --n, rp += dir, wv += dir * DP) changed |= relax_node<D>(*rp, dcol, wv);
            if (!__any_sync(FULL, changed)) break;
            if (sweep > N + 1) { capped = 1; break; }
        }
        // ---- predecessor selection: tight in-link with the smallest dist[u], then the largest u --
        for (int i = 0; i < N; ++i) {
            const uint4 r = rec1[i], r2 = rec2[i];
            const double *wv = wpad + i * DP;
            int bu = 255, be = 255;
            select_pred<D>(r, r2, dcol, wv, bu, be);
            const int v = rec_byte(r2, 7);
            pu[v * O] = (uint8_t)bu;
            pe[v * O] = (uint8_t)be;
        }
        // ---- S3: subtree sums in Kahn order (dist is dead from here) -----------------------------
        for (int v = 0; v < N; ++v) { cnt[v * O] = 0; sacc[v * O] = 0; }
        double un = 0.0;
        for (int v = 0; v < N; ++v) {
            const int u = pu[v * O];
            if (u != 255) cnt[u * O] += 1;
            else un += (double)Qc[v * O];   // unreachable destination, or demand to the origin itself (reference :496)
        }
        int top = 0;
        for (int v = 0; v < N; ++v)
            if (v != orig && pu[v * O] != 255 && cnt[v * O] == 0) stk[(top++) * O] = (uint8_t)v;
        if (!INTQ) __syncwarp();  // aux zeroing visible before the scatter-adds
        while (__any_sync(FULL, top > 0)) {
            if (top > 0) {
                const int v = stk[(--top) * O];
                QT s;
                if (INTQ) s = sacc[v * O] + Qc[v * O];
                else s = (QT)radd<REAL>((REAL)sacc[v * O], (REAL)Qc[v * O]);
                const int u = pu[v * O];
                if (INTQ) sacc[v * O] = s;                       // final subtree sum of v, reduced below
                else if (on && s != (QT)0) atomicAdd(&aux[pe[v * O]], (REAL)s);
                if (INTQ) sacc[u * O] += s;
                else sacc[u * O] = (QT)radd<REAL>((REAL)sacc[u * O], (REAL)s);
                const int cc = cnt[u * O] - 1;
                cnt[u * O] = (uint8_t)cc;
                if (cc == 0 && u != orig) stk[(top++) * O] = (uint8_t)u;
            }
        }
        if (INTQ) {
            // link load = sum over origin lanes of the subtree sum of the link's head node, for the
            // lanes whose tree uses this link: one redux.sync per (node, slot), a plain store per link
            for (int i = 0; i < N; ++i) {
                const uint4 r2 = rec2[i];
                const int v = rec_byte(r2, 7);
                const int myE = on ? (int)pe[v * O] : 255;
                const int myS = (myE != 255) ? (int)sacc[v * O] : 0;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const int ej = rec_byte(r2, 8 + j);
                    if (ej != 255) {                              // warp-uniform
                        const int tot = __reduce_add_sync(FULL, myE == ej ? myS : 0);
                        if (lane == 0) aux[ej] = (REAL)tot;
                    }
                }
            }
        }
        if (!on) un = 0.0;
        un_total = warp_sum_f64(un);
        __syncwarp();
        // ---- S4: gap (last iteration), step size, flow update, new costs ------------------------
        if (it == p.K - 1) {
            double sft = 0.0, sat = 0.0;
            for_slots([&](int, int e, int wp, unsigned) {
                sft += (double)flow[e] * wpad[wp];
                sat += (double)aux[e] * wpad[wp];
            });
            sft = warp_sum_f64(sft);
            sat = warp_sum_f64(sat);
            gap = sft > 0.0 ? (sft - sat) / sft : B200_INF_F64;
        }
        if (p.method == B200_CFW) {
            REAL bb = 0;
            if (it > 0) {
                double num = 0.0, den = 0.0;
                for (int e = lane; e < E; e += 32) {
                    const REAL dir = aux[e] - flow[e], dp = dprev_b[e];
                    num += (double)dir * (double)(dir - dp);
                    den += (double)dp * (double)dp;
                }
                num = warp_sum_f64(num);
                den = warp_sum_f64(den);
                double bq = num / (den + 1e-12);
                if (!(bq > 0.0)) bq = 0.0;
                bb = (REAL)bq;
            }
            const REAL lam = (REAL)(2.0 / (it + 2.0));
            for_slots([&](int k, int e, int wp, unsigned dm) {
                REAL dir = aux[e] - flow[e];
                if (it > 0) dir = radd<REAL>(dir, rmul<REAL>(bb, dprev_b[e]));
                REAL f = radd<REAL>(flow[e], rmul<REAL>(lam, dir));
                f = f > (REAL)0 ? f : (REAL)0;
                dprev_b[e] = dir;
                flow[e] = f;
                wpad[wp] = (double)link_cost<REAL>(p, k, f, dm != 0);
            });
        } else {
            double lam64;
            if (p.method == B200_MSA) lam64 = 1.0 / (it + 1.0);
            else if (p.method == B200_FW) lam64 = 2.0 / (it + 2.0);
            else {
                auto deriv = [&](double lam) {
                    double g = 0.0;
                    for_slots([&](int k, int e, int, unsigned dm) {
                        const double f = (double)flow[e], d = (double)aux[e] - f;
                        const double te = dm ? 1e6
                                             : Bpr<double>::eval(f + lam * d, (double)__ldg(p.cap0_csc + k),
                                                                 (double)__ldg(p.t0_csc + k), p.alpha, p.beta,
                                                                 p.beta == 4.0);
                        g += d * te;
                    });
                    return warp_sum_f64(g);
                };
                // first step = the all-or-nothing load itself (as fw's 2/(it+2)): flow = 0 is not feasible
                if (it == 0 || deriv(1.0) <= 0.0) lam64 = 1.0;
                else {
                    double lo = 0.0, hi = 1.0;
                    for (int i = 0; i < p.ls_iters; ++i) {
                        const double mid = 0.5 * (lo + hi);
                        if (deriv(mid) > 0.0) hi = mid; else lo = mid;
                    }
                    lam64 = 0.5 * (lo + hi);
                }
            }
            const REAL lam = (REAL)lam64, oml = (REAL)(1.0 - lam64);
            for_slots([&](int k, int e, int wp, unsigned dm) {
                REAL f = radd<REAL>(rmul<REAL>(oml, flow[e]), rmul<REAL>(lam, aux[e]));
                if (f != f) { f = 0; nan_seen = 1; }
                flow[e] = f;
                wpad[wp] = (double)link_cost<REAL>(p, k, f, dm != 0);
            });
        }
        __syncwarp();
    }
    // ---- TSTT (numpy pairwise order), outputs ---------------------------------------------------
    for_slots([&](int, int e, int wp, unsigned) { aux[e] = rmul<REAL>(flow[e], (REAL)wpad[wp]); });
    __syncwarp();
    const REAL base = radd<REAL>((REAL)0, warp_pairwise<REAL>(aux, E));
    for (int e = lane; e < E; e += 32) flow_b[e] = flow[e];
    nan_seen = __any_sync(FULL, nan_seen);
    if (lane == 0) {
        const double total = p.total_demand > 1.0 ? p.total_demand : 1.0;
        double v = (double)base / total;
        if (un_total > 0.0) v += p.penalty * (un_total / total);
        p.tstt[b] = v;
        p.gap[b] = gap;
        p.unassigned[b] = un_total;
        p.status[b] = (nan_seen ? B200_ST_NAN_FLOW : 0) | (un_total > 0.0 ? B200_ST_UNASSIGNED : 0) |
                      (capped ? B200_ST_SSSP_CAP : 0);
    }
}

// host: launch the instantiation with the smallest slot count D >= max in-degree
template <typename REAL, bool INTQ>
int launch_small_q(const b200_graph *g, const Params &p, cudaStream_t st)
{
    const int D = g->small_D;
    const SmallLayout L = make_small_layout(g->N, g->E, g->O, D, (int)sizeof(REAL), INTQ ? 4 : (int)sizeof(REAL));
    const int grid = (p.B + kSmallWarps - 1) / kSmallWarps;
    void (*kern)(const Params) = nullptr;
    switch (D) {
        case 3: kern = assign_small_kernel<REAL, 3, INTQ>; break;
        case 4: kern = assign_small_kernel<REAL, 4, INTQ>; break;
        case 5: kern = assign_small_kernel<REAL, 5, INTQ>; break;
        case 6: kern = assign_small_kernel<REAL, 6, INTQ>; break;
        default: kern = assign_small_kernel<REAL, 7, INTQ>; break;
    }
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    Params q = p;
    q.so.blob_bytes = (int)L.blob_bytes; q.so.off_rec2 = (int)L.off_rec2; q.so.off_q = (int)L.off_q;
    q.so.off_dist = (int)L.off_dist; q.so.off_sacc = (int)L.off_sacc; q.so.off_cnt = (int)L.off_cnt;
    q.so.off_stk = (int)L.off_stk; q.so.off_pu = (int)L.off_pu; q.so.off_pe = (int)L.off_pe; q.so.off_w = (int)L.off_w;
    q.so.off_aux = (int)L.off_aux; q.so.off_flow = (int)L.off_flow; q.so.warp_stride = (int)L.warp_stride;
    kern<<<grid, kSmallWarps * 32, L.total, st>>>(q);
    B200_CUDA(cudaGetLastError());
    return 0;
}

template <typename REAL>
int launch_small(const b200_graph *g, const Params &p, cudaStream_t st)
{
    return g->small_intq ? launch_small_q<REAL, true>(g, p, st) : launch_small_q<REAL, false>(g, p, st);
}
