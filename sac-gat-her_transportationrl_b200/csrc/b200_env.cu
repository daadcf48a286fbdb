// Env-step glue on the device: state features incl. Brandes betweenness (S5), action
// application, reward / done.  Reference: src/env/repair_env.py:207-237 (step), :244-294
// (reward, goal), :751-819 (get_state).
#include "b200_common.cuh"

#include <algorithm>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int kFeatWarps = 8;

struct FeatParams {
    int N, E, B;
    const int32_t *topo;  // in_ptr | in_src | in_eid
    int off_in_src, off_in_eid;
    const int32_t *out_ptr, *out_dst, *out_eid;
    const float *t0, *cap0;
    const void *flow;
    const uint8_t *dmg, *goal;
    const double *tstt;
    double total_demand, capacity_damage, max_t0, max_cap0;
    float *node_feat, *edge_feat, *bw_raw;
};

// One CTA per scenario.  Betweenness (networkx.betweenness_centrality(subgraph,
// normalized=True) on the edge-induced subgraph of undamaged links, reference :753-761): one warp
// per BFS source, level-synchronous pull BFS over active in-links (sigma), then level-by-level
// dependency accumulation over active out-links.  Per-warp partial sums are combined in a fixed
// order, so the result is deterministic.
//
// smem: uint8 act[E] | uint8 insub[N] | per warp { int depth[N]; double sigma[N], delta[N], bw[N] }
template <typename REAL>
__global__ void __launch_bounds__(kFeatWarps * 32) features_kernel(FeatParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = p.N, E = p.E, W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int32_t *in_ptr = p.topo, *in_src = p.topo + p.off_in_src, *in_eid = p.topo + p.off_in_eid;
    size_t o = 0;
    uint8_t *act = smem + o;           o += (size_t)((E + 15) & ~15);
    uint8_t *insub = smem + o;         o += (size_t)((N + 15) & ~15);
    double *red = (double *)(smem + o); o += 64 * 8;
    const size_t wstride = (size_t)N * (8 * 3) + (size_t)((N * 4 + 15) & ~15);
    unsigned char *wb = smem + o + wstride * warp;
    double *sigma = (double *)wb, *delta = sigma + N, *bw = delta + N;
    int *depth = (int *)(bw + N);

    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const size_t off = (size_t)b * E;
        const uint8_t *dmg_b = p.dmg + off, *goal_b = p.goal + off;
        const REAL *flow_b = (const REAL *)p.flow + off;
        __syncthreads();
        for (int v = threadIdx.x; v < N; v += blockDim.x) insub[v] = 0;
        __syncthreads();
        for (int e = threadIdx.x; e < E; e += blockDim.x) act[e] = dmg_b[e] ? 0 : 1;
        for (int k = threadIdx.x; k < E; k += blockDim.x) {
            const int e = in_eid[k];
            if (!dmg_b[e]) { insub[in_src[k]] = 1; }
        }
        for (int v = threadIdx.x; v < N; v += blockDim.x) {
            bool any_in = false;
            for (int k = in_ptr[v]; k < in_ptr[v + 1]; ++k) any_in |= (dmg_b[in_eid[k]] == 0);
            if (any_in) insub[v] = 1;
        }
        for (int v = lane; v < N; v += 32) bw[v] = 0.0;
        __syncthreads();
        // ---- Brandes, one source per warp --------------------------------------------------
        for (int s = warp; s < N; s += W) {
            if (!insub[s]) continue;
            for (int v = lane; v < N; v += 32) { depth[v] = (v == s) ? 0 : -1; sigma[v] = (v == s) ? 1.0 : 0.0; delta[v] = 0.0; }
            __syncwarp();
            int level = 0;
            for (;; ++level) {
                bool grew = false;
                for (int v = lane; v < N; v += 32) {
                    if (depth[v] >= 0) continue;
                    double sg = 0.0;
                    bool hit = false;
                    for (int k = in_ptr[v]; k < in_ptr[v + 1]; ++k) {
                        if (!act[in_eid[k]]) continue;
                        const int u = in_src[k];
                        if (depth[u] == level) { sg += sigma[u]; hit = true; }
                    }
                    if (hit) { sigma[v] = sg; depth[v] = -2 - level; grew = true; }  // mark, commit below
                }
                __syncwarp();
                for (int v = lane; v < N; v += 32) if (depth[v] == -2 - level) depth[v] = level + 1;
                __syncwarp();
                if (!__any_sync(0xffffffffu, grew)) break;
            }
            // dependencies, deepest level first (level = last non-empty level index + 1 here)
            for (int L = level - 1; L >= 0; --L) {
                for (int v = lane; v < N; v += 32) {
                    if (depth[v] != L) continue;
                    double acc = 0.0;
                    const double sv = sigma[v];
                    for (int k = __ldg(p.out_ptr + v); k < __ldg(p.out_ptr + v + 1); ++k) {
                        if (!act[__ldg(p.out_eid + k)]) continue;
                        const int x = __ldg(p.out_dst + k);
                        if (depth[x] == L + 1) acc += sv * ((1.0 + delta[x]) / sigma[x]);
                    }
                    delta[v] = acc;
                    if (v != s) bw[v] += acc;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // ---- combine warps (fixed order), scalars -------------------------------------------
        double *bw0 = (double *)(smem + o) + 2 * N;  // warp 0's bw
        for (int v = threadIdx.x; v < N; v += blockDim.x) {
            double s = 0.0;
            for (int w = 0; w < W; ++w) s += *((double *)(smem + o + wstride * w) + 2 * N + v);
            bw0[v] = s;
        }
        __syncthreads();
        // block reductions: n_sub, bw max, goal total, remaining, active count, active flow sum
        double part[6] = {0, 0, 0, 0, 0, 0};
        for (int v = threadIdx.x; v < N; v += blockDim.x) { part[0] += insub[v]; part[1] = fmax(part[1], bw0[v]); }
        for (int e = threadIdx.x; e < E; e += blockDim.x) {
            part[2] += goal_b[e];
            part[3] += (goal_b[e] && dmg_b[e]) ? 1.0 : 0.0;
            if (!dmg_b[e]) { part[4] += 1.0; part[5] += (double)flow_b[e]; }
        }
        for (int i = 0; i < 6; ++i) {
            double v = part[i];
            for (int sft = 16; sft > 0; sft >>= 1) {
                const double other = __shfl_xor_sync(0xffffffffu, v, sft);
                v = (i == 1) ? fmax(v, other) : v + other;
            }
            if (lane == 0) red[warp * 6 + i] = v;
        }
        __syncthreads();
        double tot[6];
        for (int i = 0; i < 6; ++i) {
            double v = 0.0;
            for (int w = 0; w < W; ++w) v = (i == 1) ? fmax(v, red[w * 6 + i]) : v + red[w * 6 + i];
            tot[i] = v;
        }
        const double n_sub = tot[0];
        const double scale = n_sub > 2.0 ? 1.0 / ((n_sub - 1.0) * (n_sub - 2.0)) : 1.0;
        const float bw_max = (float)(tot[1] * scale);
        const float remaining_ratio = (float)(tot[3] / fmax(tot[2], 1.0));
        const double avg_flow = tot[4] > 0.0 ? tot[5] / tot[4] : 0.0;
        const float avg_flow_norm = (float)(avg_flow / fmax(p.total_demand / (double)max(E, 1), 1.0));
        const float log_tstt = (float)log10(fmax(p.tstt[b], 1.0));
        float *nf = p.node_feat + (size_t)b * N * 4;
        for (int v = threadIdx.x; v < N; v += blockDim.x) {
            float x = (float)(bw0[v] * scale);
            if (p.bw_raw) p.bw_raw[(size_t)b * N + v] = x;
            if (bw_max > 0.0f) x = x / bw_max;
            nf[v * 4 + 0] = x; nf[v * 4 + 1] = remaining_ratio; nf[v * 4 + 2] = avg_flow_norm; nf[v * 4 + 3] = log_tstt;
        }
        float *ef = p.edge_feat + off * 6;
        const double lt0 = log10(p.max_t0 + 1.0), lcap = log10(p.max_cap0 + 1.0);
        for (int e = threadIdx.x; e < E; e += blockDim.x) {
            const bool d = dmg_b[e] != 0;
            const float cap = d ? (float)p.capacity_damage : __ldg(p.cap0 + e);
            const float t0n = (float)((double)log10f(__ldg(p.t0 + e) + 1.0f) / lt0);
            const float capn = (float)((double)log10f(cap + 1.0f) / lcap);
            float vc = 0.0f;
            if (!d) {
                vc = (float)flow_b[e] / fmaxf(cap, 1e-6f);
                vc = log1pf(vc);
                vc = fminf(fmaxf(vc, 0.0f), 10.0f);
            }
            ef[e * 6 + 0] = t0n; ef[e * 6 + 1] = capn; ef[e * 6 + 2] = vc; ef[e * 6 + 3] = d ? 1.0f : 0.0f;
            ef[e * 6 + 4] = goal_b[e] ? 1.0f : 0.0f; ef[e * 6 + 5] = (float)e / (float)max(E - 1, 1);
        }
    }
}

// Small networks (N <= 32, e.g. Sioux Falls): ONE WARP per scenario, LANE = BFS SOURCE.  The active
// out-neighbours of every node are a 32-bit mask, so a lane expands its own breadth-first search one
// (node, successor) pair per trip - every lane has about as many pairs as there are undamaged links, so the
// warp stays converged without any level synchronisation - keeping depth / path counts / BFS order /
// dependencies in shared-memory columns of its own (row = node, column = lane: conflict-free whatever node
// a lane touches).  The dependency pass walks the BFS order backwards with one division per node
// ((1 + delta[w]) / sigma[w], reused by every predecessor of w).  About 6x fewer warp-instructions per
// state than one warp per source (features_kernel), which remains the path of larger networks.
constexpr int kFeatSmallWarps = 4;

template <typename REAL>
__global__ void __launch_bounds__(kFeatSmallWarps * 32) features_small_kernel(FeatParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = p.N, E = p.E, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kFeatSmallWarps + warp;
    if (b >= p.B) return;
    const unsigned FULL = 0xffffffffu;
    // per warp: delta f64 [N][32] | coef f64 [N][32] | sigma f32 [N][32] | depth u8 [N][32] | order u8 [N][32] | adj u32 [32]
    const size_t stride = (size_t)N * 32 * (8 + 8 + 4 + 1 + 1) + 128;
    unsigned char *wb = smem + stride * warp;
    double *delta = (double *)wb, *coef = delta + N * 32;
    float *sigma = (float *)(coef + N * 32);
    uint8_t *depth = (uint8_t *)(sigma + N * 32), *order = depth + N * 32;
    unsigned *adj = (unsigned *)(order + N * 32);
    const size_t off = (size_t)b * E;
    const uint8_t *dmg_b = p.dmg + off, *goal_b = p.goal + off;
    const REAL *flow_b = (const REAL *)p.flow + off;
    // active out-neighbour masks; nodes of the edge-induced subgraph of undamaged links
    unsigned mine = 0;
    if (lane < N)
        for (int k = __ldg(p.out_ptr + lane); k < __ldg(p.out_ptr + lane + 1); ++k)
            if (!dmg_b[__ldg(p.out_eid + k)]) mine |= 1u << __ldg(p.out_dst + k);
    adj[lane] = mine;
    unsigned has_in = mine;
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) has_in |= __shfl_xor_sync(FULL, has_in, sft);
    const unsigned insub = has_in | __ballot_sync(FULL, mine != 0);
    __syncwarp();
    const int s = lane;
    const bool src_ok = s < N && ((insub >> s) & 1u);
    for (int v = 0; v < N; ++v) { depth[v * 32 + lane] = 0xff; delta[v * 32 + lane] = 0.0; }
    int qt = 0;
    if (src_ok) { depth[s * 32 + lane] = 0; sigma[s * 32 + lane] = 1.0f; order[lane] = (uint8_t)s; qt = 1; }
    // ---- forward: BFS order, depths, shortest-path counts; one (node, successor) pair per trip
    {
        int qh = 0, v = 0, dv = 0;
        unsigned rem = 0;
        float sv = 0.0f;
        bool live = src_ok;
        while (__any_sync(FULL, live)) {
            if (live && rem == 0) {
                if (qh < qt) { v = order[qh * 32 + lane]; ++qh; rem = adj[v]; dv = depth[v * 32 + lane]; sv = sigma[v * 32 + lane]; }
                else live = false;
            }
            if (live && rem) {
                const int w = __ffs(rem) - 1;
                rem &= rem - 1;
                const int dw = depth[w * 32 + lane];
                if (dw == 0xff) {
                    depth[w * 32 + lane] = (uint8_t)(dv + 1);
                    sigma[w * 32 + lane] = sv;
                    order[qt * 32 + lane] = (uint8_t)w;
                    ++qt;
                } else if (dw == dv + 1) {
                    sigma[w * 32 + lane] += sv;
                }
            }
        }
    }
    // ---- backward: dependencies in reverse BFS order; coef[w] = (1 + delta[w]) / sigma[w] once per node
    {
        int qi = qt, v = 0, dv = 0;
        unsigned rem = 0;
        double acc = 0.0;
        bool live = src_ok, open = false;
        while (__any_sync(FULL, live)) {
            if (live && rem == 0) {
                if (open) {   // node v is complete
                    const double sg = (double)sigma[v * 32 + lane];
                    const double dl = sg * acc;
                    delta[v * 32 + lane] = dl;
                    coef[v * 32 + lane] = (1.0 + dl) / sg;
                    open = false;
                }
                if (qi > 0) { --qi; v = order[qi * 32 + lane]; rem = adj[v]; dv = depth[v * 32 + lane]; acc = 0.0; open = true; }
                else live = false;
            }
            if (live && rem) {
                const int w = __ffs(rem) - 1;
                rem &= rem - 1;
                if (depth[w * 32 + lane] == dv + 1) acc += coef[w * 32 + lane];
            }
        }
        if (src_ok) delta[s * 32 + lane] = 0.0;   // the source itself is not counted (reference: networkx)
    }
    __syncwarp();
    // ---- betweenness of node `lane` = sum over sources (fixed order), then the scalars of the state
    double bw = 0.0;
    for (int v = 0; v < N; ++v) {
        double x = delta[v * 32 + lane];   // lane = source column of row v
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) x += __shfl_xor_sync(FULL, x, sft);
        if (lane == v) bw = x;
    }
    const double n_sub = (double)__popc(insub);
    const double scale = n_sub > 2.0 ? 1.0 / ((n_sub - 1.0) * (n_sub - 2.0)) : 1.0;
    double bmax = bw;
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) bmax = fmax(bmax, __shfl_xor_sync(FULL, bmax, sft));
    double part[4] = {0, 0, 0, 0};   // goal total, remaining, active count, active flow sum
    for (int e = lane; e < E; e += 32) {
        part[0] += goal_b[e];
        part[1] += (goal_b[e] && dmg_b[e]) ? 1.0 : 0.0;
        if (!dmg_b[e]) { part[2] += 1.0; part[3] += (double)flow_b[e]; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) part[i] += __shfl_xor_sync(FULL, part[i], sft);
    const float bw_max = (float)(bmax * scale);
    const float remaining_ratio = (float)(part[1] / fmax(part[0], 1.0));
    const double avg_flow = part[2] > 0.0 ? part[3] / part[2] : 0.0;
    const float avg_flow_norm = (float)(avg_flow / fmax(p.total_demand / (double)max(E, 1), 1.0));
    const float log_tstt = (float)log10(fmax(p.tstt[b], 1.0));
    if (lane < N) {
        float x = (float)(bw * scale);
        if (p.bw_raw) p.bw_raw[(size_t)b * N + lane] = x;
        if (bw_max > 0.0f) x = x / bw_max;
        *(float4 *)(p.node_feat + ((size_t)b * N + lane) * 4) = make_float4(x, remaining_ratio, avg_flow_norm, log_tstt);
    }
    float *ef = p.edge_feat + off * 6;
    const double lt0 = log10(p.max_t0 + 1.0), lcap = log10(p.max_cap0 + 1.0);
    for (int e = lane; e < E; e += 32) {
        const bool d = dmg_b[e] != 0;
        const float cap = d ? (float)p.capacity_damage : __ldg(p.cap0 + e);
        const float t0n = (float)((double)log10f(__ldg(p.t0 + e) + 1.0f) / lt0);
        const float capn = (float)((double)log10f(cap + 1.0f) / lcap);
        float vc = 0.0f;
        if (!d) {
            vc = (float)flow_b[e] / fmaxf(cap, 1e-6f);
            vc = log1pf(vc);
            vc = fminf(fmaxf(vc, 0.0f), 10.0f);
        }
        ef[e * 6 + 0] = t0n; ef[e * 6 + 1] = capn; ef[e * 6 + 2] = vc; ef[e * 6 + 3] = d ? 1.0f : 0.0f;
        ef[e * 6 + 4] = goal_b[e] ? 1.0f : 0.0f; ef[e * 6 + 5] = (float)e / (float)max(E - 1, 1);
    }
}

// ---------------------------------------------------------------------------------------------
// reset(): damage sampling on the device (reference :167-190)
// ---------------------------------------------------------------------------------------------
// One warp per scenario.  A try = `count` distinct links by a partial Fisher-Yates shuffle (lane 0, on
// a shared-memory permutation, randoms from a counter-based 64-bit mixer), accepted when the undamaged
// links form a strongly connected subgraph on the nodes they touch - forward and backward reachability
// from the lowest such node, relaxed by the whole warp to a fixed point.  Like the reference: at most
// max_tries tries, then one more draw kept unconditionally.  The random stream is not numpy's (the
// reference draws with Generator.choice); the RULE and the distribution are the reference's.
struct SampleParams {
    int N, E, B, count, max_tries;
    const int32_t *topo;  // in_ptr | in_src | in_eid
    int off_in_src, off_in_eid;
    const int32_t *out_ptr, *out_dst, *out_eid;
    unsigned long long seed;
    unsigned long long *counters;
    const uint8_t *select;
    uint8_t *dmg;
    int32_t *tries;
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

constexpr int kSampleWarps = 4;

__global__ void __launch_bounds__(kSampleWarps * 32) sample_damage_kernel(SampleParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = p.N, E = p.E, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kSampleWarps + warp;
    if (b >= p.B) return;
    if (p.select && !p.select[b]) return;
    const int32_t *in_ptr = p.topo, *in_src = p.topo + p.off_in_src, *in_eid = p.topo + p.off_in_eid;
    const size_t stride = (size_t)((E * 2 + 15) & ~15) + (size_t)((E + 15) & ~15) + 2 * (size_t)((N + 15) & ~15);
    unsigned char *wb = smem + stride * warp;
    uint16_t *perm = (uint16_t *)wb;
    uint8_t *dmg = wb + ((E * 2 + 15) & ~15);
    uint8_t *fwd = dmg + ((E + 15) & ~15), *bwd = fwd + ((N + 15) & ~15);
    const unsigned FULL = 0xffffffffu;
    unsigned long long ctr = p.counters[b];
    const unsigned long long key = p.seed ^ (0xd1342543de82ef95ull * (unsigned long long)(b + 1));
    int tries = 0;
    for (;; ++tries) {
        for (int e = lane; e < E; e += 32) { perm[e] = (uint16_t)e; dmg[e] = 0; }
        __syncwarp();
        if (lane == 0) {
            for (int i = 0; i < p.count; ++i) {
                const unsigned long long r = mix64(key + mix64(ctr + (unsigned long long)i));
                const int j = i + (int)__umul64hi(r, (unsigned long long)(E - i));
                const uint16_t t = perm[i]; perm[i] = perm[j]; perm[j] = t;
                dmg[perm[i]] = 1;
            }
        }
        ctr += (unsigned long long)p.count;
        __syncwarp();
        if (tries >= p.max_tries) break;   // the reference's unconstrained last draw
        // nodes of the edge-induced subgraph of undamaged links; its lowest node is the root
        int root = N;
        for (int v = lane; v < N; v += 32) {
            bool present = false;
            for (int k = in_ptr[v]; k < in_ptr[v + 1]; ++k) present |= dmg[in_eid[k]] == 0;
            for (int k = __ldg(p.out_ptr + v); k < __ldg(p.out_ptr + v + 1); ++k) present |= dmg[__ldg(p.out_eid + k)] == 0;
            fwd[v] = present ? 0 : 2;   // 2 = not part of the subgraph
            bwd[v] = present ? 0 : 2;
            if (present && v < root) root = v;
        }
        root = __reduce_min_sync(FULL, root);
        __syncwarp();
        bool ok = root < N;
        if (ok) {
            if (lane == 0) { fwd[root] = 1; bwd[root] = 1; }
            __syncwarp();
            for (bool changed = true; changed;) {
                changed = false;
                for (int v = lane; v < N; v += 32) {
                    if (fwd[v] == 0) {
                        bool hit = false;
                        for (int k = in_ptr[v]; k < in_ptr[v + 1]; ++k) hit |= dmg[in_eid[k]] == 0 && fwd[in_src[k]] == 1;
                        if (hit) { fwd[v] = 1; changed = true; }
                    }
                    if (bwd[v] == 0) {
                        bool hit = false;
                        for (int k = __ldg(p.out_ptr + v); k < __ldg(p.out_ptr + v + 1); ++k)
                            hit |= dmg[__ldg(p.out_eid + k)] == 0 && bwd[__ldg(p.out_dst + k)] == 1;
                        if (hit) { bwd[v] = 1; changed = true; }
                    }
                }
                __syncwarp();
                changed = __any_sync(FULL, changed);
            }
            bool all = true;
            for (int v = lane; v < N; v += 32) all &= fwd[v] != 0 && bwd[v] != 0;
            ok = __all_sync(FULL, all);
        }
        if (ok) break;
    }
    uint8_t *out = p.dmg + (size_t)b * E;
    for (int e = lane; e < E; e += 32) out[e] = dmg[e];
    if (lane == 0) {
        p.counters[b] = ctr;
        if (p.tries) p.tries[b] = tries + 1;
    }
}


__global__ void apply_actions_kernel(int B, int E, const int32_t *action, uint8_t *dmg, uint8_t *valid, int32_t *status)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int a = action[b];
    uint8_t ok = 0;
    if (a < 0 || a >= E) {
        if (status) atomicOr(status + b, 8);
    } else if (dmg[(size_t)b * E + a]) {
        dmg[(size_t)b * E + a] = 0;
        ok = 1;
    }
    valid[b] = ok;
}

// One warp per scenario: done = no goal link still damaged (reference :293-294), reward per mode
// (reference :256-291); an invalid action gives reward -1 and done = False (reference :210-212).
__global__ void reward_done_kernel(int B, int E, int mode, const double *prev, const double *curr, const double *init,
                                   const uint8_t *dmg, const uint8_t *goal, const uint8_t *valid, double ra, double rb,
                                   double rg, double clip, double *reward, uint8_t *done)
{
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    int left = 0;
    for (int e = lane; e < E; e += 32) left += (goal[(size_t)b * E + e] && dmg[(size_t)b * E + e]) ? 1 : 0;
    left = __reduce_add_sync(0xffffffffu, left);
    if (lane != 0) return;
    const bool complete = left == 0;
    if (valid && !valid[b]) { reward[b] = -1.0; done[b] = 0; return; }
    const double pv = prev[b], cv = curr[b];
    const double bonus = complete ? rb : 0.0;
    double r;
    if (mode == 3 || mode == 4) {
        const double base = fmax(init ? init[b] : pv, 1.0);
        const double ratio = cv / base;
        if (mode == 3) r = -ra * ratio;
        else r = ra * (((pv - cv) / base) * 100.0) - 1.0 * ratio;
        r = r + bonus;
    } else {
        double delta;
        if (mode == 2) delta = -cv;
        else if (mode == 1) delta = log10(fmax(pv, 1.0)) - log10(fmax(cv, 1.0));
        else delta = pv - cv;
        r = ra * delta + bonus - rg;
    }
    if (clip > 0.0) r = fmin(fmax(r, -clip), clip);
    reward[b] = r;
    done[b] = complete ? 1 : 0;
}

}  // namespace

extern "C" int b200_state_features(const b200_graph *g, int B, int prec, const void *flow, const uint8_t *dmg,
                                   const uint8_t *goal, const double *tstt, double capacity_damage, float *node_feat,
                                   float *edge_feat, float *betweenness_raw, void *stream)
{
    B200_REQUIRE(g && flow && dmg && goal && tstt && node_feat && edge_feat && B >= 0, "bad argument");
    B200_REQUIRE(prec == B200_F32 || prec == B200_F64, "prec must be B200_F32 or B200_F64");
    if (B == 0) return 0;
    FeatParams p;
    memset(&p, 0, sizeof(p));
    p.N = g->N; p.E = g->E; p.B = B; p.topo = g->topo; p.off_in_src = g->off_in_src; p.off_in_eid = g->off_in_eid;
    p.out_ptr = g->out_ptr; p.out_dst = g->out_dst; p.out_eid = g->out_eid; p.t0 = g->t0; p.cap0 = g->cap0;
    p.flow = flow; p.dmg = dmg; p.goal = goal; p.tstt = tstt; p.total_demand = g->total_demand;
    p.capacity_damage = capacity_damage; p.max_t0 = g->max_t0; p.max_cap0 = g->max_cap0;
    p.node_feat = node_feat; p.edge_feat = edge_feat; p.bw_raw = betweenness_raw;
    if (g->N <= 32 && !getenv("B200_FEATURES_GENERIC")) {   // lane-per-source kernel (Sioux Falls class)
        const size_t smem = ((size_t)g->N * 32 * (8 + 8 + 4 + 1 + 1) + 128) * kFeatSmallWarps;
        auto kern = prec == B200_F32 ? features_small_kernel<float> : features_small_kernel<double>;
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(B + kFeatSmallWarps - 1) / kFeatSmallWarps, kFeatSmallWarps * 32, smem, (cudaStream_t)stream>>>(p);
        B200_CUDA(cudaGetLastError());
        return 0;
    }
    int warps = kFeatWarps;
    auto smem_for = [&](int w) {
        return (size_t)((g->E + 15) & ~15) + (size_t)((g->N + 15) & ~15) + 64 * 8 +
               (size_t)w * ((size_t)g->N * 24 + (size_t)((g->N * 4 + 15) & ~15));
    };
    while (warps > 1 && smem_for(warps) > 200 * 1024) --warps;
    const size_t smem = smem_for(warps);
    B200_REQUIRE(smem <= 227 * 1024, "network too large for the feature kernel");
    auto kern = prec == B200_F32 ? features_kernel<float> : features_kernel<double>;
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::min(B, g->num_sms * 8);
    kern<<<grid, warps * 32, smem, (cudaStream_t)stream>>>(p);
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200_apply_actions(const b200_graph *g, int B, const int32_t *action, uint8_t *dmg, uint8_t *valid,
                                  int32_t *status, void *stream)
{
    B200_REQUIRE(g && action && dmg && valid && B >= 0, "bad argument");
    if (B == 0) return 0;
    apply_actions_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, g->E, action, dmg, valid, status);
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200_reward_done(const b200_graph *g, int B, int reward_mode, const double *prev_tstt, const double *tstt,
                                const double *initial_tstt, const uint8_t *dmg, const uint8_t *goal,
                                const uint8_t *valid, double r_alpha, double r_beta, double r_gamma, double r_clip,
                                double *reward, uint8_t *done, void *stream)
{
    B200_REQUIRE(g && prev_tstt && tstt && dmg && goal && reward && done && B >= 0, "bad argument");
    B200_REQUIRE(reward_mode >= 0 && reward_mode <= 4, "unknown reward mode");
    if (B == 0) return 0;
    const int threads = 128, per_block = threads / 32;
    reward_done_kernel<<<(B + per_block - 1) / per_block, threads, 0, (cudaStream_t)stream>>>(
        B, g->E, reward_mode, prev_tstt, tstt, initial_tstt, dmg, goal, valid, r_alpha, r_beta, r_gamma, r_clip, reward,
        done);
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200_sample_damage(const b200_graph *g, int B, int count, int max_tries, uint64_t seed, uint64_t *counters,
                                  const uint8_t *select, uint8_t *dmg, int32_t *tries, void *stream)
{
    B200_REQUIRE(g && counters && dmg && B >= 0, "bad argument");
    B200_REQUIRE(count >= 1 && count <= g->E && max_tries >= 0, "need 1 <= count <= E, max_tries >= 0");
    B200_REQUIRE(g->E < 65536, "network too large for the 16-bit permutation");
    if (B == 0) return 0;
    SampleParams p;
    p.N = g->N; p.E = g->E; p.B = B; p.count = count; p.max_tries = max_tries;
    p.topo = g->topo; p.off_in_src = g->off_in_src; p.off_in_eid = g->off_in_eid;
    p.out_ptr = g->out_ptr; p.out_dst = g->out_dst; p.out_eid = g->out_eid;
    p.seed = seed; p.counters = (unsigned long long *)counters; p.select = select; p.dmg = dmg; p.tries = tries;
    const size_t stride = (size_t)((g->E * 2 + 15) & ~15) + (size_t)((g->E + 15) & ~15) + 2 * (size_t)((g->N + 15) & ~15);
    const size_t smem = stride * kSampleWarps;
    B200_REQUIRE(smem <= 200 * 1024, "network too large for the sampler's shared-memory scratch");
    B200_CUDA(cudaFuncSetAttribute(sample_damage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sample_damage_kernel<<<(B + kSampleWarps - 1) / kSampleWarps, kSampleWarps * 32, smem, (cudaStream_t)stream>>>(p);
    B200_CUDA(cudaGetLastError());
    return 0;
}
