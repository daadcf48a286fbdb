// G1 - GAT edge-softmax + neighbour aggregation (forward and backward) over a batch of graphs
// that share one topology.  Replaces the message / softmax / aggregate part of
// torch_geometric.nn.GATConv as the reference uses it (src/models/gat_encoder.py:23-27, :32-49;
// attention semantics: softmax over the incoming links of each target node incl. the self loop,
// leaky_relu(negative_slope), coefficients = exp(e - max) / (sum + 1e-16)).  The dense projections
// (lin, lin_edge, the att_* dot products) and bias / head concat stay torch GEMMs.
//
// One CTA owns one graph at a time.  Its projected node features xw[b] ([N][H][C] fp32, one
// contiguous slab) are staged into shared memory by a TMA bulk copy, double-buffered across the
// graphs of a persistent CTA, so every xw row is read from HBM once although it is gathered
// (in-degree + 1) times.  When the slab does not fit, rows are gathered from global / L2 instead.
#include "b200_common.cuh"

#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

constexpr int kGatThreads = 1024;  // one CTA per SM (the staged slabs fill shared memory): all the warps it can hold
constexpr int kMaxDeg = 31;  // incoming links + self loop handled per (node, head)

struct GatParams {
    int B, N, E, H, C;
    const int32_t *in_ptr, *in_src, *in_eid, *out_ptr, *out_dst, *out_eid;
    const float *xw, *a_src, *a_dst, *a_edge, *alpha_in, *grad_out;
    float *out, *alpha_out, *grad_xw, *grad_a_src, *grad_a_dst, *grad_a_edge;
    float slope;
};

__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.0f ? x : slope * x; }

// Staged buffers of one graph: xw slab | a_src [N][H] | a_dst [N][H] | a_edge [E+N][H], each
// padded to 16 bytes (TMA bulk copies, all counted on the same mbarrier).
__host__ __device__ inline size_t gat_pad16(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline size_t gat_stage_bytes(int N, int E, int H, int C)
{
    return gat_pad16((size_t)N * H * C * 4) + 2 * gat_pad16((size_t)N * H * 4) + gat_pad16((size_t)(E + N) * H * 4);
}

// SW = shuffle width of the per-(node, head) softmax: 8 when in-degree + self loop <= 8 (road
// networks), else 32.
template <bool STAGED, int SW>
__global__ void __launch_bounds__(kGatThreads) gat_fwd_kernel(GatParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int N = p.N, E = p.E, H = p.H, C = p.C;
    const size_t slab = (size_t)N * H * C;           // floats per graph
    const size_t nh = (size_t)N * H, eh = (size_t)(E + N) * H;
    const size_t stage_bytes = gat_stage_bytes(N, E, H, C);
    uint64_t *bars = (uint64_t *)smem;                // 2 mbarriers
    unsigned char *stage[2] = {smem + 128, smem + 128 + stage_bytes};
    const bool vec_ok = (nh * 4) % 16 == 0 && (eh * 4) % 16 == 0;   // attention terms TMA-able too
    auto issue = [&](int buf, int b) {                // thread 0 only
        unsigned char *dst = stage[buf];
        uint32_t bytes = (uint32_t)(slab * 4);
        if (vec_ok) bytes += (uint32_t)(2 * nh * 4 + eh * 4);
        mbar_expect_tx(&bars[buf], bytes);
        tma_bulk_g2s(dst, p.xw + (size_t)b * slab, (uint32_t)(slab * 4), &bars[buf]);
        if (vec_ok) {
            unsigned char *q = dst + gat_pad16(slab * 4);
            tma_bulk_g2s(q, p.a_src + (size_t)b * nh, (uint32_t)(nh * 4), &bars[buf]);
            q += gat_pad16(nh * 4);
            tma_bulk_g2s(q, p.a_dst + (size_t)b * nh, (uint32_t)(nh * 4), &bars[buf]);
            q += gat_pad16(nh * 4);
            tma_bulk_g2s(q, p.a_edge + (size_t)b * eh, (uint32_t)(eh * 4), &bars[buf]);
        }
    };
    if (STAGED) {
        if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
        __syncthreads();
        if (threadIdx.x == 0 && blockIdx.x < p.B) issue(0, blockIdx.x);
    }
    int it = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
        const float *x;
        const float *as = p.a_src + (size_t)b * nh, *ad = p.a_dst + (size_t)b * nh;
        const float *ae = p.a_edge + (size_t)b * eh;
        if (STAGED) {
            const int s = it & 1, nb = b + gridDim.x;
            if (threadIdx.x == 0 && nb < p.B) issue(s ^ 1, nb);   // prefetch the next graph into the other buffer
            mbar_wait(&bars[s], (it >> 1) & 1);
            x = (const float *)stage[s];
            if (vec_ok) {
                as = (const float *)(stage[s] + gat_pad16(slab * 4));
                ad = (const float *)((const unsigned char *)as + gat_pad16(nh * 4));
                ae = (const float *)((const unsigned char *)ad + gat_pad16(nh * 4));
            }
        } else {
            x = p.xw + (size_t)b * slab;
        }
        float *al = p.alpha_out ? p.alpha_out + (size_t)b * eh : nullptr;
        float *o = p.out + (size_t)b * slab;
        for (int pair = warp; pair < N * H; pair += W) {
            const int v = pair / H, h = pair - v * H;
            const int k0 = __ldg(p.in_ptr + v), deg = __ldg(p.in_ptr + v + 1) - k0;
            // logits: lane j < deg handles in-link j, lane deg the self loop
            float logit = -INFINITY;
            int u = v, row = E + v;
            if (lane < deg) { u = __ldg(p.in_src + k0 + lane); row = __ldg(p.in_eid + k0 + lane); }
            if (lane <= deg) logit = leaky(as[u * H + h] + ad[v * H + h] + ae[(size_t)row * H + h], p.slope);
            float mx = logit;
#pragma unroll
            for (int sft = SW / 2; sft > 0; sft >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sft));
            mx = __shfl_sync(0xffffffffu, mx, 0);      // lanes beyond SW hold -inf groups of their own
            float ex = lane <= deg ? expf(logit - mx) : 0.0f;
            float sum = ex;
#pragma unroll
            for (int sft = SW / 2; sft > 0; sft >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sft);
            sum = __shfl_sync(0xffffffffu, sum, 0);
            const float a = ex / (sum + 1e-16f);
            if (al && lane <= deg) al[(size_t)row * H + h] = a;
            // aggregate: each lane owns float4 chunks of the C channels, 256 channels (two chunks) per
            // pass so one (alpha, source row) broadcast feeds two 16-byte loads; 32-bit offsets
            const int my_off = (u * H + h) * C;        // float offset of this lane's source row
            const int out_off = (v * H + h) * C;
            for (int c0 = 0; c0 < C; c0 += 256) {      // warp-uniform trip count: shuffles inside
                const int c = c0 + lane * 4;
                const bool act0 = c < C, act1 = c + 128 < C;
                float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
                for (int j = 0; j <= deg; ++j) {
                    const float aj = __shfl_sync(0xffffffffu, a, j);
                    const int oj = __shfl_sync(0xffffffffu, my_off, j) + c;
                    if (act0) {
                        const float4 xv = *(const float4 *)(x + oj);
                        acc0.x = fmaf(aj, xv.x, acc0.x); acc0.y = fmaf(aj, xv.y, acc0.y);
                        acc0.z = fmaf(aj, xv.z, acc0.z); acc0.w = fmaf(aj, xv.w, acc0.w);
                    }
                    if (act1) {
                        const float4 xv = *(const float4 *)(x + oj + 128);
                        acc1.x = fmaf(aj, xv.x, acc1.x); acc1.y = fmaf(aj, xv.y, acc1.y);
                        acc1.z = fmaf(aj, xv.z, acc1.z); acc1.w = fmaf(aj, xv.w, acc1.w);
                    }
                }
                if (act0) __stcs((float4 *)(o + out_off + c), acc0);   // streaming: never re-read here
                if (act1) __stcs((float4 *)(o + out_off + c + 128), acc1);
            }
        }
        if (STAGED) __syncthreads();  // everyone is done with this buffer before it is refilled
    }
}

// Backward.  Phase 1 (per target node, head): d alpha = <grad_out[v], xw[u]>, softmax and
// leaky-relu backward -> grad of the raw logit per incoming link (kept in shared memory),
// grad_a_edge, grad_a_dst.  Phase 2 (per source node, head): grad_xw[u] = sum over outgoing links
// alpha * grad_out[dst], grad_a_src[u] = sum of raw-logit grads.  Gathers only, no atomics.
template <bool STAGED>
__global__ void __launch_bounds__(kGatThreads) gat_bwd_kernel(GatParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int N = p.N, E = p.E, H = p.H, C = p.C;
    const size_t slab = (size_t)N * H * C;
    uint64_t *bar = (uint64_t *)smem;
    float *graw = (float *)(smem + 128);                                   // [(E+N)][H]
    float *xs = graw + (((size_t)(E + N) * H + 31) & ~(size_t)31);          // staged xw[b]
    float *gs = xs + slab;                                                  // staged grad_out[b]
    // staged attention terms of the graph: a_src [N][H] | a_dst [N][H] | a_edge [E+N][H] | alpha [E+N][H]
    const size_t nh = (size_t)N * H, eh = (size_t)(E + N) * H;
    const bool att_ok = STAGED && (nh * 4) % 16 == 0 && (eh * 4) % 16 == 0;
    float *att = gs + slab;
    auto issue_g = [&](int b) {   // thread 0: grad_out slab + attention terms on bar_g
        mbar_expect_tx(bar + 1, (uint32_t)(slab * 4 + (att_ok ? (2 * nh + 2 * eh) * 4 : 0)));
        tma_bulk_g2s(gs, p.grad_out + (size_t)b * slab, (uint32_t)(slab * 4), bar + 1);
        if (att_ok) {
            tma_bulk_g2s(att, p.a_src + (size_t)b * nh, (uint32_t)(nh * 4), bar + 1);
            tma_bulk_g2s(att + nh, p.a_dst + (size_t)b * nh, (uint32_t)(nh * 4), bar + 1);
            tma_bulk_g2s(att + 2 * nh, p.a_edge + (size_t)b * eh, (uint32_t)(eh * 4), bar + 1);
            tma_bulk_g2s(att + 2 * nh + eh, p.alpha_in + (size_t)b * eh, (uint32_t)(eh * 4), bar + 1);
        }
    };
    // two mbarriers: xw is only needed in phase 1, so the next graph's xw is fetched while phase 2
    // still runs on grad_out; only the grad_out copy is exposed at the top of the loop
    uint64_t *bar_x = bar, *bar_g = bar + 1;
    if (STAGED) {
        if (threadIdx.x == 0) {
            mbar_init(bar_x, 1); mbar_init(bar_g, 1); mbar_fence_init();
        }
        __syncthreads();
        if (threadIdx.x == 0 && blockIdx.x < p.B) {
            mbar_expect_tx(bar_x, (uint32_t)(slab * 4));
            tma_bulk_g2s(xs, p.xw + (size_t)blockIdx.x * slab, (uint32_t)(slab * 4), bar_x);
            issue_g(blockIdx.x);
        }
    }
    int it = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
        const float *x = p.xw + (size_t)b * slab, *go = p.grad_out + (size_t)b * slab;
        if (STAGED) {
            mbar_wait(bar_x, it & 1);
            mbar_wait(bar_g, it & 1);
            x = xs;
            go = gs;
        }
        const float *as = p.a_src + (size_t)b * N * H, *ad = p.a_dst + (size_t)b * N * H;
        const float *ae = p.a_edge + (size_t)b * (E + N) * H;
        const float *al = p.alpha_in + (size_t)b * (E + N) * H;
        if (att_ok) { as = att; ad = att + nh; ae = att + 2 * nh; al = att + 2 * nh + eh; }
        float *g_as = p.grad_a_src + (size_t)b * N * H, *g_ad = p.grad_a_dst + (size_t)b * N * H;
        float *g_ae = p.grad_a_edge + (size_t)b * (E + N) * H;
        float *gx = p.grad_xw + (size_t)b * slab;
        for (int pair = warp; pair < N * H; pair += W) {
            const int v = pair / H, h = pair - v * H;
            const int k0 = __ldg(p.in_ptr + v), deg = __ldg(p.in_ptr + v + 1) - k0;
            int u = v, row = E + v;
            if (lane < deg) { u = __ldg(p.in_src + k0 + lane); row = __ldg(p.in_eid + k0 + lane); }
            // d alpha_j = <grad_out[v,h,:], xw[u_j,h,:]>, computed cooperatively, kept by lane j.
            // grad_out[v] is loaded once per 256-channel pass; 32-bit offsets.
            float dalpha = 0.0f;
            const int my_off = (u * H + h) * C, v_off = (v * H + h) * C;
            for (int c0 = 0; c0 < C; c0 += 256) {      // warp-uniform trip count: shuffles inside
                const int c = c0 + lane * 4;
                const bool act0 = c < C, act1 = c + 128 < C;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 g0 = act0 ? *(const float4 *)(go + v_off + c) : z;
                const float4 g1 = act1 ? *(const float4 *)(go + v_off + c + 128) : z;
                for (int j = 0; j <= deg; ++j) {
                    const int oj = __shfl_sync(0xffffffffu, my_off, j) + c;
                    const float4 x0 = act0 ? *(const float4 *)(x + oj) : z;
                    const float4 x1 = act1 ? *(const float4 *)(x + oj + 128) : z;
                    float part = fmaf(g0.x, x0.x, fmaf(g0.y, x0.y, fmaf(g0.z, x0.z, g0.w * x0.w)));
                    part += fmaf(g1.x, x1.x, fmaf(g1.y, x1.y, fmaf(g1.z, x1.z, g1.w * x1.w)));
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) part += __shfl_xor_sync(0xffffffffu, part, sft);
                    if (lane == j) dalpha += part;
                }
            }
            const float a = lane <= deg ? al[(size_t)row * H + h] : 0.0f;
            float dot = a * dalpha;
            for (int sft = 16; sft > 0; sft >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, sft);
            float gr = 0.0f;
            if (lane <= deg) {
                const float raw = as[u * H + h] + ad[v * H + h] + ae[(size_t)row * H + h];
                gr = a * (dalpha - dot) * (raw > 0.0f ? 1.0f : p.slope);
                g_ae[(size_t)row * H + h] = gr;
                graw[(size_t)row * H + h] = gr;
            }
            float gsum = gr;
            for (int sft = 16; sft > 0; sft >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, sft);
            if (lane == 0) g_ad[v * H + h] = gsum;
        }
        __syncthreads();
        if (STAGED && threadIdx.x == 0 && b + (int)gridDim.x < p.B) {   // phase 1 is done with xw
            mbar_expect_tx(bar_x, (uint32_t)(slab * 4));
            tma_bulk_g2s(xs, p.xw + (size_t)(b + gridDim.x) * slab, (uint32_t)(slab * 4), bar_x);
        }
        for (int pair = warp; pair < N * H; pair += W) {
            const int u = pair / H, h = pair - u * H;
            const int k0 = __ldg(p.out_ptr + u), deg = __ldg(p.out_ptr + u + 1) - k0;
            int dst = u, row = E + u;
            if (lane < deg) { dst = __ldg(p.out_dst + k0 + lane); row = __ldg(p.out_eid + k0 + lane); }
            const float a = lane <= deg ? al[(size_t)row * H + h] : 0.0f;
            float gsum = lane <= deg ? graw[(size_t)row * H + h] : 0.0f;
            for (int sft = 16; sft > 0; sft >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, sft);
            if (lane == 0) g_as[u * H + h] = gsum;
            const int my_off = (dst * H + h) * C, u_off = (u * H + h) * C;
            for (int c0 = 0; c0 < C; c0 += 256) {      // warp-uniform trip count: shuffles inside
                const int c = c0 + lane * 4;
                const bool act0 = c < C, act1 = c + 128 < C;
                float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
                for (int j = 0; j <= deg; ++j) {
                    const float aj = __shfl_sync(0xffffffffu, a, j);
                    const int oj = __shfl_sync(0xffffffffu, my_off, j) + c;
                    if (act0) {
                        const float4 gv = *(const float4 *)(go + oj);
                        acc0.x = fmaf(aj, gv.x, acc0.x); acc0.y = fmaf(aj, gv.y, acc0.y);
                        acc0.z = fmaf(aj, gv.z, acc0.z); acc0.w = fmaf(aj, gv.w, acc0.w);
                    }
                    if (act1) {
                        const float4 gv = *(const float4 *)(go + oj + 128);
                        acc1.x = fmaf(aj, gv.x, acc1.x); acc1.y = fmaf(aj, gv.y, acc1.y);
                        acc1.z = fmaf(aj, gv.z, acc1.z); acc1.w = fmaf(aj, gv.w, acc1.w);
                    }
                }
                if (act0) __stcs((float4 *)(gx + u_off + c), acc0);
                if (act1) __stcs((float4 *)(gx + u_off + c + 128), acc1);
            }
        }
        __syncthreads();
        if (STAGED && threadIdx.x == 0 && b + (int)gridDim.x < p.B) issue_g(b + gridDim.x);   // phase 2 is done with grad_out
    }
}

// ---------------------------------------------------------------------------------------------
// Grouped staged kernels (round 2): road networks whose slabs fit shared memory, C = 128 or 256.
// ---------------------------------------------------------------------------------------------
// Differences from the kernels above, all about the instruction and latency budget the ncu captures
// showed (67 % issue slots, long-scoreboard stalls on the topology loads, an exposed grad_out copy):
//  * a staged unit is a GROUP of G consecutive graphs (one contiguous slab in HBM), sized to ~96 KB (<= 100 KB)
//    whatever the head count: one head (24 KB per graph) gets G = 4, so the per-unit barrier, TMA issue
//    and mbarrier wait are paid once per 96 (graph, node, head) pairs instead of once per 24;
//  * the topology (in-/out-link lists) is copied to shared memory once per CTA;
//  * a pair's (coefficient, source-row offset) list goes through an 8-entry shared-memory scratch of
//    the warp, read back as one broadcast LDS.64 per link: no shuffles (and no divergence checks around
//    them) in the gather loops;
//  * backward: grad_out is NOT staged ahead by TMA any more (that copy was exposed at the top of every
//    graph).  Phase 1 reads each grad_out row exactly once - the row of the pair it is working on -
//    straight from HBM (the next pair's row is already in flight in registers) and drops it into shared
//    memory for the gathers of phase 2; xw and the attention terms of the NEXT group arrive by TMA during
//    phase 2, for which alpha has been copied aside.
constexpr int kGrpThreads = 1024;

struct GrpGeom {
    int G, per_sm, threads;
    size_t slab_bytes, att_bytes, topo_bytes, scratch_bytes, total_fwd, total_bwd, graw_bytes;
};

// per_sm = 2: half-size groups, 16 warps, two CTAs per SM.  The two CTAs are in different phases most of the
// time, so one's exposed waits (the first group's copy, the phase-1 / phase-2 hand-over) are covered by the
// other's traffic; taken when a graph's slab is at most 50 KB (one head of 256 channels, four of 128 channels)
inline GrpGeom grp_geom(int N, int E, int H, int C, int per_sm)
{
    GrpGeom g;
    const size_t one = (size_t)N * H * C * 4;
    g.per_sm = per_sm;
    g.threads = per_sm == 2 ? 512 : kGrpThreads;
    g.G = (int)std::min<size_t>(8, std::max<size_t>(1, ((per_sm == 2 ? 50 : 100) * 1024) / one));
    g.slab_bytes = one * g.G;
    g.att_bytes = 2 * gat_pad16((size_t)g.G * N * H * 4) + 2 * gat_pad16((size_t)g.G * (E + N) * H * 4);   // + alpha (backward)
    g.topo_bytes = gat_pad16((size_t)(2 * (N + 1) + 4 * E) * 4);
    g.scratch_bytes = 32 * 64;
    g.graw_bytes = 2 * gat_pad16((size_t)g.G * (E + N) * H * 4);   // raw-logit grads + alpha copy
    g.total_fwd = 128 + g.topo_bytes + g.scratch_bytes + 2 * (g.slab_bytes + g.att_bytes);
    g.total_bwd = 128 + g.topo_bytes + g.scratch_bytes + g.graw_bytes + 2 * g.slab_bytes + g.att_bytes;
    return g;
}

template <int CPL>   // float4 chunks per lane: C = 128 * CPL
__global__ void __launch_bounds__(kGrpThreads) gat_fwd_grp_kernel(GatParams p, int G, size_t att_bytes, size_t topo_bytes)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int C = 128 * CPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int N = p.N, E = p.E, H = p.H, NH = N * H, EH = (E + N) * H;
    uint64_t *bars = (uint64_t *)smem;
    int *in_ptr = (int *)(smem + 128), *in_src = in_ptr + (N + 1), *in_eid = in_src + E;
    float2 *scr = (float2 *)(smem + 128 + topo_bytes) + warp * 8;
    const size_t slab_bytes = (size_t)G * NH * C * 4, stage_bytes = slab_bytes + att_bytes;
    unsigned char *stage[2] = {smem + 128 + topo_bytes + 32 * 64, smem + 128 + topo_bytes + 32 * 64 + stage_bytes};
    const int groups = (p.B + G - 1) / G;
    auto issue = [&](int buf, int group) {   // thread 0 only
        const int g0 = group * G, gc = min(G, p.B - g0);
        unsigned char *dst = stage[buf];
        mbar_expect_tx(&bars[buf], (uint32_t)((size_t)gc * (NH * C + 2 * NH + EH) * 4));
        tma_bulk_g2s(dst, p.xw + (size_t)g0 * NH * C, (uint32_t)((size_t)gc * NH * C * 4), &bars[buf]);
        unsigned char *q = dst + slab_bytes;
        tma_bulk_g2s(q, p.a_src + (size_t)g0 * NH, (uint32_t)((size_t)gc * NH * 4), &bars[buf]);
        q += gat_pad16((size_t)G * NH * 4);
        tma_bulk_g2s(q, p.a_dst + (size_t)g0 * NH, (uint32_t)((size_t)gc * NH * 4), &bars[buf]);
        q += gat_pad16((size_t)G * NH * 4);
        tma_bulk_g2s(q, p.a_edge + (size_t)g0 * EH, (uint32_t)((size_t)gc * EH * 4), &bars[buf]);
    };
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    for (int i = threadIdx.x; i < N + 1; i += blockDim.x) in_ptr[i] = __ldg(p.in_ptr + i);
    for (int i = threadIdx.x; i < E; i += blockDim.x) { in_src[i] = __ldg(p.in_src + i); in_eid[i] = __ldg(p.in_eid + i); }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x < groups) issue(0, blockIdx.x);
    int it = 0;
    for (int group = blockIdx.x; group < groups; group += gridDim.x, ++it) {
        const int s = it & 1, ng = group + gridDim.x;
        if (threadIdx.x == 0 && ng < groups) issue(s ^ 1, ng);   // prefetch the next group into the other buffer
        mbar_wait(&bars[s], (it >> 1) & 1);
        const int g0 = group * G, gc = min(G, p.B - g0), pairs = gc * NH;
        const float *x = (const float *)stage[s];
        const float *as = (const float *)(stage[s] + slab_bytes);
        const float *ad = (const float *)((const unsigned char *)as + gat_pad16((size_t)G * NH * 4));
        const float *ae = (const float *)((const unsigned char *)ad + gat_pad16((size_t)G * NH * 4));
        float *al = p.alpha_out ? p.alpha_out + (size_t)g0 * EH : nullptr;
        float *o = p.out + (size_t)g0 * NH * C + lane * 4;
        for (int pair = warp; pair < pairs; pair += W) {
            const int g = pair / NH, vh = pair - g * NH, v = vh / H, h = vh - v * H;
            const int k0 = in_ptr[v], deg = in_ptr[v + 1] - k0;
            float logit = -INFINITY;
            int u = v, row = E + v;
            if (lane < deg) { u = in_src[k0 + lane]; row = in_eid[k0 + lane]; }
            if (lane <= deg) logit = leaky(as[g * NH + u * H + h] + ad[g * NH + vh] + ae[g * EH + row * H + h], p.slope);
            float mx = logit;
#pragma unroll
            for (int sft = 4; sft > 0; sft >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sft));
            mx = __shfl_sync(0xffffffffu, mx, 0);      // lanes beyond 8 hold -inf groups of their own
            const float ex = lane <= deg ? __expf(logit - mx) : 0.0f;
            float sum = ex;
#pragma unroll
            for (int sft = 4; sft > 0; sft >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sft);
            sum = __shfl_sync(0xffffffffu, sum, 0);
            const float a = ex / (sum + 1e-16f);
            if (lane <= deg) {
                scr[lane] = make_float2(a, __int_as_float((g * NH + u * H + h) * C));
                if (al) al[g * EH + row * H + h] = a;
            }
            __syncwarp();
            float4 acc[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
            for (int j = 0; j <= deg; ++j) {
                const float2 t = scr[j];   // broadcast: (coefficient, float offset of the source row)
                const float *xr = x + __float_as_int(t.y) + lane * 4;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const float4 xv = *(const float4 *)(xr + c * 128);
                    acc[c].x = fmaf(t.x, xv.x, acc[c].x); acc[c].y = fmaf(t.x, xv.y, acc[c].y);
                    acc[c].z = fmaf(t.x, xv.z, acc[c].z); acc[c].w = fmaf(t.x, xv.w, acc[c].w);
                }
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < CPL; ++c) __stcs((float4 *)(o + (size_t)pair * C + c * 128), acc[c]);   // never re-read here
        }
        __syncthreads();  // everyone is done with this buffer before it is refilled
    }
}

template <int CPL>
__global__ void __launch_bounds__(kGrpThreads) gat_bwd_grp_kernel(GatParams p, int G, size_t topo_bytes, size_t graw_bytes)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int C = 128 * CPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int N = p.N, E = p.E, H = p.H, NH = N * H, EH = (E + N) * H;
    uint64_t *bar = (uint64_t *)smem;
    int *in_ptr = (int *)(smem + 128), *in_src = in_ptr + (N + 1), *in_eid = in_src + E;
    int *out_ptr = in_eid + E, *out_dst = out_ptr + (N + 1), *out_eid = out_dst + E;
    float2 *scr = (float2 *)(smem + 128 + topo_bytes) + warp * 8;
    float *graw = (float *)(smem + 128 + topo_bytes + 32 * 64);         // raw-logit gradients [G][(E+N)][H]
    float *al_s = graw + gat_pad16((size_t)G * EH * 4) / 4;              // alpha of this group, kept for phase 2
    const size_t slab = (size_t)G * NH * C;                              // floats
    float *xs = (float *)((unsigned char *)graw + graw_bytes);           // staged xw (TMA)
    float *gs = xs + slab;                                               // grad_out, filled by phase 1
    float *att = gs + slab;                                              // a_src | a_dst | a_edge | alpha (TMA, with xw)
    const size_t nh_pad = gat_pad16((size_t)G * NH * 4) / 4, eh_pad = gat_pad16((size_t)G * EH * 4) / 4;
    const int groups = (p.B + G - 1) / G;
    auto issue = [&](int group) {   // thread 0: xw slab + attention terms of a group
        const int g0 = group * G, gc = min(G, p.B - g0);
        mbar_expect_tx(bar, (uint32_t)((size_t)gc * (NH * C + 2 * NH + 2 * EH) * 4));
        tma_bulk_g2s(xs, p.xw + (size_t)g0 * NH * C, (uint32_t)((size_t)gc * NH * C * 4), bar);
        tma_bulk_g2s(att, p.a_src + (size_t)g0 * NH, (uint32_t)((size_t)gc * NH * 4), bar);
        tma_bulk_g2s(att + nh_pad, p.a_dst + (size_t)g0 * NH, (uint32_t)((size_t)gc * NH * 4), bar);
        tma_bulk_g2s(att + 2 * nh_pad, p.a_edge + (size_t)g0 * EH, (uint32_t)((size_t)gc * EH * 4), bar);
        tma_bulk_g2s(att + 2 * nh_pad + eh_pad, p.alpha_in + (size_t)g0 * EH, (uint32_t)((size_t)gc * EH * 4), bar);
    };
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    for (int i = threadIdx.x; i < N + 1; i += blockDim.x) { in_ptr[i] = __ldg(p.in_ptr + i); out_ptr[i] = __ldg(p.out_ptr + i); }
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
        in_src[i] = __ldg(p.in_src + i); in_eid[i] = __ldg(p.in_eid + i);
        out_dst[i] = __ldg(p.out_dst + i); out_eid[i] = __ldg(p.out_eid + i);
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x < groups) issue(blockIdx.x);
    int it = 0;
    for (int group = blockIdx.x; group < groups; group += gridDim.x, ++it) {
        const int g0 = group * G, gc = min(G, p.B - g0), pairs = gc * NH;
        const float *go = p.grad_out + (size_t)g0 * NH * C + lane * 4;
        float *g_as = p.grad_a_src + (size_t)g0 * NH, *g_ad = p.grad_a_dst + (size_t)g0 * NH;
        float *g_ae = p.grad_a_edge + (size_t)g0 * EH;
        float *gx = p.grad_xw + (size_t)g0 * NH * C + lane * 4;
        const float *as = att, *ad = att + nh_pad, *ae = att + 2 * nh_pad, *al = att + 2 * nh_pad + eh_pad;
        // ---- phase 1: per target (graph, node, head): d alpha, softmax / leaky-relu backward ----------
        float4 gr4[CPL], nx4[CPL];   // grad_out row of this pair / of the warp's next pair (in flight)
        if (warp < pairs) {
#pragma unroll
            for (int c = 0; c < CPL; ++c) nx4[c] = __ldcs((const float4 *)(go + (size_t)warp * C + c * 128));
        }
        mbar_wait(bar, it & 1);
        for (int pair = warp; pair < pairs; pair += W) {
#pragma unroll
            for (int c = 0; c < CPL; ++c) gr4[c] = nx4[c];
            if (pair + W < pairs) {
#pragma unroll
                for (int c = 0; c < CPL; ++c) nx4[c] = __ldcs((const float4 *)(go + (size_t)(pair + W) * C + c * 128));
            }
#pragma unroll
            for (int c = 0; c < CPL; ++c) *(float4 *)(gs + (size_t)pair * C + c * 128 + lane * 4) = gr4[c];   // for phase 2
            const int g = pair / NH, vh = pair - g * NH, v = vh / H, h = vh - v * H;
            const int k0 = in_ptr[v], deg = in_ptr[v + 1] - k0;
            int u = v, row = E + v;
            if (lane < deg) { u = in_src[k0 + lane]; row = in_eid[k0 + lane]; }
            if (lane <= deg) scr[lane] = make_float2(0.0f, __int_as_float((g * NH + u * H + h) * C));
            __syncwarp();
            // d alpha_j = <grad_out[v,h,:], xw[u_j,h,:]>: every lane forms its share of all (<= 8) dot
            // products, then ONE transposing butterfly leaves the total of link j in lanes 4j .. 4j+3
            float part[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                part[j] = 0.0f;
                if (j <= deg) {   // warp-uniform
                    const float *xr = xs + __float_as_int(scr[j].y) + lane * 4;
#pragma unroll
                    for (int c = 0; c < CPL; ++c) {
                        const float4 xv = *(const float4 *)(xr + c * 128);
                        part[j] = fmaf(gr4[c].x, xv.x, fmaf(gr4[c].y, xv.y, fmaf(gr4[c].z, xv.z, fmaf(gr4[c].w, xv.w, part[j]))));
                    }
                }
            }
            const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
            float q4[4], q2[2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                q4[i] = (b4 ? part[i + 4] : part[i]) + __shfl_xor_sync(0xffffffffu, b4 ? part[i] : part[i + 4], 16);
#pragma unroll
            for (int i = 0; i < 2; ++i)
                q2[i] = (b3 ? q4[i + 2] : q4[i]) + __shfl_xor_sync(0xffffffffu, b3 ? q4[i] : q4[i + 2], 8);
            float tot = (b2 ? q2[1] : q2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? q2[0] : q2[1], 4);
            tot += __shfl_xor_sync(0xffffffffu, tot, 2);
            tot += __shfl_xor_sync(0xffffffffu, tot, 1);
            __syncwarp();
            if ((lane & 3) == 0) scr[lane >> 2].x = tot;   // link j = lane / 4
            __syncwarp();
            const float dalpha = lane <= deg ? scr[lane].x : 0.0f;
            const float a = lane <= deg ? al[g * EH + row * H + h] : 0.0f;
            float dot = a * dalpha;
#pragma unroll
            for (int sft = 4; sft > 0; sft >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, sft);
            dot = __shfl_sync(0xffffffffu, dot, 0);
            float gr = 0.0f;
            if (lane <= deg) {
                const float raw = as[g * NH + u * H + h] + ad[g * NH + vh] + ae[g * EH + row * H + h];
                gr = a * (dalpha - dot) * (raw > 0.0f ? 1.0f : p.slope);
                g_ae[g * EH + row * H + h] = gr;
                graw[g * EH + row * H + h] = gr;
                al_s[g * EH + row * H + h] = a;
            }
            float gsum = gr;
#pragma unroll
            for (int sft = 4; sft > 0; sft >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, sft);
            if (lane == 0) g_ad[g * NH + vh] = gsum;
            __syncwarp();
        }
        __syncthreads();
        if (threadIdx.x == 0 && group + (int)gridDim.x < groups) issue(group + gridDim.x);   // phase 1 is done with xw / att
        // ---- phase 2: per source (graph, node, head): grad_xw = sum over out-links alpha * grad_out[dst] ----
        for (int pair = warp; pair < pairs; pair += W) {
            const int g = pair / NH, uh = pair - g * NH, u = uh / H, h = uh - u * H;
            const int k0 = out_ptr[u], deg = out_ptr[u + 1] - k0;
            int dst = u, row = E + u;
            if (lane < deg) { dst = out_dst[k0 + lane]; row = out_eid[k0 + lane]; }
            float gsum = 0.0f;
            if (lane <= deg) {
                gsum = graw[g * EH + row * H + h];
                scr[lane] = make_float2(al_s[g * EH + row * H + h], __int_as_float((g * NH + dst * H + h) * C));
            }
#pragma unroll
            for (int sft = 4; sft > 0; sft >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, sft);
            if (lane == 0) g_as[g * NH + uh] = gsum;
            __syncwarp();
            float4 acc[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
            for (int j = 0; j <= deg; ++j) {
                const float2 t = scr[j];
                const float *gr_ = gs + __float_as_int(t.y) + lane * 4;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const float4 gv = *(const float4 *)(gr_ + c * 128);
                    acc[c].x = fmaf(t.x, gv.x, acc[c].x); acc[c].y = fmaf(t.x, gv.y, acc[c].y);
                    acc[c].z = fmaf(t.x, gv.z, acc[c].z); acc[c].w = fmaf(t.x, gv.w, acc[c].w);
                }
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < CPL; ++c) __stcs((float4 *)(gx + (size_t)pair * C + c * 128), acc[c]);
        }
        __syncthreads();   // gs / graw / al_s are free for the next group
    }
}

// geometry of the grouped kernels for this network: two CTAs per SM when they fit (B200_GAT_CTAS=1|2 forces)
inline GrpGeom grp_pick(const b200_graph *g, int H, int C, bool bwd)
{
    const char *force = getenv("B200_GAT_CTAS");
    const GrpGeom two = grp_geom(g->N, g->E, H, C, 2), one = grp_geom(g->N, g->E, H, C, 1);
    const size_t need = bwd ? two.total_bwd : two.total_fwd;
    const bool fits2 = C == 256 && (size_t)g->N * H * C * 4 <= 50 * 1024 && 2 * (need + 1024) <= 227 * 1024;   // 128 channels: measured 1-3 % slower
    if (force && force[0] == '1') return one;
    if (force && force[0] == '2' && (size_t)g->N * H * C * 4 <= 50 * 1024 && 2 * (need + 1024) <= 227 * 1024) return two;
    return fits2 ? two : one;
}

// one warp per (node, head) pair of a graph, at most 32 warps: a CTA never carries idle warps, and
// small slabs (one head) leave room for a second CTA on the SM
int gat_threads(int N, int H) { return 32 * std::min(32, std::max(1, N * H)); }

int check_common(const b200_graph *g, int B, int H, int C)
{
    B200_REQUIRE(g && B >= 0 && H > 0 && C > 0, "bad argument");
    B200_REQUIRE(C % 4 == 0, "channels per head must be a multiple of 4");
    B200_REQUIRE(g->max_in_deg + 1 <= kMaxDeg, "in-degree too large for the GAT kernel");
    return 0;
}

GatParams base(const b200_graph *g, int B, int H, int C)
{
    GatParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.N = g->N; p.E = g->E; p.H = H; p.C = C;
    p.in_ptr = g->topo + g->off_in_ptr; p.in_src = g->topo + g->off_in_src; p.in_eid = g->topo + g->off_in_eid;
    p.out_ptr = g->out_ptr; p.out_dst = g->out_dst; p.out_eid = g->out_eid;
    return p;
}

}  // namespace

extern "C" int b200_gat_edge_softmax_aggregate_fwd(const b200_graph *g, int B, int H, int C, const float *xw,
                                                   const float *a_src, const float *a_dst, const float *a_edge,
                                                   float negative_slope, float *out, float *alpha, void *stream)
{
    int rc = check_common(g, B, H, C);
    if (rc) return rc;
    if (B == 0) return 0;
    B200_REQUIRE(xw && a_src && a_dst && a_edge && out, "null array argument");
    GatParams p = base(g, B, H, C);
    p.xw = xw; p.a_src = a_src; p.a_dst = a_dst; p.a_edge = a_edge; p.out = out; p.alpha_out = alpha;
    p.slope = negative_slope;
    cudaStream_t st = (cudaStream_t)stream;
    const bool narrow = g->max_in_deg + 1 <= 8;
    const bool att16 = ((size_t)g->N * H * 4) % 16 == 0 && ((size_t)(g->E + g->N) * H * 4) % 16 == 0;
    if (narrow && att16 && (C == 128 || C == 256) && !getenv("B200_GAT_SLAB")) {   // grouped staged path
        const GrpGeom gg = grp_pick(g, H, C, false);
        if (gg.total_fwd <= 226 * 1024) {
            auto kern = C == 256 ? gat_fwd_grp_kernel<2> : gat_fwd_grp_kernel<1>;
            B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gg.total_fwd));
            B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            const int groups = (B + gg.G - 1) / gg.G;
            kern<<<std::min(groups, g->num_sms * gg.per_sm), gg.threads, gg.total_fwd, st>>>(p, gg.G, gg.att_bytes, gg.topo_bytes);
            B200_CUDA(cudaGetLastError());
            return 0;
        }
    }
    const size_t smem = 128 + 2 * gat_stage_bytes(g->N, g->E, H, C);
    const int threads = gat_threads(g->N, H);
    if (smem <= 226 * 1024) {
        auto kern = narrow ? gat_fwd_kernel<true, 8> : gat_fwd_kernel<true, 32>;
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
        const int grid = std::min(B, std::max(per_sm, 1) * g->num_sms);
        kern<<<grid, threads, smem, st>>>(p);
    } else {
        auto kern = narrow ? gat_fwd_kernel<false, 8> : gat_fwd_kernel<false, 32>;
        kern<<<std::min(B, g->num_sms * 2), threads, 128, st>>>(p);
    }
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200_gat_edge_softmax_aggregate_bwd(const b200_graph *g, int B, int H, int C, const float *xw,
                                                   const float *a_src, const float *a_dst, const float *a_edge,
                                                   const float *alpha, const float *grad_out, float negative_slope,
                                                   float *grad_xw, float *grad_a_src, float *grad_a_dst,
                                                   float *grad_a_edge, void *stream)
{
    int rc = check_common(g, B, H, C);
    if (rc) return rc;
    if (B == 0) return 0;
    B200_REQUIRE(xw && a_src && a_dst && a_edge && alpha && grad_out && grad_xw && grad_a_src && grad_a_dst && grad_a_edge,
                 "null array argument");
    GatParams p = base(g, B, H, C);
    p.xw = xw; p.a_src = a_src; p.a_dst = a_dst; p.a_edge = a_edge; p.alpha_in = alpha; p.grad_out = grad_out;
    p.grad_xw = grad_xw; p.grad_a_src = grad_a_src; p.grad_a_dst = grad_a_dst; p.grad_a_edge = grad_a_edge;
    p.slope = negative_slope;
    cudaStream_t st = (cudaStream_t)stream;
    {
        const bool narrow = g->max_in_deg + 1 <= 8;
        bool narrow_out = true;   // out-degree + self loop <= 8 as well (phase 2 uses the 8-entry scratch)
        const bool att16 = ((size_t)g->N * H * 4) % 16 == 0 && ((size_t)(g->E + g->N) * H * 4) % 16 == 0;
        if (narrow && att16 && (C == 128 || C == 256) && !getenv("B200_GAT_SLAB")) {
            narrow_out = g->max_out_deg + 1 <= 8;   // host copy of the degree: no device read-back on the training path
            const GrpGeom gg = grp_pick(g, H, C, true);
            if (narrow_out && gg.total_bwd <= 226 * 1024) {
                auto kern = C == 256 ? gat_bwd_grp_kernel<2> : gat_bwd_grp_kernel<1>;
                B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gg.total_bwd));
                B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                const int groups = (B + gg.G - 1) / gg.G;
                kern<<<std::min(groups, g->num_sms * gg.per_sm), gg.threads, gg.total_bwd, st>>>(p, gg.G, gg.topo_bytes, gg.graw_bytes);
                B200_CUDA(cudaGetLastError());
                return 0;
            }
        }
    }
    const size_t slab_bytes = (size_t)g->N * H * C * 4;
    const size_t graw_bytes = ((((size_t)(g->E + g->N) * H + 31) & ~(size_t)31)) * 4;
    const size_t att_bytes = ((size_t)2 * g->N * H + 2 * (size_t)(g->E + g->N) * H) * 4 + 64;
    const size_t smem_staged = 128 + graw_bytes + 2 * slab_bytes + att_bytes;
    if (smem_staged <= 226 * 1024) {
        B200_CUDA(cudaFuncSetAttribute(gat_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_staged));
        int per_sm = 0;
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gat_bwd_kernel<true>, gat_threads(g->N, H), smem_staged));
        gat_bwd_kernel<true><<<std::min(B, std::max(per_sm, 1) * g->num_sms), gat_threads(g->N, H), smem_staged, st>>>(p);
    } else {
        const size_t smem = 128 + graw_bytes;
        B200_REQUIRE(smem <= 220 * 1024, "network too large for the GAT backward kernel");
        B200_CUDA(cudaFuncSetAttribute(gat_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gat_bwd_kernel<false><<<std::min(B, g->num_sms * 2), gat_threads(g->N, H), smem, st>>>(p);
    }
    B200_CUDA(cudaGetLastError());
    return 0;
}
