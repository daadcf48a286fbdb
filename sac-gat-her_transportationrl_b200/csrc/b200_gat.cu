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
#include <string.h>

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
    const size_t smem = 128 + 2 * gat_stage_bytes(g->N, g->E, H, C);
    cudaStream_t st = (cudaStream_t)stream;
    const bool narrow = g->max_in_deg + 1 <= 8;
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
    const size_t slab_bytes = (size_t)g->N * H * C * 4;
    const size_t graw_bytes = ((((size_t)(g->E + g->N) * H + 31) & ~(size_t)31)) * 4;
    cudaStream_t st = (cudaStream_t)stream;
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
