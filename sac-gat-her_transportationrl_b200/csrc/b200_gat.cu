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

constexpr int kGatThreads = 256;
constexpr int kMaxDeg = 31;  // incoming links + self loop handled per (node, head)

struct GatParams {
    int B, N, E, H, C;
    const int32_t *in_ptr, *in_src, *in_eid, *out_ptr, *out_dst, *out_eid;
    const float *xw, *a_src, *a_dst, *a_edge, *alpha_in, *grad_out;
    float *out, *alpha_out, *grad_xw, *grad_a_src, *grad_a_dst, *grad_a_edge;
    float slope;
};

__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.0f ? x : slope * x; }

template <bool STAGED>
__global__ void __launch_bounds__(kGatThreads) gat_fwd_kernel(GatParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int N = p.N, E = p.E, H = p.H, C = p.C;
    const size_t slab = (size_t)N * H * C;           // floats per graph
    uint64_t *bars = (uint64_t *)smem;                // 2 mbarriers
    float *stage[2] = {(float *)(smem + 128), (float *)(smem + 128) + slab};
    if (STAGED) {
        if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
        __syncthreads();
        if (threadIdx.x == 0 && blockIdx.x < p.B) {
            mbar_expect_tx(&bars[0], (uint32_t)(slab * 4));
            tma_bulk_g2s(stage[0], p.xw + (size_t)blockIdx.x * slab, (uint32_t)(slab * 4), &bars[0]);
        }
    }
    int it = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
        const float *x;
        if (STAGED) {
            const int s = it & 1, nb = b + gridDim.x;
            if (threadIdx.x == 0 && nb < p.B) {       // prefetch the next graph into the other buffer
                mbar_expect_tx(&bars[s ^ 1], (uint32_t)(slab * 4));
                tma_bulk_g2s(stage[s ^ 1], p.xw + (size_t)nb * slab, (uint32_t)(slab * 4), &bars[s ^ 1]);
            }
            mbar_wait(&bars[s], (it >> 1) & 1);
            x = stage[s];
        } else {
            x = p.xw + (size_t)b * slab;
        }
        const float *as = p.a_src + (size_t)b * N * H, *ad = p.a_dst + (size_t)b * N * H;
        const float *ae = p.a_edge + (size_t)b * (E + N) * H;
        float *al = p.alpha_out ? p.alpha_out + (size_t)b * (E + N) * H : nullptr;
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
            for (int sft = 16; sft > 0; sft >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sft));
            float ex = lane <= deg ? expf(logit - mx) : 0.0f;
            float sum = ex;
            for (int sft = 16; sft > 0; sft >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sft);
            const float a = ex / (sum + 1e-16f);
            if (al && lane <= deg) al[(size_t)row * H + h] = a;
            // aggregate: each lane owns float4 chunks of the C channels
            for (int c0 = 0; c0 < C; c0 += 128) {     // warp-uniform trip count: shuffles inside
                const int c = c0 + lane * 4;
                const bool act = c < C;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int j = 0; j <= deg; ++j) {
                    const float aj = __shfl_sync(0xffffffffu, a, j);
                    const int uj = __shfl_sync(0xffffffffu, u, j);
                    if (act) {
                        const float4 xv = *(const float4 *)(x + ((size_t)uj * H + h) * C + c);
                        acc.x = fmaf(aj, xv.x, acc.x); acc.y = fmaf(aj, xv.y, acc.y);
                        acc.z = fmaf(aj, xv.z, acc.z); acc.w = fmaf(aj, xv.w, acc.w);
                    }
                }
                if (act) *(float4 *)(o + ((size_t)v * H + h) * C + c) = acc;
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
    if (STAGED) {
        if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
        __syncthreads();
    }
    int it = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
        const float *x = p.xw + (size_t)b * slab, *go = p.grad_out + (size_t)b * slab;
        if (STAGED) {
            if (threadIdx.x == 0) {
                mbar_expect_tx(bar, (uint32_t)(slab * 8));
                tma_bulk_g2s(xs, x, (uint32_t)(slab * 4), bar);
                tma_bulk_g2s(gs, go, (uint32_t)(slab * 4), bar);
            }
            mbar_wait(bar, it & 1);
            x = xs;
            go = gs;
        }
        const float *as = p.a_src + (size_t)b * N * H, *ad = p.a_dst + (size_t)b * N * H;
        const float *ae = p.a_edge + (size_t)b * (E + N) * H;
        const float *al = p.alpha_in + (size_t)b * (E + N) * H;
        float *g_as = p.grad_a_src + (size_t)b * N * H, *g_ad = p.grad_a_dst + (size_t)b * N * H;
        float *g_ae = p.grad_a_edge + (size_t)b * (E + N) * H;
        float *gx = p.grad_xw + (size_t)b * slab;
        for (int pair = warp; pair < N * H; pair += W) {
            const int v = pair / H, h = pair - v * H;
            const int k0 = __ldg(p.in_ptr + v), deg = __ldg(p.in_ptr + v + 1) - k0;
            int u = v, row = E + v;
            if (lane < deg) { u = __ldg(p.in_src + k0 + lane); row = __ldg(p.in_eid + k0 + lane); }
            // d alpha_j = <grad_out[v,h,:], xw[u_j,h,:]>, computed cooperatively, kept by lane j
            float dalpha = 0.0f;
            for (int j = 0; j <= deg; ++j) {
                const int uj = __shfl_sync(0xffffffffu, u, j);
                float part = 0.0f;
                for (int c = lane * 4; c < C; c += 128) {
                    const float4 gv = *(const float4 *)(go + ((size_t)v * H + h) * C + c);
                    const float4 xv = *(const float4 *)(x + ((size_t)uj * H + h) * C + c);
                    part = fmaf(gv.x, xv.x, fmaf(gv.y, xv.y, fmaf(gv.z, xv.z, fmaf(gv.w, xv.w, part))));
                }
                for (int sft = 16; sft > 0; sft >>= 1) part += __shfl_xor_sync(0xffffffffu, part, sft);
                if (lane == j) dalpha = part;
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
        for (int pair = warp; pair < N * H; pair += W) {
            const int u = pair / H, h = pair - u * H;
            const int k0 = __ldg(p.out_ptr + u), deg = __ldg(p.out_ptr + u + 1) - k0;
            int dst = u, row = E + u;
            if (lane < deg) { dst = __ldg(p.out_dst + k0 + lane); row = __ldg(p.out_eid + k0 + lane); }
            const float a = lane <= deg ? al[(size_t)row * H + h] : 0.0f;
            float gsum = lane <= deg ? graw[(size_t)row * H + h] : 0.0f;
            for (int sft = 16; sft > 0; sft >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, sft);
            if (lane == 0) g_as[u * H + h] = gsum;
            for (int c0 = 0; c0 < C; c0 += 128) {     // warp-uniform trip count: shuffles inside
                const int c = c0 + lane * 4;
                const bool act = c < C;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int j = 0; j <= deg; ++j) {
                    const float aj = __shfl_sync(0xffffffffu, a, j);
                    const int dj = __shfl_sync(0xffffffffu, dst, j);
                    if (act) {
                        const float4 gv = *(const float4 *)(go + ((size_t)dj * H + h) * C + c);
                        acc.x = fmaf(aj, gv.x, acc.x); acc.y = fmaf(aj, gv.y, acc.y);
                        acc.z = fmaf(aj, gv.z, acc.z); acc.w = fmaf(aj, gv.w, acc.w);
                    }
                }
                if (act) *(float4 *)(gx + ((size_t)u * H + h) * C + c) = acc;
            }
        }
        __syncthreads();
    }
}

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
    const size_t slab_bytes = (size_t)g->N * H * C * 4;
    const size_t smem = 128 + 2 * slab_bytes;
    cudaStream_t st = (cudaStream_t)stream;
    if (smem <= 220 * 1024) {
        B200_CUDA(cudaFuncSetAttribute(gat_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gat_fwd_kernel<true>, kGatThreads, smem));
        const int grid = std::min(B, std::max(per_sm, 1) * g->num_sms);
        gat_fwd_kernel<true><<<grid, kGatThreads, smem, st>>>(p);
    } else {
        gat_fwd_kernel<false><<<std::min(B, g->num_sms * 8), kGatThreads, 128, st>>>(p);
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
    const size_t smem_staged = 128 + graw_bytes + 2 * slab_bytes;
    if (smem_staged <= 220 * 1024) {
        B200_CUDA(cudaFuncSetAttribute(gat_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_staged));
        gat_bwd_kernel<true><<<std::min(B, g->num_sms), kGatThreads, smem_staged, st>>>(p);
    } else {
        const size_t smem = 128 + graw_bytes;
        B200_REQUIRE(smem <= 220 * 1024, "network too large for the GAT backward kernel");
        B200_CUDA(cudaFuncSetAttribute(gat_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gat_bwd_kernel<false><<<std::min(B, g->num_sms * 4), kGatThreads, smem, st>>>(p);
    }
    B200_CUDA(cudaGetLastError());
    return 0;
}
