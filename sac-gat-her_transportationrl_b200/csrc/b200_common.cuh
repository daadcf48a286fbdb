// Shared declarations of the B200 (sm_100a) assignment kernels: the graph handle, error
// plumbing, TMA bulk-copy / mbarrier helpers and the BPR cost function.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdio.h>

#include "../../include/b200_assign.h"

#define B200_VERSION 200

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void b200_set_error(const char *fmt, ...);

#define B200_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t err__ = (call);                                                              \
        if (err__ != cudaSuccess) {                                                              \
            b200_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
            return (int)err__;                                                                   \
        }                                                                                        \
    } while (0)

#define B200_REQUIRE(cond, msg)                                  \
    do {                                                         \
        if (!(cond)) {                                           \
            b200_set_error("%s:%d %s", __FILE__, __LINE__, msg); \
            return B200_EINVAL;                                  \
        }                                                        \
    } while (0)

// ---------------------------------------------------------------------------------------------
// graph handle: static network resident in HBM, shared by every scenario of every batch
// ---------------------------------------------------------------------------------------------
// Topology blob (one contiguous, 16-byte aligned device allocation so a CTA can stage it into
// shared memory with ONE cp.async.bulk):
//   int32 in_ptr[N+1 (+pad)]   CSC: in-links of node v are slots in_ptr[v] .. in_ptr[v+1]-1
//   int32 in_src[E (+pad)]     tail node of the link in CSC slot k
//   int32 in_eid[E (+pad)]     original link id (TNTP row order) of CSC slot k
//   int32 out_ptr[N+1 (+pad)] | out_dst[E (+pad)]   out-links by tail node (frontier expansion)
// In-links of a node are stored in ascending tail-node order.
struct b200_graph {
    int device;
    int N, E, O, num_od, max_in_deg, max_out_deg, num_sms;
    double total_demand;
    // topology blob + offsets (in int32 elements)
    int32_t *topo;
    int topo_ints;  // padded to a multiple of 4 ints (16 B)
    int off_in_ptr, off_in_src, off_in_eid;
    // appended for the general fused kernel: out-link CSR (out_ptr[N+1] | out_dst[E]), same blob so
    // one TMA copy stages everything; gen_intq = all demands are integers (total < 2^31)
    int off_out_ptr, off_out_dst, gen_intq;
    int kahn_pack;   // integer demands < 2^28 per origin, out-degree <= 15: packed subtree words (warp_aon_chain)
    // 16-bit copy of the whole blob for the general assignment kernels (half the shared memory);
    // offsets in uint16 elements, every array padded to 16 bytes
    uint16_t *topo16;
    int topo16_bytes, o16_in_src, o16_in_eid, o16_out_ptr, o16_out_dst;
    // general fused kernel: previous-iteration trees, u16 [resident CTAs][O][N], grown on demand
    uint16_t *tree_ws;
    size_t tree_ws_bytes;
    int warm_start;   // 0 disables the warm start (env B200_NO_WARM, A/B measurements)
    int cluster_limit;   // max thread-block cluster size of the general kernel (env B200_CLUSTER; 1 = off, 0 = auto)
    int last_cluster;    // cluster size of the last fused launch (tests / tools)
    int last_warps;      // warps per CTA of the last fused launch
    int last_cch;        // 1 when the last b200_assign took its shortest paths from the contraction hierarchy
    // per-link attributes, TNTP order
    int32_t *row, *col;
    float *t0, *cap0;
    // per-CSC-slot copies (gather once at upload instead of once per iteration)
    float *t0_csc, *cap0_csc;
    // demand, CSR by origin slot: od_ptr[O+1], od_dst[num_od], od_q (double + float copies)
    int32_t *origins, *origin_slot, *od_ptr, *od_dst;
    double *od_q64;
    float *od_q32;
    // out-link CSR (used by the state-feature BFS) and GAT incoming lists incl. self loops
    int32_t *out_ptr, *out_dst, *out_eid;
    float max_t0, max_cap0, min_t0;
    int no_cch;   // env B200_NO_CCH: shortest paths by Gauss-Seidel sweeps / label correcting everywhere (A/B, parity tests)
    int last_launches;
    size_t smem_f32, smem_f64;
    int warps_f32, warps_f64;
    int32_t *origins_host;
    // small-network kernel (b200_assign_small.cuh): per-precision TMA blob (node records + dense
    // demand), per-CSC-slot link id / padded cost index; force_general = 1 routes b200_assign
    // through the general kernel (test hook, env B200_FORCE_GENERAL)
    unsigned char *small_blob32, *small_blob64, *small_blob_i32;
    uint8_t *small_eid, *small_wpos;
    int small_ok, small_D, small_intq, force_general;
    // network-specialised kernel (b200_assign_static.cuh): set when the link list equals a compiled-in
    // topology header; internal origin positions and the TMA blobs (dense demand Q[internal node][32
    // lanes] followed by per-CSC-slot t0 / cap0 / link id / padded cost index tables).  force_small = 1 (env B200_FORCE_SMALL)
    // routes b200_assign through the generic lane-per-origin kernel instead (test hook)
    int static_ok, st_intq, force_small;
    void *plugin_launch;   // launcher of the on-demand plugin that matched this topology (else built-in)
    int32_t *st_origpos;
    unsigned char *st_blob32, *st_blob64, *st_blob_i32, *st_blob_grp;
    int grp_force;   // B200_STATIC_NO_GRP=0: grouped static kernel whatever the batch size (tests)
    // customizable contraction hierarchy of the general fused kernel (b200_cch.cu): the down row stream
    // (batches of 1 + D rows of 32 words), one up row stream per origin (row pairs, up_off[O + 1]), the select stream,
    // customization tables (initial weight per filled edge, rows of 32 edges
    // with their lower triangles); W positions: [0, cch_wd) down edges + one zero-weight slot, then cch_wu up-region entries,
    // then the +inf and the write-only dummy slot.  cch_ws: per-CTA weight workspace in HBM for the
    // part of W that does not fit shared memory, grown on demand like tree_ws.
    int cch_ok, cch_F, cch_wd, cch_wu, cch_slots, cch_dn_batches, cch_cz_rows, cch_cz_levels, cch_tris, cch_wmode, cch_leaf_alias;
    double cch_exact_limit;
    uint32_t *cch_dn_rows, *cch_up_rows, *cch_sel_rows, *cch_up_off, *cch_cz_init, *cch_cz_desc, *cch_cz_tab, *cch_cz_lvl;
    int cch_D, cch_SD;   // edge rows per down batch, rows per block of the select stream
    double *cch_ws;
    size_t cch_ws_bytes, smem_cch;
    int warps_cch;
};

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#define B200_INF_F64 __longlong_as_double(0x7ff0000000000000LL)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// TMA 1-D bulk copy global -> shared (SASS: UBLKCP), completion counted on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// BPR link cost, the reference's operation order (reference: src/env/repair_env.py:670-676):
// divide, clip to [0,10], power, times alpha, plus one, times t0.  beta == 4 is two squarings;
// every product/sum is a separately rounded IEEE operation (no FMA), matching the oracle.
template <typename REAL>
struct Bpr;

template <>
struct Bpr<float> {
    __device__ __forceinline__ static float eval(float flow, float cap, float t0, float alpha, float beta, bool beta4)
    {
        float c = fmaxf(cap, 1e-6f);
        float vc = __fdiv_rn(flow, c);
        vc = (vc < 0.0f) ? 0.0f : vc;  // NaN falls through like np.clip
        vc = (vc > 10.0f) ? 10.0f : vc;
        float p;
        if (beta4) {
            float s = __fmul_rn(vc, vc);
            p = __fmul_rn(s, s);
        } else {
            p = powf(vc, beta);
        }
        float g = __fadd_rn(1.0f, __fmul_rn(alpha, p));
        return __fmul_rn(t0, g);
    }
};

template <>
struct Bpr<double> {
    __device__ __forceinline__ static double eval(double flow, double cap, double t0, double alpha, double beta, bool beta4)
    {
        double c = fmax(cap, 1e-6);
        double vc = __ddiv_rn(flow, c);
        vc = (vc < 0.0) ? 0.0 : vc;
        vc = (vc > 10.0) ? 10.0 : vc;
        double p;
        if (beta4) {
            double s = __dmul_rn(vc, vc);
            p = __dmul_rn(s, s);
        } else {
            p = pow(vc, beta);
        }
        double g = __dadd_rn(1.0, __dmul_rn(alpha, p));
        return __dmul_rn(t0, g);
    }
};

__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
