// Static network upload (replaces the arrays RepairEnv.__init__ builds, reference:
// src/env/repair_env.py:85-97) + error plumbing of the C ABI.
#include "b200_common.cuh"

#include <algorithm>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

static thread_local char g_err[512] = "";

void b200_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *b200_last_error(void) { return g_err; }
extern "C" int b200_version(void) { return B200_VERSION; }

template <typename T>
static int upload(T **dst, const std::vector<T> &src)
{
    size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    B200_CUDA(cudaMalloc((void **)dst, bytes));
    if (!src.empty()) B200_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

static int pad4(int x) { return (x + 3) & ~3; }

size_t b200_fused_smem_bytes(const b200_graph *g, int prec, int warps);  // b200_assign.cu
int b200_pick_warps(const b200_graph *g, int prec, size_t *smem_out);   // b200_assign.cu
int b200_pick_warps_cch(b200_graph *g);                                 // b200_assign.cu
int b200_cch_prepare(b200_graph *g, const int32_t *in_ptr, const int32_t *in_src);   // b200_cch.cu
void b200_cch_destroy(b200_graph *g);
int b200_small_prepare(b200_graph *g, const int32_t *in_ptr, const int32_t *in_src, const int32_t *in_eid,
                       const float *t0, int num_od, const int32_t *od_o, const int32_t *od_d, const double *od_q,
                       const int32_t *origin_slot);
int b200_static_prepare(b200_graph *g, const int32_t *row, const int32_t *col, const float *t0, const float *cap0,
                        int num_od, const int32_t *od_o, const int32_t *od_d, const double *od_q,
                        const int32_t *origin_slot);

extern "C" int b200_graph_create(int device, int N, int E, const int32_t *row, const int32_t *col, const float *t0,
                                 const float *cap0, int num_od, const int32_t *od_o, const int32_t *od_d,
                                 const double *od_q, b200_graph **out)
{
    B200_REQUIRE(out != nullptr, "out handle pointer is null");
    B200_REQUIRE(N > 0 && E > 0 && num_od >= 0, "need N > 0, E > 0, num_od >= 0");
    B200_REQUIRE(row && col && t0 && cap0, "null link arrays");
    B200_REQUIRE(num_od == 0 || (od_o && od_d && od_q), "null OD arrays");
    for (int e = 0; e < E; ++e) {
        B200_REQUIRE(row[e] >= 0 && row[e] < N && col[e] >= 0 && col[e] < N, "link endpoint out of range");
        B200_REQUIRE(row[e] != col[e], "self loops are not supported");
        B200_REQUIRE(t0[e] > 0.0f, "free-flow times must be > 0 (strictly positive link costs)");
    }
    // CSC, in-links sorted by tail node; reject parallel links (scipy would sum their costs)
    std::vector<int32_t> in_ptr(N + 1, 0), in_src(E), in_eid(E), out_ptr(N + 1, 0), out_dst(E), out_eid(E);
    std::vector<int32_t> order(E);
    for (int e = 0; e < E; ++e) order[e] = e;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return col[a] != col[b] ? col[a] < col[b] : row[a] < row[b];
    });
    for (int k = 0; k < E; ++k) {
        int e = order[k];
        in_src[k] = row[e];
        in_eid[k] = e;
        in_ptr[col[e] + 1]++;
        if (k > 0) {
            int p = order[k - 1];
            B200_REQUIRE(!(row[p] == row[e] && col[p] == col[e]), "parallel links are not supported");
        }
    }
    for (int v = 0; v < N; ++v) in_ptr[v + 1] += in_ptr[v];
    int max_in = 0;
    for (int v = 0; v < N; ++v) max_in = std::max(max_in, in_ptr[v + 1] - in_ptr[v]);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return row[a] != row[b] ? row[a] < row[b] : col[a] < col[b];
    });
    for (int k = 0; k < E; ++k) {
        int e = order[k];
        out_dst[k] = col[e];
        out_eid[k] = e;
        out_ptr[row[e] + 1]++;
    }
    for (int v = 0; v < N; ++v) out_ptr[v + 1] += out_ptr[v];

    // demand CSR by origin slot
    std::vector<int32_t> origin_slot(N, -1), origins;
    for (int k = 0; k < num_od; ++k) {
        B200_REQUIRE(od_o[k] >= 0 && od_o[k] < N && od_d[k] >= 0 && od_d[k] < N, "OD endpoint out of range");
        origin_slot[od_o[k]] = 0;
    }
    for (int v = 0; v < N; ++v)
        if (origin_slot[v] == 0) {
            origin_slot[v] = (int)origins.size();
            origins.push_back(v);
        }
    const int O = (int)origins.size();
    std::vector<int32_t> od_ptr(O + 1, 0), od_dst(num_od);
    std::vector<double> q64(num_od);
    std::vector<float> q32(num_od);
    for (int k = 0; k < num_od; ++k) od_ptr[origin_slot[od_o[k]] + 1]++;
    for (int s = 0; s < O; ++s) od_ptr[s + 1] += od_ptr[s];
    {
        std::vector<int32_t> fill(O, 0);
        for (int k = 0; k < num_od; ++k) {
            int s = origin_slot[od_o[k]];
            int p = od_ptr[s] + fill[s]++;
            od_dst[p] = od_d[k];
            q64[p] = od_q[k];
            q32[p] = (float)od_q[k];
        }
    }
    double total = 0.0;
    for (int k = 0; k < num_od; ++k) total += od_q[k];

    b200_graph *g = new b200_graph();
    memset(g, 0, sizeof(*g));
    g->device = device;
    // test hook: route b200_assign through the general (CTA-per-scenario) kernel even when the
    // network qualifies for the small-network kernel
    g->force_general = getenv("B200_FORCE_GENERAL") != nullptr;
    g->force_small = getenv("B200_FORCE_SMALL") != nullptr;
    // Warm-starting the shortest paths from the previous iteration's trees pays when frontiers are
    // wide (measured: Chicago-Sketch-size lattice, 3.2 links/node, +15 %); on maze-like networks the
    // top-down pass over a deep tree costs more than it saves (Anaheim-size, 2.2 links/node, -24 %).
    // B200_NO_WARM / B200_WARM force either (A/B measurements).
    g->warm_start = (double)E / N >= 2.6;
    if (getenv("B200_NO_WARM")) g->warm_start = 0;
    if (getenv("B200_WARM")) g->warm_start = 1;
    g->cluster_limit = getenv("B200_CLUSTER") ? atoi(getenv("B200_CLUSTER")) : 0;
    g->N = N; g->E = E; g->O = O; g->num_od = num_od; g->max_in_deg = max_in; g->total_demand = total;
    B200_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, device));
    g->num_sms = prop.multiProcessorCount;

    // topology blob
    g->off_in_ptr = 0;
    g->off_in_src = pad4(N + 1);
    g->off_in_eid = g->off_in_src + pad4(E);
    g->off_out_ptr = g->off_in_eid + pad4(E);
    g->off_out_dst = g->off_out_ptr + pad4(N + 1);
    g->topo_ints = g->off_out_dst + pad4(E);
    std::vector<int32_t> blob(g->topo_ints, 0);
    std::copy(in_ptr.begin(), in_ptr.end(), blob.begin() + g->off_in_ptr);
    std::copy(in_src.begin(), in_src.end(), blob.begin() + g->off_in_src);
    std::copy(in_eid.begin(), in_eid.end(), blob.begin() + g->off_in_eid);
    std::copy(out_ptr.begin(), out_ptr.end(), blob.begin() + g->off_out_ptr);
    std::copy(out_dst.begin(), out_dst.end(), blob.begin() + g->off_out_dst);
    {
        bool intq = total < 2147483648.0;
        for (int k = 0; k < num_od && intq; ++k) intq = od_q[k] >= 0.0 && od_q[k] == floor(od_q[k]);
        g->gen_intq = intq ? 1 : 0;
        // packed Kahn words of the general kernel: 4 bits of pending children, 28 bits of subtree demand
        int max_out = 0;
        for (int v = 0; v < N; ++v) max_out = std::max(max_out, out_ptr[v + 1] - out_ptr[v]);
        g->max_out_deg = max_out;
        double max_origin = 0.0;
        for (int s2 = 0; s2 < O; ++s2) {
            double t = 0.0;
            for (int j = od_ptr[s2]; j < od_ptr[s2 + 1]; ++j) t += q64[j];
            max_origin = std::max(max_origin, t);
        }
        g->kahn_pack = (intq && max_out <= 15 && max_origin < 268435456.0 && !getenv("B200_NO_KAHN_PACK")) ? 1 : 0;
    }
    int rc;
    if ((rc = upload(&g->topo, blob))) return rc;
    if (N < 65535 && E < 65535) {
        auto pad8 = [](int x) { return (x + 7) & ~7; };
        g->o16_in_src = pad8(N + 1);
        g->o16_in_eid = g->o16_in_src + pad8(E);
        g->o16_out_ptr = g->o16_in_eid + pad8(E);
        g->o16_out_dst = g->o16_out_ptr + pad8(N + 1);
        const int total16 = g->o16_out_dst + pad8(E);
        std::vector<uint16_t> b16(total16, 0);
        for (int v = 0; v <= N; ++v) { b16[v] = (uint16_t)in_ptr[v]; b16[g->o16_out_ptr + v] = (uint16_t)out_ptr[v]; }
        for (int k = 0; k < E; ++k) {
            b16[g->o16_in_src + k] = (uint16_t)in_src[k];
            b16[g->o16_in_eid + k] = (uint16_t)in_eid[k];
            b16[g->o16_out_dst + k] = (uint16_t)out_dst[k];
        }
        g->topo16_bytes = total16 * 2;
        if ((rc = upload(&g->topo16, b16))) return rc;
    }
    std::vector<int32_t> vrow(row, row + E), vcol(col, col + E);
    std::vector<float> vt0(t0, t0 + E), vcap(cap0, cap0 + E), t0c(E), capc(E);
    for (int k = 0; k < E; ++k) { t0c[k] = t0[in_eid[k]]; capc[k] = cap0[in_eid[k]]; }
    if ((rc = upload(&g->row, vrow)) || (rc = upload(&g->col, vcol)) || (rc = upload(&g->t0, vt0)) ||
        (rc = upload(&g->cap0, vcap)) || (rc = upload(&g->t0_csc, t0c)) || (rc = upload(&g->cap0_csc, capc)) ||
        (rc = upload(&g->origins, origins)) || (rc = upload(&g->origin_slot, origin_slot)) ||
        (rc = upload(&g->od_ptr, od_ptr)) || (rc = upload(&g->od_dst, od_dst)) || (rc = upload(&g->od_q64, q64)) ||
        (rc = upload(&g->od_q32, q32)) || (rc = upload(&g->out_ptr, out_ptr)) || (rc = upload(&g->out_dst, out_dst)) ||
        (rc = upload(&g->out_eid, out_eid)))
        return rc;
    g->max_t0 = *std::max_element(vt0.begin(), vt0.end());
    g->min_t0 = *std::min_element(vt0.begin(), vt0.end());
    g->no_cch = getenv("B200_NO_CCH") != nullptr;
    g->max_cap0 = *std::max_element(vcap.begin(), vcap.end());
    g->origins_host = (int32_t *)malloc(sizeof(int32_t) * std::max(O, 1));
    std::copy(origins.begin(), origins.end(), g->origins_host);
    b200_small_prepare(g, in_ptr.data(), in_src.data(), in_eid.data(), t0, num_od, od_o, od_d, od_q, origin_slot.data());
    b200_static_prepare(g, row, col, t0, cap0, num_od, od_o, od_d, od_q, origin_slot.data());
    if ((rc = b200_cch_prepare(g, in_ptr.data(), in_src.data()))) { b200_graph_destroy(g); return rc; }
    if (g->cch_ok && getenv("B200_CCH_LIMIT")) g->cch_exact_limit = atof(getenv("B200_CCH_LIMIT"));   // test hook: force the fallback
    b200_pick_warps_cch(g);
    g->warps_f32 = b200_pick_warps(g, B200_F32, &g->smem_f32);
    g->warps_f64 = b200_pick_warps(g, B200_F64, &g->smem_f64);
    if (g->warps_f32 <= 0 || g->warps_f64 <= 0) {
        b200_set_error("network too large for the fused kernel's shared-memory layout (N=%d E=%d)", N, E);
        b200_graph_destroy(g);
        return B200_EUNSUPPORTED;
    }
    *out = g;
    return 0;
}

extern "C" int b200_graph_destroy(b200_graph *g)
{
    if (!g) return 0;
    cudaSetDevice(g->device);
    void *ptrs[] = {g->topo, g->row, g->col, g->t0, g->cap0, g->t0_csc, g->cap0_csc, g->origins, g->origin_slot,
                    g->od_ptr, g->od_dst, g->od_q64, g->od_q32, g->out_ptr, g->out_dst, g->out_eid,
                    g->small_blob32, g->small_blob64, g->small_blob_i32, g->small_eid, g->small_wpos,
                    g->tree_ws, g->topo16, g->st_origpos, g->st_blob32, g->st_blob64, g->st_blob_i32, g->st_blob_grp};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    b200_cch_destroy(g);
    free(g->origins_host);
    delete g;
    return 0;
}

extern "C" int b200_graph_info(const b200_graph *g, int64_t info[8])
{
    B200_REQUIRE(g && info, "null argument");
    info[0] = g->N; info[1] = g->E; info[2] = g->O; info[3] = g->num_od; info[4] = g->max_in_deg;
    info[5] = (int64_t)g->smem_f32; info[6] = (int64_t)g->smem_f64; info[7] = g->num_sms;
    return 0;
}

extern "C" int b200_graph_origins(const b200_graph *g, int32_t *origins_host)
{
    B200_REQUIRE(g && origins_host, "null argument");
    memcpy(origins_host, g->origins_host, sizeof(int32_t) * g->O);
    return 0;
}

extern "C" int b200_assign_launches(const b200_graph *g) { return g ? g->last_launches : 0; }

extern "C" int b200_assign_last_config(const b200_graph *g, int32_t cfg[8])
{
    B200_REQUIRE(g && cfg, "null argument");
    const bool general = !(g->static_ok && !g->force_general && !g->force_small) && !(g->small_ok && !g->force_general);
    cfg[0] = g->last_launches;
    cfg[1] = general ? std::max(g->last_cluster, 1) : 1;
    cfg[2] = general ? g->warm_start : 0;
    cfg[3] = general ? g->last_warps : 0;
    cfg[4] = g->last_cch;
    cfg[5] = cfg[6] = cfg[7] = 0;
    return 0;
}

extern "C" const char *b200_assign_kernel_name(const b200_graph *g)
{
    if (!g) return "";
    if (g->static_ok && !g->force_general && !g->force_small) return "assign_static_kernel";
    if (g->small_ok && !g->force_general) return "assign_small_kernel";
    return "assign_fused_kernel";
}
