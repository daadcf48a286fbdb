// Generic lane-per-origin kernel for small networks: translation unit (kernel in
// b200_assign_small.cuh, host-side eligibility test + table construction here).
#include "b200_assign_internal.cuh"

using namespace b200i;

namespace {
#include "b200_assign_small.cuh"
}  // namespace

int b200_launch_small(const b200_graph *g, const Params &p, int prec, cudaStream_t st)
{
    return prec == B200_F32 ? launch_small<float>(g, p, st) : launch_small<double>(g, p, st);
}

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

