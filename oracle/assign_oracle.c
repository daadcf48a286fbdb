/* ORACLE - test infrastructure, NOT product code.
 *
 * Plain-C CPU restatement of the reference's traffic-assignment hot path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may load the library built from this
 * file; the product (CUDA) path never links or calls it.
 *
 * What it restates (reference = pop-pop-pOp-dev/SAC-GAT-HER_transportationRL, read-only):
 *   orc_bpr_*       compute_travel_time, numpy branch        src/env/repair_env.py:667-677
 *   orc_sssp        scipy.sparse.csgraph.dijkstra as called  src/env/repair_env.py:485-489
 *                   (third-party, NOT in the reference tree; requirements.txt leaves scipy
 *                   unpinned, the build container has scipy 1.18.1).  Distances: fp64 sums of
 *                   the fp32 link costs accumulated origin -> node.  Predecessors: scipy writes
 *                   pred[v] only on a strict improvement, i.e. the first-settled tight
 *                   predecessor wins; restated as "among u with dist[u]+w(u,v) == dist[v] take
 *                   the smallest dist[u], then the largest u" (SURVEY.md §7.1 / Appendix A.1,
 *                   re-checked against the installed scipy by tests/test_tie_rule.py).
 *   orc_aon_*       _all_or_nothing loading + pred walk      src/env/repair_env.py:490-503, 707-722
 *   orc_assign_*    compute_flow_assignment (msa/fw/cfw)     src/env/repair_env.py:299-345
 *                   + compute_tstt                           src/env/repair_env.py:724-735
 *                   + extensions the north star adds: bisection line search, relative gap.
 *   orc_betweenness get_state's Brandes betweenness          src/env/repair_env.py:753-766
 *                   (networkx.betweenness_centrality, third-party, unpinned; 3.6.1 here)
 *
 * Parity pin: checked against the tests/golden npz files, which were produced by running the reference
 * itself (tests/golden/make_golden.py).  _f32 mirrors the reference's dtypes (fp32 state, fp64
 * distances) and differs from it only where numpy's SIMD powf is not correctly rounded
 * (documented tolerance 1e-6); _f64 is the all-double variant.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NO_PRED (-9999)

typedef struct {
    int N, E, n_od, n_orig;
    const int *row, *col;
    const int *od_o, *od_d;
    const double *od_q;
    double total_demand;
    int *origins;      /* ascending node ids that have demand */
    int *origin_slot;  /* [N] -> index into origins or -1 */
    int *out_ptr, *out_edge; /* CSR by tail node */
    int *in_ptr, *in_edge;   /* CSC by head node */
} Net;

static void net_init(Net *n, int N, int E, const int *row, const int *col, int n_od, const int *od_o,
                     const int *od_d, const double *od_q)
{
    n->N = N; n->E = E; n->n_od = n_od; n->row = row; n->col = col;
    n->od_o = od_o; n->od_d = od_d; n->od_q = od_q;
    n->origin_slot = (int *)malloc(sizeof(int) * N);
    n->origins = (int *)malloc(sizeof(int) * (N > 0 ? N : 1));
    for (int i = 0; i < N; ++i) n->origin_slot[i] = -1;
    /* total_demand = np.sum(list(od_demand.values())) (reference: :96); demands are integer
     * valued in every shipped / generated network, so the fp64 sum is exact in any order */
    n->total_demand = 0.0;
    for (int k = 0; k < n_od; ++k) { n->origin_slot[od_o[k]] = 0; n->total_demand += od_q[k]; }
    n->n_orig = 0;
    for (int i = 0; i < N; ++i) if (n->origin_slot[i] == 0) { n->origin_slot[i] = n->n_orig; n->origins[n->n_orig++] = i; }
    n->out_ptr = (int *)calloc(N + 1, sizeof(int)); n->in_ptr = (int *)calloc(N + 1, sizeof(int));
    n->out_edge = (int *)malloc(sizeof(int) * (E > 0 ? E : 1)); n->in_edge = (int *)malloc(sizeof(int) * (E > 0 ? E : 1));
    for (int e = 0; e < E; ++e) { n->out_ptr[row[e] + 1]++; n->in_ptr[col[e] + 1]++; }
    for (int i = 0; i < N; ++i) { n->out_ptr[i + 1] += n->out_ptr[i]; n->in_ptr[i + 1] += n->in_ptr[i]; }
    int *fo = (int *)calloc(N, sizeof(int)), *fi = (int *)calloc(N, sizeof(int));
    for (int e = 0; e < E; ++e) {
        n->out_edge[n->out_ptr[row[e]] + fo[row[e]]++] = e;
        n->in_edge[n->in_ptr[col[e]] + fi[col[e]]++] = e;
    }
    free(fo); free(fi);
}

static void net_free(Net *n)
{
    free(n->origins); free(n->origin_slot); free(n->out_ptr); free(n->out_edge); free(n->in_ptr); free(n->in_edge);
}

/* ---- binary-heap Dijkstra + predecessor selection ------------------------------------------ */
typedef struct { double key; int node; } HeapItem;

static void heap_push(HeapItem *h, int *n, double key, int node)
{
    int i = (*n)++;
    while (i > 0) {
        int p = (i - 1) / 2;
        if (h[p].key <= key) break;
        h[i] = h[p]; i = p;
    }
    h[i].key = key; h[i].node = node;
}

static HeapItem heap_pop(HeapItem *h, int *n)
{
    HeapItem top = h[0], last = h[--(*n)];
    int i = 0;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= *n) break;
        if (c + 1 < *n && h[c + 1].key < h[c].key) c++;
        if (last.key <= h[c].key) break;
        h[i] = h[c]; i = c;
    }
    h[i] = last;
    return top;
}

static void sssp_tree(const Net *net, const double *w, int origin, double *dist, int *pred, int *pedge)
{
    const int N = net->N;
    HeapItem *heap = (HeapItem *)malloc(sizeof(HeapItem) * (size_t)(net->E + N + 1));
    char *done = (char *)calloc(N, 1);
    int hn = 0;
    for (int v = 0; v < N; ++v) { dist[v] = INFINITY; pred[v] = NO_PRED; pedge[v] = -1; }
    dist[origin] = 0.0;
    heap_push(heap, &hn, 0.0, origin);
    while (hn > 0) {
        HeapItem it = heap_pop(heap, &hn);
        int u = it.node;
        if (done[u]) continue;
        done[u] = 1;
        for (int k = net->out_ptr[u]; k < net->out_ptr[u + 1]; ++k) {
            int e = net->out_edge[k], v = net->col[e];
            double cand = dist[u] + w[e];
            if (cand < dist[v]) { dist[v] = cand; heap_push(heap, &hn, cand, v); }
        }
    }
    /* tie rule: smallest dist[u] among tight in-links, then largest u */
    for (int v = 0; v < N; ++v) {
        if (v == origin || !(dist[v] < INFINITY)) continue;
        int best_u = -1, best_e = -1; double best_du = INFINITY;
        for (int k = net->in_ptr[v]; k < net->in_ptr[v + 1]; ++k) {
            int e = net->in_edge[k], u = net->row[e];
            if (!(dist[u] < INFINITY)) continue;
            if (dist[u] + w[e] != dist[v]) continue;
            if (best_u < 0 || dist[u] < best_du || (dist[u] == best_du && u > best_u)) { best_u = u; best_e = e; best_du = dist[u]; }
        }
        pred[v] = best_u < 0 ? NO_PRED : best_u;
        pedge[v] = best_e;
    }
    free(heap); free(done);
}

/* S2 for a list of origins; w = link costs as double (the fp32 costs cast exactly). */
void orc_sssp(int N, int E, const int *row, const int *col, const double *w, int n_orig, const int *origins,
              double *dist, int *pred, int *pedge)
{
    Net net;
    net_init(&net, N, E, row, col, 0, NULL, NULL, NULL);
    for (int s = 0; s < n_orig; ++s)
        sssp_tree(&net, w, origins[s], dist + (long)s * N, pred + (long)s * N, pedge + (long)s * N);
    net_free(&net);
}

/* Brandes betweenness, unweighted, directed, on the edge-induced subgraph of active links
 * (networkx.betweenness_centrality(subgraph, normalized=True), reference: :753-761).  bw[N]. */
void orc_betweenness(int N, int E, const int *row, const int *col, const unsigned char *dmg, double *bw)
{
    int *optr = (int *)calloc(N + 1, sizeof(int)), *oadj = (int *)malloc(sizeof(int) * (E > 0 ? E : 1));
    char *in_sub = (char *)calloc(N, 1);
    int n_sub = 0;
    for (int e = 0; e < E; ++e) if (!dmg[e]) { optr[row[e] + 1]++; in_sub[row[e]] = 1; in_sub[col[e]] = 1; }
    for (int i = 0; i < N; ++i) { optr[i + 1] += optr[i]; n_sub += in_sub[i]; }
    int *fill = (int *)calloc(N, sizeof(int));
    for (int e = 0; e < E; ++e) if (!dmg[e]) oadj[optr[row[e]] + fill[row[e]]++] = col[e];
    free(fill);
    double *sigma = (double *)malloc(sizeof(double) * N), *delta = (double *)malloc(sizeof(double) * N);
    int *d = (int *)malloc(sizeof(int) * N), *order = (int *)malloc(sizeof(int) * N);
    for (int i = 0; i < N; ++i) bw[i] = 0.0;
    for (int s = 0; s < N; ++s) {
        if (!in_sub[s]) continue;
        for (int i = 0; i < N; ++i) { sigma[i] = 0.0; delta[i] = 0.0; d[i] = -1; }
        sigma[s] = 1.0; d[s] = 0;
        int head = 0, tail = 0;
        order[tail++] = s;
        while (head < tail) {
            int v = order[head++];
            for (int k = optr[v]; k < optr[v + 1]; ++k) {
                int x = oadj[k];
                if (d[x] < 0) { d[x] = d[v] + 1; order[tail++] = x; }
                if (d[x] == d[v] + 1) sigma[x] += sigma[v];
            }
        }
        /* accumulate by scanning each node's out-links in reverse BFS order of the tail node */
        for (int i = tail - 1; i >= 0; --i) {
            int v = order[i];
            double acc = 0.0;
            for (int k = optr[v]; k < optr[v + 1]; ++k) {
                int x = oadj[k];
                if (d[x] == d[v] + 1) acc += sigma[v] * ((1.0 + delta[x]) / sigma[x]);
            }
            delta[v] = acc;
            if (v != s) bw[v] += delta[v];
        }
    }
    if (n_sub > 2) { double sc = 1.0 / ((double)(n_sub - 1) * (double)(n_sub - 2)); for (int i = 0; i < N; ++i) bw[i] *= sc; }
    free(optr); free(oadj); free(in_sub); free(sigma); free(delta); free(d); free(order);
}

#define REAL float
#define SUF(x) x##_f32
#define POWFN powf
#include "assign_body.inc"
#undef REAL
#undef SUF
#undef POWFN

#define REAL double
#define SUF(x) x##_f64
#define POWFN pow
#include "assign_body.inc"
#undef REAL
#undef SUF
#undef POWFN
