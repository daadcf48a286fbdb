// Customizable contraction hierarchy (CCH) of the static network, planned once at graph upload
// for the general fused assignment kernel (b200_assign.cu).
//
// Why: the reference runs scipy's Dijkstra from every origin in every iteration of every scenario
// (src/env/repair_env.py:481-489).  The topology never changes - only the link costs do - so the
// search order can be fixed ahead of time: eliminate the nodes in a nested-dissection-like order,
// adding a shortcut (u,x) for every pair of remaining in-/out-neighbours of the eliminated node
// ("lower triangle" (u,v),(v,x)).  Per (scenario, iteration) the CTA first CUSTOMIZES the weights
// of the filled edges, W(u,x) = min(cost(u,x), min over lower triangles W(u,v) + W(v,x)); then the
// distances from one origin are
//   up phase    the nodes the origin reaches through edges to higher-ranked heads (on a two-way network
//               its ancestors in the elimination tree, a few dozen nodes), each relaxing those edges, and
//   down sweep  ONE pass over all nodes by decreasing rank, each pulling from its higher-ranked
//               in-neighbours - a static schedule: nodes of one level are independent, their in-edges
//               are ELL rows (32 destinations x degree) of packed words "tail slot | weight position",
//               a read-only table shared by every origin of every scenario (it lives in L1); the weights
//               themselves are stored compactly (one per filled edge) in shared memory.
// Every filled edge is relaxed once per origin (Anaheim-size: 1 272 for 914 links; Chicago-Sketch-size:
// 10 224 for 2 950) instead of the ~7 relaxations per link of frontier label correcting.
//
// Exactness: link costs are fp32 values >= min t0, so every path sum below 2^(q+52),
// q = floor(log2(min t0)) - 23, is exact in fp64 and therefore independent of the association order:
// the labels are bit for bit the ones Dijkstra (and the label-correcting kernel) produce, and the
// tie rule picks the same predecessors.  The kernel checks the bound per origin and falls back to
// Bellman-Ford on the rare origin that violates it (b200_assign.cu: warp_sssp_cch).
//
// Narrow levels of the down sweep split their high-degree destinations: chunks of in-edges go to
// "partial" distance slots first, the node then combines its partials through zero-weight entries;
// the chunk size is chosen per level by an instruction-count model of the sweep (down_layout).
#include "b200_assign_internal.cuh"

#include <map>
#include <numeric>
#include <set>
#include <unordered_map>

using namespace b200i;

namespace {

struct Plan {
    int N = 0, F = 0;
    std::vector<int> rank, order;
    std::vector<std::pair<int, int>> edges;                 // filled directed edges (tail, head)
    std::vector<int> orig_slot;                             // CSC slot of the original link or -1
    std::vector<std::vector<std::pair<int, int>>> tris;     // per edge: (edge (u,v), edge (v,x))
};

inline uint64_t key(int u, int x) { return ((uint64_t)(uint32_t)u << 32) | (uint32_t)x; }

// Elimination in rounds of independent low-degree nodes (degree <= max(min degree, 2) + slack): keeps the
// elimination tree shallow, which is what bounds the number of levels of the down sweep.  slack trades rounds
// (= levels) against fill, fill_ties orders equal degrees by the fill they cause (Chicago-Sketch-size: 21 % fewer
// lower triangles, 128 down batches instead of 136); build_tables tries a few combinations and keeps the cheapest sweep.
Plan make_plan(int N, int E, const int32_t *in_ptr, const int32_t *in_src, int slack, bool fill_ties)
{
    Plan P;
    P.N = N;
    std::vector<std::set<int>> out(N), inn(N), adj(N);
    std::unordered_map<uint64_t, int> id;
    auto add_edge = [&](int u, int x, int slot) {
        auto it = id.find(key(u, x));
        if (it != id.end()) return it->second;
        const int e = (int)P.edges.size();
        id[key(u, x)] = e;
        P.edges.push_back({u, x});
        P.orig_slot.push_back(slot);
        P.tris.emplace_back();
        out[u].insert(x); inn[x].insert(u); adj[u].insert(x); adj[x].insert(u);
        return e;
    };
    for (int v = 0; v < N; ++v)
        for (int k = in_ptr[v]; k < in_ptr[v + 1]; ++k) add_edge(in_src[k], v, k);
    (void)E;
    P.rank.assign(N, -1);
    std::vector<char> alive(N, 1);
    int remaining = N;
    std::vector<int> cand;
    while (remaining > 0) {
        cand.clear();
        int mind = 1 << 30;
        for (int v = 0; v < N; ++v)
            if (alive[v]) { cand.push_back(v); mind = std::min(mind, (int)adj[v].size()); }
        const int thr = std::max(mind, 2) + slack;
        if (fill_ties) {   // among equal degrees, the node whose elimination adds the fewest new (undirected) edges first
            std::vector<int> fill(N, 0);
            for (int v : cand) {
                if ((int)adj[v].size() > std::max(mind, 2) + slack) continue;
                int f = 0;
                for (int a : adj[v])
                    for (int b : adj[v])
                        if (a < b && !adj[a].count(b)) ++f;
                fill[v] = f;
            }
            std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) {
                return adj[a].size() != adj[b].size() ? adj[a].size() < adj[b].size() : fill[a] < fill[b];
            });
        } else {
            std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) { return adj[a].size() < adj[b].size(); });
        }
        std::vector<int> chosen;
        std::vector<char> blocked(N, 0);
        for (int v : cand) {
            if ((int)adj[v].size() > thr) break;
            if (blocked[v]) continue;
            chosen.push_back(v);
            for (int u : adj[v]) blocked[u] = 1;
        }
        for (int v : chosen) {
            alive[v] = 0;
            --remaining;
            P.rank[v] = (int)P.order.size();
            P.order.push_back(v);
            // v leaves the alive graph: its neighbours drop it, then get pairwise shortcuts
            for (int u : adj[v]) adj[u].erase(v);
            std::vector<int> ins, outs;
            for (int u : inn[v]) if (alive[u]) ins.push_back(u);
            for (int x : out[v]) if (alive[x]) outs.push_back(x);
            for (int u : ins)
                for (int x : outs)
                    if (u != x) {
                        const int a = id[key(u, v)], b = id[key(v, x)];
                        const int e = add_edge(u, x, -1);
                        P.tris[e].push_back({a, b});
                    }
        }
    }
    P.F = (int)P.edges.size();
    return P;
}

// Row streams: what a warp executes is a sequence of ROWS of 32 words (one per lane), read strictly in
// order and all of ONE fixed shape per stream, so the kernel's loops are branch-free, fully unrolled and
// prefetch the next batch through registers whatever level of the memory hierarchy serves the table.
//   down stream (shared by all origins): batches of 1 + D rows - a header row (destination slot per lane;
//   bit 16: a partial slot, which starts from +inf instead of its old content) and exactly D edge rows
//   "tail slot | weight position << 16" (padding: +inf slot | zero-weight position).  A warp barrier
//   precedes every batch.  Destinations with more than D in-edges are split into chunks of D edges over 2^span
//   adjacent lanes of the same batch (header bits 17-19 = span <= 5): the kernel reduces the group's candidates by
//   shuffles, its first lane writes - one batch whatever the in-degree up to 32 D (the dense top of the hierarchy has
//   nodes with dozens of in-edges; as partial slots combined stage by stage they cost two to three batches each).
//   Beyond 32 D, chunks go to partial slots first and the node combines them one stage later.  D is chosen per
//   network by the cost model of build_tables.
//   up stream (one per origin): pairs of rows - the pushing node (same word in every lane), then up to 32
//   of its edges to higher ranks "head | up-region position << 16" (padding: +inf slot | the +inf weight).
//   select stream (shared): for every block of 32 nodes SD rows "tail | CSC slot << 16" of their original
//   in-links in ascending tail order (padding: +inf slot | slot 0), SD = largest in-degree.
struct DownLayout {
    std::vector<uint32_t> rows;     // [batches][1 + D][32]
    std::vector<int> edge_pos;      // W position of every down edge (by edge id), -1 otherwise
    int slots = 0, wlen = 0;        // wlen: down-region positions = the down edges, compact, then ONE zero-weight slot
    int batches = 0, D = 0;
};

// in: (tail slot, edge id or -1 = zero weight).  An item with more than D in-edges takes 2^span ADJACENT lanes of a batch
// (chunks of D edges per lane); the kernel min-reduces the lanes of the group by shuffles and the first lane writes.
struct Item { int level, stage, dest, span; std::vector<std::pair<int, int>> in; };

// Layout of the down sweep for edge rows per batch D (see above).
DownLayout down_layout(const Plan &P, int D)
{
    const int N = P.N;
    DownLayout L;
    L.D = D;
    L.edge_pos.assign(P.F, -1);
    std::vector<std::vector<int>> dn(N);   // down in-edges (edge ids) per head
    int n_down = 0;
    for (int e = 0; e < P.F; ++e)
        if (P.rank[P.edges[e].first] > P.rank[P.edges[e].second]) { dn[P.edges[e].second].push_back(e); ++n_down; }
    std::vector<int> level(N, 0);
    for (int r = N - 1; r >= 0; --r) {
        const int v = P.order[r];
        int l = 0;
        for (int e : dn[v]) l = std::max(l, level[P.edges[e].first] + 1);
        level[v] = l;
    }
    std::vector<Item> items;
    std::map<int, int> next_slot;   // per level: partial slots are reused from level to level (a partial is dead
    int max_slot = N + 1;           // once its node has combined, and barriers separate the levels); slot N = +inf
    for (int v = 0; v < N; ++v) {
        if (dn[v].empty()) continue;
        std::vector<std::pair<int, int>> in;
        for (int e : dn[v]) in.push_back({P.edges[e].first, e});
        int stage = 0;
        int &ns = next_slot.emplace(level[v], N + 1).first->second;
        while ((int)in.size() > 32 * D) {   // wider than a whole batch: partial slots, combined one stage later
            std::vector<std::pair<int, int>> up;
            for (size_t i = 0; i < in.size(); i += D) {
                Item part{level[v], stage, ns, 0, {}};
                part.in.assign(in.begin() + i, in.begin() + std::min(in.size(), i + D));
                up.push_back({ns++, -1});
                items.push_back(std::move(part));
            }
            in = std::move(up);
            ++stage;
        }
        max_slot = std::max(max_slot, ns);
        int span = 0;
        while ((D << span) < (int)in.size()) ++span;
        items.push_back(Item{level[v], stage, v, span, std::move(in)});
    }
    L.slots = max_slot;
    const uint32_t ZERO = (uint32_t)n_down;   // zero-weight position: padding and partial-combine entries
    L.wlen = n_down + 1;
    int next_pos = 0;
    std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b) {
        if (a.level != b.level) return a.level < b.level;
        if (a.stage != b.stage) return a.stage < b.stage;
        return a.span > b.span;   // widest groups first: every group then starts at a multiple of its width
    });
    const uint32_t pad = (uint32_t)N | (ZERO << 16);   // tail = the +inf slot, weight 0
    for (size_t i = 0; i < items.size();) {
        // one batch: items of one (level, stage) while their lane groups fit
        const size_t hdr = L.rows.size();
        L.rows.resize(hdr + 32, (uint32_t)N);   // idle lane: min(+inf, +inf + 0) back into the +inf slot
        L.rows.resize(hdr + 32 + (size_t)D * 32, pad);
        const size_t base = hdr + 32;
        int lane = 0;
        size_t j = i;
        for (; j < items.size() && items[j].level == items[i].level && items[j].stage == items[i].stage; ++j) {
            const Item &it = items[j];
            const int width = 1 << it.span;
            if (lane + width > 32) break;
            for (int c = 0; c < width; ++c) {
                L.rows[hdr + lane + c] = (uint32_t)it.dest | (it.dest > N ? 0x10000u : 0u) | ((uint32_t)it.span << 17);
                for (int t = 0; t < D; ++t) {
                    const size_t k = (size_t)c * D + t;
                    if (k >= it.in.size()) break;
                    uint32_t wp = ZERO;
                    if (it.in[k].second >= 0) wp = (uint32_t)(L.edge_pos[it.in[k].second] = next_pos++);
                    L.rows[base + (size_t)t * 32 + lane + c] = (uint32_t)it.in[k].first | (wp << 16);
                }
            }
            lane += width;
        }
        ++L.batches;
        i = j;
    }
    return L;
}

template <typename T>
int up(T **dst, const std::vector<T> &src)
{
    const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    B200_CUDA(cudaMalloc((void **)dst, bytes));
    if (!src.empty()) B200_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// Everything the kernel reads, as host vectors (uploaded by b200_cch_prepare, replayed on the host by
// b200_cch_host_sssp).
struct Tables {
    // up_off[O + 1]: first PAIR of origin slot s; cz_lvl[L]: first row of customization level L (+ end);
    // cz_desc[r] = triangles per edge of row r (12 bits) | first word of the row in cz_tab / 32 << 12
    std::vector<uint32_t> dn_rows, up_rows, up_off, sel_rows, cz_init, cz_desc, cz_tab, cz_lvl;
    int F = 0, wd = 0, wu = 0, slots = 0, tris = 0, dn_batches = 0, D = 0, SD = 0;
};

// false: the network is outside the 16-bit limits of the tables
// model_origins: origin count the plan search assumes (the host replay of one origin passes the network's scale)
bool build_tables(int N, int E, const int32_t *in_ptr, const int32_t *in_src, int O, const int32_t *origins, Tables &T,
                  int model_origins = 0)
{
    if (N >= 65000 || E >= 65000) return false;
    // Plan search: elimination slack x edge rows per batch.  Cost model from the phase timings of the kernel
    // (tools/phase_timing.py): a down batch costs a warp ~300 clocks of dependent latency + ~55 per edge row, every
    // origin runs all batches (16 warps share the origins), a lower triangle costs the CTA ~9 clocks of customization.
    const char *force_d = getenv("B200_CCH_D"), *force_slack = getenv("B200_CCH_SLACK");   // tuning: force either
    Plan P;
    DownLayout D;
    double best_cost = -1.0;
    for (int variant = 0; variant < 4; ++variant) {
        const int slack = variant >> 1;
        const bool fill_ties = variant & 1;
        if (force_slack && atoi(force_slack) != slack) continue;
        Plan cand = make_plan(N, E, in_ptr, in_src, slack, fill_ties);
        size_t tr = 0;
        for (auto &t : cand.tris) tr += t.size();
        DownLayout best_d;
        double cost_d = -1.0;
        for (int d : {2, 3, 4, 5, 6, 8}) {
            if (force_d && atoi(force_d) != d) continue;
            DownLayout c = down_layout(cand, d);
            const double cost = (double)std::max(model_origins > 0 ? model_origins : O, 1) * c.batches * (300.0 + 55.0 * d) / 16.0 + 9.0 * (double)tr;
            if (getenv("B200_CCH_DEBUG"))
                fprintf(stderr, "[b200 cch] slack %d fill-ties %d D=%d: %d batches, %d slots; %d filled edges, %zu triangles; model %.0f\n",
                        slack, (int)fill_ties, d, c.batches, c.slots, cand.F, tr, cost);
            if (cost_d < 0.0 || cost < cost_d) { cost_d = cost; best_d = std::move(c); }
        }
        if (cost_d >= 0.0 && (best_cost < 0.0 || cost_d < best_cost * 0.98)) {   // a later variant must win by 2 %
            best_cost = cost_d;
            P = std::move(cand);
            D = std::move(best_d);
        }
    }
    if (best_cost < 0.0) return false;
    // up region: out-edges to higher-ranked heads, CSR by tail
    std::vector<int> pos(P.F, -1);
    std::vector<std::vector<int>> uo(N);
    for (int e = 0; e < P.F; ++e)
        if (P.rank[P.edges[e].first] < P.rank[P.edges[e].second]) uo[P.edges[e].first].push_back(e);
    int nu = 0;
    for (int v = 0; v < N; ++v) nu += (int)uo[v].size();
    const long wtot = (long)D.wlen + nu;
    if (wtot + 2 >= 65535 || D.slots >= 65535) return false;
    // up-region positions: out-edges to higher-ranked heads, grouped by tail
    std::vector<int> uo_first(N + 1, 0);
    {
        int next = 0;
        for (int v = 0; v < N; ++v) {
            uo_first[v] = next;
            for (int e : uo[v]) pos[e] = D.wlen + next++;
        }
        uo_first[N] = next;
    }
    // one up stream per origin: the nodes it reaches through edges to higher ranks, in rising rank, each
    // pushing along its own edges to higher ranks.  (On a two-way network these are the origin's ancestors
    // in the elimination tree; with one-way links the directed fill is sparser than the undirected one the
    // tree is defined on, so the set is taken from the directed edges themselves.)
    const uint32_t up_pad = (uint32_t)N | ((uint32_t)nu << 16);   // +inf slot | the +inf weight (first slot after the up region)
    T.up_off.assign(O + 1, 0);
    std::vector<char> seen(N, 0);
    for (int s = 0; s < O; ++s) {
        T.up_off[s] = (uint32_t)(T.up_rows.size() / 64);
        std::vector<int> reach{origins[s]}, stack{origins[s]};
        seen[origins[s]] = 1;
        while (!stack.empty()) {
            const int a = stack.back();
            stack.pop_back();
            for (int e : uo[a]) {
                const int x = P.edges[e].second;
                if (!seen[x]) { seen[x] = 1; reach.push_back(x); stack.push_back(x); }
            }
        }
        std::sort(reach.begin(), reach.end(), [&](int a, int b) { return P.rank[a] < P.rank[b]; });
        for (int a : reach) {
            seen[a] = 0;
            const int d = (int)uo[a].size();
            for (int r = 0; r < d; r += 32) {
                T.up_rows.insert(T.up_rows.end(), 32, (uint32_t)a);
                for (int l = 0; l < 32; ++l)
                    T.up_rows.push_back(r + l < d ? (uint32_t)P.edges[uo[a][r + l]].second | ((uint32_t)(uo_first[a] + r + l) << 16)
                                                  : up_pad);
            }
        }
    }
    T.up_off[O] = (uint32_t)(T.up_rows.size() / 64);
    // select stream: original in-links per block of 32 nodes
    T.SD = 1;
    for (int v = 0; v < N; ++v) T.SD = std::max(T.SD, in_ptr[v + 1] - in_ptr[v]);
    T.sel_rows.assign((size_t)((N + 31) / 32) * T.SD * 32, (uint32_t)N);
    for (int v = 0; v < N; ++v)
        for (int k = in_ptr[v]; k < in_ptr[v + 1]; ++k)
            T.sel_rows[((size_t)(v / 32) * T.SD + (k - in_ptr[v])) * 32 + (v % 32)] = (uint32_t)in_src[k] | ((uint32_t)k << 16);
    for (int e = 0; e < P.F; ++e)
        if (D.edge_pos[e] >= 0) pos[e] = D.edge_pos[e];
    const uint32_t INFPOS = (uint32_t)wtot, DUMMY = (uint32_t)wtot + 1;
    // customization: initial value of every filled edge, then rows of 32 edges per level
    T.cz_init.resize(P.F);
    for (int e = 0; e < P.F; ++e)
        T.cz_init[e] = (uint32_t)pos[e] | ((uint32_t)(P.orig_slot[e] >= 0 ? P.orig_slot[e] : 0xFFFF) << 16);
    std::vector<int> lvl(P.F, 0), by_low(P.F);
    std::iota(by_low.begin(), by_low.end(), 0);
    auto low = [&](int e) { return std::min(P.rank[P.edges[e].first], P.rank[P.edges[e].second]); };
    std::stable_sort(by_low.begin(), by_low.end(), [&](int a, int b) { return low(a) < low(b); });
    int max_lvl = 0;
    for (int e : by_low) {
        int l = 0;
        for (auto &t : P.tris[e]) l = std::max(l, 1 + std::max(lvl[t.first], lvl[t.second]));
        lvl[e] = l;
        max_lvl = std::max(max_lvl, l);
    }
    for (int L = 1; L <= max_lvl; ++L) {   // rows of one level are independent; a CTA barrier separates the levels
        T.cz_lvl.push_back((uint32_t)T.cz_desc.size());
        std::vector<int> es;
        for (int e = 0; e < P.F; ++e)
            if (lvl[e] == L) es.push_back(e);
        std::stable_sort(es.begin(), es.end(), [&](int a, int b) { return P.tris[a].size() > P.tris[b].size(); });
        for (size_t i = 0; i < es.size(); i += 32) {
            const size_t j = std::min(es.size(), i + 32);
            const int nt = (int)P.tris[es[i]].size();
            if (nt > 4095 || T.cz_tab.size() / 32 >= (1u << 20)) return false;   // descriptor fields
            T.cz_desc.push_back((uint32_t)nt | (uint32_t)(T.cz_tab.size() / 32) << 12);
            for (int l = 0; l < 32; ++l) T.cz_tab.push_back(i + l < j ? (uint32_t)pos[es[i + l]] : DUMMY);
            for (int t = 0; t < nt; ++t)
                for (int l = 0; l < 32; ++l) {
                    uint32_t w = INFPOS | (INFPOS << 16);
                    if (i + l < j && t < (int)P.tris[es[i + l]].size()) {
                        const auto &tr = P.tris[es[i + l]][t];
                        w = (uint32_t)pos[tr.first] | ((uint32_t)pos[tr.second] << 16);
                        ++T.tris;
                    }
                    T.cz_tab.push_back(w);
                }
        }
    }
    T.cz_lvl.push_back((uint32_t)T.cz_desc.size());
    T.dn_rows = std::move(D.rows);
    T.F = P.F; T.wd = D.wlen; T.wu = nu; T.slots = D.slots; T.dn_batches = D.batches; T.D = D.D;
    return true;
}

// CSC (in-links by head, ascending tail) of a link list, as b200_graph_create builds it
void make_csc(int N, int E, const int32_t *row, const int32_t *col, std::vector<int32_t> &in_ptr,
              std::vector<int32_t> &in_src, std::vector<int32_t> &in_eid)
{
    std::vector<int32_t> order(E);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return col[a] != col[b] ? col[a] < col[b] : row[a] < row[b]; });
    in_ptr.assign(N + 1, 0); in_src.resize(E); in_eid.resize(E);
    for (int k = 0; k < E; ++k) { in_src[k] = row[order[k]]; in_eid[k] = order[k]; in_ptr[col[order[k]] + 1]++; }
    for (int v = 0; v < N; ++v) in_ptr[v + 1] += in_ptr[v];
}

}  // namespace

// Host replay of exactly the schedule the kernel executes (customization rows, up phase, down-sweep
// batches) for one cost vector and one origin: the planner's self-check, runs without a GPU.
// cost: fp64 [E] by link id; dist_out: fp64 [N].  stats (optional, int64[8]): filled edges, lower
// triangles, down-region entries, up-region entries, distance slots, down batches, customization rows, 0.
extern "C" int b200_cch_host_sssp(int N, int E, const int32_t *row, const int32_t *col, const double *cost, int origin,
                                  double *dist_out, int64_t *stats)
{
    B200_REQUIRE(N > 0 && E > 0 && row && col && cost && dist_out && origin >= 0 && origin < N, "bad argument");
    std::vector<int32_t> in_ptr, in_src, in_eid;
    make_csc(N, E, row, col, in_ptr, in_src, in_eid);
    Tables T;
    const int32_t one_origin[1] = {origin};
    if (!build_tables(N, E, in_ptr.data(), in_src.data(), 1, one_origin, T, std::max(1, N / 4))) { b200_set_error("network too large for the hierarchy tables"); return B200_EUNSUPPORTED; }
    const double INF = INFINITY;
    std::vector<double> W((size_t)T.wd + T.wu + 2, 0.0);   // position wd - 1 keeps 0.0: the zero-weight slot
    W[T.wd + T.wu] = INF;
    for (int e = 0; e < T.F; ++e) {
        const uint32_t wi = T.cz_init[e];
        W[wi & 0xffffu] = (wi >> 16) == 0xffffu ? INF : cost[in_eid[wi >> 16]];
    }
    for (size_t r = 0; r < T.cz_desc.size(); ++r) {
        const int nt = (int)(T.cz_desc[r] & 0xfffu);
        const size_t off = (size_t)(T.cz_desc[r] >> 12) * 32;
        std::vector<double> acc(32);
        for (int l = 0; l < 32; ++l) {
            acc[l] = W[T.cz_tab[off + l]];
            for (int t = 0; t < nt; ++t) {
                const uint32_t tri = T.cz_tab[off + 32 * (t + 1) + l];
                const double c = W[tri & 0xffffu] + W[tri >> 16];
                if (c < acc[l]) acc[l] = c;
            }
        }
        for (int l = 0; l < 32; ++l) W[T.cz_tab[off + l]] = acc[l];
    }
    std::vector<double> d(T.slots, INF);
    d[origin] = 0.0;
    for (uint32_t q = T.up_off[0]; q < T.up_off[1]; ++q) {   // up stream of this origin: (node, edges) row pairs
        const uint32_t *row = &T.up_rows[(size_t)q * 64];
        const double da = d[row[0] & 0xffffu];
        for (int l = 0; l < 32; ++l) {
            const double c = da + W[T.wd + (row[32 + l] >> 16)];
            if (c < d[row[32 + l] & 0xffffu]) d[row[32 + l] & 0xffffu] = c;
        }
    }
    for (int b = 0; b < T.dn_batches; ++b) {   // down stream: header row + D edge rows per batch
        const uint32_t *row = &T.dn_rows[(size_t)b * (1 + T.D) * 32];
        double best[32];
        for (int l = 0; l < 32; ++l) {
            best[l] = (row[l] & 0x10000u) ? INF : d[row[l] & 0xffffu];
            for (int j = 1; j <= T.D; ++j) {
                const uint32_t w = row[j * 32 + l];
                const double c = d[w & 0xffffu] + W[w >> 16];
                if (c < best[l]) best[l] = c;
            }
        }
        for (int l = 0; l < 32; ++l) {   // lane groups: the first lane writes the minimum of the group
            const int width = 1 << ((row[l] >> 17) & 7u);
            if (l % width) continue;
            double m = best[l];
            for (int c = 1; c < width; ++c) m = std::min(m, best[l + c]);
            d[row[l] & 0xffffu] = m;
        }
    }
    for (int v = 0; v < N; ++v) dist_out[v] = d[v];
    if (stats) {
        stats[0] = T.F; stats[1] = T.tris; stats[2] = T.wd; stats[3] = T.wu; stats[4] = T.slots;
        stats[5] = (int64_t)T.dn_batches; stats[6] = (int64_t)T.cz_desc.size(); stats[7] = (int64_t)T.D;
    }
    return 0;
}

// Plans the hierarchy and uploads its tables; sets g->cch_ok.  Returns 0 also when the network is
// outside the 16-bit limits of the tables (the label-correcting kernel then serves it alone).
int b200_cch_prepare(b200_graph *g, const int32_t *in_ptr, const int32_t *in_src)
{
    g->cch_ok = 0;
    Tables T;
    if (!build_tables(g->N, g->E, in_ptr, in_src, g->O, g->origins_host, T)) return 0;
    int rc;
    if ((rc = up(&g->cch_dn_rows, T.dn_rows)) || (rc = up(&g->cch_up_rows, T.up_rows)) ||
        (rc = up(&g->cch_sel_rows, T.sel_rows)) || (rc = up(&g->cch_up_off, T.up_off)) ||
        (rc = up(&g->cch_cz_init, T.cz_init)) || (rc = up(&g->cch_cz_desc, T.cz_desc)) || (rc = up(&g->cch_cz_tab, T.cz_tab)) ||
        (rc = up(&g->cch_cz_lvl, T.cz_lvl)))
        return rc;
    g->cch_D = T.D;
    g->cch_SD = T.SD;
    g->cch_F = T.F;
    g->cch_wd = T.wd;
    g->cch_wu = T.wu;
    g->cch_slots = T.slots;
    g->cch_dn_batches = T.dn_batches;
    g->cch_cz_rows = (int)T.cz_desc.size();
    g->cch_cz_levels = (int)T.cz_lvl.size() - 1;
    g->cch_tris = T.tris;
    int e2 = 0;
    frexp((double)g->min_t0, &e2);
    g->cch_exact_limit = ldexp(1.0, (e2 - 1) - 23 + 52);   // labels below this are exact sums (file header)
    g->cch_ok = 1;
    return 0;
}

void b200_cch_destroy(b200_graph *g)
{
    void *ptrs[] = {g->cch_dn_rows, g->cch_up_rows, g->cch_sel_rows, g->cch_up_off,
                    g->cch_cz_init, g->cch_cz_desc, g->cch_cz_tab, g->cch_cz_lvl, g->cch_ws};
    for (void *p : ptrs)
        if (p) cudaFree(p);
}
