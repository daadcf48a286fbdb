// Grouped variant of the network-specialised kernel (included after b200_assign_static.cuh, same
// anonymous namespace): NO PHANTOM LANES.
//
// assign_static_kernel gives a scenario one warp with lane = origin; a network with fewer than 32
// nodes leaves lanes idle through every instruction (Sioux Falls: 24 origins, 8 of 32 lanes, 25 % of the
// sweeps, the predecessor selection and the subtree peel).  Here a CTA of GW warps hosts GS scenarios with
// OL = N lanes each, thread t = scenario t / OL, origin t % OL (Sioux Falls: 3 warps = 96 lanes = 4
// scenarios x 24 origins, no idle lane).  Scenarios straddle warp boundaries, so what one warp used to
// keep to itself is shared by the CTA:
//   * per-thread columns (packed subtree word, Kahn stack, tree parent / in-link slot) are [N][threads];
//   * per-scenario arrays (link costs, hierarchy weights, link loads, flows) are indexed by the thread's
//     scenario: a weight load of a warp touches at most two scenarios' copies;
//   * customization runs on all scenarios at once, lane o of a scenario taking word o of a table row
//     (codegen: CZG table, N edges per batch); batches are separated by CTA barriers;
//   * link loads are summed by shared-memory atomics (int32: exact, order-independent), the unassigned
//     demand likewise; S4 deals the GS x E (scenario, link) pairs over the threads; TSTT (numpy pairwise
//     order) and the gap are taken by one warp per scenario.
// Same arithmetic in the same order per origin and per link as assign_static_kernel, hence the same bits.
// Instantiated for what the bench and the env run: fp32 link state, integer demands, fixed-step msa / fw
// with beta = 4 (FAST), exactly summable costs (CCH); everything else keeps assign_static_kernel.
#pragma once

template <class T>
struct GrpShape {
    static constexpr int OL = T::N;   // lanes per scenario
    static constexpr int util(int gw) { return (gw * 32 / OL) * OL * 1000 / (gw * 32); }
    static constexpr int GW = util(3) > util(2) && util(3) > util(1) && util(3) >= util(4) ? 3
                              : util(4) > util(2) && util(4) > util(1) ? 4
                              : util(2) > util(1) ? 2 : 1;
    static constexpr int GS = GW * 32 / OL, TH = GW * 32;
    static constexpr bool worthwhile = T::N <= 32 && util(GW) > util(1) + 100;   // > 10 points fewer idle lanes than one warp per scenario
};

template <class T>
struct GrpLayout {
    using S = GrpShape<T>;
    static constexpr int N = T::N, E = T::E, TH = S::TH, GS = S::GS;
    static constexpr int EP = (E + 31) & ~31;
    // blob: Q int32 [N][32] | t0[EP] f32 | cap0[EP] f32 | link id[EP] u8 | padded cost index[EP] u8 |
    //       byte offset of the link's hierarchy weight slot [EP] u16 || CZG table u32 [CZG_LEN]
    // The part before || is staged in shared memory (BLOB bytes, one TMA copy); the customization table stays
    // in global memory and is read through L1 (4.6 KB shared by every CTA of the SM): without it a CTA is
    // 30.7 KB and SEVEN fit an SM, which makes 4096 scenarios a single wave of 148 x 7 x 4 = 4144 slots.
    static constexpr int TAB_T0 = N * 32 * 4, TAB_CAP = TAB_T0 + EP * 4, TAB_EID = TAB_CAP + EP * 4,
                         TAB_WPOS = TAB_EID + EP, TAB_CPOS = TAB_WPOS + EP, TAB_CZ = calign16(TAB_CPOS + EP * 2);
    static constexpr int BLOB = TAB_CZ, BLOB_ALL = calign16(TAB_CZ + T::CZG_LEN * 4);
    // columns [N][TH]
    // packed subtree word u32 | tree parent + in-link slot (parent | slot << 8) u16 | Kahn stack u8
    static constexpr int WORD = 0, PUK = WORD + N * TH * 4, STK = PUK + N * TH * 2, COLS = calign16(STK + N * TH);
    // per scenario
    static constexpr int WPAD = woff<T>(N);
    static constexpr int S_W = 0, S_CW = S_W + calign16(WPAD * 8), S_AUX = S_CW + calign16((T::CW + 2) * 8),
                         S_FLOW = S_AUX + calign16(E * 4), SCEN = S_FLOW + calign16(E * 4);
    static constexpr int RED = COLS + SCEN * GS;   // per scenario: unassigned (int), nan flag
    static constexpr int TOTAL = 16 + BLOB + RED + calign16(GS * 8);
};

// first word of customization batch b in T::CZG_TAB
template <class T>
__host__ __device__ constexpr int czg_off(int b)
{
    int o = 0;
    for (int i = 0; i < b; ++i) o += 32 * (1 + T::CZG_T[i]);
    return o;
}

__device__ __forceinline__ void grp_bar(int threads)
{
    asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory");
}

// Registers: 7 CTAs of 3 warps put 6 warps on one of the SM's four sub-partitions (16 K registers each), so a thread
// may use at most 16384 / (6 x 32) = 85 -> 80 registers; ptxas derives that from the bounds below (96 or 88 registers
// drop the occupancy to 6 CTAs per SM and the 4096-scenario launch to two waves: measured 0.44-0.46 ms against 0.33).
// Tried and dropped: keeping the tree (parent, in-link, children counts) in registers until the Kahn scan and packing
// parent | slot | sum into one word - 15 % fewer shared-memory wavefronts, but more instructions and spills: 0.36 ms.
template <class T>
__global__ void __launch_bounds__(GrpShape<T>::TH, 65536 / (GrpShape<T>::TH * 96) > 0 ? 65536 / (GrpShape<T>::TH * 96) : 1)
    assign_static_grp_kernel(const Params p)
{
    using S = GrpShape<T>;
    using L = GrpLayout<T>;
    constexpr int N = T::N, E = T::E, TH = S::TH, GS = S::GS, OL = S::OL;
    constexpr int LPT = (GS * E + TH - 1) / TH;   // (scenario, link) pairs per thread in S4
    static_assert(N <= 32 && E <= 254, "one origin lane per node (orphan mask is 32 bits), link slots are bytes");
    extern __shared__ __align__(16) unsigned char smem[];
    const int t = threadIdx.x, lane = t & 31;
    uint64_t *bar = (uint64_t *)smem;
    unsigned char *blob = smem + 16;
    if (t == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (t == 0) {
        mbar_expect_tx(bar, (uint32_t)L::BLOB);
        tma_bulk_g2s(blob, p.st_blob, (uint32_t)L::BLOB, bar);
    }
    mbar_wait(bar, 0);

    const bool lane_live = t < GS * OL;           // threads beyond the last whole scenario only attend the barriers
    const int s = lane_live ? t / OL : 0, o = lane_live ? t - s * OL : 0;
    unsigned char *cols = smem + 16 + L::BLOB;
    unsigned char *scen0 = cols + L::COLS;
    int *red = (int *)(scen0 + L::SCEN * GS);      // [GS][2]: unassigned demand, NaN flag
    unsigned *word = (unsigned *)(cols + L::WORD) + t;
    uint8_t *stk = cols + L::STK + t;
    uint16_t *puk = (uint16_t *)(cols + L::PUK) + t;
    unsigned char *sb = scen0 + L::SCEN * s;
    const uint32_t wbase = smem_u32(sb + L::S_W), wcb = smem_u32(sb + L::S_CW);
    int *auxi = (int *)(sb + L::S_AUX);
    const unsigned *cz = (const unsigned *)(p.st_blob + L::TAB_CZ);   // global, L1-resident
    const int *Qc = (const int *)blob + o;
    const int orig = __ldg(p.st_origpos + o);
    const unsigned FULL = 0xffffffffu;
    const float *t0k = (const float *)(blob + L::TAB_T0), *capk = (const float *)(blob + L::TAB_CAP);
    const uint8_t *eidk = blob + L::TAB_EID, *wposk = blob + L::TAB_WPOS;
    const uint16_t *cposk = (const uint16_t *)(blob + L::TAB_CPOS);

    // scenario of this thread's origin lane, and the (scenario, link) pairs this thread updates in S4
    const int b0 = blockIdx.x * GS;
    if (t < GS) { red[2 * t] = 0; red[2 * t + 1] = 0; }
    if (t < GS) sts_f64(smem_u32(scen0 + L::SCEN * t + L::S_CW) + T::CW * 8, B200_INF_F64);   // +inf slot of every scenario
    int pair_s[LPT], pair_k[LPT];
    unsigned dmg_bits = 0, pair_ok = 0;
#pragma unroll
    for (int j = 0; j < LPT; ++j) {
        const int idx = t + TH * j;
        pair_s[j] = idx / E;
        pair_k[j] = idx - pair_s[j] * E;
        if (idx < GS * E) {
            pair_ok |= 1u << j;
            const int bb = min(b0 + pair_s[j], p.B - 1);   // a group past the end of the batch recomputes the last scenario, unwritten
            if (p.dmg[(size_t)bb * E + eidk[pair_k[j]]]) dmg_bits |= 1u << j;
        }
    }
    auto cost = [&](int j, float f) -> double {
        if ((dmg_bits >> j) & 1u) return 1e6;
        const int k = pair_k[j];
        return (double)Bpr<float>::eval(f, capk[k], t0k[k], (float)p.alpha, (float)p.beta, true);
    };
#pragma unroll
    for (int j = 0; j < LPT; ++j)
        if ((pair_ok >> j) & 1u) {
            unsigned char *ps = scen0 + L::SCEN * pair_s[j];
            const int k = pair_k[j], bb = min(b0 + pair_s[j], p.B - 1);
            const float f = ((const float *)p.flow)[(size_t)bb * E + eidk[k]];
            ((float *)(ps + L::S_FLOW))[k] = f;
            ((int *)(ps + L::S_AUX))[k] = 0;
            const double ck = cost(j, f);
            ((double *)(ps + L::S_W))[wposk[k]] = ck;
            sts_f64(smem_u32(ps + L::S_CW) + cposk[k], ck);
        }
    constexpr unsigned kChild = 1u << 24;
    for (int it = 0; it < p.K; ++it) {
        if (it == 0) grp_bar(TH);   // costs, hierarchy weights of the links and zeroed link loads of every scenario are in place
        // ---- S2: customize the filled-edge weights of all scenarios at once ------------------------
        static_for<0, T::CZG_NB>([&](auto Bi) {
            constexpr int b = decltype(Bi)::value, TT = T::CZG_T[b], base = czg_off<T>(b);
            const unsigned ent = __ldg(cz + base + o);
            const uint32_t ea = wcb + (ent & 0x7fffu);
            double acc = lds_f64(ea);
            acc = (ent & 0x8000u) ? B200_INF_F64 : acc;   // pure shortcut: no link of its own
            static_for<0, TT>([&](auto Ti) {
                const unsigned tri = __ldg(cz + base + 32 * (1 + decltype(Ti)::value) + o);
                const double c = lds_f64(wcb + (tri & 0xffffu)) + lds_f64(wcb + (tri >> 16));
                acc = c < acc ? c : acc;
            });
            if (lane_live) sts_f64(ea, acc);
            grp_bar(TH);   // a scenario's edges of one level sit in two warps
        });
        // ---- one sweep up the elimination order, one sweep down; labels in registers ----------------
        double d[N];
        static_for<0, N>([&](auto R) {
            constexpr int r = decltype(R)::value, v = T::CORD[r], cnt = T::CUP_CNT[r], off = T::CUP_OFF[r];
            const double at_origin = __hiloint2double(v == orig ? 0 : 0x7ff00000, 0);
            if constexpr (cnt == 0) {
                d[v] = at_origin;
            } else {
                double c[cnt + 1];
                static_for<0, (cnt + 1) / 2>([&](auto J) {
                    constexpr int j = decltype(J)::value;
                    const double2 ww = lds_cost2(wcb, (off + 2 * j) * 8);
                    c[2 * j] = ww.x;
                    c[2 * j + 1] = ww.y;
                });
                static_for<0, cnt>([&](auto J) {
                    constexpr int j = decltype(J)::value, su = T::CSRC[off + j];
                    c[j] = d[su] + c[j];
                });
                min_tree<cnt>(c);
                d[v] = v == orig ? 0.0 : c[0];
            }
        });
        static_rfor<N, 0>([&](auto R) {
            constexpr int r = decltype(R)::value, v = T::CORD[r], cnt = T::CDN_CNT[r], off = T::CDN_OFF[r];
            if constexpr (cnt > 0) {
                double c[cnt + 1];
                static_for<0, (cnt + 1) / 2>([&](auto J) {
                    constexpr int j = decltype(J)::value;
                    const double2 ww = lds_cost2(wcb, (off + 2 * j) * 8);
                    c[2 * j] = ww.x;
                    c[2 * j + 1] = ww.y;
                });
                static_for<0, cnt>([&](auto J) {
                    constexpr int j = decltype(J)::value, su = T::CSRC[off + j];
                    c[j] = d[su] + c[j];
                });
                min_tree<cnt>(c);
                d[v] = c[0] < d[v] ? c[0] : d[v];
            }
        });
        // ---- predecessor selection + children counts (packed word: pending children << 24 | subtree sum) ----
        static_for<0, N>([&](auto I) { word[decltype(I)::value * TH] = (unsigned)Qc[decltype(I)::value * 32]; });
        unsigned orphan = 0;   // bit i: node i has no tree parent (the origin, or unreachable)
        static_for<0, N>([&](auto I) {
            constexpr int i = decltype(I)::value, k0 = T::KOFF[i], deg = T::KOFF[i + 1] - k0;
            const double dv = d[i];
            const bool reach = dv < B200_INF_F64;
            int bu = 255, bk = 255;
            double bdu = B200_INF_F64;
            double c[deg + 1];
            static_for<0, (deg + 1) / 2>([&](auto J) {
                constexpr int j = decltype(J)::value, wo = (woff<T>(i) + 2 * j) * 8;
                const double2 ww = lds_cost2(wbase, wo);
                c[2 * j] = ww.x;
                c[2 * j + 1] = ww.y;
            });
            static_for<0, deg>([&](auto J) {
                constexpr int j = decltype(J)::value, k = k0 + j, sk = T::SRC[k];
                const double du = d[sk];
                if (du + c[j] == dv && du <= bdu && reach) { bu = sk; bk = k; bdu = du; }
            });
            puk[i * TH] = (uint16_t)(bu | (bk << 8));
            const bool has = bu != 255;
            orphan |= has ? 0u : 1u << i;
            word[(has ? bu : i) * TH] += has ? kChild : 0u;
        });
        // ---- S3: subtree sums in Kahn order; link loads by atomics on the scenario's int32 array ------
        unsigned unq = 0;
        int top = 0;
#pragma unroll 4
        for (int i = 0; i < N; ++i) {
            const bool none = (orphan >> i) & 1u;
            const unsigned wi = word[i * TH];
            const bool tree = (i != orig) & !none;
            unq += none ? (wi & (kChild - 1u)) : 0u;
            const bool leaf = tree & (wi < kChild);
            if (leaf) stk[top * TH] = (uint8_t)i;
            top += leaf ? 1 : 0;
        }
        while (__any_sync(FULL, top > 0)) {
            if (top > 0) {
                const int v = stk[(--top) * TH];
                const unsigned uk = puk[v * TH];
                const int u = uk & 255u;
                const unsigned sv = word[v * TH];
                if (sv != 0u && lane_live) atomicAdd(auxi + (uk >> 8), (int)sv);
                const unsigned wu = word[u * TH] + sv - kChild;
                word[u * TH] = wu;
                if (wu < kChild && u != orig) stk[(top++) * TH] = (uint8_t)u;
            }
        }
        if (it == p.K - 1 && unq != 0u && lane_live) atomicAdd(red + 2 * s, (int)unq);
        grp_bar(TH);   // every scenario's link loads are complete
        // ---- S4: gap (last iteration), step size, flow update, new costs ----------------------------
        if (it == p.K - 1) {
            // per scenario sum f t and sum aux t: one warp per scenario, links dealt over its lanes
            for (int ss = t >> 5; ss < GS; ss += S::GW) {
                unsigned char *ps = scen0 + L::SCEN * ss;
                const float *fl = (const float *)(ps + L::S_FLOW);
                const int *ax = (const int *)(ps + L::S_AUX);
                const double *ww = (const double *)(ps + L::S_W);
                double sft = 0.0, sat = 0.0;
                for (int k = lane; k < E; k += 32) {
                    sft += (double)fl[k] * ww[wposk[k]];
                    sat += (double)(float)ax[k] * ww[wposk[k]];
                }
                sft = warp_sum_f64(sft);
                sat = warp_sum_f64(sat);
                if (lane == 0 && b0 + ss < p.B && (!p.active || p.active[b0 + ss]))
                    p.gap[b0 + ss] = sft > 0.0 ? (sft - sat) / sft : B200_INF_F64;
            }
            grp_bar(TH);
        }
        const double lam64 = p.method == B200_MSA ? 1.0 / (it + 1.0) : 2.0 / (it + 2.0);
        const float lam = (float)lam64, oml = (float)(1.0 - lam64);
#pragma unroll
        for (int j = 0; j < LPT; ++j)
            if ((pair_ok >> j) & 1u) {
                unsigned char *ps = scen0 + L::SCEN * pair_s[j];
                const int k = pair_k[j];
                float *fl = (float *)(ps + L::S_FLOW);
                int *ax = (int *)(ps + L::S_AUX);
                float f = radd<float>(rmul<float>(oml, fl[k]), rmul<float>(lam, (float)ax[k]));
                ax[k] = 0;   // the next iteration's link loads start from zero (same thread, same word)
                if (f != f) { f = 0; red[2 * pair_s[j] + 1] = 1; }
                fl[k] = f;
                const double ck = cost(j, f);
                ((double *)(ps + L::S_W))[wposk[k]] = ck;
                sts_f64(smem_u32(ps + L::S_CW) + cposk[k], ck);
            }
        grp_bar(TH);
    }
    // ---- outputs: flows, TSTT (numpy pairwise order over link ids), status ---------------------------
#pragma unroll
    for (int j = 0; j < LPT; ++j)
        if ((pair_ok >> j) & 1u) {
            unsigned char *ps = scen0 + L::SCEN * pair_s[j];
            const int k = pair_k[j], bb = b0 + pair_s[j];
            const float f = ((const float *)(ps + L::S_FLOW))[k];
            // link loads are dead: the array becomes the link-id-ordered scratch of the TSTT products
            ((float *)(ps + L::S_AUX))[eidk[k]] = rmul<float>(f, (float)((const double *)(ps + L::S_W))[wposk[k]]);
            if (bb < p.B && (!p.active || p.active[bb])) ((float *)p.flow)[(size_t)bb * E + eidk[k]] = f;
        }
    grp_bar(TH);
    for (int ss = t >> 5; ss < GS; ss += S::GW) {
        const int bb = b0 + ss;
        const float base = radd<float>(0.0f, warp_pairwise<float>((const float *)(scen0 + L::SCEN * ss + L::S_AUX), E));
        if (lane == 0 && bb < p.B && (!p.active || p.active[bb])) {
            const double total = p.total_demand > 1.0 ? p.total_demand : 1.0;
            const double un_total = (double)red[2 * ss];
            double v = (double)base / total;
            if (un_total > 0.0) v += p.penalty * (un_total / total);
            p.tstt[bb] = v;
            p.unassigned[bb] = un_total;
            p.status[bb] = (red[2 * ss + 1] ? B200_ST_NAN_FLOW : 0) | (un_total > 0.0 ? B200_ST_UNASSIGNED : 0);
        }
    }
}

// host: blob of the grouped kernel (integer demands), launch
template <class T>
int static_grp_prepare(b200_graph *g, const float *t0, const float *cap0, int num_od, const int32_t *od_o,
                       const int32_t *od_d, const double *od_q, const int32_t *origin_slot)
{
    using L = GrpLayout<T>;
    const int N = T::N, E = T::E;
    std::vector<int> pos(N);
    for (int i = 0; i < N; ++i) pos[T::NODE[i]] = i;
    std::vector<unsigned char> blob(L::BLOB_ALL, 0);
    for (int k = 0; k < num_od; ++k)
        ((int32_t *)blob.data())[(size_t)pos[od_d[k]] * 32 + origin_slot[od_o[k]]] += (int32_t)od_q[k];
    float *t0p = (float *)(blob.data() + L::TAB_T0), *capp = (float *)(blob.data() + L::TAB_CAP);
    uint8_t *eidp = blob.data() + L::TAB_EID, *wpp = blob.data() + L::TAB_WPOS;
    uint16_t *cpp = (uint16_t *)(blob.data() + L::TAB_CPOS);
    unsigned *czp = (unsigned *)(blob.data() + L::TAB_CZ);
    for (int k = 0; k < L::EP; ++k) { t0p[k] = 1.0f; capp[k] = 1.0f; cpp[k] = (uint16_t)((T::CW + 1) * 8); }
    for (int i = 0, o = 0; i < N; ++i) {
        for (int k = T::KOFF[i]; k < T::KOFF[i + 1]; ++k) {
            t0p[k] = t0[T::EID[k]];
            capp[k] = cap0[T::EID[k]];
            eidp[k] = (uint8_t)T::EID[k];
            wpp[k] = (uint8_t)(o + k - T::KOFF[i]);
            cpp[k] = (uint16_t)(T::CORIG[k] * 8);
        }
        o += ((T::KOFF[i + 1] - T::KOFF[i]) + 1) & ~1;
    }
    for (int i = 0; i < T::CZG_LEN; ++i) czp[i] = T::CZG_TAB[i];
    if (cudaMalloc((void **)&g->st_blob_grp, blob.size()) != cudaSuccess ||
        cudaMemcpy(g->st_blob_grp, blob.data(), blob.size(), cudaMemcpyHostToDevice) != cudaSuccess)
        return 0;
    return 1;
}

template <class T>
int launch_static_grp(const b200_graph *g, const Params &p, cudaStream_t st)
{
    using S = GrpShape<T>;
    using L = GrpLayout<T>;
    auto kern = assign_static_grp_kernel<T>;
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (getenv("B200_DEBUG_OCC")) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, S::TH, L::TOTAL);
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, kern);
        fprintf(stderr, "[b200] grouped static kernel: %d CTAs/SM (%d threads, %d B dynamic smem, %d regs)\n", nb, S::TH,
                L::TOTAL, fa.numRegs);
    }
    Params q = p;
    q.st_blob = g->st_blob_grp;
    kern<<<(p.B + S::GS - 1) / S::GS, S::TH, L::TOTAL, st>>>(q);
    B200_CUDA(cudaGetLastError());
    return 0;
}
