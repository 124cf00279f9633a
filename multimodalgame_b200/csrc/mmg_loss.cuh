// K_baseline_fwd, K_stats, K_lossgrad.
//
// K_baseline_fwd: both Baseline MLPs (model.py:496-516) for all T*B rows at once.  The baselines only consume
//   detached by-products of the conversation (model.py:835-843), so they leave the recurrent critical path and run
//   as two tiled GEMMs with a fused relu + linear2 epilogue.
// K_stats (1 CTA): get_rec_outp's stop-step selection (model.py:879-904), log_softmax / NLL / loglikelihood /
//   argmax / top-k (1264-1275, 1333-1338), and the batch-global statistics every REINFORCE term needs: per
//   (loss, step) the active-row count n_t (model.py:947,981) and the mean/variance of `logs - baseline` for
//   torch.std (915).  In a data-parallel run these statistics are what the ranks all-reduce.
// K_lossgrad: closed-form dLoss/d(probabilities) of calculate_loss_binary (907-927) and calculate_loss_bas (971-973)
//   with the multistep weighting (930-988) and the mask wiring (1248-1262); CTA 0 also evaluates the loss values.
#pragma once
#include "mmg_kernels.cuh"

namespace mmg {

// `dyn` / `dyn_floats`: dynamic shared memory for the small-K row tile (gemm_rows_tile); the generic tile is the fallback.
MMG_DEVICE void baseline_fwd_tile(const Dims& d, const ParamPtrs& P, const WsPtrs& W, const float* desc, int n_bas_tiles, int use_u,
                                  float* dyn, int dyn_floats) {
    MMG_SHARED __attribute__((aligned(16))) float gs[kGemmSmemFloats];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int ntm = cdiv(d.R, kTile), ntn = W.ntb;
    int t = blockIdx.x;
    if (t >= n_bas_tiles) {
        // extra tiles (fast path): wd = q . desc for all T*B rows (model.py:442-449), saved for the w_d gradient
        t -= n_bas_tiles;
        const int nwv = cdiv(d.WV, kTile);
        const int nt = t % nwv, mt = t / nwv;
        // -desc_attn: rows (q_class(n) a_n) over the NW words and desc = desc_set (model.py:444-449)
        const int Kw = d.A ? d.NW : d.D;
        const Operand A = Operand{d.A ? W.qa : W.q, nullptr, nullptr, nullptr, Kw, 0, 0, 0, 0, OP_PLAIN};
        const Operand Bo = Operand{desc, nullptr, nullptr, nullptr, d.WV, 0, 1, 0, 0, OP_PLAIN};
        float acc[4][4];
        gemm_tile_deep(A, Bo, d.R, d.WV, mt * kTile, nt * kTile, 0, Kw, acc, gs);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int r = mt * kTile + ty * 4 + a;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int v = nt * kTile + tx * 4 + c;
                if (r < d.R && v < d.WV) W.wd[(size_t)r * d.WV + v] = acc[a][c];
            }
        }
        return;
    }
    const int which = t / (ntm * ntn);     // 0: baseline_sen, 1: baseline_rec
    t %= ntm * ntn;
    const int nt = t % ntn, mt = t / ntn;
    Operand A, Bo;
    const float *b1, *w2;
    float *hid, *part;
    int K;
    const float* u = nullptr;     // per-example partial pre-activation (already includes the bias)
    if (which == 0 && use_u) {   // fast path: the h_x half of every row comes from U[b] (computed beside the exchange loop)
        A = Operand{W.rec_feats, nullptr, nullptr, nullptr, d.M, 0, 0, 0, 0, OP_PLAIN};
        Bo = Operand{P.p[MMG_P_BS_L1_W] + d.Hi, nullptr, nullptr, nullptr, d.Hi + d.M, 0, 0, 0, 0, OP_PLAIN};
        b1 = P.p[MMG_P_BS_L1_B]; w2 = P.p[MMG_P_BS_L2_W]; hid = W.h1s; part = W.bs_part; K = d.M; u = W.ubs;
    } else if (which == 0) {   // rows [h_x[b] ; z_r[t,b]]  (model.py:835-836), z_r[t] = rec_feats slot t
        A = Operand{W.h_x, W.rec_feats, nullptr, nullptr, d.Hi, d.M, 0, d.B, d.Hi, OP_PLAIN};
        Bo = Operand{P.p[MMG_P_BS_L1_W], nullptr, nullptr, nullptr, d.Hi + d.M, 0, 0, 0, 0, OP_PLAIN};
        b1 = P.p[MMG_P_BS_L1_B]; w2 = P.p[MMG_P_BS_L2_W]; hid = W.h1s; part = W.bs_part; K = d.Hi + d.M;
    } else {            // rows [z[t,b] ; h_z after step t]  (model.py:842-843)
        A = Operand{W.sen_feats, W.h_z + (size_t)d.B * d.Hr, nullptr, nullptr, d.M, d.Hr, 0, 0, d.M, OP_PLAIN};
        Bo = Operand{P.p[MMG_P_BR_L1_W], nullptr, nullptr, nullptr, d.M + d.Hr, 0, 0, 0, 0, OP_PLAIN};
        b1 = P.p[MMG_P_BR_L1_B]; w2 = P.p[MMG_P_BR_L2_W]; hid = W.h1r; part = W.br_part; K = d.M + d.Hr;
    }
    float acc[4][4];
    {
        // rows with the reduction index contiguous on both sides: the staged small-K tile when shapes and alignment allow
        const float *a1 = nullptr, *a2 = nullptr, *bw = nullptr;
        int lda1 = 0, ka1 = 0, lda2 = 0, ka2 = 0, ldb = 0;
        if (which == 0 && use_u) { a1 = W.rec_feats; lda1 = d.M; ka1 = d.M; bw = P.p[MMG_P_BS_L1_W] + d.Hi; ldb = d.Hi + d.M; }
        else if (which == 1) { a1 = W.sen_feats; lda1 = d.M; ka1 = d.M; a2 = W.h_z + (size_t)d.B * d.Hr; lda2 = d.Hr; ka2 = d.Hr;
                               bw = P.p[MMG_P_BR_L1_W]; ldb = d.M + d.Hr; }
        if (a1 != nullptr && rows_tile_ok(a1, lda1, ka1, a2, lda2, ka2, bw, ldb) && rows_tile_smem_floats(ka1 + ka2) <= dyn_floats)
            gemm_rows_tile(a1, lda1, ka1, a2, lda2, ka2, bw, ldb, d.R, d.Hb, mt * kTile, nt * kTile, acc, dyn);
        else
            gemm_tile_deep(A, Bo, d.R, d.Hb, mt * kTile, nt * kTile, 0, K, acc, gs);
    }
    MMG_TRACE_AT(2, 5);
    // epilogue: every load (bias or U row, linear2 weights) is issued before the first store, through the read-only path, so the
    // 16 elements of a thread cost one memory round trip instead of one per element (the stores may alias ordinary loads)
    const int n0 = nt * kTile + tx * 4;
    const bool vec = (d.Hb & 3) == 0 && n0 + 3 < d.Hb;
    float add[4][4], w2v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) w2v[c] = n0 + c < d.Hb ? ldg(w2 + n0 + c) : 0.f;
    if (u == nullptr) {
        float bv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) bv[c] = n0 + c < d.Hb ? ldg(b1 + n0 + c) : 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) add[a][c] = bv[c];
    } else {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int r = mt * kTile + ty * 4 + a;
            const float* ur = u + (size_t)((r < d.R ? r : 0) % d.B) * d.Hb + n0;
            if (vec) {
                const float4 t4 = ldg4(reinterpret_cast<const float4*>(ur));
                add[a][0] = t4.x; add[a][1] = t4.y; add[a][2] = t4.z; add[a][3] = t4.w;
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) add[a][c] = n0 + c < d.Hb ? ldg(ur + c) : 0.f;
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int r = mt * kTile + ty * 4 + a;
        float v[4], dot = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            v[c] = fmaxf(0.f, acc[a][c] + add[a][c]);
            if (n0 + c < d.Hb) dot = fmaf(v[c], w2v[c], dot);
        }
        if (r < d.R) {
            if (vec) *reinterpret_cast<float4*>(hid + (size_t)r * d.Hb + n0) = make_float4(v[0], v[1], v[2], v[3]);
            else {
#pragma unroll
                for (int c = 0; c < 4; ++c) if (n0 + c < d.Hb) hid[(size_t)r * d.Hb + n0 + c] = v[c];
            }
        }
        dot = half_warp_sum(dot);
        if (tx == 0 && r < d.R) part[(size_t)r * ntn + nt] = dot;
    }
}

// Baseline scores = linear2 bias + the per-tile partial dot products of K_baseline_fwd (summed in tile order).
MMG_GLOBAL void __launch_bounds__(256)
k_baseline_finish(Dims d, ParamPtrs P, WsPtrs W) {
    pdl_wait();                 // PDL: the previous kernel of the stream has completed and flushed
    pdl_launch_dependents();    // let the next kernel's CTAs be scheduled behind this grid
    const float b2s = ldg(P.p[MMG_P_BS_L2_B]), b2r = ldg(P.p[MMG_P_BR_L2_B]);
    for (int r = blockIdx.x * 256 + threadIdx.x; r < d.R; r += gridDim.x * 256) {
        float s = b2s, q = b2r;
        for (int j0 = 0; j0 < W.ntb; j0 += 8) {          // 16 partial loads in flight, added in tile order
            float ps[8], pq[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                ps[u] = j0 + u < W.ntb ? W.bs_part[(size_t)r * W.ntb + j0 + u] : 0.f;
                pq[u] = j0 + u < W.ntb ? W.br_part[(size_t)r * W.ntb + j0 + u] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { s += ps[u]; q += pq[u]; }
        }
        W.bs[r] = s; W.br[r] = q;
    }
}

enum { kStatsThreadsMax = 1024 };

MMG_DEVICE unsigned char mask_at(const Dims& d, const WsPtrs& W, int slot, int b) {
    return d.fixed ? (unsigned char)1 : W.stop_mask[(size_t)slot * d.B + b];
}

// Body of K_stats for a CTA of NT threads.  `per_example_done`: the conversation kernel's epilogue has already produced the
// per-example results (ystep, outp, g_outp, logs, argmax, hit); only the batch sums remain.
template <int NT>
MMG_DEVICE void stats_body(const Dims& d, const ParamPtrs& P, const WsPtrs& W, const ExchangeInputs& in, const PeerView& pv,
                           bool per_example_done) {
    constexpr int kStatsThreads = NT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kStatsThreads / 32;
    // ---- finalize baseline scores: per-tile partial dots added in tile order; two rows per pass, all loads in flight --------
    const float b2s = ldg(P.p[MMG_P_BS_L2_B]), b2r = ldg(P.p[MMG_P_BR_L2_B]);
    for (int r0 = tid; r0 < d.R; r0 += 2 * kStatsThreads) {
        const int r1 = r0 + kStatsThreads;
        const bool two = r1 < d.R;
        float s0 = b2s, q0 = b2r, s1 = b2s, q1 = b2r;
        for (int j0 = 0; j0 < W.ntb; j0 += 8) {
            float ps[2][8], pq[2][8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool ok = j0 + u < W.ntb;
                ps[0][u] = ok ? W.bs_part[(size_t)r0 * W.ntb + j0 + u] : 0.f;
                pq[0][u] = ok ? W.br_part[(size_t)r0 * W.ntb + j0 + u] : 0.f;
                ps[1][u] = ok && two ? W.bs_part[(size_t)r1 * W.ntb + j0 + u] : 0.f;
                pq[1][u] = ok && two ? W.br_part[(size_t)r1 * W.ntb + j0 + u] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { s0 += ps[0][u]; q0 += pq[0][u]; s1 += ps[1][u]; q1 += pq[1][u]; }
        }
        W.bs[r0] = s0; W.br[r0] = q0;
        if (two) { W.bs[r1] = s1; W.br[r1] = q1; }
    }
    // ---- per example (one HALF-warp each, lanes over classes; 64 examples per round): prediction step, log-softmax,
    //      log-likelihood, argmax, top-k, dNLL/d outp -----------------------------------------------------------------
    MMG_SHARED double s_red[2][kStatsThreadsMax / 16];
    double nll_local = 0.0, correct_local = 0.0;
    const float invB = 1.0f / (float)d.Bg;
    const int hl = tid & 15, hwid = tid >> 4, nhw = kStatsThreads / 16;
    if (per_example_done) {     // the conversation kernel left logs[] / hit[]: every thread adds its rows, tree below
        for (int b = tid; b < d.B; b += kStatsThreads) { nll_local -= (double)W.logs[b]; correct_local += (double)W.hit[b]; }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {      // half-warp tree so that the hl == 0 lanes hold the partials the code below expects
            nll_local += shfl_xor_d(nll_local, o); correct_local += shfl_xor_d(correct_local, o);
        }
    } else
    for (int b0 = 0; b0 < d.B; b0 += nhw) {
        const int b = b0 + hwid;
        const bool ok = b < d.B;
        const int bb = ok ? b : 0;
        int ts = d.T - 1;
        if (!d.fixed) {   // first step whose outgoing mask is 0 (model.py:893-896); the last mask is forced to 0 (870)
            float first = (float)(d.T - 1);
            for (int t = hl; t < d.T; t += 16)
                if (W.stop_mask[(size_t)(t + 1) * d.B + bb] == 0) first = fminf(first, (float)t);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) first = fminf(first, shfl_xor_f(first, o));
            ts = (int)first;
        }
        const float* yy = W.y + ((size_t)ts * d.B + bb) * d.D;
        const int tg = (int)in.target[bb];
        const float ytg = yy[tg];
        float mx = -INFINITY;
        for (int dd = hl; dd < d.D; dd += 16) mx = fmaxf(mx, yy[dd]);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, shfl_xor_f(mx, o));
        float am = 3.0e9f, se = 0.f, rank = 0.f;
        for (int dd = hl; dd < d.D; dd += 16) {
            const float v = yy[dd];
            if (v == mx) am = fminf(am, (float)dd);          // first index of the maximum, like a serial `>` scan
            se += expf(v - mx);
            if (v > ytg) rank += 1.f;
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) am = fminf(am, shfl_xor_f(am, o));
        se = group_sum<16>(se);
        rank = group_sum<16>(rank);
        const float lse = mx + logf(se);
        const float lt = ytg - lse;
        if (ok) {
            for (int dd = hl; dd < d.D; dd += 16) {
                const float v = yy[dd];
                W.outp[(size_t)b * d.D + dd] = v;
                W.g_outp[(size_t)b * d.D + dd] = (expf(v - lse) - (dd == tg ? 1.f : 0.f)) * invB;   // d nll / d outp
            }
            if (hl == 0) {
                W.ystep[b] = ts;
                W.logs[b] = lt;
                W.argmax[b] = (int)am;
                nll_local -= (double)lt;
                if ((int)rank < in.top_k) correct_local += 1.0;
            }
        }
    }
    if (hl == 0) { s_red[0][hwid] = nll_local; s_red[1][hwid] = correct_local; }
    MMG_SYNCTHREADS();
    if (tid == 0) {
        double a = 0, c = 0;
        for (int w = 0; w < nhw; ++w) { a += s_red[0][w]; c += s_red[1][w]; }
        W.stats[stat_scalar(d, 0)] = a;
        W.stats[stat_scalar(d, 1)] = c;
        W.stats[stat_scalar(d, 2)] = 0.0;
        W.stats[stat_scalar(d, 3)] = 0.0;
    }
    // ---- batch statistics per (loss kind, step): one warp per STEP, all kinds in one pass over the rows (shared loads) ------
    // kind 0: sender messages  (baseline bs[t], mask s_masks[t])      model.py:1258,1291
    // kind 1: receiver message (baseline br[t], mask s_masks[t+1], t <= T-2)   model.py:1257,1285-1286
    // kind 2: stop bit         (baseline br[t], mask s_masks[t])      model.py:1256,1279-1280
    // baselines: SSE of both over s_masks[t]; n_mask = rows still active after step t
    for (int t = warp; t < d.T; t += nwarps) {
        double n0 = 0, a0 = 0, c0 = 0, n1 = 0, a1 = 0, c1 = 0, n2 = 0, a2 = 0, c2 = 0, er2 = 0, es2 = 0, nm = 0;
#pragma unroll 2
        for (int b = lane; b < d.B; b += 32) {
            const float lg = W.logs[b];
            const float bsv = W.bs[(size_t)t * d.B + b], brv = W.br[(size_t)t * d.B + b];
            const bool m_in = mask_at(d, W, t, b) != 0, m_out = mask_at(d, W, t + 1, b) != 0;
            const double ws = (double)(lg - bsv), wr = (double)(lg - brv);
            if (m_in) {
                n0 += 1.0; a0 += ws; c0 += ws * ws;
                if (!d.fixed) { n2 += 1.0; a2 += wr; c2 += wr * wr; }
                er2 += wr * wr; es2 += ws * ws;
            }
            if (m_out && t < d.T - 1) { n1 += 1.0; a1 += wr; c1 += wr * wr; }
            if (d.fixed || m_out) nm += 1.0;
        }
        n0 = warp_sum_d(n0); a0 = warp_sum_d(a0); c0 = warp_sum_d(c0);
        n1 = warp_sum_d(n1); a1 = warp_sum_d(a1); c1 = warp_sum_d(c1);
        n2 = warp_sum_d(n2); a2 = warp_sum_d(a2); c2 = warp_sum_d(c2);
        er2 = warp_sum_d(er2); es2 = warp_sum_d(es2); nm = warp_sum_d(nm);
        if (lane == 0) {
            W.stats[stat_idx(d, 0, t, 0)] = n0; W.stats[stat_idx(d, 0, t, 1)] = a0; W.stats[stat_idx(d, 0, t, 2)] = c0;
            W.stats[stat_idx(d, 1, t, 0)] = n1; W.stats[stat_idx(d, 1, t, 1)] = a1; W.stats[stat_idx(d, 1, t, 2)] = c1;
            W.stats[stat_idx(d, 2, t, 0)] = n2; W.stats[stat_idx(d, 2, t, 1)] = a2; W.stats[stat_idx(d, 2, t, 2)] = c2;
            W.stats[stat_bas(d, t, 0)] = er2; W.stats[stat_bas(d, t, 1)] = es2; W.stats[stat_bas(d, t, 2)] = nm;
        }
    }
    if (pv.world > 1) {
        MMG_SYNCTHREADS();
        // PUSH as self-validating packets (ll_store): this rank's statistics go into slot `rank` of EVERY rank's symmetric statistics
        // area; no flag and no fence, the consumers spin on the packets' tags in their own memory
        const int cnt = stats_count(d);
        for (int i = tid; i < cnt * pv.world; i += kStatsThreads)
            ll_store(pv.stats[i / cnt] + 2 * (size_t)pv.rank * cnt, i % cnt, W.stats[i % cnt], pv.iter);
    }
}

// Per (kind, t) coefficients derived from the (all-reduced) statistics.
struct LossCoef { float cA, cE; };   // REINFORCE scale (step weight / n / max(1, std)), entropy scale (lambda * step weight / n)

MMG_DEVICE void loss_coefs(const Dims& d, const mmg_config& cfg, const double* st, int kind, int t, double tot_kind,
                           LossCoef& out) {
    const double n = st[stat_idx(d, kind, t, 0)];
    out.cA = 0.f; out.cE = 0.f;
    if (n <= 0.0) return;
    const int steps = (kind == 1) ? d.T - 1 : d.T;                 // list lengths, model.py:967
    const double sw = d.fixed ? 1.0 / (double)steps : n / tot_kind;   // model.py:960-961, tot_kind = sum_t n_t of this loss
    double inv = 1.0;
    if (n > 1.0) {                                                 // model.py:914-915, unbiased std
        const double s1 = st[stat_idx(d, kind, t, 1)], s2 = st[stat_idx(d, kind, t, 2)];
        double var = (s2 - s1 * s1 / n) / (n - 1.0);
        if (var < 0) var = 0;
        const double sd = (double)(float)sqrt(var);
        inv = 1.0 / (sd > 1.0 ? sd : 1.0);
    }
    const int has_ent = kind == 0 ? cfg.has_entropy_sen : (kind == 1 ? cfg.has_entropy_rec : cfg.has_entropy_s);
    const float lam = kind == 0 ? cfg.entropy_sen : (kind == 1 ? cfg.entropy_rec : cfg.entropy_s);
    out.cA = (float)(sw / n * inv);
    out.cE = has_ent ? (float)((double)lam * sw / n) : 0.f;
}

enum { kLossThreads = 256 };

// d/dp of  -w * [f log(p+e) + (1-f) log(1-p+e)] * cA  +  cE * [p log(p+e) + (1-p) log(1-p+e)]
MMG_DEVICE float binary_grad(float p, float f, float wcA, float cE) {
    const float e = 1e-8f;
    const float a = p + e, bb = 1.f - p + e;
    float g = -wcA * (f / a - (1.f - f) / bb);
    if (cE != 0.f) g += cE * (logf(a) + p / a - logf(bb) - (1.f - p) / bb);
    return g;
}

// Single-rank runs: the statistics are final as soon as they are computed, so the CTA that computes them also derives the
// loss coefficients once (W.coefs: [3][T] LossCoef, then 1 / denominator of the baseline MSE) instead of every consumer CTA.
MMG_DEVICE void loss_coefs_store(const Dims& d, const mmg_config& cfg, const WsPtrs& W, int nthreads) {
    MMG_SHARED double tot_c[3];
    const int tid = threadIdx.x;
    const double* st = W.stats;
    if (tid < 3) {
        const int steps = (tid == 1) ? d.T - 1 : d.T;
        double v = 0.0;
        for (int t = 0; t < steps; ++t) v += st[stat_idx(d, tid, t, 0)];
        tot_c[tid] = v;
    }
    MMG_SYNCTHREADS();
    LossCoef* out = reinterpret_cast<LossCoef*>(W.coefs);
    for (int i = tid; i < 3 * d.T; i += nthreads) loss_coefs(d, cfg, st, i / d.T, i % d.T, d.fixed ? 0.0 : tot_c[i / d.T], out[i]);
    if (tid == 0) {
        const double tot = d.fixed ? (double)d.Bg * d.T : tot_c[0];
        W.coefs[6 * d.T] = tot > 0 ? (float)(1.0 / tot) : 0.f;
    }
}

// ---- fused statistics, two levels (B a multiple of the tile height, so a 64-row tile lies inside one exchange step) -------------
// Level 1, by the last of the 2 * ntb CTAs that finish row tile `mt`: baseline scores of its 64 rows (per-tile partial dots added
// in tile order) and the rows' share of the 12 per-step sums -> loss_part[mt][12] (double; the array is free until the backward).
MMG_DEVICE void row_tile_stats(const Dims& d, const ParamPtrs& P, const WsPtrs& W, int mt) {
    // compact on purpose: this runs once per row tile, and straight-line code executes at instruction-fetch speed (12 unrolled
    // double-precision shuffle trees were ~400 instructions); the 64 rows meet in shared memory and 12 threads add them in row order
    MMG_SHARED double red[12][kTile + 1];
    const int tid = threadIdx.x;
    if (tid < kTile) {
        const int r = mt * kTile + tid;
        const int t = r / d.B, b = r - t * d.B;
        float s = ldg(P.p[MMG_P_BS_L2_B]), q = ldg(P.p[MMG_P_BR_L2_B]);
        for (int j0 = 0; j0 < W.ntb; j0 += 8) {
            float ps[8], pq[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool ok = j0 + u < W.ntb;
                ps[u] = ok ? ld_cg(W.bs_part + (size_t)r * W.ntb + j0 + u) : 0.f;
                pq[u] = ok ? ld_cg(W.br_part + (size_t)r * W.ntb + j0 + u) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { s += ps[u]; q += pq[u]; }
        }
        W.bs[r] = s; W.br[r] = q;
        const float lg = W.logs[b];
        const bool m_in = mask_at(d, W, t, b) != 0, m_out = mask_at(d, W, t + 1, b) != 0;
        const double ws = (double)(lg - s), wr = (double)(lg - q);
        const double on_in = m_in ? 1.0 : 0.0, on_rec = (m_out && t < d.T - 1) ? 1.0 : 0.0, on_s = (m_in && !d.fixed) ? 1.0 : 0.0;
        // n0 a0 c0 | n1 a1 c1 | n2 a2 c2 | er2 es2 nm  (see stats_body)
        red[0][tid] = on_in;  red[1][tid] = on_in * ws;   red[2][tid] = on_in * ws * ws;
        red[3][tid] = on_rec; red[4][tid] = on_rec * wr;  red[5][tid] = on_rec * wr * wr;
        red[6][tid] = on_s;   red[7][tid] = on_s * wr;    red[8][tid] = on_s * wr * wr;
        red[9][tid] = on_in * wr * wr; red[10][tid] = on_in * ws * ws; red[11][tid] = (d.fixed || m_out) ? 1.0 : 0.0;
    }
    MMG_SYNCTHREADS();
    if (tid < 12) {
        double sum = 0.0;
#pragma unroll 4
        for (int i = 0; i < kTile; ++i) sum += red[tid][i];
        W.loss_part[(size_t)mt * 12 + tid] = sum;
    }
}

// Level 2, by the last row tile to finish: per step, the row tiles' shares added in tile order; the batch sums of the
// per-example results; then the peers are told (data-parallel) or the loss coefficients derived (single rank).
MMG_DEVICE void final_stats(const Dims& d, const WsPtrs& W, const PeerView& pv, const mmg_config& cfg) {
    MMG_SHARED double red2[2][kGemmThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per_step = d.B / kTile;
    for (int i = tid; i < d.T * 12; i += kGemmThreads) {
        const int t = i / 12, k = i - 12 * t;
        double s = 0.0;
        for (int j = 0; j < per_step; ++j) s += ld_cg_d(W.loss_part + (size_t)(t * per_step + j) * 12 + k);
        W.stats[k < 9 ? stat_idx(d, k / 3, t, k % 3) : stat_bas(d, t, k - 9)] = s;
    }
    double nl = 0.0, co = 0.0;
    for (int b = tid; b < d.B; b += kGemmThreads) { nl -= (double)W.logs[b]; co += (double)W.hit[b]; }
    nl = warp_sum_d(nl); co = warp_sum_d(co);
    if (lane == 0) { red2[0][warp] = nl; red2[1][warp] = co; }
    MMG_SYNCTHREADS();
    if (tid == 0) {
        double a = 0, c = 0;
        for (int w = 0; w < kGemmThreads / 32; ++w) { a += red2[0][w]; c += red2[1][w]; }
        W.stats[stat_scalar(d, 0)] = a;
        W.stats[stat_scalar(d, 1)] = c;
        W.stats[stat_scalar(d, 2)] = 0.0;
        W.stats[stat_scalar(d, 3)] = 0.0;
    }
    MMG_SYNCTHREADS();
    if (pv.world > 1) {
        // PUSH as self-validating packets (ll_store): this rank's statistics go into slot `rank` of EVERY rank's symmetric statistics
        // area; no flag and no fence, the consumers spin on the packets' tags in their own memory
        const int cnt = stats_count(d);
        for (int i = tid; i < cnt * pv.world; i += kGemmThreads)
            ll_store(pv.stats[i / cnt] + 2 * (size_t)pv.rank * cnt, i % cnt, W.stats[i % cnt], pv.iter);
    } else {
        loss_coefs_store(d, cfg, W, kGemmThreads);
    }
}

// `fuse_stats`: the last CTA to finish its tile (ticket) also runs the batch statistics, so K_stats is not launched: the
// per-example half was done by the conversation kernel's epilogue.
MMG_GLOBAL void __launch_bounds__(kGemmThreads)
k_baseline_fwd(Dims d, ParamPtrs P, WsPtrs W, const float* desc, int n_bas_tiles, int use_u, int fuse_stats, ExchangeInputs in,
               PeerView pv, mmg_config cfg, int dyn_floats) {
    pdl_wait();                 // PDL: the previous kernel of the stream has completed and flushed
    pdl_launch_dependents();    // let the next kernel's CTAs be scheduled behind this grid
    MMG_TRACE_AT(2, 0);
    MMG_DYN_SMEM(dyn_raw);
    baseline_fwd_tile(d, P, W, desc, n_bas_tiles, use_u, reinterpret_cast<float*>(dyn_raw), dyn_floats);
    MMG_TRACE_AT(2, 1);
    if (!fuse_stats) return;
    MMG_SHARED int s_last;
    {
        const int ntm = cdiv(d.R, kTile), ntn = W.ntb;
        if (d.B % kTile == 0 && ntm * 12 <= kLossCtasMax * 8) {
            if ((int)blockIdx.x >= n_bas_tiles) return;                   // the wd tiles take no part in the statistics
            const int mt = ((int)blockIdx.x % (ntm * ntn)) / ntn;
            MMG_SYNCTHREADS();                                             // the barrier orders every thread's stores before thread 0's fence + ticket
            if (threadIdx.x == 0) s_last = (ticket_take(W.tile_tickets + mt) == 2u * ntn - 1) ? 1 : 0;
            MMG_SYNCTHREADS();
            if (!s_last) return;
            fence_acquire();
            if (threadIdx.x == 0) W.tile_tickets[mt] = 0;                 // zero between launches (K_wgrad uses the same counters)
            MMG_TRACE_AT(2, 2);
            row_tile_stats(d, P, W, mt);
            MMG_SYNCTHREADS();
            if (threadIdx.x == 0) s_last = (ticket_take(W.tickets + 7) == (unsigned)ntm - 1) ? 1 : 0;
            MMG_SYNCTHREADS();
            if (!s_last) return;
            fence_acquire();
            if (threadIdx.x == 0) W.tickets[7] = 0;
            MMG_TRACE_AT(2, 3);
            final_stats(d, W, pv, cfg);
            MMG_TRACE_AT(2, 4);
            return;
        }
    }
    fence_acquire();            // every thread: its tile results are visible device-wide before the CTA is counted
    MMG_SYNCTHREADS();
    if (threadIdx.x == 0) s_last = (ticket_take(W.tickets + 7) == gridDim.x - 1) ? 1 : 0;
    MMG_SYNCTHREADS();
    if (!s_last) return;
    fence_acquire();
    if (threadIdx.x == 0) W.tickets[7] = 0;
    MMG_TRACE_AT(2, 2);
    stats_body<kGemmThreads>(d, P, W, in, pv, true);
    MMG_TRACE_AT(2, 3);
    if (pv.world <= 1) {
        MMG_SYNCTHREADS();
        loss_coefs_store(d, cfg, W, kGemmThreads);
    }
    MMG_TRACE_AT(2, 4);
}

MMG_GLOBAL void __launch_bounds__(kStatsThreadsMax)
k_stats(Dims d, ParamPtrs P, WsPtrs W, ExchangeInputs in, PeerView pv) {
    pdl_wait();                 // PDL: the previous kernel of the stream has completed and flushed
    pdl_launch_dependents();    // let the next kernel's CTAs be scheduled behind this grid
    stats_body<kStatsThreadsMax>(d, P, W, in, pv, false);
}

MMG_HOST_DEVICE int loss_smem_bytes(const Dims& d, int world) {
    return 3 * d.T * (int)sizeof(LossCoef) + 16 + (world > 1 ? stats_count(d) * 8 + 16 : 0);
}

// Shared by K_lossgrad and the fused backward kernel (256 threads): global statistics (summed over the peers when
// data-parallel), per-(loss, step) coefficients and the baseline-MSE scale.  Returns the statistics to use.
MMG_DEVICE const double* loss_prologue(const Dims& d, const mmg_config& cfg, const WsPtrs& W, const PeerView& pv,
                                       unsigned char* smem, LossCoef*& coef, float*& bas_scale) {
    coef = reinterpret_cast<LossCoef*>(smem);                          // [3][T]
    bas_scale = reinterpret_cast<float*>(coef + 3 * d.T);              // [2]: 1 / denominator of the baseline MSE
    const int tid = threadIdx.x;
    const double* st = W.stats;
    if (pv.world > 1) {
        // global batch statistics = sum over ranks (rank order) of the slots the peers pushed into THIS rank's memory
        double* st_s = reinterpret_cast<double*>(bas_scale + 4);
        const int cnt = stats_count(d);
        for (int i = tid; i < cnt; i += kLossThreads) {
            double v = 0.0;
            for (int r = 0; r < pv.world; ++r) v += ll_load(pv.stats[pv.rank] + 2 * (size_t)r * cnt, i, pv.iter, pv.error);
            st_s[i] = v;
            if (blockIdx.x == 0) W.stats[i] = v;      // later kernels (K_update) read the global numbers from here
        }
        MMG_SYNCTHREADS();
        st = st_s;
    }
    // active-row totals per loss kind: one warp each, lanes over steps (a serial accumulate would pay one memory round
    // trip per step); only needed for adaptive-length conversations
    MMG_SHARED double tot_s[3];
    if (!d.fixed && tid < 96) {
        const int kind = tid >> 5, lane = tid & 31;
        const int steps = (kind == 1) ? d.T - 1 : d.T;
        double v = 0.0;
        for (int t = lane; t < steps; t += 32) v += st[stat_idx(d, kind, t, 0)];
        v = warp_sum_d(v);
        if (lane == 0) tot_s[kind] = v;
    }
    MMG_SYNCTHREADS();
    for (int i = tid; i < 3 * d.T; i += kLossThreads)
        loss_coefs(d, cfg, st, i / d.T, i % d.T, d.fixed ? 0.0 : tot_s[i / d.T], coef[i]);
    if (tid == 0) {
        // sum_t n_t over s_masks[:-1] (model.py:983); fixed: B_global * T (mean over steps of batch means)
        const double tot = d.fixed ? (double)d.Bg * d.T : tot_s[0];
        bas_scale[0] = tot > 0 ? (float)(1.0 / tot) : 0.f;
    }
    MMG_SYNCTHREADS();
    return st;
}

// Loss values: per-CTA partials (acc: 0 binary_sen, 1 binary_rec, 2 binary_s, 3 bas_rec, 4 bas_sen — rank-local
// contributions, any thread may hold a share), summed in CTA order by the last CTA to finish (deterministic).
// Per-CTA partial of the five loss sums -> loss_part[cta][0..4] (256 threads).
MMG_DEVICE void loss_partials(const WsPtrs& W, const double (&acc)[5]) {
    MMG_SHARED double red[5][kLossThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = 0; i < 5; ++i) {
        const double v = warp_sum_d(acc[i]);
        if (lane == 0) red[i][warp] = v;
    }
    MMG_SYNCTHREADS();
    if (tid < 5) {
        double v = 0;
        for (int w = 0; w < kLossThreads / 32; ++w) v += red[tid][w];
        W.loss_part[(size_t)blockIdx.x * 8 + tid] = v;
    }
}

// The loss values from `nparts` per-CTA partials (added in CTA order: deterministic) and the global statistics in W.stats;
// one CTA of 256 threads, after every partial is visible.
// `bump_counter`: count this iteration in the receiver message head's own update counter (torch.optim.Adam keeps one step count
// per parameter, and that head only steps when it received a gradient).  False when an earlier kernel of the fused sequence has
// already done it (K_wgrad), so that K_update reads a value nobody writes while it runs.
MMG_DEVICE void loss_finalize(const Dims& d, const WsPtrs& W, int nparts, bool bump_counter = true, float* host_out = nullptr) {
    MMG_SHARED double red[5][kLossThreads / 32];
    MMG_SHARED double nll_red[kLossThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* st = W.stats;
    double nl = 0;
    for (int b = tid; b < d.B; b += kLossThreads) nl -= (double)W.logs[b] / (double)d.Bg;
    nl = warp_sum_d(nl);
    if (lane == 0) nll_red[warp] = nl;
    // per-CTA partials: thread c sums CTAs c, c + 256, ...; then a fixed shuffle tree and 8 warp totals
    double part[5] = {0, 0, 0, 0, 0};
    for (int c = tid; c < nparts; c += kLossThreads)
        for (int i = 0; i < 5; ++i) part[i] += W.loss_part[(size_t)c * 8 + i];
    for (int i = 0; i < 5; ++i) {
        const double v = warp_sum_d(part[i]);
        if (lane == 0) red[i][warp] = v;
    }
    MMG_SYNCTHREADS();
    if (tid == 0) {
        double v[6];
        v[0] = 0;
        for (int w = 0; w < kLossThreads / 32; ++w) v[0] += nll_red[w];
        for (int i = 0; i < 5; ++i) {
            double a = 0;
            for (int w = 0; w < kLossThreads / 32; ++w) a += red[i][w];
            v[i + 1] = a;
        }
        float* L = W.losses;
        L[MMG_LOSS_NLL] = (float)v[0];
        L[MMG_LOSS_BINARY_SEN] = (float)v[1];
        L[MMG_LOSS_BINARY_REC] = (float)v[2];
        L[MMG_LOSS_BINARY_S] = (float)v[3];
        L[MMG_LOSS_BAS_REC] = (float)v[4];
        L[MMG_LOSS_BAS_SEN] = (float)v[5];
        L[MMG_LOSS_REC] = (float)(v[0] + v[2] + (d.fixed ? 0.0 : v[3]));      // model.py:1296-1300
        L[MMG_LOSS_SEN] = (float)v[1];                                          // model.py:1301
        L[MMG_LOSS_TOPK_CORRECT] = (float)st[stat_scalar(d, 1)];
        int tp = d.T;
        if (!d.fixed) for (int t = 0; t < d.T; ++t) if (st[stat_bas(d, t, 2)] <= 0.0) { tp = t + 1; break; }
        L[MMG_LOSS_ACTIVE_STEPS] = (float)tp;
        if (bump_counter && st[stat_idx(d, 1, 0, 0)] > 0.0) W.opt_counters[0] += 1;   // updates seen by the receiver message head
        for (int i = MMG_LOSS_ACTIVE_STEPS + 1; i < MMG_LOSS_COUNT; ++i) L[i] = 0.f;
        if (host_out != nullptr) {        // mapped pinned host memory (mmg_inputs.h_losses_out): four 16-byte stores over PCIe
#pragma unroll
            for (int i = 0; i < MMG_LOSS_COUNT; i += 4)
                *reinterpret_cast<float4*>(host_out + i) = make_float4(L[i], L[i + 1], L[i + 2], L[i + 3]);
        }
    }
}

// Loss values: per-CTA partials (acc: 0 binary_sen, 1 binary_rec, 2 binary_s, 3 bas_rec, 4 bas_sen — rank-local
// contributions, any thread may hold a share), summed in CTA order by the last CTA to finish (deterministic).
MMG_DEVICE void loss_epilogue(const Dims& d, const WsPtrs& W, const double (&acc)[5]) {
    MMG_SHARED int s_last;
    loss_partials(W, acc);
    fence_acquire();
    MMG_SYNCTHREADS();
    if (threadIdx.x == 0) s_last = (ticket_take(W.tickets) == gridDim.x - 1) ? 1 : 0;
    MMG_SYNCTHREADS();
    if (!s_last) return;
    fence_acquire();
    loss_finalize(d, W, (int)gridDim.x);
    if (threadIdx.x == 0) W.tickets[0] = 0;                              // ready for the next launch
}

MMG_GLOBAL void __launch_bounds__(kLossThreads)
k_lossgrad(Dims d, mmg_config cfg, WsPtrs W, PeerView pv) {
    pdl_wait();                 // PDL: the previous kernel of the stream has completed and flushed
    pdl_launch_dependents();    // let the next kernel's CTAs be scheduled behind this grid
    MMG_DYN_SMEM(smem_raw);
    LossCoef* coef;
    float* bas_scale;
    const double* st = loss_prologue(d, cfg, W, pv, smem_raw, coef, bas_scale);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    (void)tid;
    const bool binary = d.use_binary != 0;
    // ---- one warp per (t, b) row: upstream gradients + this row's share of every loss value ----------------------
    double acc[5] = {0, 0, 0, 0, 0};
    const int rows_per_cta = kLossThreads / 32;
    for (int row = blockIdx.x * rows_per_cta + warp; row < d.R; row += gridDim.x * rows_per_cta) {
        const int t = row / d.B, b = row % d.B;
        const float lg = W.logs[b];
        const bool m_in = mask_at(d, W, t, b) != 0;           // active entering step t
        const bool m_out = mask_at(d, W, t + 1, b) != 0;      // still active after step t
        const float bsv = W.bs[row], brv = W.br[row];
        if (binary) {
            const LossCoef c0 = coef[0 * d.T + t], c1 = coef[1 * d.T + t];
            const float w0 = m_in ? (lg - bsv) * c0.cA : 0.f, e0 = m_in ? c0.cE : 0.f;
            const bool rec_on = m_out && (t < d.T - 1);
            const float w1 = rec_on ? (lg - brv) * c1.cA : 0.f, e1 = rec_on ? c1.cE : 0.f;
            // log-likelihood and (neg)entropy sums of the two messages: the rows of calculate_loss_binary (model.py:907-921)
            float lp0 = 0.f, hh0 = 0.f, lp1 = 0.f, hh1 = 0.f;
            for (int j = lane; j < d.M; j += 32) {
                const size_t i = (size_t)row * d.M + j;
                float gs = 0.f, gr = 0.f;
                if (m_in) {
                    const float p = W.sen_probs[i], f = W.sen_feats[i];
                    const float l1 = logf(p + 1e-8f), l0 = logf(1.f - p + 1e-8f);
                    lp0 += f * l1 + (1.f - f) * l0;
                    hh0 += p * l1 + (1.f - p) * l0;
                    gs = binary_grad(p, f, w0, e0);
                }
                if (rec_on) {
                    const float p = W.rec_probs[i], f = W.rec_feats[i + (size_t)d.B * d.M];
                    const float l1 = logf(p + 1e-8f), l0 = logf(1.f - p + 1e-8f);
                    lp1 += f * l1 + (1.f - f) * l0;
                    hh1 += p * l1 + (1.f - p) * l0;
                    gr = binary_grad(p, f, w1, e1);
                }
                W.g_sen_probs[i] = gs;
                W.g_rec_probs[i] = gr;
            }
            lp0 = warp_sum(lp0); hh0 = warp_sum(hh0); lp1 = warp_sum(lp1); hh1 = warp_sum(hh1);
            if (lane == 0) {
                float gs = 0.f;
                if (m_in) {
                    acc[0] += (double)(-(lg - bsv) * c0.cA) * lp0 + (double)c0.cE * hh0;
                    if (!d.fixed) {
                        const LossCoef c2 = coef[2 * d.T + t];
                        const float sp = W.stop_prob[row], sf = W.stop_feat[row];
                        const float l1 = logf(sp + 1e-8f), l0 = logf(1.f - sp + 1e-8f);
                        acc[2] += (double)(-(lg - brv) * c2.cA) * (sf * l1 + (1.f - sf) * l0) + (double)c2.cE * (sp * l1 + (1.f - sp) * l0);
                        gs = binary_grad(sp, sf, (lg - brv) * c2.cA, c2.cE);
                    }
                    acc[3] += (double)(brv - lg) * (double)(brv - lg) * bas_scale[0];
                    acc[4] += (double)(bsv - lg) * (double)(bsv - lg) * bas_scale[0];
                }
                if (rec_on) acc[1] += (double)(-(lg - brv) * c1.cA) * lp1 + (double)c1.cE * hh1;
                W.g_stop_prob[row] = gs;
                W.g_bs[row] = m_in ? 2.f * (bsv - lg) * bas_scale[0] : 0.f;     // model.py:971-988
                W.g_br[row] = m_in ? 2.f * (brv - lg) * bas_scale[0] : 0.f;
            }
        } else if (lane == 0) {
            W.g_stop_prob[row] = 0.f; W.g_bs[row] = 0.f; W.g_br[row] = 0.f;
        }
    }
    (void)st;
    loss_epilogue(d, W, acc);
}

}  // namespace mmg
