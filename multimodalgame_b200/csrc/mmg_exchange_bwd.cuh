// K_exchange_bwd — backward of the conversation (the reference's 4 autograd traversals, model.py:1309-1328) for the
// two agents, given dLoss/d(probabilities, scores).  Two CTA roles in one launch:
//   receiver CTAs: BPTT through the T GRU steps (h_z is never detached between steps, model.py:340) with three
//       gradient injections per step — receiver message head (w, w_h, w_d), STOP head (s) and, at each example's
//       prediction step, the class-score head (y1, y2).  softmax(y) into the message head is detached
//       (model.py:441) so non-prediction steps contribute nothing to y1/y2.
//   sender CTAs: the Sender has no recurrence (every inter-agent tensor is cut with .data, model.py:807-811), so its
//       per-step MLP backward runs concurrently on other SMs and only accumulates d h_x over steps.
// The kernels emit the pre-activation gradients ("deltas"); all weight gradients are then batched GEMMs over the
// T*B rows (K_wgrad).  Transposed, packed weight copies come from the backward image (TMA bulk copy into smem).
#pragma once
#include "mmg_exchange_fwd.cuh"

namespace mmg {

MMG_HOST_DEVICE int bwd_rec_state_floats(const Dims& d, int BT) {
    const int MP = d.M4 * 4, HrP = d.Hr4 * 4, G3P = align4(d.G3), H2P = align4(2 * d.Hr + d.A), DP = align4(d.D);
    int n = BT * (HrP + MP + H2P + G3P + DP) + 4 * BT;
    n += 2 * BT * kLoopThreads + 8;
    if (d.A) n += BT * (align4(d.D * d.Hr) + align4(d.NW) + 2 * (kLoopThreads / 32) * align4(d.A));
    (void)HrP;
    return n;
}
MMG_HOST_DEVICE int bwd_sen_state_floats(const Dims& d, int BT) {
    const int MP = d.M4 * 4, HiP = align4(d.Hi);
    int pmax = round_up(d.Hi, 32);
    if (pmax < kLoopThreads) pmax = kLoopThreads;
    return BT * (MP + HiP) + BT * pmax + 8;
}

template <int BT>
MMG_GLOBAL void __launch_bounds__(kLoopThreads, 1)
k_exchange_bwd(Dims d, WsPtrs W, int n_rec_ctas, AttnArgs aa) {
    MMG_DYN_SMEM(smem_raw);
    float* sm = reinterpret_cast<float*>(smem_raw);
    const BwdImage im = make_bwd_image(d);
    const int tid = threadIdx.x;
    const int MP = d.M4 * 4, HrP = d.Hr4 * 4, HiP = align4(d.Hi), G3P = align4(d.G3), H2P = align4(2 * d.Hr + d.A), DP = align4(d.D);
    const bool binary = d.use_binary != 0;

    if ((int)blockIdx.x >= n_rec_ctas) {
        // =================================== sender role ======================================================
        const int b0 = ((int)blockIdx.x - n_rec_ctas) * BT;
        int o = im.sender_end;
        float* dlz = sm + o; o += BT * MP;
        float* dhx = sm + o; o += BT * HiP;
        int pmax = round_up(d.Hi, 32);
        if (pmax < kLoopThreads) pmax = kLoopThreads;
        float* part = sm + o; o += BT * pmax;
        o = align4(o); o += (o & 1);
        uint64_t* bar = reinterpret_cast<uint64_t*>(sm + o);
        if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
        MMG_SYNCTHREADS();
        pdl_wait(); pdl_launch_dependents();
        if (tid == 0) tma_stage(sm, W.bwd_image, (uint32_t)im.sender_end * 4u, bar);
        for (int idx = tid; idx < BT * MP; idx += kLoopThreads) dlz[idx] = 0.f;
        for (int idx = tid; idx < BT * HiP; idx += kLoopThreads) dhx[idx] = 0.f;
#ifdef MMG_CPU_EMU
        MMG_SYNCTHREADS();
#endif
        mbar_wait(bar, 0);
        MMG_SYNCTHREADS();
        const float* WbT = sm + im.wbT;
        const SplitPlan sp = make_split(d.Hi, d.M4);
        for (int t = 0; t < d.T; ++t) {
            for (int idx = tid; idx < BT * d.M; idx += kLoopThreads) {
                const int bt = idx / d.M, j = idx % d.M, b = b0 + bt;
                float dl = 0.f;
                if (b < d.B) {
                    const size_t i = ((size_t)t * d.B + b) * d.M + j;
                    const float p = W.sen_probs[i];
                    dl = W.g_sen_probs[i] * p * (1.f - p);                  // through the sigmoid (model.py:223)
                    W.d_lz[i] = dl;
                }
                dlz[bt * MP + j] = dl;
            }
            MMG_SYNCTHREADS();
            split_matvec<BT, false>(WbT, d.Hi, d.M4, dlz, MP, part, sp);   // d a = W_b^T . d logits
            MMG_SYNCTHREADS();
            for (int idx = tid; idx < BT * d.Hi; idx += kLoopThreads) {
                const int bt = idx / d.Hi, n = idx % d.Hi, b = b0 + bt;
                if (b < d.B) {
                    const size_t i = ((size_t)t * d.B + b) * d.Hi + n;
                    const float a = W.a_s[i];
                    const float dpre = gather_part<BT>(part, sp, bt, n) * (1.f - a * a);  // through tanh (model.py:216)
                    // d_as = gradient w.r.t. the code term h_w; sum: d pre, prod: d pre * h_x, ignore_code: none
                    float dhw = dpre, dx = dpre;
                    if (d.mix_prod && !d.ignore_code) { dhw = dpre * W.h_x[(size_t)b * d.Hi + n]; dx = dpre * W.hw_s[i]; }
                    if (d.ignore_code) dhw = 0.f;
                    W.d_as[i] = dhw;
                    dhx[bt * HiP + n] += dx;                                // h_x is shared by all steps (model.py:195)
                }
            }
            // next iteration's first stage only writes dlz (already consumed) -> the barrier after it orders `part`
        }
        for (int idx = tid; idx < BT * d.Hi; idx += kLoopThreads) {
            const int bt = idx / d.Hi, n = idx % d.Hi, b = b0 + bt;
            if (b < d.B) W.dhx[(size_t)b * d.Hi + n] = dhx[bt * HiP + n];
        }
        return;
    }

    // ===================================== receiver role ======================================================
    const int b0 = (int)blockIdx.x * BT;
    const int img0 = im.sender_end;
    float* img = sm - img0;
    int o = im.total - img0;
    float* dh = sm + o;    o += BT * HrP;      // carried d h (direct GRU path)
    float* dlw = sm + o;   o += BT * MP;
    float* dvec = sm + o;  o += BT * H2P;      // [d_hw (Hr) ; G_h (Hr)]
    float* dghs = sm + o;  o += BT * G3P;
    o += BT * DP;
    float* dls = sm + o;   o += BT;
    float* yflag = sm + o; o += BT;
    o = align4(o);
    float* partA = sm + o; o += BT * kLoopThreads;
    float* partB = sm + o; o += BT * kLoopThreads;
    const int NWP = align4(d.NW), AP = align4(d.A), DH = align4(d.D * d.Hr), lane = tid & 31, warp = tid >> 5;
    float* y1e = sm + o;   o += d.A ? BT * DH : 0;                          // attended description half of y1 (prediction step)
    float* dav = sm + o;   o += d.A ? BT * NWP : 0;                         // d attention weights, then d scores
    float* pddh = sm + o;  o += d.A ? BT * (kLoopThreads / 32) * AP : 0;    // per-warp partials of d (d_h(h))
    float* pdva = sm + o;  o += d.A ? BT * (kLoopThreads / 32) * AP : 0;    // per-warp partials of d d_attn.weight
    o += (o & 1);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + o);

    const float* WwT = img + im.wwT;
    const float* HeadT = img + im.headT;
    const float* WhhT = img + im.whhT;
    const float* ws = img + im.ws;
    const float* w2 = img + im.w2;
    const float* y1d = img + im.y1d;
    const SplitPlan sp_w = make_split(d.Hr, d.M4), sp_head = make_split(d.Hr, cdiv(2 * d.Hr + d.A, 4)),
                    sp_hh = make_split(d.Hr, cdiv(d.G3, 4));

    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    MMG_SYNCTHREADS();
    pdl_wait(); pdl_launch_dependents();
    if (tid == 0) tma_stage(sm, W.bwd_image + img0, (uint32_t)(im.total - img0) * 4u, bar);
    for (int idx = tid; idx < BT * HrP; idx += kLoopThreads) dh[idx] = 0.f;
    for (int idx = tid; idx < BT * MP; idx += kLoopThreads) dlw[idx] = 0.f;
    for (int idx = tid; idx < BT * H2P; idx += kLoopThreads) dvec[idx] = 0.f;
    for (int idx = tid; idx < BT * G3P; idx += kLoopThreads) dghs[idx] = 0.f;
    for (int idx = tid; idx < BT * kLoopThreads; idx += kLoopThreads) { partA[idx] = 0.f; partB[idx] = 0.f; }
#ifdef MMG_CPU_EMU
    MMG_SYNCTHREADS();
#endif
    mbar_wait(bar, 0);
    MMG_SYNCTHREADS();

    bool first = true;
    for (int t = d.T - 1; t >= 0; --t) {
        // ---- R1: logits of the message / STOP heads; fetch the class-score gradient at the prediction step -------
        for (int idx = tid; idx < BT * d.M; idx += kLoopThreads) {
            const int bt = idx / d.M, j = idx % d.M, b = b0 + bt;
            float dl = 0.f;
            if (b < d.B) {
                const size_t i = ((size_t)t * d.B + b) * d.M + j;
                if (binary) {
                    const float p = W.rec_probs[i];
                    dl = W.g_rec_probs[i] * p * (1.f - p);
                }
                W.d_lw[i] = dl;
            }
            dlw[bt * MP + j] = dl;
        }
        if (tid < BT) {
            const int b = b0 + tid;
            float v = 0.f, fl = 0.f;
            if (b < d.B) {
                const size_t row = (size_t)t * d.B + b;
                const float sp = W.stop_prob[row];
                v = W.g_stop_prob[row] * sp * (1.f - sp);
                W.d_ls[row] = v;
                fl = (W.ystep[b] == t) ? 1.f : 0.f;
            }
            dls[tid] = v; yflag[tid] = fl;
        }
        MMG_SYNCTHREADS();
        if (d.A) {
            // ---- -desc_attn, prediction step only: rebuild the attended y1 half from the saved attention weights, and the
            //      attended description itself (the `weighted_desc` rows that multiply d y1, model.py:383-410)
            for (int idx = tid; idx < BT * d.D * d.Hr; idx += kLoopThreads) {
                const int bt = idx / (d.D * d.Hr), r = idx % (d.D * d.Hr), dd = r / d.Hr, k = r % d.Hr, b = b0 + bt;
                if (yflag[bt] == 0.f) continue;
                const float* arow = W.attn + ((size_t)t * d.B + b) * d.NW;
                float s = ldg(aa.b1 + k);
                const int s1 = W.seg[dd + 1];
                for (int n = W.seg[dd]; n < s1; ++n) s = fmaf(arow[n], ldg(W.wtab_y1 + (size_t)n * d.Hr + k), s);
                y1e[bt * DH + r] = s;
            }
            for (int idx = tid; idx < BT * d.D * d.WV; idx += kLoopThreads) {
                const int bt = idx / (d.D * d.WV), r = idx % (d.D * d.WV), dd = r / d.WV, v = r % d.WV, b = b0 + bt;
                if (yflag[bt] == 0.f) continue;
                const float* arow = W.attn + ((size_t)t * d.B + b) * d.NW;
                float s = 0.f;
                const int s1 = W.seg[dd + 1];
                for (int n = W.seg[dd]; n < s1; ++n) s = fmaf(arow[n], ldg(aa.desc_set + (size_t)n * d.WV + v), s);
                W.wdsel[(size_t)b * d.D * d.WV + r] = s;
            }
            MMG_SYNCTHREADS();
        }
        // ---- R2: d h_w = W_w^T . d logits_w ;  class-score head at the prediction step ---------------------------
        split_matvec<BT, false>(WwT, d.Hr, d.M4, dlw, MP, partA, sp_w);
        for (int idx = tid; idx < BT * d.Hr; idx += kLoopThreads) {
            const int bt = idx / d.Hr, k = idx % d.Hr, b = b0 + bt;
            float G = 0.f;
            if (yflag[bt] != 0.f) {
                // y[d] = y2.bias + sum_k w2[k] relu(y1h[k] + y1d[d][k])   (model.py:432-433)
                const size_t row = (size_t)t * d.B + b;
                const float yh = W.y1h[row * d.Hr + k], wk = w2[k];
                float dw2 = 0.f;
                for (int dd = 0; dd < d.D; ++dd) {
                    const float g = W.g_outp[(size_t)b * d.D + dd];
                    const float pre = yh + (d.A ? y1e[bt * DH + dd * d.Hr + k] : y1d[dd * d.Hr + k]);
                    const float v = pre > 0.f ? g * wk : 0.f;
                    W.dy1[((size_t)b * d.D + dd) * d.Hr + k] = v;
                    G += v;
                    dw2 = fmaf(g, fmaxf(pre, 0.f), dw2);
                }
                W.g_h[(size_t)b * d.Hr + k] = G;
                W.dw2p[(size_t)b * d.Hr + k] = dw2;
                W.hsel[(size_t)b * d.Hr + k] = W.h_z[((size_t)(t + 1) * d.B + b) * d.Hr + k];
            }
            dvec[bt * H2P + d.Hr + k] = G;
        }
        MMG_SYNCTHREADS();
        // ---- R3: through tanh of the message hidden (model.py:452) ---------------------------------------------
        for (int idx = tid; idx < BT * d.Hr; idx += kLoopThreads) {
            const int bt = idx / d.Hr, k = idx % d.Hr, b = b0 + bt;
            float v = 0.f;
            if (b < d.B) {
                const size_t i = ((size_t)t * d.B + b) * d.Hr + k;
                const float hw = W.h_w[i];
                v = gather_part<BT>(partA, sp_w, bt, k) * (1.f - hw * hw);
                W.d_hw[i] = v;
            }
            dvec[bt * H2P + k] = v;
        }
        MMG_SYNCTHREADS();
        if (d.A) {
            // ---- -desc_attn backward.  a_n enters (1) the message hidden through q_class(n) a_n (desc_set . w_d^T)[n]
            //      at every step (weighted_desc is NOT detached, model.py:444-449) and (2) the class scores through
            //      a_n (desc_set . y1^T)[n] at the prediction step.
            for (int o2 = warp; o2 < BT * d.NW; o2 += kLoopThreads / 32) {
                const int bt = o2 / d.NW, n = o2 % d.NW, b = b0 + bt;
                float s = 0.f;
                if (b < d.B) {
                    const int cls = W.wcls[n];
                    const float qd = W.q[((size_t)t * d.B + b) * d.D + cls];
                    const float* dyr = W.dy1 + ((size_t)b * d.D + cls) * d.Hr;
                    const bool yf = yflag[bt] != 0.f;
                    for (int k = lane; k < d.Hr; k += 32) {
                        s = fmaf(qd * dvec[bt * H2P + k], ldg(W.wtab_wd + (size_t)n * d.Hr + k), s);
                        if (yf) s = fmaf(dyr[k], ldg(W.wtab_y1 + (size_t)n * d.Hr + k), s);
                    }
                }
                s = warp_sum(s);
                if (lane == 0) dav[bt * NWP + n] = s;
            }
            MMG_SYNCTHREADS();
            for (int o2 = warp; o2 < BT * d.D; o2 += kLoopThreads / 32) {          // through the segment softmax (model.py:378)
                const int bt = o2 / d.D, dd = o2 % d.D, b = b0 + bt;
                if (b >= d.B) continue;
                const float* arow = W.attn + ((size_t)t * d.B + b) * d.NW;
                const int s0 = W.seg[dd], s1 = W.seg[dd + 1];
                float s = 0.f;
                for (int n = s0 + lane; n < s1; n += 32) s = fmaf(arow[n], dav[bt * NWP + n], s);
                s = warp_sum(s);
                for (int n = s0 + lane; n < s1; n += 32) dav[bt * NWP + n] = arow[n] * (dav[bt * NWP + n] - s);
            }
            MMG_SYNCTHREADS();
            // through score = d_attn(tanh(d_d(word) + d_h(h))) (model.py:366): each thread owns fixed (word, unit) pairs for the
            // whole kernel, so its running sum of d (d_d(word)) needs no atomics
            float* dslab = W.ddd_part + (size_t)blockIdx.x * d.NW * d.A;
            for (int a = lane; a < d.A; a += 32) {
                const float va = ldg(aa.va + a);
                for (int bt = 0; bt < BT; ++bt) {
                    const int b = b0 + bt;
                    float accd = 0.f, accv = 0.f;
                    if (b < d.B) {
                        const float dha = W.dh_s[((size_t)t * d.B + b) * d.A + a];
                        for (int n = warp; n < d.NW; n += kLoopThreads / 32) {
                            const float th = tanhf(ldg(W.wtab_dd + (size_t)n * d.A + a) + dha);
                            const float de = dav[bt * NWP + n];
                            const float du = de * va * (1.f - th * th);
                            accd += du;
                            accv = fmaf(de, th, accv);
                            float* slot = dslab + (size_t)n * d.A + a;
                            *slot = (first && bt == 0) ? du : *slot + du;
                        }
                    } else if (first && bt == 0) {
                        for (int n = warp; n < d.NW; n += kLoopThreads / 32) dslab[(size_t)n * d.A + a] = 0.f;
                    }
                    pddh[(bt * (kLoopThreads / 32) + warp) * AP + a] = accd;
                    pdva[(bt * (kLoopThreads / 32) + warp) * AP + a] = accv;
                }
            }
            MMG_SYNCTHREADS();
            for (int idx = tid; idx < BT * d.A; idx += kLoopThreads) {
                const int bt = idx / d.A, a = idx % d.A, b = b0 + bt;
                float sd = 0.f, sv = 0.f;
                for (int w = 0; w < kLoopThreads / 32; ++w) {
                    sd += pddh[(bt * (kLoopThreads / 32) + w) * AP + a];
                    sv += pdva[(bt * (kLoopThreads / 32) + w) * AP + a];
                }
                dvec[bt * H2P + 2 * d.Hr + a] = sd;
                if (b < d.B) {
                    W.ddh[((size_t)t * d.B + b) * d.A + a] = sd;
                    W.dva[((size_t)t * d.B + b) * d.A + a] = sv;
                }
            }
            for (int bt = warp; bt < BT; bt += kLoopThreads / 32) {               // d d_attn.bias (zero up to rounding)
                const int b = b0 + bt;
                float s = 0.f;
                if (b < d.B) for (int n = lane; n < d.NW; n += 32) s += dav[bt * NWP + n];
                s = warp_sum(s);
                if (lane == 0 && b < d.B) W.dba[(size_t)t * d.B + b] = s;
            }
            MMG_SYNCTHREADS();
        }
        // ---- R4: d h' += W_h^T . d_hw + W_1h^T . G_h (+ d_h^T . d (d_h(h))) ---------------------------------------
        split_matvec<BT, false>(HeadT, d.Hr, cdiv(2 * d.Hr + d.A, 4), dvec, H2P, partA, sp_head);
        MMG_SYNCTHREADS();
        // ---- R5: total d h', GRU gate gradients -----------------------------------------------------------------
        for (int idx = tid; idx < BT * d.Hr; idx += kLoopThreads) {
            const int bt = idx / d.Hr, k = idx % d.Hr, b = b0 + bt;
            float dr_pre = 0.f, du_pre = 0.f, dn_pre = 0.f, dghn = 0.f, direct = 0.f;
            if (b < d.B) {
                const size_t row = (size_t)t * d.B + b;
                float dht = dh[bt * HrP + k] + (first ? 0.f : gather_part<BT>(partB, sp_hh, bt, k))
                            + gather_part<BT>(partA, sp_head, bt, k) + ws[k] * dls[bt];
                const float* g = W.gates + row * 4 * d.Hr;
                const float r = g[k], u = g[d.Hr + k], nn = g[2 * d.Hr + k], ghn = g[3 * d.Hr + k];
                const float hp = W.h_z[row * d.Hr + k];                  // slot t = state entering step t
                // h' = n + u (h - n)
                const float du = dht * (hp - nn);
                const float dn = dht * (1.f - u);
                direct = dht * u;
                dn_pre = dn * (1.f - nn * nn);
                dghn = dn_pre * r;
                dr_pre = dn_pre * ghn * r * (1.f - r);
                du_pre = du * u * (1.f - u);
                float* gi = W.dgi + row * d.G3;
                float* gh = W.dgh + row * d.G3;
                gi[k] = dr_pre; gi[d.Hr + k] = du_pre; gi[2 * d.Hr + k] = dn_pre;
                gh[k] = dr_pre; gh[d.Hr + k] = du_pre; gh[2 * d.Hr + k] = dghn;
            }
            dghs[bt * G3P + k] = dr_pre; dghs[bt * G3P + d.Hr + k] = du_pre; dghs[bt * G3P + 2 * d.Hr + k] = dghn;
            dh[bt * HrP + k] = direct;
        }
        MMG_SYNCTHREADS();
        // ---- R6: d h_prev += W_hh^T . d gh (consumed by the next R5) ---------------------------------------------
        split_matvec<BT, false>(WhhT, d.Hr, cdiv(d.G3, 4), dghs, G3P, partB, sp_hh);
        first = false;
        // no barrier needed here: R1 of the next step touches dlw/dls/yflag only, and the barrier after it orders partB
    }
}

}  // namespace mmg
