// K_exchange_bwd — backward of the conversation (the reference's 4 autograd traversals, model.py:1309-1328) for the
// two agents, given dLoss/d(probabilities, scores).  Two CTA roles in one launch:
//   receiver CTAs: BPTT through the T GRU steps (h_z is never detached between steps, model.py:340) with three
//       gradient injections per step — receiver message head (w, w_h, w_d), STOP head (s) and, at each example's
//       prediction step, the class-score head (y1, y2).  softmax(y) into the message head is detached
//       (model.py:441) so non-prediction steps contribute nothing to y1/y2.
//   sender CTAs: the Sender has no recurrence (every inter-agent tensor is cut with .data, model.py:807-811), so its
//       per-step MLP backward runs concurrently on other SMs and only accumulates d h_x over steps.
// With -desc_attn the receiver CTAs also run the attention backward of every step (scores, segment softmax, the two
// places the attention weights enter) and keep per-CTA sums of the per-word gradients, see the blocks marked desc_attn.
// The kernels emit the pre-activation gradients ("deltas"); all weight gradients are then batched GEMMs over the
// T*B rows (K_wgrad).  Transposed, packed weight copies come from the backward image (TMA bulk copy into smem).
#pragma once
#include "mmg_exchange_fwd.cuh"

namespace mmg {

MMG_HOST_DEVICE int bwd_rec_state_floats(const Dims& d, int BT) {
    const int MP = d.M4 * 4, HrP = d.Hr4 * 4, G3P = align4(d.G3), H2P = align4(2 * d.Hr + d.A), DP = align4(d.D);
    int n = BT * (HrP + MP + H2P + G3P + DP) + 4 * BT;
    n += 2 * BT * kLoopThreads + 8;
    if (d.A) n += BT * (d.D * HrP + 2 * align4(d.NW) + align4(d.A) + 2 * (kLoopThreads / 32) * align4(d.A)) + align4(d.A) + HrP;
    (void)HrP;
    return n;
}
MMG_HOST_DEVICE int bwd_sen_state_floats(const Dims& d, int BT) {
    const int MP = d.M4 * 4, HiP = align4(d.Hi);
    int pmax = round_up(d.Ha, 32);
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
        int pmax = round_up(d.Ha, 32);
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
        const SplitPlan sp = make_split(d.Ha, d.M4);
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
            split_matvec<BT, false>(WbT, d.Ha, d.M4, dlz, MP, part, sp);   // d a = W_b^T . d logits
            MMG_SYNCTHREADS();
            for (int idx = tid; idx < BT * d.Hi; idx += kLoopThreads) {
                const int bt = idx / d.Hi, n = idx % d.Hi, b = b0 + bt;
                if (b < d.B && d.mix_mou) {
                    // four blocks a = tanh([h_x ; h_w ; h_x - h_w ; h_x * h_w]): d h_x = d0 + d2 + d3 h_w, d h_w = d1 - d2 + d3 h_x
                    const size_t i = ((size_t)t * d.B + b) * d.Hi + n;
                    const float* as = W.a_s + ((size_t)t * d.B + b) * d.Ha;
                    float dp[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { const float a = as[k * d.Hi + n]; dp[k] = gather_part<BT>(part, sp, bt, k * d.Hi + n) * (1.f - a * a); }
                    const float hxv = W.h_x[(size_t)b * d.Hi + n], hwv = W.hw_s[i];
                    W.d_as[i] = dp[1] - dp[2] + dp[3] * hxv;
                    dhx[bt * HiP + n] += dp[0] + dp[2] + dp[3] * hwv;
                } else if (b < d.B) {
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
    float* qs = sm + o;    o += BT * DP;       // softmax(y) of this step (desc_attn)
    float* dls = sm + o;   o += BT;
    float* yflag = sm + o; o += BT;
    o = align4(o);
    float* partA = sm + o; o += BT * kLoopThreads;
    float* partB = sm + o; o += BT * kLoopThreads;
    const int NWP = align4(d.NW), AP = align4(d.A), DH = d.D * HrP, lane = tid & 31, warp = tid >> 5;
    float* y1e = sm + o;   o += d.A ? BT * DH : 0;                          // attended description half of y1, then d y1 (prediction step)
    float* dav = sm + o;   o += d.A ? BT * NWP : 0;                         // d attention weights, then d scores
    float* att = sm + o;   o += d.A ? BT * NWP : 0;                         // attention weights of this step
    float* dhs = sm + o;   o += d.A ? BT * AP : 0;                          // d_h(h) of this step
    float* vas = sm + o;   o += d.A ? AP : 0;                               // d_attn.weight, zero padded
    float* b1s = sm + o;   o += d.A ? HrP : 0;                              // y1.bias
    float* ddacc = sm + o; o += (d.A && aa.acc_smem) ? d.NW * AP : 0;       // running d (d_d(word)) sums of this CTA
    const size_t slab_floats = (size_t)d.NW * (AP + HrP);                   // per-CTA slab: [NW][AP] d (d_d(word)) ; [NW][HrP] Z
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
    if (d.A) {
        for (int idx = tid; idx < AP; idx += kLoopThreads) vas[idx] = idx < d.A ? ldg(aa.va + idx) : 0.f;
        for (int idx = tid; idx < HrP; idx += kLoopThreads) b1s[idx] = idx < d.Hr ? ldg(aa.b1 + idx) : 0.f;
        for (int idx = tid; idx < BT * AP; idx += kLoopThreads) dhs[idx] = 0.f;
        for (int idx = tid; idx < BT * DH; idx += kLoopThreads) y1e[idx] = 0.f;
        if (aa.acc_smem) for (int idx = tid; idx < d.NW * AP; idx += kLoopThreads) ddacc[idx] = 0.f;
        float4* zs = reinterpret_cast<float4*>(W.ddd_part + (size_t)blockIdx.x * slab_floats + (size_t)d.NW * AP);
        for (int idx = tid; idx < d.NW * (HrP >> 2); idx += kLoopThreads) zs[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
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
        if (d.A) {      // this step's attention weights, d_h(h) and softmax(y) rows -> shared memory
            for (int idx = tid; idx < BT * d.NW; idx += kLoopThreads) {
                const int bt = idx / d.NW, n = idx % d.NW, b = b0 + bt;
                att[bt * NWP + n] = b < d.B ? W.attn[((size_t)t * d.B + b) * d.NW + n] : 0.f;
            }
            for (int idx = tid; idx < BT * d.A; idx += kLoopThreads) {
                const int bt = idx / d.A, a = idx % d.A, b = b0 + bt;
                dhs[bt * AP + a] = b < d.B ? W.dh_s[((size_t)t * d.B + b) * d.A + a] : 0.f;
            }
            for (int idx = tid; idx < BT * d.D; idx += kLoopThreads) {
                const int bt = idx / d.D, dd = idx % d.D, b = b0 + bt;
                qs[bt * DP + dd] = b < d.B ? W.q[((size_t)t * d.B + b) * d.D + dd] : 0.f;
            }
        }
        MMG_SYNCTHREADS();
        if (d.A) {
            // ---- -desc_attn, prediction step only: rebuild the attended y1 half from the saved attention weights, and the
            //      (the weight gradient of y1's description columns is formed per WORD below, after R2)
            const int K4 = HrP >> 2;
            for (int idx = tid; idx < BT * d.D * K4; idx += kLoopThreads) {
                const int bt = idx / (d.D * K4), r = idx % (d.D * K4), dd = r / K4, k4 = r % K4;
                if (yflag[bt] == 0.f) continue;
                float4 s4 = *reinterpret_cast<const float4*>(b1s + 4 * k4);
                const int s1 = W.seg[dd + 1];
                const float4* tab = reinterpret_cast<const float4*>(W.wtab_y1) + k4;
#pragma unroll 8
                for (int n = W.seg[dd]; n < s1; ++n) {
                    const float4 w = ldg4(tab + (size_t)n * K4);
                    const float a = att[bt * NWP + n];
                    s4.x = fmaf(a, w.x, s4.x); s4.y = fmaf(a, w.y, s4.y); s4.z = fmaf(a, w.z, s4.z); s4.w = fmaf(a, w.w, s4.w);
                }
                *reinterpret_cast<float4*>(y1e + bt * DH + dd * HrP + 4 * k4) = s4;
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
                    const float pre = yh + (d.A ? y1e[bt * DH + dd * HrP + k] : y1d[dd * d.Hr + k]);
                    const float v = pre > 0.f ? g * wk : 0.f;
                    W.dy1[((size_t)b * d.D + dd) * d.Hr + k] = v;
                    if (d.A) y1e[bt * DH + dd * HrP + k] = v;       // the attention backward reads d y1 from shared memory
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
        if (d.A) {
            // d y1.weight[:, desc columns] = sum_{b,class} d y1[b,class]^T (x) weighted_desc[b,class]  (model.py:383-410)
            //                              = sum_words ( sum_b a[b,n] d y1[b,class(n)] )^T (x) desc_set[n]:
            // accumulate the per-word factor Z[n] in this CTA's slab (fixed element -> thread mapping, no atomics);
            // K_attn_reduce sums the slabs and K_wgrad multiplies by desc_set (K = NW instead of B*D attended rows)
            const int K4 = HrP >> 2;
            float4* zs = reinterpret_cast<float4*>(W.ddd_part + (size_t)blockIdx.x * slab_floats + (size_t)d.NW * AP);
            for (int bt = 0; bt < BT; ++bt) {
                if (yflag[bt] == 0.f) continue;
                for (int idx = tid; idx < d.NW * K4; idx += kLoopThreads) {
                    const int n = idx / K4, k4 = idx % K4;
                    const float a = att[bt * NWP + n];
                    const float4 g = *reinterpret_cast<const float4*>(y1e + bt * DH + W.wcls[n] * HrP + 4 * k4);
                    float4 z = zs[idx];
                    z.x = fmaf(a, g.x, z.x); z.y = fmaf(a, g.y, z.y); z.z = fmaf(a, g.z, z.z); z.w = fmaf(a, g.w, z.w);
                    zs[idx] = z;
                }
            }
        }
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
            const int q4 = tid & 3, g4 = tid >> 2, K4 = HrP >> 2, A4 = AP >> 2;
            for (int base = 0; base < BT * d.NW; base += kLoopThreads / 4) {
                const int o2 = base + g4;
                const bool ok = o2 < BT * d.NW;
                const int bt = ok ? o2 / d.NW : 0, n = ok ? o2 % d.NW : 0, b = b0 + bt;
                float s = 0.f;
                if (ok && b < d.B) {
                    const int cls = W.wcls[n];
                    const float qd = qs[bt * DP + cls];
                    const bool yf = yflag[bt] != 0.f;
                    const float4* twd = reinterpret_cast<const float4*>(W.wtab_wd + (size_t)n * HrP);
                    const float4* ty1 = reinterpret_cast<const float4*>(W.wtab_y1 + (size_t)n * HrP);
                    const float4* dhw = reinterpret_cast<const float4*>(dvec + bt * H2P);
                    const float4* dyr = reinterpret_cast<const float4*>(y1e + bt * DH + cls * HrP);
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
                    for (int i = q4; i < K4; i += 4) {
                        const float4 w = ldg4(twd + i), g = dhw[i];
                        s1 = fmaf(g.x, w.x, s1); s1 = fmaf(g.y, w.y, s1); s1 = fmaf(g.z, w.z, s1); s1 = fmaf(g.w, w.w, s1);
                        if (yf) {
                            const float4 u = ldg4(ty1 + i), e = dyr[i];
                            s2 = fmaf(e.x, u.x, s2); s2 = fmaf(e.y, u.y, s2); s2 = fmaf(e.z, u.z, s2); s2 = fmaf(e.w, u.w, s2);
                        }
                    }
                    s = fmaf(qd, s1, s2);
                }
                s = group_sum<4>(s);
                if (ok && q4 == 0) dav[bt * NWP + n] = s;
            }
            MMG_SYNCTHREADS();
            for (int base = 0; base < BT * d.D; base += kLoopThreads / 8) {        // through the segment softmax (model.py:378)
                const int o2 = base + (tid >> 3), l8 = tid & 7;
                const bool ok = o2 < BT * d.D;
                const int bt = ok ? o2 / d.D : 0, dd = ok ? o2 % d.D : 0;
                const int s0 = ok ? W.seg[dd] : 0, s1 = ok ? W.seg[dd + 1] : 0;
                float s = 0.f;
                for (int n = s0 + l8; n < s1; n += 8) s = fmaf(att[bt * NWP + n], dav[bt * NWP + n], s);
                s = group_sum<8>(s);
                for (int n = s0 + l8; n < s1; n += 8) dav[bt * NWP + n] = att[bt * NWP + n] * (dav[bt * NWP + n] - s);
            }
            MMG_SYNCTHREADS();
            // through score = d_attn(tanh(d_d(word) + d_h(h))) (model.py:366): thread = (word group, float4 unit group); every
            // thread owns fixed (word, unit) elements for the whole kernel, so its running sum of d (d_d(word)) needs no atomics
            float* dslab = aa.acc_smem ? ddacc : W.ddd_part + (size_t)blockIdx.x * slab_floats;
            for (int ib = 0; ib < A4; ib += 4) {        // uniform trip count: the partial sums meet through warp shuffles
                const int i4 = ib + q4 < A4 ? ib + q4 : A4 - 1;
                const bool act = ib + q4 < A4;
                const float4 va = *reinterpret_cast<const float4*>(vas + 4 * i4);
                for (int bt = 0; bt < BT; ++bt) {
                    const int b = b0 + bt;
                    float4 accd = make_float4(0.f, 0.f, 0.f, 0.f), accv = accd;
                    if (b < d.B && act) {
                        const float4 dha = *reinterpret_cast<const float4*>(dhs + bt * AP + 4 * i4);
                        const bool init = first && bt == 0 && !aa.acc_smem;
#pragma unroll 4
                        for (int n = g4; n < d.NW; n += kLoopThreads / 4) {
                            const float4 w = ldg4(reinterpret_cast<const float4*>(W.wtab_dd + (size_t)n * AP) + i4);
                            float4* slot = reinterpret_cast<float4*>(dslab + (size_t)n * AP) + i4;
                            float4 acc = init ? make_float4(0.f, 0.f, 0.f, 0.f) : *slot;
                            const float de = dav[bt * NWP + n];
                            const float tx = attn_tanh(w.x, dha.x), ty = attn_tanh(w.y, dha.y), tz = attn_tanh(w.z, dha.z), tw = attn_tanh(w.w, dha.w);
                            const float ux = de * va.x * (1.f - tx * tx), uy = de * va.y * (1.f - ty * ty),
                                        uz = de * va.z * (1.f - tz * tz), uw = de * va.w * (1.f - tw * tw);
                            accd.x += ux; accd.y += uy; accd.z += uz; accd.w += uw;
                            accv.x = fmaf(de, tx, accv.x); accv.y = fmaf(de, ty, accv.y); accv.z = fmaf(de, tz, accv.z); accv.w = fmaf(de, tw, accv.w);
                            acc.x += ux; acc.y += uy; acc.z += uz; acc.w += uw;
                            *slot = acc;
                        }
                    } else if (first && bt == 0 && act && !aa.acc_smem) {
                        for (int n = g4; n < d.NW; n += kLoopThreads / 4)
                            *(reinterpret_cast<float4*>(dslab + (size_t)n * AP) + i4) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    // sum over the 8 word groups of this warp (lanes with equal q4), then one partial per warp
#pragma unroll
                    for (int m = 4; m < 32; m <<= 1) {
                        accd.x += shfl_xor_f(accd.x, m); accd.y += shfl_xor_f(accd.y, m); accd.z += shfl_xor_f(accd.z, m); accd.w += shfl_xor_f(accd.w, m);
                        accv.x += shfl_xor_f(accv.x, m); accv.y += shfl_xor_f(accv.y, m); accv.z += shfl_xor_f(accv.z, m); accv.w += shfl_xor_f(accv.w, m);
                    }
                    if (lane < 4 && act) {
                        *reinterpret_cast<float4*>(pddh + (bt * (kLoopThreads / 32) + warp) * AP + 4 * i4) = accd;
                        *reinterpret_cast<float4*>(pdva + (bt * (kLoopThreads / 32) + warp) * AP + 4 * i4) = accv;
                    }
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
    if (d.A && aa.acc_smem) {       // publish this CTA's d (d_d(word)) sums (summed over CTAs by K_attn_reduce)
        MMG_SYNCTHREADS();
        float4* out = reinterpret_cast<float4*>(W.ddd_part + (size_t)blockIdx.x * slab_floats);
        for (int idx = tid; idx < d.NW * (AP >> 2); idx += kLoopThreads) out[idx] = reinterpret_cast<const float4*>(ddacc)[idx];
    }
}

// Sum of the per-CTA d (d_d(word)) slabs into slab 0, slab order (deterministic): K = NW rows for the d_d.weight GEMM
// instead of n_ctas * NW.
MMG_GLOBAL void __launch_bounds__(256)
k_attn_reduce(float* slabs, int n_slabs, int slab_f4) {
    // CTA = 32 float4 columns x 8 slab groups; the groups meet in shared memory in a fixed order
    MMG_SHARED float4 part[8][32];
    pdl_wait(); pdl_launch_dependents();
    float4* s4 = reinterpret_cast<float4*>(slabs);
    const int col = threadIdx.x & 31, grp = threadIdx.x >> 5;
    for (int base = blockIdx.x * 32; base < slab_f4; base += gridDim.x * 32) {
        const int i = base + col;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < slab_f4) {
#pragma unroll 8
            for (int c = grp; c < n_slabs; c += 8) {
                const float4 v = s4[(size_t)c * slab_f4 + i];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        part[grp][col] = acc;
        MMG_SYNCTHREADS();
        if (grp == 0 && i < slab_f4) {
            for (int g = 1; g < 8; ++g) { const float4 v = part[g][col]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
            s4[i] = acc;
        }
        MMG_SYNCTHREADS();
    }
}

}  // namespace mmg
