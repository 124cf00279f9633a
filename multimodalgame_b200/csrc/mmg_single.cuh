// Single-turn module forwards: Sender.forward (model.py:144-238), Receiver.forward (model.py:303-477, incl. the
// -desc_attn branch 344-410) and Baseline.forward (model.py:496-516) as stand-alone kernels behind the C-ABI
// (mmg_sender_forward / mmg_receiver_forward / mmg_baseline_forward), so that callers that drive the agents turn by turn
// (the reference's sample printer, model.py:1464, or any custom loop) have the module-level entry points the reference has.
// They are NOT the hot path: the fused conversation (mmg_exchange_forward) is.  One CTA per example, any shape, libm-accurate
// transcendentals, no saved activations / autograd.
#pragma once
#include "mmg_exchange_fwd.cuh"

namespace mmg {

enum { kSingleThreads = 256 };

// out[r] = bias[r] + W[r][0:K] . v (+ W2[r][0:K2] . v2), rows strided over the CTA's warps, lanes over the reduction.
// All loops are warp-uniform (the shuffles of warp_sum need the whole warp).
MMG_DEVICE void block_matvec(float* out, const float* Wm, int ld, int rows, const float* v, int K, const float* bias,
                             const float* W2 = nullptr, int ld2 = 0, const float* v2 = nullptr, int K2 = 0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int r = warp; r < rows; r += nwarps) {
        float s = 0.f;
        for (int k = lane; k < K; k += 32) s = fmaf(ldg(Wm + (size_t)r * ld + k), v[k], s);
        if (W2 != nullptr) for (int k = lane; k < K2; k += 32) s = fmaf(ldg(W2 + (size_t)r * ld2 + k), v2[k], s);
        s = warp_sum(s);
        if (lane == 0) out[r] = s + (bias != nullptr ? ldg(bias + r) : 0.f);
    }
}

MMG_HOST_DEVICE int sender_step_smem_floats(const Dims& d) { return align4(d.F) + align4(d.Hi) + align4(d.Ha) + 2 * align4(d.M); }

// Sender.forward default path (model.py:195-238): h_x = image_layer(x); code = sigmoid(code_bias) at t == 0 else w;
// a = tanh(mix(h_x, code_layer(code))); logits = binary_layer(a); binary: p = sigmoid, message = 1[u < p] (train) or
// round(p) (eval) [+ flipout]; continuous: message = logits, no probabilities.
MMG_GLOBAL void __launch_bounds__(kSingleThreads)
k_sender_step(Dims d, ParamPtrs P, const float* x, const float* w, int t, int train, const double* u, const double* u_flip,
              unsigned long long seed, unsigned long long counter, float* msg, float* probs, float* h_x) {
    MMG_DYN_SMEM(smem_raw);
    float* sm = reinterpret_cast<float*>(smem_raw);
    float* xs = sm;
    float* hx = xs + align4(d.F);
    float* av = hx + align4(d.Hi);
    float* code = av + align4(d.Ha);
    float* lg = code + align4(d.M);
    const int tid = threadIdx.x, b = blockIdx.x;
    for (int f = tid; f < d.F; f += kSingleThreads) xs[f] = x[(size_t)b * d.F + f];
    for (int j = tid; j < d.M; j += kSingleThreads)
        code[j] = (t == 0) ? sigmoidf_(ldg(P.p[MMG_P_SEN_CODE_BIAS] + j))                              // model.py:199-207
                : ((d.mix_mou && d.ignore_code) ? sigmoidf_(ldg(P.p[MMG_P_SEN_CODE_BIAS_MOU] + j)) : w[(size_t)b * d.M + j]);
    MMG_SYNCTHREADS();
    block_matvec(hx, P.p[MMG_P_SEN_IMG_W], d.F, d.Hi, xs, d.F, P.p[MMG_P_SEN_IMG_B]);                 // model.py:195
    block_matvec(av, P.p[MMG_P_SEN_CODE_W], d.M, d.Hi, code, d.M, P.p[MMG_P_SEN_CODE_B]);
    MMG_SYNCTHREADS();
    for (int n = tid; n < d.Hi; n += kSingleThreads) {
        const float hxv = hx[n], hwv = av[n];
        h_x[(size_t)b * d.Hi + n] = hxv;
        if (d.mix_mou) {                                                                            // model.py:211-213, 219-221
            av[n] = tanhf(hxv); av[d.Hi + n] = tanhf(hwv); av[2 * d.Hi + n] = tanhf(hxv - hwv); av[3 * d.Hi + n] = tanhf(hxv * hwv);
            continue;
        }
        const float pre = d.ignore_code ? hxv : (d.mix_prod ? hxv * hwv : hxv + hwv);               // model.py:208-221
        av[n] = tanhf(pre);
    }
    MMG_SYNCTHREADS();
    block_matvec(lg, P.p[MMG_P_SEN_BIN_W], d.Ha, d.M, av, d.Ha, P.p[MMG_P_SEN_BIN_B]);
    MMG_SYNCTHREADS();
    for (int j = tid; j < d.M; j += kSingleThreads) {
        const size_t i = (size_t)b * d.M + j;
        float z = lg[j];
        if (d.use_binary) {
            const float p = sigmoidf_(lg[j]);
            probs[i] = p;
            z = train ? draw_bit(u, i, p, seed, counter, 0, (unsigned)(b + d.row0), (unsigned)j) : rintf(p);
            if (d.flip_sen >= 0.f && (train || d.flipout_dev))
                z = flip_bit(z, d.flip_sen, u_flip, i, seed, counter, 0, 0, (unsigned)(b + d.row0), (unsigned)j);
        }
        msg[i] = z;
    }
}

struct ReceiverStepIO {
    const float* z;            // (B,M)   message from the sender
    const float* desc;         // (D,WV)
    const float* desc_set;     // (NW,WV) -desc_attn
    const int* desc_set_lens;  // (D)
    float* h_z;                // (B,Hr)  in: previous state (ignored when first); out: new state
    float* s_prob_prod;        // (B)     eval: running product of the STOP probabilities (in/out)
    const double *u_stop, *u_rec, *u_flip;
    float *s, *s_prob, *w, *w_probs, *y, *h_w;    // (B) (B) (B,M) (B,M) (B,D) (B,Hr)
    int first, train;
    unsigned long long seed, counter;
};

MMG_HOST_DEVICE int receiver_step_smem_floats(const Dims& d) {
    int n = align4(d.M) + 2 * align4(d.Hr) + 2 * align4(d.G3) + 2 * align4(d.Hr) + 2 * align4(d.D) + align4(d.WV) + align4(d.M) + 8;
    if (d.A) n += align4(d.A) + 2 * align4(d.NW) + d.D * align4(d.WV) + align4(d.D + 1);
    return n;
}

MMG_GLOBAL void __launch_bounds__(kSingleThreads)
k_receiver_step(Dims d, ParamPtrs P, ReceiverStepIO io) {
    MMG_DYN_SMEM(smem_raw);
    float* sm = reinterpret_cast<float*>(smem_raw);
    int o = 0;
    float* zs = sm + o;   o += align4(d.M);
    float* hp = sm + o;   o += align4(d.Hr);
    float* hn = sm + o;   o += align4(d.Hr);
    float* gi = sm + o;   o += align4(d.G3);
    float* gh = sm + o;   o += align4(d.G3);
    float* y1h = sm + o;  o += align4(d.Hr);
    float* whv = sm + o;  o += align4(d.Hr);
    float* yv = sm + o;   o += align4(d.D);
    float* qv = sm + o;   o += align4(d.D);
    float* wdv = sm + o;  o += align4(d.WV);
    float* lg = sm + o;   o += align4(d.M);
    float* misc = sm + o; o += 8;
    const int WVP = align4(d.WV);
    float* dh = sm + o;   o += d.A ? align4(d.A) : 0;
    float* ev = sm + o;   o += d.A ? align4(d.NW) : 0;
    float* att = sm + o;  o += d.A ? align4(d.NW) : 0;
    float* ad = sm + o;   o += d.A ? d.D * WVP : 0;             // attended description of every class (model.py:383-397)
    int* segs = reinterpret_cast<int*>(sm + o);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kSingleThreads / 32, b = blockIdx.x;

    for (int j = tid; j < d.M; j += kSingleThreads) zs[j] = io.z[(size_t)b * d.M + j];
    for (int k = tid; k < d.Hr; k += kSingleThreads) hp[k] = io.first ? 0.f : io.h_z[(size_t)b * d.Hr + k];   // model.py:336-337
    if (d.A && tid == 0) {
        int start = 0;
        for (int dd = 0; dd < d.D; ++dd) { segs[dd] = start; start += io.desc_set_lens[dd]; if (start > d.NW) start = d.NW; }
        segs[d.D] = start;
    }
    MMG_SYNCTHREADS();
    // ---- GRU cell, gate order r,z,n (model.py:340) ---------------------------------------------------------------
    block_matvec(gi, P.p[MMG_P_REC_RNN_WIH], d.M, d.G3, zs, d.M, P.p[MMG_P_REC_RNN_BIH]);
    block_matvec(gh, P.p[MMG_P_REC_RNN_WHH], d.Hr, d.G3, hp, d.Hr, P.p[MMG_P_REC_RNN_BHH]);
    MMG_SYNCTHREADS();
    for (int k = tid; k < d.Hr; k += kSingleThreads) {
        const float r = sigmoidf_(gi[k] + gh[k]);
        const float uu = sigmoidf_(gi[d.Hr + k] + gh[d.Hr + k]);
        const float nn = tanhf(gi[2 * d.Hr + k] + r * gh[2 * d.Hr + k]);
        const float h = nn + uu * (hp[k] - nn);
        hn[k] = h;
        io.h_z[(size_t)b * d.Hr + k] = h;
    }
    MMG_SYNCTHREADS();
    // ---- heads that only need h': STOP (model.py:414-429), y1's hidden half, w_h --------------------------------------
    block_matvec(y1h, P.p[MMG_P_REC_Y1_W] + d.y1_hcol, d.Hr + d.WV, d.Hr, hn, d.Hr, P.p[MMG_P_REC_Y1_B]);
    block_matvec(whv, P.p[MMG_P_REC_WH_W], d.Hr, d.Hr, hn, d.Hr, P.p[MMG_P_REC_WH_B]);
    block_matvec(misc, P.p[MMG_P_REC_S_W], d.Hr, 1, hn, d.Hr, P.p[MMG_P_REC_S_B]);
    if (d.A) block_matvec(dh, P.p[MMG_P_REC_DH_W], d.Hr, d.A, hn, d.Hr, P.p[MMG_P_REC_DH_B]);                   // model.py:359
    MMG_SYNCTHREADS();
    if (tid == 0) {
        const float sp = sigmoidf_(misc[0]);
        float sbit;
        if (io.train) {
            sbit = draw_bit(io.u_stop, (size_t)b, sp, io.seed, io.counter, 1, (unsigned)(b + d.row0), 0);
        } else {
            const float prod = (io.first || !d.s_prob_prod) ? sp : io.s_prob_prod[b] * sp;          // model.py:423-427
            io.s_prob_prod[b] = prod;
            sbit = rintf(prod);
        }
        io.s[b] = sbit;
        io.s_prob[b] = sp;
    }
    const float* dsrc = io.desc;      // rows mixed by softmax(y): class descriptions, or the attended bags of words
    int dld = d.WV;
    if (d.A) {
        // ---- -desc_attn (model.py:344-410): scores over all words, softmax inside each class, attended descriptions ----
        for (int n = warp; n < d.NW; n += nwarps) {
            float s = 0.f;
            for (int a = lane; a < d.A; a += 32) {
                float dd = ldg(P.p[MMG_P_REC_DD_B] + a);
                const float* wr = P.p[MMG_P_REC_DD_W] + (size_t)a * d.WV;
                const float* xr = io.desc_set + (size_t)n * d.WV;
                for (int v = 0; v < d.WV; ++v) dd = fmaf(ldg(wr + v), ldg(xr + v), dd);
                s = fmaf(ldg(P.p[MMG_P_REC_DA_W] + a), tanhf(dd + dh[a]), s);
            }
            s = warp_sum(s);
            if (lane == 0) ev[n] = s + ldg(P.p[MMG_P_REC_DA_B]);
        }
        MMG_SYNCTHREADS();
        for (int dd = warp; dd < d.D; dd += nwarps) {
            const int s0 = segs[dd], s1 = segs[dd + 1];
            float mx = -INFINITY;
            for (int n = s0 + lane; n < s1; n += 32) mx = fmaxf(mx, ev[n]);
            mx = warp_max(mx);
            float se = 0.f;
            for (int n = s0 + lane; n < s1; n += 32) se += expf(ev[n] - mx);
            se = warp_sum(se);
            for (int n = s0 + lane; n < s1; n += 32) att[n] = expf(ev[n] - mx) / se;
        }
        MMG_SYNCTHREADS();
        for (int idx = tid; idx < d.D * d.WV; idx += kSingleThreads) {
            const int dd = idx / d.WV, v = idx % d.WV;
            float s = 0.f;
            for (int n = segs[dd]; n < segs[dd + 1]; ++n) s = fmaf(att[n], ldg(io.desc_set + (size_t)n * d.WV + v), s);
            ad[dd * WVP + v] = s;
        }
        MMG_SYNCTHREADS();
        dsrc = ad;
        dld = WVP;
    }
    // ---- class scores y[d] = y2(relu(y1([h' ; desc_d]))) (model.py:412,432-433), one warp per class -----------------------
    {
        const float* w1d = P.p[MMG_P_REC_Y1_W] + d.y1_dcol;
        for (int dd = warp; dd < d.D; dd += nwarps) {
            float s = 0.f;
            for (int k = lane; k < d.Hr; k += 32) {
                float pre = y1h[k];
                const float* wr = w1d + (size_t)k * (d.Hr + d.WV);
                const float* xr = dsrc + (size_t)dd * dld;
                for (int v = 0; v < d.WV; ++v) pre = fmaf(ldg(wr + v), xr[v], pre);
                s = fmaf(ldg(P.p[MMG_P_REC_Y2_W] + k), fmaxf(pre, 0.f), s);
            }
            s = warp_sum(s);
            if (lane == 0) {
                const float v = s + ldg(P.p[MMG_P_REC_Y2_B]);
                yv[dd] = v;
                io.y[(size_t)b * d.D + dd] = v;
            }
        }
    }
    MMG_SYNCTHREADS();
    // ---- q = softmax(y) (detached, model.py:441), wd = q . descriptions (442-449) ---------------------------------------
    if (warp == 0) {
        float mx = -INFINITY;
        for (int dd = lane; dd < d.D; dd += 32) mx = fmaxf(mx, yv[dd]);
        mx = warp_max(mx);
        float se = 0.f;
        for (int dd = lane; dd < d.D; dd += 32) se += expf(yv[dd] - mx);
        se = warp_sum(se);
        for (int dd = lane; dd < d.D; dd += 32) qv[dd] = expf(yv[dd] - mx) / se;
    }
    MMG_SYNCTHREADS();
    for (int v = tid; v < d.WV; v += kSingleThreads) {
        float s = 0.f;
        for (int dd = 0; dd < d.D; ++dd) s = fmaf(qv[dd], dsrc[(size_t)dd * dld + v], s);
        wdv[v] = s;
    }
    MMG_SYNCTHREADS();
    // ---- h_w = tanh(w_h(h') + w_d(wd)) (model.py:452), message head (454-475) ------------------------------------------
    block_matvec(gi, P.p[MMG_P_REC_WD_W], d.WV, d.Hr, wdv, d.WV, nullptr);
    MMG_SYNCTHREADS();
    for (int k = tid; k < d.Hr; k += kSingleThreads) {
        const float hw = tanhf(whv[k] + gi[k]);
        gh[k] = hw;
        io.h_w[(size_t)b * d.Hr + k] = hw;
    }
    MMG_SYNCTHREADS();
    block_matvec(lg, P.p[MMG_P_REC_W_W], d.Hr, d.M, gh, d.Hr, P.p[MMG_P_REC_W_B]);
    MMG_SYNCTHREADS();
    for (int j = tid; j < d.M; j += kSingleThreads) {
        const size_t i = (size_t)b * d.M + j;
        float wv = lg[j];
        if (d.use_binary) {
            const float p = sigmoidf_(lg[j]);
            io.w_probs[i] = p;
            wv = io.train ? draw_bit(io.u_rec, i, p, io.seed, io.counter, 2, (unsigned)(b + d.row0), (unsigned)j) : rintf(p);
            if (d.flip_rec >= 0.f && (io.train || d.flipout_dev))
                wv = flip_bit(wv, d.flip_rec, io.u_flip, i, io.seed, io.counter, 0, 1, (unsigned)(b + d.row0), (unsigned)j);
            if (d.ignore_receiver) wv = 0.f;                                                        // model.py:470-472
        }
        io.w[i] = wv;
    }
}

// Baseline.forward (model.py:496-516): linear2(relu(linear1(cat(x, binary, inp)))), any of the three pieces may be absent.
MMG_GLOBAL void __launch_bounds__(kSingleThreads)
k_baseline_step(const float* w1, const float* b1, const float* w2, const float* b2, int Hb, const float* x, int nx,
                const float* binary, int nb, const float* inp, int ni, float* out) {
    MMG_DYN_SMEM(smem_raw);
    float* in = reinterpret_cast<float*>(smem_raw);
    MMG_SHARED float red[kSingleThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kSingleThreads / 32, r = blockIdx.x;
    const int K = nx + nb + ni;
    for (int k = tid; k < K; k += kSingleThreads)
        in[k] = k < nx ? x[(size_t)r * nx + k] : (k < nx + nb ? binary[(size_t)r * nb + k - nx] : inp[(size_t)r * ni + k - nx - nb]);
    MMG_SYNCTHREADS();
    float acc = 0.f;
    for (int n = warp; n < Hb; n += nwarps) {
        float s = 0.f;
        for (int k = lane; k < K; k += 32) s = fmaf(ldg(w1 + (size_t)n * K + k), in[k], s);
        s = warp_sum(s);
        acc = fmaf(fmaxf(s + ldg(b1 + n), 0.f), ldg(w2 + n), acc);      // identical in every lane
    }
    if (lane == 0) red[warp] = acc;
    MMG_SYNCTHREADS();
    if (tid == 0) {
        float s = ldg(b2);
        for (int wq = 0; wq < nwarps; ++wq) s += red[wq];
        out[r] = s;
    }
}

}  // namespace mmg
