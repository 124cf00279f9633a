// Fast path of the conversation kernels for the dimensions every BASELINE.json configuration uses
// (img_h_dim = 256, rec_hidden = 64, rec_w_dim = sender_out_dim in {32, 64}); other shapes run the generic kernels in
// mmg_exchange_fwd.cuh / mmg_exchange_bwd.cuh.  Same math, same outputs, same saved activations.
//
// The recurrence is latency-bound (one example's step is ~50k MACs), so the design goal is the shortest dependent
// chain per exchange step, not throughput:
//   * all shapes are compile-time: every loop is unrolled, no integer division in the step loop;
//   * every matrix sits in shared memory in the order its phase reads it (a warp's float4 loads are 512 contiguous
//     bytes, conflict-free), staged once per kernel by the TMA unit (cp.async.bulk + mbarrier);
//   * reductions along K use warp shuffles instead of a shared-memory exchange, which removes a barrier per mat-vec:
//     7 barriers per step (the generic kernel needs 12);
//   * W_hh . h' for step t+1 rides in the same phase as the heads of step t (both only need h');
//   * the backward pass splits into a t-parallel part (message/STOP/class-score head deltas for all T steps at once)
//     and the true BPTT chain, which is reduced to gate algebra + one 192x64 mat-vec per step on shared memory.
#pragma once
#include "mmg_exchange_fwd.cuh"

namespace mmg {

MMG_DEVICE float dot4(const float4& a, const float4& b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
    return acc;
}

// ---- image packing (called by k_pre) ---------------------------------------------------------------------------
MMG_DEVICE float fast_fwd_image_elem(const Dims& d, const FastFwdImage& im, const ParamPtrs& P, int e) {
    const int M = d.M;
    if (e < im.wb) { const int q = e - im.wc; const int c = q & 3, f4 = q >> 2, half = f4 & 1, n = (f4 >> 1) & 255, qq = f4 >> 9;
        return ldg(P.p[MMG_P_SEN_CODE_W] + (size_t)n * M + half * (M / 2) + 4 * qq + c); }
    if (e < im.b_code) return ldg(P.p[MMG_P_SEN_BIN_W] + (e - im.wb));
    if (e < im.hw0) return ldg(P.p[MMG_P_SEN_CODE_B] + (e - im.b_code));
    if (e < im.b_b) return 0.f;                                   // hw0: dot role
    if (e < im.sender_end) return ldg(P.p[MMG_P_SEN_BIN_B] + (e - im.b_b));
    if (e < im.wfull) { const int q = e - im.wih; const int c = q & 3, f4 = q >> 2, part = f4 & 7, k = (f4 >> 3) & 63, gq = f4 >> 9;
        const int MQ = M / 32, g = gq / MQ, qq = gq % MQ;
        return ldg(P.p[MMG_P_REC_RNN_WIH] + (size_t)(g * 64 + k) * M + part * (M / 8) + 4 * qq + c); }
    if (e < im.wghn) { const int q = e - im.wfull; const int c = q & 3, f4 = q >> 2, half = f4 & 1, o = (f4 >> 1) & 255, qq = f4 >> 9;
        const int col = half * 32 + 4 * qq + c;
        if (o < 64) return ldg(P.p[MMG_P_REC_Y1_W] + (size_t)o * (64 + d.WV) + col);
        if (o < 128) return ldg(P.p[MMG_P_REC_WH_W] + (size_t)(o - 64) * 64 + col);
        return ldg(P.p[MMG_P_REC_RNN_WHH] + (size_t)(o - 128) * 64 + col); }
    if (e < im.ww) { const int q = e - im.wghn; const int c = q & 3, f4 = q >> 2, part = f4 & 7, k = (f4 >> 3) & 63, qq = f4 >> 9;
        return ldg(P.p[MMG_P_REC_RNN_WHH] + (size_t)(128 + k) * 64 + part * 8 + 4 * qq + c); }
    if (e < im.b_ih) { const int q = e - im.ww; const int c = q & 3, f4 = q >> 2, TPO = kFastThreads / M, KPT = 64 / TPO;
        const int part = f4 % TPO, j = (f4 / TPO) % M, qq = f4 / (TPO * M);
        return ldg(P.p[MMG_P_REC_W_W] + (size_t)j * 64 + part * KPT + 4 * qq + c); }
    if (e < im.b_full) return ldg(P.p[MMG_P_REC_RNN_BIH] + (e - im.b_ih));
    if (e < im.b_ghn) { const int o = e - im.b_full;
        if (o < 64) return 0.f;
        if (o < 128) return ldg(P.p[MMG_P_REC_WH_B] + (o - 64));
        return ldg(P.p[MMG_P_REC_RNN_BHH] + (o - 128)); }
    if (e < im.ws) return ldg(P.p[MMG_P_REC_RNN_BHH] + 128 + (e - im.b_ghn));
    if (e < im.b_w) return ldg(P.p[MMG_P_REC_S_W] + (e - im.ws));
    if (e < im.w2) return ldg(P.p[MMG_P_REC_W_B] + (e - im.b_w));
    if (e < im.misc) return ldg(P.p[MMG_P_REC_Y2_W] + (e - im.w2));
    if (e < im.y1d) { const int i = e - im.misc;
        return i == 0 ? ldg(P.p[MMG_P_REC_Y2_B]) : (i == 1 ? ldg(P.p[MMG_P_REC_S_B]) : 0.f); }
    return 0.f;                                                   // y1d / wdd: dot role
}

MMG_DEVICE float fast_bwd_image_elem(const Dims& d, const FastBwdImage& im, const ParamPtrs& P, int e) {
    if (e < im.whT) { const int q = e - im.wwT; const int c = q & 3, f4 = q >> 2, k = f4 & 63, j4 = f4 >> 6;
        return ldg(P.p[MMG_P_REC_W_W] + (size_t)(4 * j4 + c) * 64 + k); }
    if (e < im.w1hT) { const int q = e - im.whT; const int c = q & 3, f4 = q >> 2, k = f4 & 63, k4 = f4 >> 6;
        return ldg(P.p[MMG_P_REC_WH_W] + (size_t)(4 * k4 + c) * 64 + k); }
    if (e < im.whhT) { const int q = e - im.w1hT; const int c = q & 3, f4 = q >> 2, k = f4 & 63, k4 = f4 >> 6;
        return ldg(P.p[MMG_P_REC_Y1_W] + (size_t)(4 * k4 + c) * (64 + d.WV) + k); }
    if (e < im.ws) { const int q = e - im.whhT; const int c = q & 3, f4 = q >> 2, part = f4 & 3, k = (f4 >> 2) & 63, qq = f4 >> 8;
        return ldg(P.p[MMG_P_REC_RNN_WHH] + (size_t)(part * 48 + 4 * qq + c) * 64 + k); }
    if (e < im.w2) return ldg(P.p[MMG_P_REC_S_W] + (e - im.ws));
    if (e < im.y1d) return ldg(P.p[MMG_P_REC_Y2_W] + (e - im.w2));
    return 0.f;                                                   // y1d: dot role
}

// ---- forward -----------------------------------------------------------------------------------------------------
// 512 threads (16 warps) per CTA: the step is instruction-issue bound, so each phase is spread over all four
// schedulers with 4 warps each; K-splits inside a phase meet through 1-4 shuffles.
MMG_HOST_DEVICE int fast_uni_stride(int M) { return 2 * M + 4; }       // per (example, step): z draws, w draws, stop draw
MMG_HOST_DEVICE int fast_fwd_state_floats(int BT, int M, int D, int T) {
    // hx, av (256 each) | win, zv (M each) | hv, y1hv, whv, hwv (64 each) | ghv (192) | yv (DP) | ev (16 warps x DP)
    // | uniforms (T x stride) | sprod, smask | barrier
    return BT * (2 * kFastHi + 2 * M + 4 * kFastHr + 3 * kFastHr + align4(D) + (kFastThreads / 32) * align4(D) +
                 T * fast_uni_stride(M)) + align4(2 * BT) + 8;
}

template <int BT, int M, bool kSenderSmem>
MMG_GLOBAL void __launch_bounds__(kFastThreads, 1)
k_exchange_fwd_fast(Dims d, WsPtrs W, ExchangeInputs in, const float* b_img, int row_offset) {
    constexpr int HI = kFastHi, HR = kFastHr, NT = kFastThreads, NW = NT / 32;
    constexpr int MQ = M / 32, TPO = NT / M, KPT = HR / TPO, OPW = M / NW, UST = 2 * M + 4;
    MMG_DYN_SMEM(smem_raw);
    float* sm = reinterpret_cast<float*>(smem_raw);
    const FastFwdImage im = make_fast_fwd_image(M, d.D);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = blockIdx.x * BT;
    const int D = d.D, DP = align4(D), B = d.B, T = d.T;
    const int img0 = kSenderSmem ? 0 : im.sender_end;
    float* img = sm - img0;
    int o = im.total - img0;
    float* hx = sm + o;    o += BT * HI;
    float* av = sm + o;    o += BT * HI;
    float* win = sm + o;   o += BT * M;
    float* zv = sm + o;    o += BT * M;
    float* hv = sm + o;    o += BT * HR;
    float* y1hv = sm + o;  o += BT * HR;
    float* whv = sm + o;   o += BT * HR;
    float* hwv = sm + o;   o += BT * HR;
    float* ghv = sm + o;   o += BT * 3 * HR;
    float* yv = sm + o;    o += BT * DP;
    float* ev = sm + o;    o += BT * NW * DP;
    float* uni = sm + o;   o += BT * T * UST;
    float* sprod = sm + o; o += BT;
    float* smask = sm + o; o += BT;
    o = align4(o);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + o);

    const float* gimg = W.fwd_image;
    const float* simg = kSenderSmem ? img : gimg;                 // sender sections: shared memory or L2
    const float4* Wc4 = reinterpret_cast<const float4*>(simg + im.wc);
    const float4* Wb4 = reinterpret_cast<const float4*>(simg + im.wb);
    const float* b_code = simg + im.b_code;
    const float* hw0 = simg + im.hw0;
    const float* b_b = simg + im.b_b;
    const float4* Wih4 = reinterpret_cast<const float4*>(img + im.wih);
    const float4* Wfull4 = reinterpret_cast<const float4*>(img + im.wfull);
    const float4* Wghn4 = reinterpret_cast<const float4*>(img + im.wghn);
    const float4* Ww4 = reinterpret_cast<const float4*>(img + im.ww);
    const float* b_ih = img + im.b_ih;
    const float* b_full = img + im.b_full;
    const float* b_ghn = img + im.b_ghn;
    const float* wsv = img + im.ws;
    const float* b_w = img + im.b_w;
    const float4* w2_4 = reinterpret_cast<const float4*>(img + im.w2);
    const float4* y1d4 = reinterpret_cast<const float4*>(img + im.y1d);
    const float* wdd = img + im.wdd;
    const bool train = in.train != 0;
    const bool binary = d.use_binary != 0;
    const bool own_draws = train && in.u_sen == nullptr;   // on-device Philox stream (else: injected float64 uniforms)

    // ---- prologue ----------------------------------------------------------------------------------------------
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    MMG_SYNCTHREADS();
    pdl_wait();
    if (tid == 0) tma_stage(sm, gimg + img0, (uint32_t)(im.total - img0) * 4u, bar);
    {   // h_x rows of this CTA: split-K partials of K_pre summed in a fixed order + bias (model.py:195)
        const int n = tid >> 1, half = tid & 1;
#pragma unroll
        for (int bt = 0; bt < BT; ++bt) {
            const int b = b0 + bt;
            float v = 0.f;
            if (b < B) for (int s = half; s < W.hx_split; s += 2) v += W.hx_part[((size_t)s * B + b) * HI + n];
            v = group_sum<2>(v);
            if (b < B) v += ldg(b_img + n);
            if (half == 0) {
                hx[bt * HI + n] = v;
                if (b < B) W.h_x[(size_t)b * HI + n] = v;
            }
            if (tid < HR) {
                float h = 0.f;
                if (b < B) {
                    if (in.h0 != nullptr) h = in.h0[(size_t)b * HR + tid];
                    W.h_z[(size_t)b * HR + tid] = h;                    // slot 0 = state entering step 0
                }
                hv[bt * HR + tid] = h;
            }
            if (tid < M) {
                win[bt * M + tid] = d.first_rec;                        // model.py:786
                if (b < B) W.rec_feats[(size_t)b * M + tid] = d.first_rec;   // slot 0
            }
            if (tid == 0) { sprod[bt] = 1.f; smask[bt] = 1.f; if (b < B) W.stop_mask[b] = 1; }
        }
    }
    if (own_draws) {
        // every Bernoulli draw of this CTA's conversations up front (same Philox counters as the generic kernel:
        // stream = 4 t + {0 sender, 1 stop, 2 receiver}, counter = row * 65536 + column / 4)
        const unsigned long long seed = W.rng_state[0], iter = W.rng_state[1];
        constexpr int GPS = 2 * (M / 4) + 1;                             // 4-wide groups per (example, step)
        for (int idx = tid; idx < BT * T * GPS; idx += NT) {
            const int g = idx % GPS, t = (idx / GPS) % T, bt = idx / (GPS * T);
            const unsigned row = (unsigned)(b0 + bt + row_offset);
            float r[4];
            float* dst = uni + (bt * T + t) * UST;
            if (g < M / 4) {
                philox_uniform4(seed, iter, t * 4 + 0, row * 65536u + g, r);
                dst += 4 * g;
            } else if (g < 2 * (M / 4)) {
                philox_uniform4(seed, iter, t * 4 + 2, row * 65536u + (g - M / 4), r);
                dst += M + 4 * (g - M / 4);
            } else {
                philox_uniform4(seed, iter, t * 4 + 1, row * 65536u, r);
                dst += 2 * M;
            }
            dst[0] = r[0]; dst[1] = r[1]; dst[2] = r[2]; dst[3] = r[3];
        }
    }
#ifdef MMG_CPU_EMU
    MMG_SYNCTHREADS();
#endif
    mbar_wait(bar, 0);
    MMG_SYNCTHREADS();

    // Heads phase: everything that only needs h' — class-score / message-head pre-activations, the STOP bit and
    // W_hh . h' + b_hh for the NEXT step's gates.  `t < 0`: prologue call, only the W_hh part is kept.
    auto heads = [&](int t) {
        {   // rows [y1h ; w_h ; gh_r ; gh_u]: two threads per output, 32 reduction elements each
            const int oo = tid >> 1, half = tid & 1;
            float acc[BT];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) acc[bt] = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 w = Wfull4[(q * 256 + oo) * 2 + half];
#pragma unroll
                for (int bt = 0; bt < BT; ++bt)
                    acc[bt] = dot4(w, *reinterpret_cast<const float4*>(hv + bt * HR + half * 32 + 4 * q), acc[bt]);
            }
            const float bias = b_full[oo];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const float v = group_sum<2>(acc[bt]) + bias;
                const int b = b0 + bt;
                if (half == 0) {
                    if (oo < HR) {
                        if (t >= 0) {
                            y1hv[bt * HR + oo] = v;
                            if (b < B) W.y1h[((size_t)t * B + b) * HR + oo] = v;
                        }
                    } else if (oo < 2 * HR) {
                        whv[bt * HR + oo - HR] = v;
                    } else {
                        ghv[bt * 3 * HR + oo - 2 * HR] = v;
                    }
                }
            }
        }
        {   // rows gh_n: 8 threads per output, 8 reduction elements each
            const int k = tid >> 3, part = tid & 7;
            float acc[BT];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) acc[bt] = 0.f;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float4 w = Wghn4[(q * HR + k) * 8 + part];
#pragma unroll
                for (int bt = 0; bt < BT; ++bt)
                    acc[bt] = dot4(w, *reinterpret_cast<const float4*>(hv + bt * HR + part * 8 + 4 * q), acc[bt]);
            }
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const float v = group_sum<8>(acc[bt]);
                if (part == 0) ghv[bt * 3 * HR + 2 * HR + k] = b_ghn[k] + v;
            }
        }
        if (t >= 0 && warp == NW - 1) {   // STOP bit (model.py:414-429, 852)
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                float v = wsv[lane] * hv[bt * HR + lane] + wsv[lane + 32] * hv[bt * HR + lane + 32];
                v = warp_sum(v);
                const int b = b0 + bt;
                if (lane == 0) {
                    const size_t row = (size_t)t * B + b;
                    const float sp = sigmoidf_(v + img[im.misc + 1]);
                    float sbit;
                    if (train) {
                        if (b >= B) sbit = 0.f;
                        else if (own_draws) sbit = (uni[(bt * T + t) * UST + 2 * M] < sp) ? 1.f : 0.f;
                        else sbit = (in.u_stop[row] < (double)sp) ? 1.f : 0.f;
                    } else {
                        const float prod = (t == 0 || !d.s_prob_prod) ? sp : sprod[bt] * sp;
                        sprod[bt] = prod;
                        sbit = rintf(prod);
                    }
                    const float m = fminf(smask[bt], sbit);
                    smask[bt] = m;
                    if (b < B) {
                        W.stop_feat[row] = sbit;
                        W.stop_prob[row] = sp;
                        W.stop_mask[(size_t)(t + 1) * B + b] = (unsigned char)(m != 0.f);
                    }
                }
            }
        }
    };

    heads(-1);                                   // gh for step 0 from the initial state
    const float4 w2q = w2_4[lane & 15];
    const float y2b = img[im.misc];
    MMG_SYNCTHREADS();

    for (int t = 0; t < T; ++t) {
        // ---- P1: sender hidden a = tanh(h_x + code_layer(w_prev)) (model.py:199-216): 2 threads per unit ----------
        {
            const int n = tid >> 1, half = tid & 1;
            float acc[BT];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) acc[bt] = 0.f;
            if (t > 0) {
#pragma unroll
                for (int q = 0; q < M / 8; ++q) {
                    const float4 w = kSenderSmem ? Wc4[(q * HI + n) * 2 + half] : ldg4(Wc4 + (q * HI + n) * 2 + half);
#pragma unroll
                    for (int bt = 0; bt < BT; ++bt)
                        acc[bt] = dot4(w, *reinterpret_cast<const float4*>(win + bt * M + half * (M / 2) + 4 * q), acc[bt]);
                }
            }
            const float base = (t == 0) ? hw0[n] : b_code[n];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const int b = b0 + bt;
                const float a = tanhf(hx[bt * HI + n] + (base + group_sum<2>(acc[bt])));
                if (half == 0) {
                    av[bt * HI + n] = a;
                    if (b < B) W.a_s[((size_t)t * B + b) * HI + n] = a;
                }
                if (t > 0 && tid < M && b < B) W.code_in[((size_t)t * B + b) * M + tid] = win[bt * M + tid];
            }
        }
        MMG_SYNCTHREADS();
        // ---- P2: binary_layer + sender message (model.py:216-238, 814-820): warp w owns outputs [w*OPW, (w+1)*OPW) --
        {
            float4 a0[BT], a1[BT];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                a0[bt] = *reinterpret_cast<const float4*>(av + bt * HI + 4 * lane);
                a1[bt] = *reinterpret_cast<const float4*>(av + bt * HI + 128 + 4 * lane);
            }
            float s[BT][OPW];
#pragma unroll
            for (int oo = 0; oo < OPW; ++oo) {
                const int j = warp * OPW + oo;
                const float4 w0 = kSenderSmem ? Wb4[j * (HI / 4) + lane] : ldg4(Wb4 + j * (HI / 4) + lane);
                const float4 w1 = kSenderSmem ? Wb4[j * (HI / 4) + 32 + lane] : ldg4(Wb4 + j * (HI / 4) + 32 + lane);
#pragma unroll
                for (int bt = 0; bt < BT; ++bt) s[bt][oo] = dot4(w1, a1[bt], dot4(w0, a0[bt], 0.f));
            }
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                float mine = 0.f;
#pragma unroll
                for (int oo = 0; oo < OPW; ++oo) {
                    const float v = warp_sum(s[bt][oo]);
                    if (lane == oo) mine = v;
                }
                if (lane < OPW) {
                    const int j = warp * OPW + lane, b = b0 + bt;
                    const float logit = b_b[j] + mine;
                    const size_t row = (size_t)t * B + b;
                    float p = 0.f, zval;
                    if (binary) {
                        p = sigmoidf_(logit);
                        if (train) {
                            if (b >= B) zval = 0.f;
                            else if (own_draws) zval = (uni[(bt * T + t) * UST + j] < p) ? 1.f : 0.f;
                            else zval = (in.u_sen[row * M + j] < (double)p) ? 1.f : 0.f;
                        } else {
                            zval = rintf(p);
                        }
                    } else {
                        zval = logit;
                    }
                    if (in.corrupt_mask != nullptr) zval = fabsf(zval - in.corrupt_mask[j]);
                    zv[bt * M + j] = zval;
                    if (b < B) {
                        W.sen_feats[row * M + j] = zval;
                        if (binary) W.sen_probs[row * M + j] = p;
                    }
                }
            }
        }
        MMG_SYNCTHREADS();
        // ---- P3: GRU step (model.py:340), gate order r,z,n; W_hh . h + b_hh is already in ghv ----------------------
        {
            const int k = tid >> 3, part = tid & 7;
            float g[BT][3];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) { g[bt][0] = 0.f; g[bt][1] = 0.f; g[bt][2] = 0.f; }
#pragma unroll
            for (int gate = 0; gate < 3; ++gate)
#pragma unroll
                for (int q = 0; q < MQ; ++q) {
                    const float4 w = Wih4[((gate * MQ + q) * HR + k) * 8 + part];
#pragma unroll
                    for (int bt = 0; bt < BT; ++bt)
                        g[bt][gate] = dot4(w, *reinterpret_cast<const float4*>(zv + bt * M + part * (M / 8) + 4 * q), g[bt][gate]);
                }
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const float gi_r = group_sum<8>(g[bt][0]) + b_ih[k];
                const float gi_u = group_sum<8>(g[bt][1]) + b_ih[HR + k];
                const float gi_n = group_sum<8>(g[bt][2]) + b_ih[2 * HR + k];
                const float gh_r = ghv[bt * 3 * HR + k], gh_u = ghv[bt * 3 * HR + HR + k], gh_n = ghv[bt * 3 * HR + 2 * HR + k];
                // one sigmoid instruction stream serves both gates: lane part 0 evaluates r, lane part 1 evaluates u
                const float sg = sigmoidf_((part & 1) ? (gi_u + gh_u) : (gi_r + gh_r));
                const float r = shfl_xor_f(sg, part), u = shfl_xor_f(sg, part ^ 1);
                const float nn = tanhf(gi_n + r * gh_n);
                const float hp = hv[bt * HR + k];
                const float hn = nn + u * (hp - nn);
                MMG_SYNCWARP();                      // every lane of the group has read hv[k] before it is overwritten
                if (part == 0) {
                    const int b = b0 + bt;
                    hv[bt * HR + k] = hn;
                    if (b < B) {
                        const size_t row = (size_t)t * B + b;
                        float* gg = W.gates + row * 4 * HR;
                        gg[k] = r; gg[HR + k] = u; gg[2 * HR + k] = nn; gg[3 * HR + k] = gh_n;
                        W.h_z[((size_t)(t + 1) * B + b) * HR + k] = hn;
                    }
                }
            }
        }
        MMG_SYNCTHREADS();
        // ---- P4: heads of h' + STOP bit + W_hh . h' for the next step ------------------------------------------------
        heads(t);
        MMG_SYNCTHREADS();
        // ---- P5: class scores y[d] = y2(relu(y1h + y1d[d])) (model.py:432-433): 16 lanes per class ----------------
        {
            const int sub = lane & 15, cw = lane >> 4;
            float4 ya[BT];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) ya[bt] = *reinterpret_cast<const float4*>(y1hv + bt * HR + 4 * sub);
            for (int c0 = 0; c0 < D; c0 += 2 * NW) {
                const int cls = c0 + warp * 2 + cw;
                const bool valid = cls < D;
                float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) r0 = y1d4[cls * 16 + sub];
#pragma unroll
                for (int bt = 0; bt < BT; ++bt) {
                    float s = 0.f;
                    s = fmaf(w2q.x, fmaxf(0.f, ya[bt].x + r0.x), s);
                    s = fmaf(w2q.y, fmaxf(0.f, ya[bt].y + r0.y), s);
                    s = fmaf(w2q.z, fmaxf(0.f, ya[bt].z + r0.z), s);
                    s = fmaf(w2q.w, fmaxf(0.f, ya[bt].w + r0.w), s);
                    s = group_sum<16>(s);
                    if (sub == 0 && valid) {
                        s += y2b;
                        yv[bt * DP + cls] = s;
                        const int b = b0 + bt;
                        if (b < B) W.y[((size_t)t * B + b) * D + cls] = s;
                    }
                }
            }
        }
        MMG_SYNCTHREADS();
        // ---- P6: q = softmax(y) (detached, model.py:441); h_w = tanh(w_h(h') + sum_d q_d wdd[d]) (442-452) ----------
        // every warp evaluates the D exponentials once into its own strip of shared memory (no block barrier)
        {
            const int k = tid >> 3, part = tid & 7;
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const int b = b0 + bt;
                float* e = ev + (bt * NW + warp) * DP;
                float mx = -INFINITY;
                for (int dd = lane; dd < D; dd += 32) mx = fmaxf(mx, yv[bt * DP + dd]);
                mx = warp_max(mx);
                float se = 0.f;
                for (int dd = lane; dd < D; dd += 32) {
                    const float x = expf(yv[bt * DP + dd] - mx);
                    e[dd] = x;
                    se += x;
                }
                se = warp_sum(se);
                const float inv = 1.f / se;
                MMG_SYNCWARP();
                if (train && warp == 0 && b < B)
                    for (int dd = lane; dd < D; dd += 32) W.q[((size_t)t * B + b) * D + dd] = e[dd] * inv;
                float acc = 0.f;
                for (int dd = part; dd < D; dd += 8) acc = fmaf(e[dd], wdd[((dd >> 3) * HR + k) * 8 + part], acc);
                acc = group_sum<8>(acc);
                if (part == 0) {
                    const float hw = tanhf(whv[bt * HR + k] + acc * inv);
                    hwv[bt * HR + k] = hw;
                    if (b < B) W.h_w[((size_t)t * B + b) * HR + k] = hw;
                }
                MMG_SYNCWARP();
            }
        }
        MMG_SYNCTHREADS();
        // ---- P8: receiver message w(h_w) (model.py:454-475): TPO threads per output ---------------------------------
        {
            const int j = tid / TPO, part = tid % TPO;
            float acc[BT];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) acc[bt] = 0.f;
#pragma unroll
            for (int q = 0; q < KPT / 4; ++q) {
                const float4 w = Ww4[(q * M + j) * TPO + part];
#pragma unroll
                for (int bt = 0; bt < BT; ++bt)
                    acc[bt] = dot4(w, *reinterpret_cast<const float4*>(hwv + bt * HR + part * KPT + 4 * q), acc[bt]);
            }
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const float v = group_sum<TPO>(acc[bt]);
                if (part == 0) {
                    const int b = b0 + bt;
                    const float logit = b_w[j] + v;
                    const size_t row = (size_t)t * B + b;
                    float p = 0.f, wv;
                    if (binary) {
                        p = sigmoidf_(logit);
                        if (train) {
                            if (b >= B) wv = 0.f;
                            else if (own_draws) wv = (uni[(bt * T + t) * UST + M + j] < p) ? 1.f : 0.f;
                            else wv = (in.u_rec[row * M + j] < (double)p) ? 1.f : 0.f;
                        } else {
                            wv = rintf(p);
                        }
                        if (d.ignore_receiver) wv = 0.f;
                    } else {
                        wv = logit;
                    }
                    win[bt * M + j] = wv;
                    if (b < B) {
                        W.rec_feats[((size_t)(t + 1) * B + b) * M + j] = wv;
                        if (binary) W.rec_probs[row * M + j] = p;
                    }
                }
            }
        }
        MMG_SYNCTHREADS();
    }
    pdl_launch_dependents();
}

// ---- backward ------------------------------------------------------------------------------------------------------
// One example per CTA.  CTAs [0, n_rec): receiver (BPTT); the rest: sender (no recurrence, model.py:807-811).
MMG_HOST_DEVICE int fast_bwd_rec_state_floats(int T, int M, int D) {
    // dlw (T,M) | dhw (T,64) | inj (T,64) | gat (T,5,64) | hws, y1hs (T,64 each) | dgh (2,192) | gv (64) | gout (DP) | dls (T) | barrier
    return T * M + 2 * T * kFastHr + 5 * T * kFastHr + 2 * T * kFastHr + 2 * 3 * kFastHr + kFastHr + align4(D) + align4(T) + 8;
}
MMG_HOST_DEVICE int fast_bwd_sen_state_floats(int T, int M) { return T * M + 2 * kFastHi + T * kFastHi + 8; }

template <int M>
MMG_GLOBAL void __launch_bounds__(kFastBwdThreads, 1)
k_exchange_bwd_fast(Dims d, WsPtrs W, const float* bin_w, const float* code_w, int n_rec_ctas) {
    constexpr int HI = kFastHi, HR = kFastHr, NT = kFastBwdThreads, M4 = M / 4;
    MMG_DYN_SMEM(smem_raw);
    float* sm = reinterpret_cast<float*>(smem_raw);
    const int tid = threadIdx.x;
    const int T = d.T, B = d.B, D = d.D;
    const bool binary = d.use_binary != 0;

    if ((int)blockIdx.x >= n_rec_ctas) {
        // =================================== sender ==============================================================
        const int b = (int)blockIdx.x - n_rec_ctas, n = tid;
        float* dlz = sm;                                              // (T, M)
        pdl_wait();
        float wb[M];                                                   // column n of binary_layer.weight
#pragma unroll
        for (int j = 0; j < M; ++j) wb[j] = ldg(bin_w + (size_t)j * HI + n);
        float* das0 = sm + T * M;                                     // (256) d a at t = 0
        float* cred = das0 + HI;                                      // (256 / M, M) partial sums
        float* as_s = cred + HI;                                      // (T, 256) saved tanh outputs of this example
        for (int t = 0; t < T; ++t) as_s[t * HI + n] = W.a_s[((size_t)t * B + b) * HI + n];
        for (int idx = tid; idx < T * M; idx += NT) {
            const int t = idx / M, j = idx % M;
            const size_t i = ((size_t)t * B + b) * M + j;
            const float p = W.sen_probs[i];
            const float dl = W.g_sen_probs[i] * p * (1.f - p);        // through the sigmoid (model.py:223)
            W.d_lz[i] = dl;
            dlz[idx] = dl;
        }
        MMG_SYNCTHREADS();
        float dhx = 0.f;
#pragma unroll 2
        for (int t = 0; t < T; ++t) {
            const size_t i = ((size_t)t * B + b) * HI + n;
            const float a = as_s[t * HI + n];
            float acc = 0.f;
#pragma unroll
            for (int j4 = 0; j4 < M4; ++j4) {
                const float4 g = *reinterpret_cast<const float4*>(dlz + t * M + 4 * j4);
                acc = fmaf(g.x, wb[4 * j4], acc); acc = fmaf(g.y, wb[4 * j4 + 1], acc);
                acc = fmaf(g.z, wb[4 * j4 + 2], acc); acc = fmaf(g.w, wb[4 * j4 + 3], acc);
            }
            const float das = acc * (1.f - a * a);                    // through tanh (model.py:216)
            W.d_as[i] = das;
            if (t == 0) das0[n] = das;
            dhx += das;                                               // h_x is shared by all steps (model.py:195)
        }
        W.dhx[(size_t)b * HI + n] = dhx;
        // d code_layer-input at t = 0 (the code is sigmoid(code_bias) for every example, model.py:199-200):
        // dcode_part[b][j] = sum_n d_a[0][n] code_layer.weight[n][j]; K_wgrad sums over b and applies d sigmoid.
        MMG_SYNCTHREADS();
        {
            const int j = tid % M, part = tid / M;                    // NT / M parts, each HI * M / NT rows
            constexpr int RPP = HI * M / NT;
            float acc = 0.f;
#pragma unroll 8
            for (int r = 0; r < RPP; ++r) {
                const int nn = part * RPP + r;
                acc = fmaf(das0[nn], ldg(code_w + (size_t)nn * M + j), acc);
            }
            cred[part * M + j] = acc;
        }
        MMG_SYNCTHREADS();
        if (tid < M) {
            float v = 0.f;
#pragma unroll
            for (int p = 0; p < NT / M; ++p) v += cred[p * M + tid];
            W.dcode_part[(size_t)b * M + tid] = v;
        }
        return;
    }

    // ===================================== receiver ==============================================================
    const FastBwdImage im = make_fast_bwd_image(M, D);
    const int b = blockIdx.x, lane = tid & 31;
    const int DP = align4(D);
    int o = im.total;
    float* dlw = sm + o;  o += T * M;
    float* dhw = sm + o;  o += T * HR;
    float* inj = sm + o;  o += T * HR;
    float* gat = sm + o;  o += 5 * T * HR;          // per step: r, u, n, gh_n, h_prev
    float* hws = sm + o;  o += T * HR;              // h_w of every step
    float* y1hs = sm + o; o += T * HR;              // y1h of every step (the prediction step is only known after the load)
    float* dghv = sm + o; o += 2 * 3 * HR;
    float* gv = sm + o;   o += HR;
    float* gout = sm + o; o += DP;
    float* dls = sm + o;  o += align4(T);
    o += (o & 1);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + o);
    const float4* WwT4 = reinterpret_cast<const float4*>(sm + im.wwT);
    const float4* WhT4 = reinterpret_cast<const float4*>(sm + im.whT);
    const float4* W1hT4 = reinterpret_cast<const float4*>(sm + im.w1hT);
    const float4* WhhT4 = reinterpret_cast<const float4*>(sm + im.whhT);
    const float* wsv = sm + im.ws;
    const float* w2 = sm + im.w2;
    const float* y1d = sm + im.y1d;

    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    MMG_SYNCTHREADS();
    pdl_wait();
    if (tid == 0) tma_stage(sm, W.bwd_image, (uint32_t)im.total * 4u, bar);
    // ---- A0: everything the chain needs, into shared memory -----------------------------------------------------
    for (int idx = tid; idx < T * M; idx += NT) {
        const int t = idx / M, j = idx % M;
        const size_t i = ((size_t)t * B + b) * M + j;
        float dl = 0.f;
        if (binary) {
            const float p = W.rec_probs[i];
            dl = W.g_rec_probs[i] * p * (1.f - p);
        }
        W.d_lw[i] = dl;
        dlw[idx] = dl;
    }
    for (int idx = tid; idx < T * 4 * HR; idx += NT) {
        const int t = idx >> 8, c = idx & 255;
        gat[t * 5 * HR + c] = W.gates[((size_t)t * B + b) * 4 * HR + c];
    }
    for (int idx = tid; idx < T * HR; idx += NT) {
        const int t = idx >> 6, k = idx & 63;
        gat[t * 5 * HR + 4 * HR + k] = W.h_z[((size_t)t * B + b) * HR + k];   // slot t = state entering step t
        hws[idx] = W.h_w[((size_t)t * B + b) * HR + k];
        y1hs[idx] = W.y1h[((size_t)t * B + b) * HR + k];
    }
    for (int t = tid; t < T; t += NT) {
        const size_t row = (size_t)t * B + b;
        const float sp = W.stop_prob[row];
        const float v = W.g_stop_prob[row] * sp * (1.f - sp);
        W.d_ls[row] = v;
        dls[t] = v;
    }
    for (int dd = tid; dd < D; dd += NT) gout[dd] = W.g_outp[(size_t)b * D + dd];
    const int ys = W.ystep[b];
#ifdef MMG_CPU_EMU
    MMG_SYNCTHREADS();
#endif
    mbar_wait(bar, 0);
    MMG_SYNCTHREADS();
    const int k = tid >> 2, part = tid & 3;
    // ---- A1: d h_w for every step (t-parallel); class-score head at the prediction step ----------------------------
    for (int t = part; t < T; t += 4) {
        float acc = 0.f;
#pragma unroll
        for (int j4 = 0; j4 < M4; ++j4) acc = dot4(WwT4[j4 * HR + k], *reinterpret_cast<const float4*>(dlw + t * M + 4 * j4), acc);
        const size_t i = ((size_t)t * B + b) * HR + k;
        const float hw = hws[t * HR + k];
        const float v = acc * (1.f - hw * hw);                        // through tanh (model.py:452)
        W.d_hw[i] = v;
        dhw[t * HR + k] = v;
    }
    {
        // y[d] = y2.bias + sum_k w2[k] relu(y1h[k] + y1d[d][k])   (model.py:432-433)
        const float yh = y1hs[ys * HR + k], wk = w2[k];
        float G = 0.f, dw2 = 0.f;
        for (int dd = part; dd < D; dd += 4) {
            const float g = gout[dd];
            const float pre = yh + y1d[((dd >> 2) * HR + k) * 4 + part];
            const float v = pre > 0.f ? g * wk : 0.f;
            W.dy1[((size_t)b * D + dd) * HR + k] = v;
            G += v;
            dw2 = fmaf(g, fmaxf(pre, 0.f), dw2);
        }
        G = group_sum<4>(G);
        dw2 = group_sum<4>(dw2);
        if (part == 0) {
            gv[k] = G;
            W.g_h[(size_t)b * HR + k] = G;
            W.dw2p[(size_t)b * HR + k] = dw2;
            W.hsel[(size_t)b * HR + k] = W.h_z[((size_t)(ys + 1) * B + b) * HR + k];
        }
    }
    MMG_SYNCTHREADS();
    // ---- A2: per-step injection into d h': W_h^T d_hw + s.weight * d_ls (+ W_1h^T G at the prediction step) --------
    for (int t = part; t < T; t += 4) {
        float acc = wsv[k] * dls[t];
#pragma unroll
        for (int k4 = 0; k4 < HR / 4; ++k4) acc = dot4(WhT4[k4 * HR + k], *reinterpret_cast<const float4*>(dhw + t * HR + 4 * k4), acc);
        if (t == ys) {
#pragma unroll
            for (int k4 = 0; k4 < HR / 4; ++k4) acc = dot4(W1hT4[k4 * HR + k], *reinterpret_cast<const float4*>(gv + 4 * k4), acc);
        }
        inj[t * HR + k] = acc;
    }
    MMG_SYNCTHREADS();
    // ---- B: the BPTT chain (h_z is never detached between steps, model.py:340) --------------------------------------
    float direct = 0.f, rec = 0.f;
    int buf = 0;
    for (int t = T - 1; t >= 0; --t) {
        const float* g = gat + t * 5 * HR;
        const float r = g[k], u = g[HR + k], nn = g[2 * HR + k], ghn = g[3 * HR + k], hp = g[4 * HR + k];
        const float dht = direct + rec + inj[t * HR + k];
        // h' = n + u (h - n)
        const float du = dht * (hp - nn);
        const float dn = dht * (1.f - u);
        direct = dht * u;
        const float dn_pre = dn * (1.f - nn * nn);
        const float dghn = dn_pre * r;
        const float dr_pre = dn_pre * ghn * r * (1.f - r);
        const float du_pre = du * u * (1.f - u);
        if (part == 0) {
            const size_t row = (size_t)t * B + b;
            float* gi = W.dgi + row * 3 * HR;
            float* gh = W.dgh + row * 3 * HR;
            gi[k] = dr_pre; gi[HR + k] = du_pre; gi[2 * HR + k] = dn_pre;
            gh[k] = dr_pre; gh[HR + k] = du_pre; gh[2 * HR + k] = dghn;
            float* dg = dghv + buf * 3 * HR;
            dg[k] = dr_pre; dg[HR + k] = du_pre; dg[2 * HR + k] = dghn;
        }
        if (t == 0) break;
        MMG_SYNCTHREADS();
        {   // d h_prev += W_hh^T . d gh: 4 threads per output, 48 reduction elements each
            const float* dg = dghv + buf * 3 * HR;
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < 12; ++q) acc = dot4(WhhT4[(q * HR + k) * 4 + part], *reinterpret_cast<const float4*>(dg + part * 48 + 4 * q), acc);
            rec = group_sum<4>(acc);
        }
        buf ^= 1;
    }
    (void)lane;
}

}  // namespace mmg
