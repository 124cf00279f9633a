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
#include "mmg_loss.cuh"

#if defined(MMG_PHASE_TIMING) && !defined(MMG_CPU_EMU)
// debug build only (scripts/phase_timing.py): per-phase clock stamps of CTA 0 into the g_sen_probs scratch
#define MMG_STAMP(p) do { if (lane == 0) stamps[(t * 8 + (p)) * 8 + warp] = (unsigned)clock64(); } while (0)
#define MMG_BSTAMP(slot) do { if (threadIdx.x == 0 && (blockIdx.x == 0 || (int)blockIdx.x == n_rec_ctas)) reinterpret_cast<unsigned*>(W.g_bs)[((int)blockIdx.x == 0 ? 0 : 32) + (slot)] = (unsigned)clock64(); } while (0)
#else
#define MMG_STAMP(p) do { } while (0)
#define MMG_BSTAMP(slot) MMG_TRACE_AT(3, slot)
#endif

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
    if (e < im.wb) { const int q = e - im.wc; const int k4 = q >> 10, n = (q >> 2) & 255, c = q & 3;
        return ldg(P.p[MMG_P_SEN_CODE_W] + (size_t)n * M + 4 * k4 + c); }
    if (e < im.b_code) return ldg(P.p[MMG_P_SEN_BIN_W] + (e - im.wb));
    if (e < im.hw0) return ldg(P.p[MMG_P_SEN_CODE_B] + (e - im.b_code));
    if (e < im.b_b) return 0.f;                                   // hw0: dot role
    if (e < im.sender_end) return ldg(P.p[MMG_P_SEN_BIN_B] + (e - im.b_b));
    if (e < im.whead) { const int q = e - im.wih; const int c = q & 3, f4 = q >> 2, part = f4 & 3, k = (f4 >> 2) & 63, gq = f4 >> 8;
        const int MQ = M / 16, g = gq / MQ, qq = gq % MQ;
        return ldg(P.p[MMG_P_REC_RNN_WIH] + (size_t)(g * 64 + k) * M + part * (M / 4) + 4 * qq + c); }
    if (e < im.wgh) { const int q = e - im.whead; const int c = q & 3, f4 = q >> 2, half = f4 & 1, o = (f4 >> 1) & 127, qq = f4 >> 8;
        const int col = half * 32 + 4 * qq + c;
        if (o < 64) return ldg(P.p[MMG_P_REC_Y1_W] + (size_t)o * (64 + d.WV) + d.y1_hcol + col);
        return ldg(P.p[MMG_P_REC_WH_W] + (size_t)(o - 64) * 64 + col); }
    if (e < im.ww) { const int q = e - im.wgh; const int c = q & 3, f4 = q >> 2, part = f4 & 3, k = (f4 >> 2) & 63, gq = f4 >> 8;
        const int g = gq >> 2, qq = gq & 3;
        return ldg(P.p[MMG_P_REC_RNN_WHH] + (size_t)(g * 64 + k) * 64 + part * 16 + 4 * qq + c); }
    if (e < im.b_ih) { const int q = e - im.ww; const int c = q & 3, f4 = q >> 2, LPO = kFastThreads / M, KPT = 64 / LPO;
        const int part = f4 % LPO, j = (f4 / LPO) % M, qq = f4 / (LPO * M);
        return ldg(P.p[MMG_P_REC_W_W] + (size_t)j * 64 + part * KPT + 4 * qq + c); }
    if (e < im.b_hh) return ldg(P.p[MMG_P_REC_RNN_BIH] + (e - im.b_ih));
    if (e < im.b_wh) return ldg(P.p[MMG_P_REC_RNN_BHH] + (e - im.b_hh));
    if (e < im.ws) return ldg(P.p[MMG_P_REC_WH_B] + (e - im.b_wh));
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
// Phase timing on B200 showed the step is bound by the DEPENDENT chain inside each phase (every instruction of a warp
// waits for the previous one), not by issue slots, so this kernel keeps every chain short: 4 independent accumulators
// per dot product, at most 3 shuffle stages per reduction, MUFU-based sigmoid/tanh/exp, loop-invariant addresses
// hoisted into registers, and the two products that are not on the critical path (W_hh . h' for the NEXT step, the
// STOP head) issued inside phases whose own chain leaves the pipes idle.
// With the chains short, the next wall is shared-memory bandwidth: ~200 KB of weights per step through a 128 B/clk
// pipe is ~1600 cycles.  The weights are loop constants, so each thread keeps ITS slice of every matrix in registers
// (kRegRecv: GRU + heads + message head, 112 registers; kRegSend: sender, 64 registers at msg_dim 32) and shared memory
// only carries the per-step vectors and the class tables.
MMG_HOST_DEVICE int fast_uni_stride(int M) { return 2 * M + 4; }       // per (example, step): z draws, w draws, stop draw
MMG_HOST_DEVICE int fast_fwd_state_floats(int BT, int M, int D, int T) {
    // hx, av (256 each) | win, zv (M each) | hv, y1hv, whv, hwv (64 each) | ghv (192) | yv (DP) | wmax (8)
    // | uniforms (T x stride) | sprod, smask | barrier
    // | ysel (DP): class scores of the prediction step (kept for the per-example epilogue)
    return BT * (2 * kFastHi + 2 * M + 4 * kFastHr + 3 * kFastHr + 2 * align4(D) + 8 + T * fast_uni_stride(M)) +
           align4(2 * BT) + 8;
}

// -desc_attn on the fast forward kernel (one example per CTA, any batch, msg_dim 32, attention width 64): the two word tables read
// every step by all threads (d_d(desc_set) and desc_set . w_d^T) live in shared memory with rows padded to 72 floats;
// desc_set . y1^T is read from L2 once per step by the (class, column group) threads.
enum { kFastAttnLd = 72, kFastAttnA = 64 };   // row stride 72: the per-unit reads of 4 consecutive words hit 32 distinct banks
MMG_HOST_DEVICE int fast_fwd_attn_floats(int D, int NW) {
    // tdd, twd (NW x 72) | ev, att (NWP) | dhv, vas, b1s (64) | y1e (D x 64) | seg (D+1), wcls (NW) as ints
    return 2 * NW * kFastAttnLd + 2 * align4(NW) + 3 * 64 + D * 64 + 64 * 64 + align4(D + 1) + align4(NW);
}
MMG_HOST_DEVICE bool fast_fwd_attn_dims(const Dims& d) {
    // any batch: one example per CTA, in waves when B exceeds the SM count (conversation CTAs never wait on anything)
    return d.Hi == kFastHi && d.Hr == kFastHr && d.M == 32 && d.T <= kFastMaxT && d.A == kFastAttnA && !d.mix_mou;
}

MMG_DEVICE void fma4(const float4& w, const float4& x, float4& acc) {      // two packed fp32x2 FMAs (FFMA2)
    const float2 lo = ffma2(make_float2(w.x, w.y), make_float2(x.x, x.y), make_float2(acc.x, acc.y));
    const float2 hi = ffma2(make_float2(w.z, w.w), make_float2(x.z, x.w), make_float2(acc.z, acc.w));
    acc.x = lo.x; acc.y = lo.y; acc.z = hi.x; acc.w = hi.y;
}
MMG_DEVICE float hsum4(const float4& a) { return (a.x + a.y) + (a.z + a.w); }
MMG_DEVICE float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
MMG_DEVICE float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

#ifdef MMG_NO_SAVE
#define MMG_SAVE_OK(b) ((b) < 0)
#else
#define MMG_SAVE_OK(b) (kPerf || (b) < B)
#endif

// kPerf: the training configuration bench.py measures (binary messages, on-device draws, no corruption mask, batch a
// multiple of BT) with every mode flag a compile-time constant: no mode branches and no row guards in the step loop.
template <int BT, int M, bool kRegSend, bool kPerf, bool kAttn = false>
MMG_GLOBAL void __launch_bounds__(kFastThreads, 1)
k_exchange_fwd_fast(Dims d, WsPtrs W, ExchangeInputs in, const float* b_img, int row_offset, const float* bs_w1,
                    const float* bs_b1, int n_conv_ctas, AttnArgs aa, int epilogue) {
    static_assert(!kAttn || (BT == 1 && M == 32 && kRegSend), "attention: one example per CTA, msg_dim 32");
    constexpr int HI = kFastHi, HR = kFastHr, NT = kFastThreads, NW = NT / 32;
    constexpr int M4 = M / 4, MQ = M / 16, LPO = NT / M, KPT = HR / LPO, KB = HI / LPO, UST = 2 * M + 4;
    MMG_DYN_SMEM(smem_raw);
    float* sm = reinterpret_cast<float*>(smem_raw);
    if ((int)blockIdx.x >= n_conv_ctas) {
        // Side role on SMs the conversations leave idle: U[b] = baseline_sen.linear1[:, :Hi] . h_x[b] + bias
        // (model.py:835-836).  h_x is the same for all T steps of an example, so this 9/10 of the sender-side baseline
        // GEMM is done once per example, concurrently with the exchange loop.
        pdl_wait(); pdl_launch_dependents();
        MMG_TRACE_AT(1, 4);
        const int ntn = cdiv(d.Hb, kTile);
        const int tile = (int)blockIdx.x - n_conv_ctas, nt = tile % ntn, mt = tile / ntn;
        // h_x rows are finalised by the conversation CTAs in their prologue (lower block indices, never blocked)
        if (threadIdx.x == 0) flag_wait(W.tickets + 1, (unsigned)n_conv_ctas);
        MMG_SYNCTHREADS();
        MMG_TRACE_AT(1, 5);
        Operand A = Operand{W.h_x, nullptr, nullptr, nullptr, HI, 0, 0, 0, 0, OP_PLAIN};
        Operand Bo = Operand{bs_w1, nullptr, nullptr, nullptr, HI + M, 0, 0, 0, 0, OP_PLAIN};
        float acc[4][4];
        gemm_tile(A, Bo, d.B, d.Hb, mt * kTile, nt * kTile, 0, HI, acc, nullptr, sm);
        const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int b = mt * kTile + ty * 4 + a;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int n = nt * kTile + tx * 4 + c;
                if (b < d.B && n < d.Hb) W.ubs[(size_t)b * d.Hb + n] = acc[a][c] + ldg(bs_b1 + n);
            }
        }
        MMG_TRACE_AT(1, 6);
        return;
    }
    const FastFwdImage im = make_fast_fwd_image(M, d.D);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = blockIdx.x * BT;
    const int D = d.D, DP = align4(D), B = d.B, T = d.T;
    // shared-memory image: [sender section (unless it lives in registers)] [tail: biases, class tables]
    const int snd = kRegSend ? 0 : im.sender_end;
    float* simg_s = sm;                                           // sender section at its image offsets
    float* img = sm + snd - im.b_ih;                              // img[off] valid for off >= im.b_ih
    int o = snd + im.total - im.b_ih;
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
    float* ysel = sm + o;  o += BT * DP;
    float* wmax = sm + o;  o += BT * 8;
    float* uni = sm + o;   o += BT * T * UST;
    float* sprod = sm + o; o += BT;
    float* smask = sm + o; o += BT;
    o = align4(o);
    // -desc_attn state (kAttn)
    constexpr int LDT = kFastAttnLd;
    const int NWD = kAttn ? d.NW : 0, NWP = align4(NWD);
    float* tdd = sm + o;   o += NWD * LDT;          // d_d(desc_set), rows padded
    float* twd = sm + o;   o += NWD * LDT;          // desc_set . w_d^T
    float* ev = sm + o;    o += NWP;                // scores
    float* att = sm + o;   o += NWP;                // attention weights
    float* dhv = sm + o;   o += kAttn ? 64 : 0;     // d_h(h')
    float* vas = sm + o;   o += kAttn ? 64 : 0;     // d_attn.weight
    float* b1s = sm + o;   o += kAttn ? 64 : 0;     // y1.bias
    float* y1e = sm + o;   o += kAttn ? D * 64 : 0; // attended description half of y1, [class][64] like the y1d table
    float* wdh = sm + o;   o += kAttn ? 64 * 64 : 0; // d_h.weight
    int* segs = reinterpret_cast<int*>(sm + o); o += kAttn ? align4(D + 1) : 0;
    int* wcl = reinterpret_cast<int*>(sm + o);  o += NWP;
    o = align4(o);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + o);
#if defined(MMG_PHASE_TIMING) && !defined(MMG_CPU_EMU)
    unsigned* stamps = reinterpret_cast<unsigned*>(sm + o + 4);     // (T, 8 phases, 8 warps), debug build only
#endif

    const float* gimg = W.fwd_image;
    const float* simg = kRegSend ? gimg : simg_s;                 // sender section: read once from L2, or shared memory
    const bool train = kPerf ? true : in.train != 0;
    const bool binary = kPerf ? true : d.use_binary != 0;
    const bool own_draws = kPerf ? true : (train && in.u_sen == nullptr);   // on-device Philox stream (else: injected float64 uniforms)
    const float* corrupt_mask = kPerf ? nullptr : in.corrupt_mask;
    const bool ignore_receiver = kPerf ? false : d.ignore_receiver != 0;
    // flipout noise (model.py:233-234,467-468): never part of the kPerf instantiation
    const bool flip_sen = !kPerf && binary && d.flip_sen >= 0.f && (train || d.flipout_dev);
    const bool flip_rec = !kPerf && binary && d.flip_rec >= 0.f && (train || d.flipout_dev);
    unsigned long long fseed = 0, fiter = 0;
    if (flip_sen || flip_rec) { fseed = W.rng_state[0]; fiter = W.rng_state[1]; }

    // per-thread roles, fixed for the whole kernel (addresses hoisted out of the step loop)
    const int k4t = tid >> 2, p4 = tid & 3;                       // (hidden unit, K-quarter) pairs: GRU, W_hh, mix
    const int jo = tid / LPO, po = tid % LPO;                     // (message bit, K-slice) pairs: both message heads
    const int oh = tid >> 1, hh = tid & 1;                        // (head row, K-half) pairs: y1h / w_h rows
    const int sub = lane & 7, cw = lane >> 3;                     // (K-eighth, class within warp): class scores
    const float4* wc_t = reinterpret_cast<const float4*>(simg + im.wc) + tid;                 // + k4 * HI
    const float4* wb_t = reinterpret_cast<const float4*>(simg + im.wb) + jo * (HI / 4) + po;  // + i * LPO
    const float4* wih_t = reinterpret_cast<const float4*>(gimg + im.wih) + k4t * 4 + p4;      // + (g*MQ+q) * 256
    const float4* whd_t = reinterpret_cast<const float4*>(gimg + im.whead) + oh * 2 + hh;     // + q * 256
    const float4* wgh_t = reinterpret_cast<const float4*>(gimg + im.wgh) + k4t * 4 + p4;      // + (g*4+q) * 256
    const float4* ww_t = reinterpret_cast<const float4*>(gimg + im.ww) + jo * LPO + po;       // + q * M * LPO
    const float4* y1d_t = reinterpret_cast<const float4*>(kAttn ? y1e : img + im.y1d) + sub;  // + cls*16 (+8)
    const float* wdd_t = img + im.wdd + k4t * 4 + p4;                                          // + (d/4) * 256

    // ---- prologue ----------------------------------------------------------------------------------------------
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    MMG_SYNCTHREADS();
    pdl_wait(); pdl_launch_dependents();
    MMG_TRACE_AT(1, 0);
    if (tid == 0) tma_stage2(sm, gimg, (uint32_t)snd * 4u, sm + snd, gimg + im.b_ih, (uint32_t)(im.total - im.b_ih) * 4u, bar);
    // this thread's slice of every loop matrix -> registers (coalesced 16-byte loads from the L2-resident image)
    float4 rc[kRegSend ? M4 : 1], rb[kRegSend ? KB / 4 : 1], ri[3 * MQ], rh[8], rg[12], rw[KPT / 4];
    if constexpr (kRegSend) {
#pragma unroll
        for (int k4 = 0; k4 < M4; ++k4) rc[k4] = ldg4(wc_t + k4 * HI);
#pragma unroll
        for (int i = 0; i < KB / 4; ++i) rb[i] = ldg4(wb_t + i * LPO);
    }
#pragma unroll
    for (int i = 0; i < 3 * MQ; ++i) ri[i] = ldg4(wih_t + i * NT);
#pragma unroll
    for (int q = 0; q < 8; ++q) rh[q] = ldg4(whd_t + q * NT);
#pragma unroll
    for (int i = 0; i < 12; ++i) rg[i] = ldg4(wgh_t + i * NT);
#pragma unroll
    for (int q = 0; q < KPT / 4; ++q) rw[q] = ldg4(ww_t + q * M * LPO);
    float bdh_t = 0.f;
    if constexpr (kAttn) {
        // d_h.weight (64, 64) in shared memory, stored [q][row][half][4] so that thread (row, K-half) of P4 reads float4 q at
        // a unit-stride address (the register file is full: GRU, heads and sender already live there)
        if (tid < 128) bdh_t = ldg(aa.dh_b + oh);
        for (int idx = tid; idx < 64 * 16; idx += NT) {
            const int row = idx >> 4, c4 = idx & 15, half = c4 >> 3, q = c4 & 7;
            *reinterpret_cast<float4*>(wdh + ((q * 64 + row) * 2 + half) * 4) = ldg4(reinterpret_cast<const float4*>(aa.dh_w + (size_t)row * 64) + c4);
        }
        for (int idx = tid; idx < NWD * 16; idx += NT) {
            const int n = idx >> 4, c = idx & 15;
            // the table holds the word factor e^{2 d_d(word)} of tanh(w + h) = 1 - 2 / (e^{2w} e^{2h} + 1) (K_pre, attn_tanh)
            *reinterpret_cast<float4*>(tdd + n * LDT + 4 * c) = ldg4(reinterpret_cast<const float4*>(W.wtab_dd + (size_t)n * 64) + c);
            *reinterpret_cast<float4*>(twd + n * LDT + 4 * c) = ldg4(reinterpret_cast<const float4*>(W.wtab_wd + (size_t)n * 64) + c);
        }
        if (tid < 64) { vas[tid] = ldg(aa.va + tid); b1s[tid] = ldg(aa.b1 + tid); }
        for (int idx = tid; idx <= D; idx += NT) segs[idx] = W.seg[idx];
        for (int idx = tid; idx < NWD; idx += NT) wcl[idx] = W.wcls[idx];
    }
    // h_x rows of this CTA: split-K partials of K_pre summed in a fixed order + bias (model.py:195)
#pragma unroll
    for (int bt = 0; bt < BT; ++bt) {
        const int b = b0 + bt, n = tid;
        float v = 0.f;
        if (b < B) {
            v = ldg(b_img + n);
            for (int s = 0; s < W.hx_split; s += 8) {      // 8 partial loads in flight, added in slab order
                float pv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) pv[u] = s + u < W.hx_split ? W.hx_part[((size_t)(s + u) * B + b) * HI + n] : 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) v += pv[u];
            }
            W.h_x[(size_t)b * HI + n] = v;
        }
        hx[bt * HI + n] = v;
        if (n < HR) {
            float h = 0.f;
            if (b < B) {
                if (in.h0 != nullptr) h = in.h0[(size_t)b * HR + n];
                W.h_z[(size_t)b * HR + n] = h;                        // slot 0 = state entering step 0
            }
            hv[bt * HR + n] = h;
        }
        if (n < M) {
            win[bt * M + n] = d.first_rec;                            // model.py:786
            if (b < B) W.rec_feats[(size_t)b * M + n] = d.first_rec;  // slot 0
        }
        if (n == 0) { sprod[bt] = 1.f; smask[bt] = 1.f; if (b < B) W.stop_mask[b] = 1; }
        for (int dd = D + n; dd < DP; dd += NT) yv[bt * DP + dd] = -INFINITY;   // padded classes never win the softmax
    }
    if (n_conv_ctas < (int)gridDim.x) {     // side-role CTAs wait for every h_x row
        MMG_SYNCTHREADS();
        if (tid == 0) flag_arrive(W.tickets + 1);
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

    const float b_b_t = (simg + im.b_b)[jo], b_w_t = (img + im.b_w)[jo];
    const float hw0_v = (simg + im.hw0)[tid], b_code_t = (simg + im.b_code)[tid];
    const float bih_r = (img + im.b_ih)[k4t], bih_u = (img + im.b_ih)[HR + k4t], bih_n = (img + im.b_ih)[2 * HR + k4t];
    const float bhh_r = (img + im.b_hh)[k4t], bhh_u = (img + im.b_hh)[HR + k4t], bhh_n = (img + im.b_hh)[2 * HR + k4t];
    const float bhd_t = oh < HR ? 0.f : (img + im.b_wh)[oh - HR];
    const float ws_a = (img + im.ws)[lane], ws_b = (img + im.ws)[lane + 32];
    const float4 w2a = lds4(img + im.w2 + 4 * sub), w2b = lds4(img + im.w2 + 32 + 4 * sub);
    const float y2b = img[im.misc], s_bias = img[im.misc + 1];

    // W_hh . h + b_hh for the NEXT GRU step (needs only h): thread (k, quarter) -> three gate rows, 16 columns each
    auto gh_phase = [&]() {
        float4 acc[BT][3];
#pragma unroll
        for (int bt = 0; bt < BT; ++bt) { acc[bt][0] = zero4(); acc[bt][1] = zero4(); acc[bt][2] = zero4(); }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 w0 = rg[q], w1 = rg[4 + q], w2 = rg[8 + q];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const float4 h = lds4(hv + bt * HR + p4 * 16 + 4 * q);
                fma4(w0, h, acc[bt][0]); fma4(w1, h, acc[bt][1]); fma4(w2, h, acc[bt][2]);
            }
        }
#pragma unroll
        for (int bt = 0; bt < BT; ++bt) {
            const float r = group_sum<4>(hsum4(acc[bt][0])), u = group_sum<4>(hsum4(acc[bt][1])), n = group_sum<4>(hsum4(acc[bt][2]));
            if (p4 == 0) {
                ghv[bt * 3 * HR + k4t] = r + bhh_r;
                ghv[bt * 3 * HR + HR + k4t] = u + bhh_u;
                ghv[bt * 3 * HR + 2 * HR + k4t] = n + bhh_n;
            }
        }
    };

    gh_phase();                                   // gates' recurrent half for step 0 from the initial state
    MMG_SYNCTHREADS();
    MMG_TRACE_AT(1, 1);
    // prediction step of every example (model.py:893-896): the first step whose outgoing stop mask is 0, else the last one.
    // Every thread tracks it in registers (smask only changes between barriers), so the scores can be kept when they appear.
    int ystep_r[BT];
#pragma unroll
    for (int bt = 0; bt < BT; ++bt) ystep_r[bt] = -1;
    // the label of the example this warp finishes in the epilogue, fetched now (first touch of `target`: a DRAM round trip)
    long long tg_pref = 0;
    if (epilogue && warp < BT && b0 + warp < B) tg_pref = in.target[b0 + warp];

    for (int t = 0; t < T; ++t) {
        // ---- P1: sender hidden a = tanh(h_x + code_layer(w_prev)) (model.py:199-216): one thread per unit -----------
        MMG_STAMP(0);
        {
            float4 acc[BT];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) acc[bt] = zero4();
            if (t > 0) {
#pragma unroll
                for (int k4 = 0; k4 < M4; ++k4) {
                    float4 w;
                    if constexpr (kRegSend) w = rc[k4]; else w = wc_t[k4 * HI];
#pragma unroll
                    for (int bt = 0; bt < BT; ++bt) fma4(w, lds4(win + bt * M + 4 * k4), acc[bt]);
                }
            }
            const float base = (t == 0) ? hw0_v : b_code_t;
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const int b = b0 + bt;
                const float hwv = base + hsum4(acc[bt]), hxv = hx[bt * HI + tid];
                float pre = hxv + hwv;                                        // model.py:208-221
                if constexpr (!kPerf) { if (d.ignore_code) pre = hxv; else if (d.mix_prod) pre = hxv * hwv; }
                const float a = fast_tanh(pre);
                av[bt * HI + tid] = a;
                if (MMG_SAVE_OK(b)) {
                    W.a_s[((size_t)t * B + b) * HI + tid] = a;
                    if constexpr (!kPerf) { if (d.mix_prod) W.hw_s[((size_t)t * B + b) * HI + tid] = hwv; }
                    if (t > 0 && tid < M) W.code_in[((size_t)t * B + b) * M + tid] = win[bt * M + tid];
                }
            }
        }
        MMG_STAMP(1);
        MMG_SYNCTHREADS();
        // ---- P2: binary_layer + sender message (model.py:216-238, 814-820): LPO lanes per message bit ---------------
        {
            float4 acc[BT];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) acc[bt] = zero4();
#pragma unroll
            for (int i = 0; i < KB / 4; ++i) {
                float4 w;
                if constexpr (kRegSend) w = rb[i]; else w = wb_t[i * LPO];
#pragma unroll
                for (int bt = 0; bt < BT; ++bt) fma4(w, lds4(av + bt * HI + 4 * (i * LPO + po)), acc[bt]);
            }
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const float logit = b_b_t + group_sum<LPO>(hsum4(acc[bt]));
                if (po == 0) {
                    const int j = jo, b = b0 + bt;
                    const size_t row = (size_t)t * B + b;
                    float p = 0.f, zval;
                    if (binary) {
                        p = fast_sigmoid(logit);
                        if (train) {
                            if (!kPerf && b >= B) zval = 0.f;
                            else if (own_draws) zval = (uni[(bt * T + t) * UST + j] < p) ? 1.f : 0.f;
                            else zval = (in.u_sen[row * M + j] < (double)p) ? 1.f : 0.f;
                        } else {
                            zval = rintf(p);
                        }
                    } else {
                        zval = logit;
                    }
                    if (flip_sen && b < B)
                        zval = flip_bit(zval, d.flip_sen, in.u_flip_sen, row * M + j, fseed, fiter, t, 0, b + row_offset, j);
                    if (corrupt_mask != nullptr) zval = fabsf(zval - corrupt_mask[j]);
                    zv[bt * M + j] = zval;
                    if (MMG_SAVE_OK(b)) {
                        W.sen_feats[row * M + j] = zval;
                        if (binary) W.sen_probs[row * M + j] = p;
                    }
                }
            }
        }
        MMG_STAMP(2);
        MMG_SYNCTHREADS();
        // ---- P3: GRU step (model.py:340), gate order r,z,n; W_hh . h + b_hh is already in ghv ----------------------
        {
            float4 g[BT][3];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) { g[bt][0] = zero4(); g[bt][1] = zero4(); g[bt][2] = zero4(); }
#pragma unroll
            for (int q = 0; q < MQ; ++q) {
                const float4 w0 = ri[q], w1 = ri[MQ + q], w2 = ri[2 * MQ + q];
#pragma unroll
                for (int bt = 0; bt < BT; ++bt) {
                    const float4 z = lds4(zv + bt * M + p4 * (M / 4) + 4 * q);
                    fma4(w0, z, g[bt][0]); fma4(w1, z, g[bt][1]); fma4(w2, z, g[bt][2]);
                }
            }
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const float gh_r = ghv[bt * 3 * HR + k4t], gh_u = ghv[bt * 3 * HR + HR + k4t], gh_n = ghv[bt * 3 * HR + 2 * HR + k4t];
                const float hp = hv[bt * HR + k4t];
                const float gi_r = group_sum<4>(hsum4(g[bt][0])) + bih_r;
                const float gi_u = group_sum<4>(hsum4(g[bt][1])) + bih_u;
                const float gi_n = group_sum<4>(hsum4(g[bt][2])) + bih_n;
                const float r = fast_sigmoid(gi_r + gh_r);
                const float u = fast_sigmoid(gi_u + gh_u);
                const float nn = fast_tanh(gi_n + r * gh_n);
                const float hn = nn + u * (hp - nn);
                MMG_SYNCWARP();                      // every lane of the group has read hv[k] before it is overwritten
                if (p4 == 0) {
                    const int b = b0 + bt;
                    hv[bt * HR + k4t] = hn;
                    if (MMG_SAVE_OK(b)) {
                        const size_t row = (size_t)t * B + b;
                        float* gg = W.gates + row * 4 * HR;
                        gg[k4t] = r; gg[HR + k4t] = u; gg[2 * HR + k4t] = nn; gg[3 * HR + k4t] = gh_n;
                        W.h_z[((size_t)(t + 1) * B + b) * HR + k4t] = hn;
                    }
                }
            }
        }
        MMG_STAMP(3);
        MMG_SYNCTHREADS();
        // ---- P4: rows [y1.weight[:, :64] ; w_h.weight] . h' (model.py:432,452): 2 threads per row --------------------
        {
            float4 acc[BT];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) acc[bt] = zero4();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 w = rh[q];
#pragma unroll
                for (int bt = 0; bt < BT; ++bt) fma4(w, lds4(hv + bt * HR + hh * 32 + 4 * q), acc[bt]);
            }
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const float v = group_sum<2>(hsum4(acc[bt])) + bhd_t;
                if (hh == 0) {
                    const int b = b0 + bt;
                    if (oh < HR) {
                        y1hv[bt * HR + oh] = v;
                        if (MMG_SAVE_OK(b)) W.y1h[((size_t)t * B + b) * HR + oh] = v;
                    } else {
                        whv[bt * HR + oh - HR] = v;
                    }
                }
            }
            if constexpr (kAttn) {       // d_h(h') (model.py:359): 64 more rows, threads < 128 (whole warps)
                if (tid < 128) {
                    float4 a4 = zero4();
#pragma unroll
                    for (int q = 0; q < 8; ++q) fma4(lds4(wdh + (q * 128 + tid) * 4), lds4(hv + hh * 32 + 4 * q), a4);
                    const float v = group_sum<2>(hsum4(a4)) + bdh_t;
                    if (hh == 0) {
                        const float eh = attn_e2(v);        // kept as e^{2 d_h(h')}, see attn_tanh
                        dhv[oh] = eh;
                        if (train && MMG_SAVE_OK(b0)) W.dh_s[((size_t)t * B + b0) * 64 + oh] = eh;
                    }
                }
            }
            gh_phase();          // W_hh . h' for the NEXT step: independent of the head rows, interleaves with them
        }
        MMG_STAMP(4);
        MMG_SYNCTHREADS();
        if constexpr (kAttn) {
            // ---- PA: description attention (model.py:344-410) from the shared-memory word tables ------------------------
            {   // scores: 4 lanes per word (16 units each), 64 words per pass
                const float4 va0 = lds4(vas + 4 * p4), va1 = lds4(vas + 16 + 4 * p4), va2 = lds4(vas + 32 + 4 * p4), va3 = lds4(vas + 48 + 4 * p4);
                const float4 h0 = lds4(dhv + 4 * p4), h1 = lds4(dhv + 16 + 4 * p4), h2 = lds4(dhv + 32 + 4 * p4), h3 = lds4(dhv + 48 + 4 * p4);
                auto th = [](float ew, float eh) { return attn_tanh(ew, eh); };
                const float ba = ldg(aa.ba);
                for (int base = 0; base < NWD; base += NT / 4) {
                    const int n = base + k4t;
                    const bool ok = n < NWD;
                    const float* row = tdd + (ok ? n : 0) * LDT + 4 * p4;
                    const float4 w0 = lds4(row), w1 = lds4(row + 16), w2 = lds4(row + 32), w3 = lds4(row + 48);
                    float s0 = va0.x * th(w0.x, h0.x), s1 = va0.y * th(w0.y, h0.y), s2 = va0.z * th(w0.z, h0.z), s3 = va0.w * th(w0.w, h0.w);
                    s0 = fmaf(va1.x, th(w1.x, h1.x), s0); s1 = fmaf(va1.y, th(w1.y, h1.y), s1); s2 = fmaf(va1.z, th(w1.z, h1.z), s2); s3 = fmaf(va1.w, th(w1.w, h1.w), s3);
                    s0 = fmaf(va2.x, th(w2.x, h2.x), s0); s1 = fmaf(va2.y, th(w2.y, h2.y), s1); s2 = fmaf(va2.z, th(w2.z, h2.z), s2); s3 = fmaf(va2.w, th(w2.w, h2.w), s3);
                    s0 = fmaf(va3.x, th(w3.x, h3.x), s0); s1 = fmaf(va3.y, th(w3.y, h3.y), s1); s2 = fmaf(va3.z, th(w3.z, h3.z), s2); s3 = fmaf(va3.w, th(w3.w, h3.w), s3);
                    const float sc = group_sum<4>((s0 + s1) + (s2 + s3)) + ba;
                    if (ok && p4 == 0) ev[n] = sc;
                }
            }
            MMG_SYNCTHREADS();
            {   // softmax inside each class's word segment: 8 lanes per class
                const int l8 = tid & 7;
                for (int base = 0; base < D; base += NT / 8) {
                    const int dd = base + (tid >> 3);
                    const bool ok = dd < D;
                    const int s0 = ok ? segs[dd] : 0, s1 = ok ? segs[dd + 1] : 0;
                    float mx = -INFINITY;
                    for (int n = s0 + l8; n < s1; n += 8) mx = fmaxf(mx, ev[n]);
                    mx = group_max<8>(mx);
                    float se = 0.f;
                    for (int n = s0 + l8; n < s1; n += 8) se += fast_exp(ev[n] - mx);
                    se = group_sum<8>(se);
                    const float inv = fast_rcp(se);
                    for (int n = s0 + l8; n < s1; n += 8) {
                        const float a = fast_exp(ev[n] - mx) * inv;
                        att[n] = a;
                        if (train && MMG_SAVE_OK(b0)) W.attn[((size_t)t * B + b0) * NWD + n] = a;
                    }
                }
            }
            MMG_SYNCTHREADS();
            {   // attended description half of y1: thread = (class, float4 column group), two items in flight, table in L2
                const float4* tab = reinterpret_cast<const float4*>(W.wtab_y1);
                for (int base = 0; base < D * 16; base += 2 * NT) {
                    const int ia = base + tid, ib = ia + NT;
                    const bool oka = ia < D * 16, okb = ib < D * 16;
                    const int da = oka ? ia >> 4 : 0, ka = ia & 15, db = okb ? ib >> 4 : 0, kb = ib & 15;
                    const int a0 = segs[da], a1 = segs[da + 1], c0 = segs[db], c1 = segs[db + 1];
                    const int len = max(a1 - a0, c1 - c0);
                    float4 sa = lds4(b1s + 4 * ka), sb = lds4(b1s + 4 * kb);
#pragma unroll 8
                    for (int i = 0; i < len; ++i) {
                        const int na = min(a0 + i, a1 - 1), nb = min(c0 + i, c1 - 1);
                        const float4 wa = ldg4(tab + (size_t)na * 16 + ka), wb = ldg4(tab + (size_t)nb * 16 + kb);
                        const float fa = a0 + i < a1 ? att[na] : 0.f, fb = c0 + i < c1 ? att[nb] : 0.f;
                        sa.x = fmaf(fa, wa.x, sa.x); sa.y = fmaf(fa, wa.y, sa.y); sa.z = fmaf(fa, wa.z, sa.z); sa.w = fmaf(fa, wa.w, sa.w);
                        sb.x = fmaf(fb, wb.x, sb.x); sb.y = fmaf(fb, wb.y, sb.y); sb.z = fmaf(fb, wb.z, sb.z); sb.w = fmaf(fb, wb.w, sb.w);
                    }
                    if (oka) *reinterpret_cast<float4*>(y1e + da * 64 + 4 * ka) = sa;
                    if (okb) *reinterpret_cast<float4*>(y1e + db * 64 + 4 * kb) = sb;
                }
            }
            MMG_SYNCTHREADS();
        }
        // ---- P5: class scores y[d] = y2(relu(y1h + y1d[d])) (model.py:432-433), 8 lanes per class; the next step's
        //      W_hh . h' rides along (independent chain) ---------------------------------------------------------------
        {
            float4 ya[BT], yb[BT];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) { ya[bt] = lds4(y1hv + bt * HR + 4 * sub); yb[bt] = lds4(y1hv + bt * HR + 32 + 4 * sub); }
            float wm[BT];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) wm[bt] = -INFINITY;
            for (int c0 = 0; c0 < D; c0 += 4 * NW) {
                const int cls = c0 + warp * 4 + cw;
                const bool valid = cls < D;
                float4 r0 = zero4(), r1 = zero4();
                if (valid) { r0 = y1d_t[cls * 16]; r1 = y1d_t[cls * 16 + 8]; }
#pragma unroll
                for (int bt = 0; bt < BT; ++bt) {
                    float s0 = w2a.x * fmaxf(0.f, ya[bt].x + r0.x), s1 = w2a.y * fmaxf(0.f, ya[bt].y + r0.y);
                    float s2 = w2a.z * fmaxf(0.f, ya[bt].z + r0.z), s3 = w2a.w * fmaxf(0.f, ya[bt].w + r0.w);
                    s0 = fmaf(w2b.x, fmaxf(0.f, yb[bt].x + r1.x), s0); s1 = fmaf(w2b.y, fmaxf(0.f, yb[bt].y + r1.y), s1);
                    s2 = fmaf(w2b.z, fmaxf(0.f, yb[bt].z + r1.z), s2); s3 = fmaf(w2b.w, fmaxf(0.f, yb[bt].w + r1.w), s3);
                    const float s = group_sum<8>((s0 + s1) + (s2 + s3)) + y2b;
                    if (valid) wm[bt] = fmaxf(wm[bt], s);
                    if (sub == 0 && valid) {
                        yv[bt * DP + cls] = s;
                        const int b = b0 + bt;
                        if (MMG_SAVE_OK(b)) W.y[((size_t)t * B + b) * D + cls] = s;
                    }
                }
            }
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {    // maximum over this warp's classes (softmax shift for P6)
                float m = wm[bt];
                m = fmaxf(m, shfl_xor_f(m, 8));
                m = fmaxf(m, shfl_xor_f(m, 16));
                if (lane == 0) wmax[bt * 8 + warp] = m;
            }
        }
        MMG_STAMP(5);
        MMG_SYNCTHREADS();
        // ---- P6: q = softmax(y) (detached, model.py:441); h_w = tanh(w_h(h') + sum_d q_d wdd[d]) (442-452); the STOP
        //      head (model.py:414-429, 852) rides along ------------------------------------------------------------------
        {
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const int b = b0 + bt;
                const float4 m0 = lds4(wmax + bt * 8), m1 = lds4(wmax + bt * 8 + 4);
                const float mx = fmaxf(fmaxf(fmaxf(m0.x, m0.y), fmaxf(m0.z, m0.w)), fmaxf(fmaxf(m1.x, m1.y), fmaxf(m1.z, m1.w)));
                float acc0 = 0.f, acc1 = 0.f, se0 = 0.f, se1 = 0.f;
                // this thread's classes: c = dd0 + 4 u + p4; 8 independent exponentials in flight per 32-class chunk
                // (padded classes hold y = -inf and wdd = 0)
                for (int dd0 = 0; dd0 < DP; dd0 += 32) {
                    float e[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int c = dd0 + 4 * u + p4;
                        e[u] = c < DP ? fast_exp(yv[bt * DP + c] - mx) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; u += 2) {
                        const int g0 = (dd0 >> 2) + u;
                        se0 += e[u]; se1 += e[u + 1];
                        if (dd0 + 4 * u < DP) acc0 = fmaf(e[u], wdd_t[g0 * NT], acc0);
                        if (dd0 + 4 * u + 4 < DP) acc1 = fmaf(e[u + 1], wdd_t[(g0 + 1) * NT], acc1);
                    }
                }
                // STOP head: every warp evaluates the 64-wide dot product (no divergent block), one lane commits
                float sv = fmaf(ws_a, hv[bt * HR + lane], ws_b * hv[bt * HR + lane + 32]);
                const float se = group_sum<4>(se0 + se1);
                if constexpr (kAttn) {
                    // confidence-weighted ATTENDED description through w_d (model.py:441-452): sum over words of
                    // q_class(n) a_n (desc_set . w_d^T)[n].  The word weights are formed once per word (into the score
                    // buffer, no longer needed), then thread (unit k4t, quarter p4) takes words p4, p4 + 4, ...
                    const float inv_se = fast_rcp(se);
                    for (int n = tid; n < NWD; n += NT) {
                        const float qa = fast_exp(yv[bt * DP + wcl[n]] - mx) * att[n];
                        ev[n] = qa;
                        if (train && MMG_SAVE_OK(b)) W.qa[((size_t)t * B + b) * NWD + n] = qa * inv_se;
                    }
                    MMG_SYNCTHREADS();
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                    const float* tw = twd + k4t;
                    int n = p4;
                    for (; n + 12 < NWD; n += 16) {
                        a0 = fmaf(ev[n], tw[n * LDT], a0);
                        a1 = fmaf(ev[n + 4], tw[(n + 4) * LDT], a1);
                        a2 = fmaf(ev[n + 8], tw[(n + 8) * LDT], a2);
                        a3 = fmaf(ev[n + 12], tw[(n + 12) * LDT], a3);
                    }
                    for (; n < NWD; n += 4) a0 = fmaf(ev[n], tw[n * LDT], a0);
                    acc0 = (a0 + a1) + (a2 + a3); acc1 = 0.f;
                }
                const float acc = group_sum<4>(acc0 + acc1);
                sv = warp_sum(sv);
                const float inv = fast_rcp(se);
                if (train && MMG_SAVE_OK(b))
                    for (int c = tid; c < D; c += NT) W.q[((size_t)t * B + b) * D + c] = fast_exp(yv[bt * DP + c] - mx) * inv;
                if (p4 == 0) {
                    const float hw = fast_tanh(whv[bt * HR + k4t] + acc * inv);
                    hwv[bt * HR + k4t] = hw;
                    if (MMG_SAVE_OK(b)) W.h_w[((size_t)t * B + b) * HR + k4t] = hw;
                }
                if (tid == NT - 1) {
                    const size_t row = (size_t)t * B + b;
                    const float sp = fast_sigmoid(sv + s_bias);
                    float sbit;
                    if (train) {
                        if (!kPerf && b >= B) sbit = 0.f;
                        else if (own_draws) sbit = (uni[(bt * T + t) * UST + 2 * M] < sp) ? 1.f : 0.f;
                        else sbit = (in.u_stop[row] < (double)sp) ? 1.f : 0.f;
                    } else {
                        const float prod = (t == 0 || !d.s_prob_prod) ? sp : sprod[bt] * sp;
                        sprod[bt] = prod;
                        sbit = rintf(prod);
                    }
                    const float m = fminf(smask[bt], sbit);
                    smask[bt] = m;
                    if (MMG_SAVE_OK(b)) {
                        W.stop_feat[row] = sbit;
                        W.stop_prob[row] = sp;
                        W.stop_mask[(size_t)(t + 1) * B + b] = (unsigned char)(m != 0.f);
                    }
                }
            }
        }
        MMG_STAMP(6);
        MMG_SYNCTHREADS();
        // ---- P8: receiver message w(h_w) (model.py:454-475): LPO lanes per message bit -----------------------------
        if (epilogue) {
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                if (ystep_r[bt] < 0 && (t == T - 1 || (!d.fixed && smask[bt] == 0.f))) {
                    ystep_r[bt] = t;
                    for (int c = tid; c < DP; c += NT) ysel[bt * DP + c] = yv[bt * DP + c];
                }
            }
        }
        {
            float4 acc[BT];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) acc[bt] = zero4();
#pragma unroll
            for (int q = 0; q < KPT / 4; ++q) {
                const float4 w = rw[q];
#pragma unroll
                for (int bt = 0; bt < BT; ++bt) fma4(w, lds4(hwv + bt * HR + po * KPT + 4 * q), acc[bt]);
            }
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const float logit = b_w_t + group_sum<LPO>(hsum4(acc[bt]));
                if (po == 0) {
                    const int j = jo, b = b0 + bt;
                    const size_t row = (size_t)t * B + b;
                    float p = 0.f, wv;
                    if (binary) {
                        p = fast_sigmoid(logit);
                        if (train) {
                            if (!kPerf && b >= B) wv = 0.f;
                            else if (own_draws) wv = (uni[(bt * T + t) * UST + M + j] < p) ? 1.f : 0.f;
                            else wv = (in.u_rec[row * M + j] < (double)p) ? 1.f : 0.f;
                        } else {
                            wv = rintf(p);
                        }
                        if (flip_rec && b < B)
                            wv = flip_bit(wv, d.flip_rec, in.u_flip_rec, row * M + j, fseed, fiter, t, 1, b + row_offset, j);
                        if (ignore_receiver) wv = 0.f;
                    } else {
                        wv = logit;
                    }
                    win[bt * M + j] = wv;
                    if (MMG_SAVE_OK(b)) {
                        W.rec_feats[((size_t)(t + 1) * B + b) * M + j] = wv;
                        if (binary) W.rec_probs[row * M + j] = p;
                    }
                }
            }
        }
        MMG_STAMP(7);
        MMG_SYNCTHREADS();
    }
    MMG_TRACE_AT(1, 2);
    if (epilogue) {
        // ---- per-example results (get_rec_outp, model.py:879-904; log_softmax / NLL / loglikelihood / argmax, 1264-1275;
        //      top-k, 1333-1338; d nll / d outp): one warp per example, lanes over the classes.  This is the per-example half
        //      of K_stats; the batch statistics follow in the last CTA of K_baseline_fwd (they need the baselines).
        const float invB = 1.0f / (float)d.Bg;
        for (int bt = warp; bt < BT; bt += NW) {
            const int b = b0 + bt;
            if (b >= B) continue;
            const float* yy = ysel + bt * DP;
            const int tg = (int)tg_pref;               // bt == warp: BT <= 4 < 8 warps
            const float ytg = yy[tg];
            float mx = -INFINITY;
            for (int dd = lane; dd < D; dd += 32) mx = fmaxf(mx, yy[dd]);
            mx = warp_max(mx);
            float am = 3.0e9f, se = 0.f, rank = 0.f;
            for (int dd = lane; dd < D; dd += 32) {
                const float v = yy[dd];
                if (v == mx) am = fminf(am, (float)dd);          // first index of the maximum, like a serial `>` scan
                se += expf(v - mx);
                if (v > ytg) rank += 1.f;
            }
            am = -warp_max(-am);
            se = warp_sum(se);
            rank = warp_sum(rank);
            const float lse = mx + logf(se);
            const float lt = ytg - lse;
            for (int dd = lane; dd < D; dd += 32) {
                const float v = yy[dd];
                W.outp[(size_t)b * D + dd] = v;
                W.g_outp[(size_t)b * D + dd] = (expf(v - lse) - (dd == tg ? 1.f : 0.f)) * invB;   // d nll / d outp
            }
            if (lane == 0) {
                W.ystep[b] = ystep_r[bt];
                W.logs[b] = lt;
                W.argmax[b] = (int)am;
                W.hit[b] = ((int)rank < in.top_k) ? 1.f : 0.f;
            }
        }
    }
#if defined(MMG_PHASE_TIMING) && !defined(MMG_CPU_EMU)
    if (blockIdx.x == 0) for (int i = tid; i < T * 64; i += NT) reinterpret_cast<unsigned*>(W.g_sen_probs)[i] = stamps[i];
#endif
    MMG_TRACE_AT(1, 3);
}

// ---- backward ------------------------------------------------------------------------------------------------------
// One example per CTA.  CTAs [0, n_rec): receiver (BPTT); the rest: sender (no recurrence, model.py:807-811).
MMG_HOST_DEVICE int fast_bwd_rec_state_floats(int T, int M, int D) {
    // dlw (T,M) | dhw (T,64) | inj (T,64) | gat (T,5,64) | hws, y1hs (T,64 each) | dgh (2,192) | gv (64) | gout (DP) | dls (T) | barrier
    return T * (M + 4) + 2 * T * (kFastHr + 4) + 5 * T * kFastHr + 2 * T * kFastHr + 4 * T * kFastHr + 2 * 3 * kFastHr + kFastHr + align4(D) + align4(T) + 8;
}
MMG_HOST_DEVICE int fast_bwd_sen_state_floats(int T, int M) { return T * M + 2 * kFastHi + T * kFastHi + align4(T) + 8; }

// Loss coefficients for the fused backward: single rank -> the table K_baseline_fwd's last CTA derived (one load per thread);
// with peers -> the full prologue (wait for the peers' statistics, sum them in rank order, derive the coefficients).
MMG_DEVICE void fused_loss_coefs(const Dims& d, const mmg_config& cfg, const WsPtrs& W, const PeerView& pv, unsigned char* smem,
                                 LossCoef*& coef, float*& bas_scale) {
    if (pv.world > 1) { loss_prologue(d, cfg, W, pv, smem, coef, bas_scale); return; }
    coef = reinterpret_cast<LossCoef*>(smem);
    bas_scale = reinterpret_cast<float*>(coef + 3 * d.T);
    float* dst = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < 6 * d.T + 1; i += kFastBwdThreads) dst[i] = W.coefs[i];
    MMG_SYNCTHREADS();
}

// kFuse: the closed-form loss gradients of K_lossgrad (calculate_loss_binary / calculate_loss_bas, model.py:907-988, with the
// mask wiring of 1248-1262) are evaluated HERE, where the backward pass consumes them, from the batch statistics the forward
// sequence left in the workspace; the loss values come out of the same pass (per-CTA partials, summed in CTA order by the
// last CTA).  `loss_off`: float offset of the coefficient scratch inside the dynamic shared memory.
template <int M, bool kFuse>
MMG_GLOBAL void __launch_bounds__(kFastBwdThreads, 1)
k_exchange_bwd_fast(Dims d, WsPtrs W, const float* bin_w, const float* code_w, int n_rec_ctas, mmg_config cfg, PeerView pv,
                    int loss_off, const float* bs_w2) {
    constexpr int HI = kFastHi, HR = kFastHr, NT = kFastBwdThreads, M4 = M / 4;
    MMG_DYN_SMEM(smem_raw);
    float* sm = reinterpret_cast<float*>(smem_raw);
    const int tid = threadIdx.x;
    const int T = d.T, B = d.B, D = d.D;
    const bool binary = d.use_binary != 0;
    LossCoef* coef = nullptr;
    float* bas_scale = nullptr;
    double lacc[5] = {0, 0, 0, 0, 0};      // this thread's share of: binary_sen, binary_rec, binary_s, bas_rec, bas_sen

    if ((int)blockIdx.x >= n_rec_ctas) {
        // =================================== sender ==============================================================
        const int b = (int)blockIdx.x - n_rec_ctas, n = tid;
        float* dlz = sm;                                              // (T, M)
        pdl_wait(); pdl_launch_dependents();
        if constexpr (kFuse) fused_loss_coefs(d, cfg, W, pv, smem_raw + (size_t)loss_off * 4, coef, bas_scale);
        MMG_BSTAMP(0);
        float wb[M];                                                   // column n of binary_layer.weight
#pragma unroll
        for (int j = 0; j < M; ++j) wb[j] = ldg(bin_w + (size_t)j * HI + n);
        float* das0 = sm + T * M;                                     // (256) d a at t = 0
        float* cred = das0 + HI;                                      // (256 / M, M) partial sums
        float* as_s = cred + HI;                                      // (T, 256) saved tanh outputs of this example
        float* gbs_s = as_s + T * HI;                                 // (T) d loss / d bs[t, b]
#pragma unroll 4
        for (int t = 0; t < T; ++t) as_s[t * HI + n] = W.a_s[((size_t)t * B + b) * HI + n];
        for (int idx = tid; idx < T * M; idx += NT) {
            const int t = idx / M, j = idx % M;
            const size_t i = ((size_t)t * B + b) * M + j;
            const float p = W.sen_probs[i];
            float g;
            if constexpr (kFuse) {
                // sender REINFORCE + entropy term of step t (kind 0: baseline bs[t], mask s_masks[t], model.py:1258,1291)
                g = 0.f;
                const size_t row = (size_t)t * B + b;
                const bool m_in = binary && mask_at(d, W, t, b) != 0;
                const float lg = W.logs[b], bsv = W.bs[row];
                if (m_in) {
                    const LossCoef c0 = coef[t];
                    const float f = W.sen_feats[i];
                    const float l1 = logf(p + 1e-8f), l0 = logf(1.f - p + 1e-8f);
                    lacc[0] += (double)(-(lg - bsv) * c0.cA) * (double)(f * l1 + (1.f - f) * l0) + (double)c0.cE * (double)(p * l1 + (1.f - p) * l0);
                    g = binary_grad(p, f, (lg - bsv) * c0.cA, c0.cE);
                    if (j == 0) lacc[4] += (double)(bsv - lg) * (double)(bsv - lg) * bas_scale[0];
                }
                if (j == 0) { const float gb = m_in ? 2.f * (bsv - lg) * bas_scale[0] : 0.f; W.g_bs[row] = gb; gbs_s[t] = gb; }   // model.py:971-988
            } else {
                g = W.g_sen_probs[i];
                if (j == 0) gbs_s[t] = W.g_bs[(size_t)t * B + b];
            }
            const float dl = g * p * (1.f - p);                       // through the sigmoid (model.py:223)
            W.d_lz[i] = dl;
            dlz[idx] = dl;
        }
        MMG_SYNCTHREADS();
        MMG_BSTAMP(1);
        float dhx = 0.f;
#pragma unroll 2
        for (int t = 0; t < T; ++t) {
            const size_t i = ((size_t)t * B + b) * HI + n;
            const float a = as_s[t * HI + n];
            float4 a4 = zero4();
#pragma unroll
            for (int j4 = 0; j4 < M4; ++j4)
                fma4(lds4(dlz + t * M + 4 * j4), make_float4(wb[4 * j4], wb[4 * j4 + 1], wb[4 * j4 + 2], wb[4 * j4 + 3]), a4);
            const float dpre = hsum4(a4) * (1.f - a * a);             // through tanh (model.py:216)
            // d_as = gradient w.r.t. the code term h_w; sum: d pre, prod: d pre * h_x, ignore_code: none
            float das = dpre, dx = dpre;
            if (d.mix_prod && !d.ignore_code) { das = dpre * W.h_x[(size_t)b * HI + n]; dx = dpre * W.hw_s[i]; }
            if (d.ignore_code) das = 0.f;
            W.d_as[i] = das;
            if (t == 0) das0[n] = das;
            dhx += dx;                                                // h_x is shared by all steps (model.py:195)
        }
        W.dhx[(size_t)b * HI + n] = dhx;
        MMG_BSTAMP(2);
        // d code_layer-input at t = 0 (the code is sigmoid(code_bias) for every example, model.py:199-200):
        // dcode_part[b][j] = sum_n d_a[0][n] code_layer.weight[n][j]; K_wgrad sums over b and applies d sigmoid.
        MMG_SYNCTHREADS();
        {
            const int j = tid % M, part = tid / M;                    // NT / M parts, each HI * M / NT rows
            constexpr int RPP = HI * M / NT;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int r = 0; r < RPP; ++r) {
                const int nn = part * RPP + r;
                acc[r & 3] = fmaf(das0[nn], ldg(code_w + (size_t)nn * M + j), acc[r & 3]);
            }
            cred[part * M + j] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        }
        MMG_SYNCTHREADS();
        if (tid < M) {
            float v = 0.f;
#pragma unroll
            for (int p = 0; p < NT / M; ++p) v += cred[p * M + tid];
            W.dcode_part[(size_t)b * M + tid] = v;
        }
        // t-summed relu gradient of the sender-side baseline's hidden layer for this example:
        //   S[b][n] = linear2.weight[n] * sum_t g_bs[t, b] * (h1s[t, b, n] > 0)
        // h_x is shared by the T rows of an example, so the h_x half of d linear1.weight is S^T . h_x with K = B (and the bias
        // gradient its column sum).  Done HERE, on CTAs that are idle while the receiver CTAs walk the BPTT chain, so that K_wgrad
        // stages a plain matrix (the array U[b] of the forward pass is free by now and has the right shape).
        for (int n2 = tid; n2 < d.Hb; n2 += NT) {
            float acc = 0.f;
            for (int t0 = 0; t0 < T; t0 += 8) {           // 8 steps' loads in flight, added in step order
                float hv8[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) hv8[u] = t0 + u < T ? ldg(W.h1s + ((size_t)(t0 + u) * B + b) * d.Hb + n2) : 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) if (t0 + u < T && hv8[u] > 0.f) acc += gbs_s[t0 + u];
            }
            W.ubs[(size_t)b * d.Hb + n2] = acc * ldg(bs_w2 + n2);
        }
        if constexpr (kFuse) loss_partials(W, lacc);     // this CTA's share of the loss values (summed by K_wgrad's last CTA)
        MMG_BSTAMP(3);
        return;
    }

    // ===================================== receiver ==============================================================
    const FastBwdImage im = make_fast_bwd_image(M, D);
    const int b = blockIdx.x, lane = tid & 31;
    const int DP = align4(D);
    int o = im.total;
    constexpr int LDM = M + 4, LDH = HR + 4;      // padded row strides: rows t, t+1, t+2, t+3 are read by one quarter-warp
    float* dlw = sm + o;  o += T * LDM;
    float* dhw = sm + o;  o += T * LDH;
    float* inj = sm + o;  o += T * LDH;
    float* gat = sm + o;  o += 5 * T * HR;          // per step: r, u, n, gh_n, h_prev
    float* hws = sm + o;  o += T * HR;              // h_w of every step
    float* y1hs = sm + o; o += T * HR;              // y1h of every step (the prediction step is only known after the load)
    float* dgs = sm + o;  o += 4 * T * HR;          // per step: d r_pre, d u_pre, d n_pre, d gh_n (stored to HBM after the chain)
    float* dghv = sm + o; o += 2 * 3 * HR;
    float* gv = sm + o;   o += HR;
    float* gout = sm + o; o += DP;
    float* dls = sm + o;  o += align4(T);
    o += (o & 1);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + o);
    const float4* WwT4 = reinterpret_cast<const float4*>(sm + im.wwT);
    const float4* WhT4 = reinterpret_cast<const float4*>(sm + im.whT);
    const float4* W1hT4 = reinterpret_cast<const float4*>(sm + im.w1hT);
    const float* wsv = sm + im.ws;
    const float* w2 = sm + im.w2;
    const float* y1d = sm + im.y1d;

    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    MMG_SYNCTHREADS();
    pdl_wait(); pdl_launch_dependents();
    MMG_BSTAMP(0);
    // W_hh^T (48 KB) never touches shared memory: each thread keeps its 12 float4 of the BPTT mat-vec in registers
    if (tid == 0) tma_stage2(sm, W.bwd_image, (uint32_t)im.whhT * 4u, sm + im.ws, W.bwd_image + im.ws, (uint32_t)(im.total - im.ws) * 4u, bar);
    if constexpr (kFuse) fused_loss_coefs(d, cfg, W, pv, smem_raw + (size_t)loss_off * 4, coef, bas_scale);
    float4 rwhh[12];
    {
        const float4* src = reinterpret_cast<const float4*>(W.bwd_image + im.whhT) + (tid >> 2) * 4 + (tid & 3);
#pragma unroll
        for (int q = 0; q < 12; ++q) rwhh[q] = ldg4(src + q * HR * 4);
    }
    // ---- A0: everything the chain needs, into shared memory -----------------------------------------------------
    for (int idx = tid; idx < T * M; idx += NT) {
        const int t = idx / M, j = idx % M;
        const size_t i = ((size_t)t * B + b) * M + j;
        float dl = 0.f;
        if (binary) {
            const float p = W.rec_probs[i];
            float g;
            if constexpr (kFuse) {
                // receiver-message term of step t (kind 1: baseline br[t], mask s_masks[t+1], none at the last step;
                // model.py:1257,1284-1286)
                g = 0.f;
                if (t < T - 1 && mask_at(d, W, t + 1, b) != 0) {
                    const LossCoef c1 = coef[T + t];
                    const float lg = W.logs[b], brv = W.br[(size_t)t * B + b];
                    const float f = W.rec_feats[i + (size_t)B * M];
                    const float l1 = logf(p + 1e-8f), l0 = logf(1.f - p + 1e-8f);
                    lacc[1] += (double)(-(lg - brv) * c1.cA) * (double)(f * l1 + (1.f - f) * l0) + (double)c1.cE * (double)(p * l1 + (1.f - p) * l0);
                    g = binary_grad(p, f, (lg - brv) * c1.cA, c1.cE);
                }
            } else {
                g = W.g_rec_probs[i];
            }
            dl = g * p * (1.f - p);
        }
        W.d_lw[i] = dl;
        dlw[t * LDM + j] = dl;
    }
    // saved activations of this example: asynchronous 16-byte copies, all in flight together
    for (int idx = tid; idx < T * HR; idx += NT) {             // gates: T rows of 256 floats = 64 chunks each
        const int t = idx >> 6, c = idx & 63;
        cp_async16(gat + t * 5 * HR + 4 * c, W.gates + ((size_t)t * B + b) * 4 * HR + 4 * c);
    }
    for (int idx = tid; idx < T * 16 * 3; idx += NT) {         // h_prev, h_w, y1h: T rows of 64 floats = 16 chunks each
        const int which = idx / (T * 16), r = idx % (T * 16), t = r >> 4, c = r & 15;
        const size_t row = ((size_t)t * B + b) * HR + 4 * c;
        if (which == 0)      cp_async16(gat + t * 5 * HR + 4 * HR + 4 * c, W.h_z + row);     // slot t = state entering step t
        else if (which == 1) cp_async16(hws + t * HR + 4 * c, W.h_w + row);
        else                 cp_async16(y1hs + t * HR + 4 * c, W.y1h + row);
    }
    for (int t = tid; t < T; t += NT) {
        const size_t row = (size_t)t * B + b;
        const float sp = W.stop_prob[row];
        float g;
        if constexpr (kFuse) {
            // STOP bit (kind 2: baseline br[t], mask s_masks[t], adaptive length only; model.py:1256,1279-1280) and the
            // receiver-side baseline's MSE (model.py:971-988)
            g = 0.f;
            const bool m_in = binary && mask_at(d, W, t, b) != 0;
            const float lg = W.logs[b], brv = W.br[row];
            if (m_in) {
                if (!d.fixed) {
                    const LossCoef c2 = coef[2 * T + t];
                    const float sf = W.stop_feat[row];
                    const float l1 = logf(sp + 1e-8f), l0 = logf(1.f - sp + 1e-8f);
                    lacc[2] += (double)(-(lg - brv) * c2.cA) * (double)(sf * l1 + (1.f - sf) * l0) + (double)c2.cE * (double)(sp * l1 + (1.f - sp) * l0);
                    g = binary_grad(sp, sf, (lg - brv) * c2.cA, c2.cE);
                }
                lacc[3] += (double)(brv - lg) * (double)(brv - lg) * bas_scale[0];
            }
            W.g_br[row] = m_in ? 2.f * (brv - lg) * bas_scale[0] : 0.f;
        } else {
            g = W.g_stop_prob[row];
        }
        const float v = g * sp * (1.f - sp);
        W.d_ls[row] = v;
        dls[t] = v;
    }
    for (int dd = tid; dd < D; dd += NT) gout[dd] = W.g_outp[(size_t)b * D + dd];
    const int ys = W.ystep[b];
    MMG_BSTAMP(1);
    cp_async_wait_all();
#ifdef MMG_CPU_EMU
    MMG_SYNCTHREADS();
#endif
    mbar_wait(bar, 0);
    MMG_SYNCTHREADS();
    MMG_BSTAMP(2);
    const int k = tid >> 2, part = tid & 3;
    // ---- A1: d h_w for every step (t-parallel); class-score head at the prediction step ----------------------------
    for (int t = part; t < T; t += 4) {
        float4 a4 = zero4();
#pragma unroll
        for (int j4 = 0; j4 < M4; ++j4) fma4(WwT4[j4 * HR + k], lds4(dlw + t * LDM + 4 * j4), a4);
        const float acc = hsum4(a4);
        const size_t i = ((size_t)t * B + b) * HR + k;
        const float hw = hws[t * HR + k];
        const float v = acc * (1.f - hw * hw);                        // through tanh (model.py:452)
        W.d_hw[i] = v;
        dhw[t * LDH + k] = v;
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
    MMG_BSTAMP(3);
    // ---- A2: per-step injection into d h': W_h^T d_hw + s.weight * d_ls (+ W_1h^T G at the prediction step) --------
    for (int t = part; t < T; t += 4) {
        float4 a4 = zero4();
#pragma unroll
        for (int k4 = 0; k4 < HR / 4; ++k4) fma4(WhT4[k4 * HR + k], lds4(dhw + t * LDH + 4 * k4), a4);
        if (t == ys) {
#pragma unroll
            for (int k4 = 0; k4 < HR / 4; ++k4) fma4(W1hT4[k4 * HR + k], lds4(gv + 4 * k4), a4);
        }
        inj[t * LDH + k] = hsum4(a4) + wsv[k] * dls[t];
    }
    MMG_SYNCTHREADS();
    MMG_BSTAMP(4);
    // ---- B: the BPTT chain (h_z is never detached between steps, model.py:340) --------------------------------------
    float direct = 0.f, rec = 0.f;
    int buf = 0;
    for (int t = T - 1; t >= 0; --t) {
        const float* g = gat + t * 5 * HR;
        const float r = g[k], u = g[HR + k], nn = g[2 * HR + k], ghn = g[3 * HR + k], hp = g[4 * HR + k];
        const float dht = direct + rec + inj[t * LDH + k];
        // h' = n + u (h - n)
        const float du = dht * (hp - nn);
        const float dn = dht * (1.f - u);
        direct = dht * u;
        const float dn_pre = dn * (1.f - nn * nn);
        const float dghn = dn_pre * r;
        const float dr_pre = dn_pre * ghn * r * (1.f - r);
        const float du_pre = du * u * (1.f - u);
        if (part == 0) {
            float* ds = dgs + t * 4 * HR;
            ds[k] = dr_pre; ds[HR + k] = du_pre; ds[2 * HR + k] = dn_pre; ds[3 * HR + k] = dghn;
            float* dg = dghv + buf * 3 * HR;
            dg[k] = dr_pre; dg[HR + k] = du_pre; dg[2 * HR + k] = dghn;
        }
        if (t == 0) break;
        MMG_SYNCTHREADS();
        {   // d h_prev += W_hh^T . d gh: 4 threads per output, 48 reduction elements each
            const float* dg = dghv + buf * 3 * HR;
            float4 a4 = zero4();
#pragma unroll
            for (int q = 0; q < 12; ++q) fma4(rwhh[q], lds4(dg + part * 48 + 4 * q), a4);
            rec = group_sum<4>(hsum4(a4));
        }
        buf ^= 1;
    }
    // gate deltas of all steps -> HBM in one coalesced sweep (d gi = [r, u, n], d gh = [r, u, gh_n])
    MMG_SYNCTHREADS();
    for (int idx = tid; idx < T * 3 * HR; idx += NT) {
        const int t = idx / (3 * HR), c = idx % (3 * HR);
        const size_t row = (size_t)t * B + b;
        const float* ds = dgs + t * 4 * HR;
        W.dgi[row * 3 * HR + c] = ds[c];
        W.dgh[row * 3 * HR + c] = c < 2 * HR ? ds[c] : ds[c + HR];
    }
    if constexpr (kFuse) loss_partials(W, lacc);         // off the BPTT critical path
    MMG_BSTAMP(5);
    (void)lane;
}

}  // namespace mmg
