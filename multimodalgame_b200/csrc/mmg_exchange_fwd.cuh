// K_exchange_fwd — the whole T-step conversation (model.py:801-867) in ONE persistent kernel.
//
// One CTA owns BT examples for all T steps; examples are independent in the forward pass (desc and parameters are
// shared), so no inter-CTA communication is needed.  The kernel-layout weight image written by K_pre (~190 KB at the
// headline configuration) is staged ONCE into shared memory by the TMA unit (cp.async.bulk + mbarrier) and reused
// for every step: no weight byte is re-read from HBM/L2 inside the recurrence and there is no host round trip for
// sampling (the reference does 3 D2H + 3 H2D copies per step, model.py:225-231,418-420,458-464).
//
// Per step:  sender   a = tanh(h_x + code_layer(w_prev));  p_z = sigmoid(binary_layer(a));  z ~ Bernoulli(p_z)
//            receiver h' = GRU(z, h);  s_p = sigmoid(s(h'));  y[d] = y2(relu(y1h(h') + y1d[d]))   (y1 split in halves)
//                     q = softmax(y);  h_w = tanh(w_h(h') + sum_d q_d wdd[d]);  p_w = sigmoid(w(h_w));  w ~ Bernoulli
// Mat-vecs read "packed" weights (see mmg_layout.h) as float4, split along the reduction dimension across warps when
// the output is narrower than the CTA, and meet in shared memory.
// -desc_attn (model.py:344-410): between the heads and the class scores the receiver attends over the words of every
// class description; the per-word linear halves are loop-invariant tables written by K_pre, see the block in the loop.
#pragma once
#include "mmg_kernels.cuh"

namespace mmg {

enum { kLoopThreads = 256 };

struct SplitPlan { int Nr, KS; };
MMG_HOST_DEVICE SplitPlan make_split(int N, int K4) {
    SplitPlan s;
    s.Nr = round_up(N, 32);
    s.KS = kLoopThreads / s.Nr;
    if (s.KS < 1) s.KS = 1;
    if (s.KS > K4) s.KS = K4;
    if (s.KS < 1) s.KS = 1;
    return s;
}

// part[(ks * BT + bt) * Nr + n] = sum_{k4 in slice ks} W4[k4][n] . v[bt][4 k4 .. 4 k4 + 3]
template <int BT, bool kGlobalW>
MMG_DEVICE void split_matvec(const float* Wp, int N, int K4, const float* v, int ldv, float* part, SplitPlan sp) {
    const float4* W4 = reinterpret_cast<const float4*>(Wp);
    for (int idx = threadIdx.x; idx < sp.Nr * sp.KS; idx += kLoopThreads) {
        const int n = idx % sp.Nr, ks = idx / sp.Nr;
        if (n >= N) continue;
        const int kb = (ks * K4) / sp.KS, ke = ((ks + 1) * K4) / sp.KS;
        float acc[BT];
#pragma unroll
        for (int bt = 0; bt < BT; ++bt) acc[bt] = 0.f;
#pragma unroll 4
        for (int k4 = kb; k4 < ke; ++k4) {
            const float4 w = kGlobalW ? ldg4(W4 + (size_t)k4 * N + n) : W4[(size_t)k4 * N + n];
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const float4 x = *reinterpret_cast<const float4*>(v + bt * ldv + 4 * k4);
                acc[bt] = fmaf(w.x, x.x, acc[bt]);
                acc[bt] = fmaf(w.y, x.y, acc[bt]);
                acc[bt] = fmaf(w.z, x.z, acc[bt]);
                acc[bt] = fmaf(w.w, x.w, acc[bt]);
            }
        }
#pragma unroll
        for (int bt = 0; bt < BT; ++bt) part[(ks * BT + bt) * sp.Nr + n] = acc[bt];
    }
}
template <int BT>
MMG_DEVICE float gather_part(const float* part, SplitPlan sp, int bt, int n) {
    float s = 0.f;
    for (int ks = 0; ks < sp.KS; ++ks) s += part[(ks * BT + bt) * sp.Nr + n];
    return s;
}

// Bernoulli draw `u < p` (model.py:227): injected float64 uniform, or the on-device Philox stream.
MMG_DEVICE float draw_bit(const double* u, size_t uidx, float p, unsigned long long seed, unsigned long long iter,
                          unsigned stream_id, unsigned row, unsigned col) {
    if (u != nullptr) return (u[uidx] < (double)p) ? 1.f : 0.f;
    float r[4];
    philox_uniform4(seed, iter, stream_id, row * 65536u + (col >> 2), r);
    return (r[col & 3] < p) ? 1.f : 0.f;
}

// flipout (model.py:554-568): |bit - 1[u < p_flip]| with a second uniform per bit, drawn right after the message draw.
// `agent`: 0 sender, 1 receiver (separate halves of Philox stream 4 t + 3).
MMG_DEVICE float flip_bit(float bit, float p_flip, const double* u, size_t uidx, unsigned long long seed,
                          unsigned long long iter, int t, int agent, unsigned row, unsigned col) {
    float m;
    if (u != nullptr) m = (u[uidx] < (double)p_flip) ? 1.f : 0.f;
    else {
        float r[4];
        philox_uniform4(seed, iter, t * 4 + 3, row * 65536u + (agent ? 32768u : 0u) + (col >> 2), r);
        m = (r[col & 3] < p_flip) ? 1.f : 0.f;
    }
    return fabsf(bit - m);
}

MMG_HOST_DEVICE int fwd_state_floats(const Dims& d, int BT) {
    // hx, a, win, z, pz, h, head, yv, q, hwr, partA, partB, misc
    const int HiP = align4(d.Hi), HaP = align4(d.Ha), MP = d.M4 * 4, HrP = d.Hr4 * 4;
    int n = BT * (HiP + HaP) + BT * MP * 3 + BT * HrP * 2 + BT * align4(d.NH) + BT * align4(d.D) * 2;
    int pmax = round_up(d.Hi, 32);
    if (round_up(d.G3, 32) > pmax) pmax = round_up(d.G3, 32);
    if (round_up(d.NH, 32) > pmax) pmax = round_up(d.NH, 32);
    if (pmax < kLoopThreads) pmax = kLoopThreads;
    n += 2 * BT * pmax;
    n += 4 * BT + 8;   // sprod, active, barrier
    if (d.A) n += BT * (2 * align4(d.NW) + align4(d.A) + d.D * HrP + (kLoopThreads / d.Hr4) * HrP) + align4(d.A) + HrP;
    return n;
}

template <int BT>
MMG_GLOBAL void __launch_bounds__(kLoopThreads, 1)
k_exchange_fwd(Dims d, WsPtrs W, ExchangeInputs in, const float* b_img, int sender_smem, int row_offset, AttnArgs aa) {
    MMG_DYN_SMEM(smem_raw);
    float* sm = reinterpret_cast<float*>(smem_raw);
    const FwdImage im = make_fwd_image(d);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = blockIdx.x * BT;
    const int HiP = align4(d.Hi), HaP = align4(d.Ha), MP = d.M4 * 4, HrP = d.Hr4 * 4, NHP = align4(d.NH), DP = align4(d.D);
    const int img0 = sender_smem ? 0 : im.sender_end;
    // ---- shared-memory carve-up ---------------------------------------------------------------------------
    float* img = sm - img0;                       // img[off] valid for off >= img0
    int o = im.total - img0;
    float* hx = sm + o;   o += BT * HiP;
    float* av = sm + o;   o += BT * HaP;
    float* win = sm + o;  o += BT * MP;
    float* zv = sm + o;   o += BT * MP;
    float* pv = sm + o;   o += BT * MP;
    float* hv = sm + o;   o += BT * HrP;
    float* hwr = sm + o;  o += BT * HrP;
    float* head = sm + o; o += BT * NHP;
    float* yv = sm + o;   o += BT * DP;
    float* qv = sm + o;   o += BT * DP;
    int pmax = round_up(d.Hi, 32);
    if (round_up(d.G3, 32) > pmax) pmax = round_up(d.G3, 32);
    if (round_up(d.NH, 32) > pmax) pmax = round_up(d.NH, 32);
    if (pmax < kLoopThreads) pmax = kLoopThreads;
    float* partA = sm + o; o += BT * pmax;
    float* partB = sm + o; o += BT * pmax;
    float* sprod = sm + o; o += BT;
    float* smask = sm + o; o += BT;
    o = align4(o);
    const int NWP = align4(d.NW), AP = align4(d.A);
    float* ev = sm + o;   o += d.A ? BT * NWP : 0;                    // attention scores, later q_d(n) * a_n
    float* att = sm + o;  o += d.A ? BT * NWP : 0;                    // attention weights
    float* dhv = sm + o;  o += d.A ? BT * AP : 0;                     // d_h(h')
    const int DH = d.D * HrP;
    float* y1e = sm + o;  o += d.A ? BT * DH : 0;                     // attended description half of y1, rows padded to float4
    float* vas = sm + o;  o += d.A ? AP : 0;                          // d_attn.weight, zero padded
    float* b1s = sm + o;  o += d.A ? HrP : 0;                         // y1.bias
    float* wpart = sm + o; o += d.A ? BT * (kLoopThreads / (HrP >> 2)) * HrP : 0;   // word-slice partials of the w_d mix
    o += (o & 1);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + o);

    const float* gimg = W.fwd_image;
    const float* Wc = sender_smem ? img + im.wc : gimg + im.wc;
    const float* Wb = sender_smem ? img + im.wb : gimg + im.wb;
    const float* b_code = sender_smem ? img + im.b_code : gimg + im.b_code;
    const float* hw0 = sender_smem ? img + im.hw0 : gimg + im.hw0;
    const float* hw0m = sender_smem ? img + im.hw0m : gimg + im.hw0m;
    const bool code_const = d.mix_mou && d.ignore_code;     // the code term after step 0 is a constant vector (model.py:201-205)
    const float* b_b = sender_smem ? img + im.b_b : gimg + im.b_b;
    const float* Wih = img + im.wih;
    const float* Whh = img + im.whh;
    const float* Whead = img + im.whead;
    const float* Ww = img + im.ww;
    const float* b_ih = img + im.b_ih;
    const float* b_hh = img + im.b_hh;
    const float* b_head = img + im.b_head;
    const float* w2 = img + im.w2;
    const float* b_w = img + im.b_w;
    const float* y1d = img + im.y1d;
    const float* wdd = img + im.wdd;

    const SplitPlan sp_code = make_split(d.Hi, d.M4), sp_bin = make_split(d.M, d.Ha4), sp_gi = make_split(d.G3, d.M4),
                    sp_gh = make_split(d.G3, d.Hr4), sp_head = make_split(d.NH, d.Hr4), sp_w = make_split(d.M, d.Hr4);
    const bool train = in.train != 0;
    const bool binary = d.use_binary != 0;

    // ---- prologue -------------------------------------------------------------------------------------------
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    MMG_SYNCTHREADS();
    pdl_wait(); pdl_launch_dependents();
    if (tid == 0) tma_stage(sm, gimg + img0, (uint32_t)(im.total - img0) * 4u, bar);
    // h_x rows of this CTA: sum the split-K partials of K_pre in a fixed order, add the bias
    for (int idx = tid; idx < BT * d.Hi; idx += kLoopThreads) {
        const int bt = idx / d.Hi, n = idx % d.Hi, b = b0 + bt;
        float v = 0.f;
        if (b < d.B) {
            v = ldg(b_img + n);
            for (int s = 0; s < W.hx_split; ++s) v += W.hx_part[((size_t)s * d.B + b) * d.Hi + n];
            W.h_x[(size_t)b * d.Hi + n] = v;
        }
        hx[bt * HiP + n] = v;
    }
    for (int idx = tid; idx < BT * HrP; idx += kLoopThreads) {
        const int bt = idx / HrP, k = idx % HrP, b = b0 + bt;
        float v = 0.f;
        if (b < d.B && k < d.Hr) {
            if (in.h0 != nullptr) v = in.h0[(size_t)b * d.Hr + k];
            W.h_z[(size_t)b * d.Hr + k] = v;                         // slot 0 = state entering step 0
        }
        hv[idx] = v;
        hwr[idx] = 0.f;
    }
    for (int idx = tid; idx < BT * MP; idx += kLoopThreads) {
        const int bt = idx / MP, j = idx % MP, b = b0 + bt;
        const float v = (j < d.M) ? d.first_rec : 0.f;                 // model.py:786
        win[idx] = v; zv[idx] = 0.f; pv[idx] = 0.f;
        if (b < d.B && j < d.M) W.rec_feats[(size_t)b * d.M + j] = v;  // slot 0
    }
    if (tid < BT) { sprod[tid] = 1.f; smask[tid] = 1.f; if (b0 + tid < d.B) W.stop_mask[b0 + tid] = 1; }
    for (int idx = tid; idx < BT * HaP; idx += kLoopThreads) av[idx] = 0.f;
    if (d.A) {
        for (int idx = tid; idx < AP; idx += kLoopThreads) vas[idx] = idx < d.A ? ldg(aa.va + idx) : 0.f;
        for (int idx = tid; idx < HrP; idx += kLoopThreads) b1s[idx] = idx < d.Hr ? ldg(aa.b1 + idx) : 0.f;
        for (int idx = tid; idx < BT * AP; idx += kLoopThreads) dhv[idx] = 0.f;
    }
#ifdef MMG_CPU_EMU
    MMG_SYNCTHREADS();
#endif
    mbar_wait(bar, 0);
    MMG_SYNCTHREADS();

    unsigned long long seed = 0, iter = 0;
    if ((train && in.u_sen == nullptr) || d.flip_sen >= 0.f || d.flip_rec >= 0.f) { seed = W.rng_state[0]; iter = W.rng_state[1]; }

    for (int t = 0; t < d.T; ++t) {
        // ---- S1: sender code term W_code . w_prev (model.py:207); step 0 uses the constant hw0 (199-200) ----
        if (t > 0 && !code_const) {
            if (sender_smem) split_matvec<BT, false>(Wc, d.Hi, d.M4, win, MP, partA, sp_code);
            else             split_matvec<BT, true>(Wc, d.Hi, d.M4, win, MP, partA, sp_code);
            MMG_SYNCTHREADS();
        }
        // ---- S2: a = tanh(h_x + h_w) (model.py:216) -------------------------------------------------------
        for (int idx = tid; idx < BT * d.Hi; idx += kLoopThreads) {
            const int bt = idx / d.Hi, n = idx % d.Hi, b = b0 + bt;
            const float hw = (t == 0) ? hw0[n] : (code_const ? hw0m[n] : b_code[n] + gather_part<BT>(partA, sp_code, bt, n));
            const float hxv = hx[bt * HiP + n];
            if (d.mix_mou) {      // a = tanh([h_x ; h_w ; h_x - h_w ; h_x * h_w]) (model.py:211-213, 219-221)
                const float a0 = tanhf(hxv), a1 = tanhf(hw), a2 = tanhf(hxv - hw), a3 = tanhf(hxv * hw);
                float* ab = av + bt * HaP;
                ab[n] = a0; ab[d.Hi + n] = a1; ab[2 * d.Hi + n] = a2; ab[3 * d.Hi + n] = a3;
                if (b < d.B) {
                    float* as = W.a_s + ((size_t)t * d.B + b) * d.Ha;
                    as[n] = a0; as[d.Hi + n] = a1; as[2 * d.Hi + n] = a2; as[3 * d.Hi + n] = a3;
                    W.hw_s[((size_t)t * d.B + b) * d.Hi + n] = hw;
                }
                continue;
            }
            const float a = tanhf(d.ignore_code ? hxv : (d.mix_prod ? hxv * hw : hxv + hw));     // model.py:208-221
            av[bt * HaP + n] = a;
            if (b < d.B) {
                W.a_s[((size_t)t * d.B + b) * d.Hi + n] = a;
                if (d.mix_prod) W.hw_s[((size_t)t * d.B + b) * d.Hi + n] = hw;
            }
        }
        if (t > 0 && !code_const) {
            for (int idx = tid; idx < BT * d.M; idx += kLoopThreads) {
                const int bt = idx / d.M, j = idx % d.M, b = b0 + bt;
                if (b < d.B) W.code_in[((size_t)t * d.B + b) * d.M + j] = win[bt * MP + j];
            }
        }
        MMG_SYNCTHREADS();
        // ---- S3: binary_layer (model.py:216) -----------------------------------------------------------------
        if (sender_smem) split_matvec<BT, false>(Wb, d.M, d.Ha4, av, HaP, partA, sp_bin);
        else             split_matvec<BT, true>(Wb, d.M, d.Ha4, av, HaP, partA, sp_bin);
        MMG_SYNCTHREADS();
        // ---- S4: sender message (model.py:222-238, 814-820) ------------------------------------------------
        for (int idx = tid; idx < BT * d.M; idx += kLoopThreads) {
            const int bt = idx / d.M, j = idx % d.M, b = b0 + bt;
            const float logit = b_b[j] + gather_part<BT>(partA, sp_bin, bt, j);
            float p = 0.f, zval;
            const size_t row = (size_t)t * d.B + b;
            if (binary) {
                p = sigmoidf_(logit);
                if (train) zval = (b < d.B) ? draw_bit(in.u_sen, row * d.M + j, p, seed, iter, t * 4 + 0, b + row_offset, j) : 0.f;
                else       zval = rintf(p);
            } else {
                zval = logit;
            }
            if (binary && d.flip_sen >= 0.f && (train || d.flipout_dev) && b < d.B)
                zval = flip_bit(zval, d.flip_sen, in.u_flip_sen, row * d.M + j, seed, iter, t, 0, b + row_offset, j);
            if (in.corrupt_mask != nullptr) zval = fabsf(zval - in.corrupt_mask[j]);
            zv[bt * MP + j] = zval;
            pv[bt * MP + j] = p;
            if (b < d.B) {
                W.sen_feats[row * d.M + j] = zval;
                if (binary) W.sen_probs[row * d.M + j] = p;
            }
        }
        MMG_SYNCTHREADS();
        // ---- S5: GRU mat-vecs (model.py:340) -------------------------------------------------------------------
        split_matvec<BT, false>(Wih, d.G3, d.M4, zv, MP, partA, sp_gi);
        split_matvec<BT, false>(Whh, d.G3, d.Hr4, hv, HrP, partB, sp_gh);
        MMG_SYNCTHREADS();
        // ---- S6: GRU gates, gate order r,z,n; h' = n + u (h - n) --------------------------------------------
        for (int idx = tid; idx < BT * d.Hr; idx += kLoopThreads) {
            const int bt = idx / d.Hr, k = idx % d.Hr, b = b0 + bt;
            const float gi_r = b_ih[k] + gather_part<BT>(partA, sp_gi, bt, k);
            const float gi_u = b_ih[d.Hr + k] + gather_part<BT>(partA, sp_gi, bt, d.Hr + k);
            const float gi_n = b_ih[2 * d.Hr + k] + gather_part<BT>(partA, sp_gi, bt, 2 * d.Hr + k);
            const float gh_r = b_hh[k] + gather_part<BT>(partB, sp_gh, bt, k);
            const float gh_u = b_hh[d.Hr + k] + gather_part<BT>(partB, sp_gh, bt, d.Hr + k);
            const float gh_n = b_hh[2 * d.Hr + k] + gather_part<BT>(partB, sp_gh, bt, 2 * d.Hr + k);
            const float r = sigmoidf_(gi_r + gh_r);
            const float u = sigmoidf_(gi_u + gh_u);
            const float nn = tanhf(gi_n + r * gh_n);
            const float hp = hv[bt * HrP + k];
            const float hn = nn + u * (hp - nn);
            hv[bt * HrP + k] = hn;
            if (b < d.B) {
                const size_t row = (size_t)t * d.B + b;
                float* g = W.gates + row * 4 * d.Hr;
                g[k] = r; g[d.Hr + k] = u; g[2 * d.Hr + k] = nn; g[3 * d.Hr + k] = gh_n;
                W.h_z[((size_t)(t + 1) * d.B + b) * d.Hr + k] = hn;
            }
        }
        MMG_SYNCTHREADS();
        // ---- S7: stacked heads [y1.weight[:, :Hr] ; w_h ; s] . h' (model.py:414,432,452) --------------------
        split_matvec<BT, false>(Whead, d.NH, d.Hr4, hv, HrP, partA, sp_head);
        MMG_SYNCTHREADS();
        // ---- S8a: finalize heads, STOP bit (model.py:414-429, 852) --------------------------------------------
        for (int idx = tid; idx < BT * d.NH; idx += kLoopThreads) {
            const int bt = idx / d.NH, oo = idx % d.NH, b = b0 + bt;
            const float v = b_head[oo] + gather_part<BT>(partA, sp_head, bt, oo);
            head[bt * NHP + oo] = v;
            const size_t row = (size_t)t * d.B + b;
            if (oo < d.Hr) {
                if (b < d.B) W.y1h[row * d.Hr + oo] = v;
            } else if (oo > 2 * d.Hr) {        // d_h(h') of the description attention (model.py:359), rides in the stacked heads
                const int a = oo - 2 * d.Hr - 1;
                const float eh = attn_e2(v);            // kept as e^{2 d_h(h')}, see attn_tanh
                dhv[bt * AP + a] = eh;
                if (train && b < d.B) W.dh_s[row * d.A + a] = eh;
            } else if (oo == 2 * d.Hr) {
                const float sp = sigmoidf_(v);
                float sbit;
                if (train) {
                    sbit = (b < d.B) ? draw_bit(in.u_stop, row, sp, seed, iter, t * 4 + 1, b + row_offset, 0) : 0.f;
                } else {
                    const float prod = (t == 0 || !d.s_prob_prod) ? sp : sprod[bt] * sp;
                    sprod[bt] = prod;
                    sbit = rintf(prod);
                }
                const float m = fminf(smask[bt], sbit);
                smask[bt] = m;
                if (b < d.B) {
                    W.stop_feat[row] = sbit;
                    W.stop_prob[row] = sp;
                    W.stop_mask[(size_t)(t + 1) * d.B + b] = (unsigned char)(m != 0.f);
                }
            }
        }
        MMG_SYNCTHREADS();
        if (d.A) {
            // ---- -desc_attn (model.py:344-410): additive attention of h' over the words, softmax inside each class's
            //      segment.  The per-word halves (d_d(desc_set), desc_set . y1^T, desc_set . w_d^T) are loop invariant
            //      tables written by K_pre (rows padded to float4), so the (B, NW, WV) broadcasts of the reference never
            //      exist.  The tables live in L2; every loop below keeps several independent 16-byte loads in flight per
            //      thread and reduces inside 4- or 8-lane groups.
            const int q4 = tid & 3, g4 = tid >> 2, A4 = AP >> 2;
            // scores (model.py:366): 4 lanes per word, two words per lane group in flight
            for (int base = 0; base < BT * d.NW; base += kLoopThreads / 2) {
                const int oa = base + g4, ob = oa + kLoopThreads / 4;
                const bool oka = oa < BT * d.NW, okb = ob < BT * d.NW;
                const int bta = oka ? oa / d.NW : 0, na = oka ? oa % d.NW : 0;
                const int btb = okb ? ob / d.NW : 0, nb = okb ? ob % d.NW : 0;
                const float4* rowa = reinterpret_cast<const float4*>(W.wtab_dd + (size_t)na * AP);
                const float4* rowb = reinterpret_cast<const float4*>(W.wtab_dd + (size_t)nb * AP);
                const float4* dha = reinterpret_cast<const float4*>(dhv + bta * AP);
                const float4* dhb = reinterpret_cast<const float4*>(dhv + btb * AP);
                const float4* va4 = reinterpret_cast<const float4*>(vas);
                float sa = 0.f, sb = 0.f;
#pragma unroll 4
                for (int i = q4; i < A4; i += 4) {
                    const float4 wa = ldg4(rowa + i), wb = ldg4(rowb + i), ha = dha[i], hb = dhb[i], v = va4[i];
                    sa = fmaf(v.x, attn_tanh(wa.x, ha.x), sa); sa = fmaf(v.y, attn_tanh(wa.y, ha.y), sa);
                    sa = fmaf(v.z, attn_tanh(wa.z, ha.z), sa); sa = fmaf(v.w, attn_tanh(wa.w, ha.w), sa);
                    sb = fmaf(v.x, attn_tanh(wb.x, hb.x), sb); sb = fmaf(v.y, attn_tanh(wb.y, hb.y), sb);
                    sb = fmaf(v.z, attn_tanh(wb.z, hb.z), sb); sb = fmaf(v.w, attn_tanh(wb.w, hb.w), sb);
                }
                sa = group_sum<4>(sa); sb = group_sum<4>(sb);
                if (q4 == 0) {
                    const float ba = ldg(aa.ba);
                    if (oka) ev[bta * NWP + na] = sa + ba;
                    if (okb) ev[btb * NWP + nb] = sb + ba;
                }
            }
            MMG_SYNCTHREADS();
            for (int base = 0; base < BT * d.D; base += kLoopThreads / 8) {                    // segment softmax  model.py:372-381
                const int o2 = base + (tid >> 3), l8 = tid & 7;
                const bool ok = o2 < BT * d.D;
                const int bt = ok ? o2 / d.D : 0, dd = ok ? o2 % d.D : 0, b = b0 + bt;
                const int s0 = ok ? W.seg[dd] : 0, s1 = ok ? W.seg[dd + 1] : 0;
                float mx = -INFINITY;
                for (int n = s0 + l8; n < s1; n += 8) mx = fmaxf(mx, ev[bt * NWP + n]);
                mx = group_max<8>(mx);
                float se = 0.f;
                for (int n = s0 + l8; n < s1; n += 8) se += expf(ev[bt * NWP + n] - mx);
                se = group_sum<8>(se);
                const float inv = 1.f / se;
                for (int n = s0 + l8; n < s1; n += 8) {
                    const float a = expf(ev[bt * NWP + n] - mx) * inv;
                    att[bt * NWP + n] = a;
                    if (train && b < d.B) W.attn[((size_t)t * d.B + b) * d.NW + n] = a;
                }
            }
            MMG_SYNCTHREADS();
            const int K4 = HrP >> 2;
            // y1 . [attended desc ; .] (model.py:383-410,432): thread = (class, float4 column group), two items in flight
            for (int base = 0; base < BT * d.D * K4; base += 2 * kLoopThreads) {
                const int ia = base + tid, ib = ia + kLoopThreads, tot = BT * d.D * K4;
                const bool oka = ia < tot, okb = ib < tot;
                const int ra = oka ? ia : 0, rb = okb ? ib : 0;
                const int bta = ra / (d.D * K4), da = (ra % (d.D * K4)) / K4, ka = ra % K4;
                const int btb = rb / (d.D * K4), db = (rb % (d.D * K4)) / K4, kb = rb % K4;
                const int a0 = W.seg[da], a1 = W.seg[da + 1], c0 = W.seg[db], c1 = W.seg[db + 1];
                const int len = max(a1 - a0, c1 - c0);
                float4 sa = *reinterpret_cast<const float4*>(b1s + 4 * ka), sb = *reinterpret_cast<const float4*>(b1s + 4 * kb);
                const float4* taba = reinterpret_cast<const float4*>(W.wtab_y1) + ka;
                const float4* tabb = reinterpret_cast<const float4*>(W.wtab_y1) + kb;
#pragma unroll 4
                for (int i = 0; i < len; ++i) {
                    const int na = min(a0 + i, a1 - 1), nb = min(c0 + i, c1 - 1);      // clamped: the load is always legal
                    const float4 wa = ldg4(taba + (size_t)na * K4), wb = ldg4(tabb + (size_t)nb * K4);
                    const float fa = a0 + i < a1 ? att[bta * NWP + na] : 0.f, fb = c0 + i < c1 ? att[btb * NWP + nb] : 0.f;
                    sa.x = fmaf(fa, wa.x, sa.x); sa.y = fmaf(fa, wa.y, sa.y); sa.z = fmaf(fa, wa.z, sa.z); sa.w = fmaf(fa, wa.w, sa.w);
                    sb.x = fmaf(fb, wb.x, sb.x); sb.y = fmaf(fb, wb.y, sb.y); sb.z = fmaf(fb, wb.z, sb.z); sb.w = fmaf(fb, wb.w, sb.w);
                }
                if (oka) *reinterpret_cast<float4*>(y1e + bta * DH + da * HrP + 4 * ka) = sa;
                if (okb) *reinterpret_cast<float4*>(y1e + btb * DH + db * HrP + 4 * kb) = sb;
            }
            MMG_SYNCTHREADS();
        }
        // ---- S8b: class scores y[d] = y2(relu(y1h + y1d[d])) (model.py:432-433) — one warp per (example, class)
        for (int pair = warp; pair < BT * d.D; pair += kLoopThreads / 32) {
            const int bt = pair / d.D, dd = pair % d.D, b = b0 + bt;
            const float* yd = d.A ? y1e + bt * DH + dd * HrP : y1d + dd * d.Hr;
            float s = 0.f;
            for (int k = lane; k < d.Hr; k += 32)
                s = fmaf(w2[k], fmaxf(0.f, head[bt * NHP + k] + yd[k]), s);
            s = warp_sum(s);
            if (lane == 0) {
                s += img[im.misc];
                yv[bt * DP + dd] = s;
                if (b < d.B) W.y[((size_t)t * d.B + b) * d.D + dd] = s;
            }
        }
        MMG_SYNCTHREADS();
        // ---- S9: q = softmax(y) (model.py:441) — one warp per example -------------------------------------------
        for (int bt = warp; bt < BT; bt += kLoopThreads / 32) {
            const int b = b0 + bt;
            float mx = -INFINITY;
            for (int dd = lane; dd < d.D; dd += 32) mx = fmaxf(mx, yv[bt * DP + dd]);
            mx = warp_max(mx);
            float se = 0.f;
            for (int dd = lane; dd < d.D; dd += 32) se += expf(yv[bt * DP + dd] - mx);
            se = warp_sum(se);
            const float inv = 1.f / se;
            for (int dd = lane; dd < d.D; dd += 32) {
                const float qq = expf(yv[bt * DP + dd] - mx) * inv;
                qv[bt * DP + dd] = qq;
                if (b < d.B) W.q[((size_t)t * d.B + b) * d.D + dd] = qq;
            }
        }
        MMG_SYNCTHREADS();
        if (d.A) {
            // word weights of the confidence-weighted description: q_class(n) * a_n (model.py:441-449); the rows are also the
            // A operand of the wd = (q a) . desc_set GEMM tiles that run beside the baselines (w_d gradient)
            for (int idx = tid; idx < BT * d.NW; idx += kLoopThreads) {
                const int bt = idx / d.NW, n = idx % d.NW, b = b0 + bt;
                const float v = qv[bt * DP + W.wcls[n]] * att[bt * NWP + n];
                ev[bt * NWP + n] = v;
                if (train && b < d.B) W.qa[((size_t)t * d.B + b) * d.NW + n] = v;
            }
            MMG_SYNCTHREADS();
            // partial sums of (q a) . (desc_set . w_d^T): thread = (word slice, float4 column group)
            const int K4 = HrP >> 2, NS = kLoopThreads / K4;
            const int k4 = tid % K4, sl = tid / K4;
            if (sl < NS) {
                const float4* tab = reinterpret_cast<const float4*>(W.wtab_wd) + k4;
                for (int bt = 0; bt < BT; ++bt) {
                    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
                    for (int n = sl; n < d.NW; n += NS) {
                        const float4 w = ldg4(tab + (size_t)n * K4);
                        const float a = ev[bt * NWP + n];
                        s.x = fmaf(a, w.x, s.x); s.y = fmaf(a, w.y, s.y); s.z = fmaf(a, w.z, s.z); s.w = fmaf(a, w.w, s.w);
                    }
                    *reinterpret_cast<float4*>(wpart + (bt * NS + sl) * HrP + 4 * k4) = s;
                }
            }
            MMG_SYNCTHREADS();
        }
        // ---- S10: h_w = tanh(w_h(h') + w_d(q . desc)) (model.py:442-452); wd = q . desc saved for the backward --
        for (int idx = tid; idx < BT * (d.Hr + d.WV); idx += kLoopThreads) {
            if (idx < BT * d.Hr) {
                const int bt = idx / d.Hr, k = idx % d.Hr, b = b0 + bt;
                float s = 0.f;
                if (d.A) { const int NS = kLoopThreads / (HrP >> 2); for (int sl = 0; sl < NS; ++sl) s += wpart[(bt * NS + sl) * HrP + k]; }
                else for (int dd = 0; dd < d.D; ++dd) s = fmaf(qv[bt * DP + dd], wdd[dd * d.Hr + k], s);
                const float hw = tanhf(head[bt * NHP + d.Hr + k] + s);
                hwr[bt * HrP + k] = hw;
                if (b < d.B) W.h_w[((size_t)t * d.B + b) * d.Hr + k] = hw;
            } else if (train && !d.A) {        // -desc_attn: wd comes from GEMM tiles beside the baselines
                const int i2 = idx - BT * d.Hr;
                const int bt = i2 / d.WV, v = i2 % d.WV, b = b0 + bt;
                if (b < d.B) {
                    float s = 0.f;
                    for (int dd = 0; dd < d.D; ++dd) s = fmaf(qv[bt * DP + dd], ldg(in.desc + (size_t)dd * d.WV + v), s);
                    W.wd[((size_t)t * d.B + b) * d.WV + v] = s;
                }
            }
        }
        MMG_SYNCTHREADS();
        // ---- S11: w(h_w) (model.py:454) ------------------------------------------------------------------------
        split_matvec<BT, false>(Ww, d.M, d.Hr4, hwr, HrP, partA, sp_w);
        MMG_SYNCTHREADS();
        // ---- S12: receiver message (model.py:455-475) ---------------------------------------------------------
        for (int idx = tid; idx < BT * d.M; idx += kLoopThreads) {
            const int bt = idx / d.M, j = idx % d.M, b = b0 + bt;
            const float logit = b_w[j] + gather_part<BT>(partA, sp_w, bt, j);
            const size_t row = (size_t)t * d.B + b;
            float p = 0.f, wv;
            if (binary) {
                p = sigmoidf_(logit);
                if (train) wv = (b < d.B) ? draw_bit(in.u_rec, row * d.M + j, p, seed, iter, t * 4 + 2, b + row_offset, j) : 0.f;
                else       wv = rintf(p);
                if (d.flip_rec >= 0.f && (train || d.flipout_dev) && b < d.B)
                    wv = flip_bit(wv, d.flip_rec, in.u_flip_rec, row * d.M + j, seed, iter, t, 1, b + row_offset, j);
                if (d.ignore_receiver) wv = 0.f;
            } else {
                wv = logit;
            }
            win[bt * MP + j] = wv;
            pv[bt * MP + j] = p;
            if (b < d.B) {
                W.rec_feats[((size_t)(t + 1) * d.B + b) * d.M + j] = wv;
                if (binary) W.rec_probs[row * d.M + j] = p;
            }
        }
        MMG_SYNCTHREADS();
    }
}

}  // namespace mmg
