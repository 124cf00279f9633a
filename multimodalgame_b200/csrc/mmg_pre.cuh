// K_pre — per-iteration pre-pass (one launch):
//   role A (blocks [0, n_hx_tiles)): split-K tiles of the image-layer GEMM h_x = x . W_img^T (model.py:195).  x and
//          W_img are loop invariant inside one conversation, so the reference's T recomputations collapse to one.
//   role B (remaining blocks): (1) re-pack the loop weights from their state_dict layout into the kernel-layout
//          images the exchange kernels stage into shared memory with TMA bulk copies; (2) the loop-invariant
//          class halves  y1d = desc . y1.weight[:, Hr:]^T + y1.bias  and  wdd = desc . w_d.weight^T  (this is what
//          removes build_inp's B*D x (Hr+WV) cartesian product, model.py:412,519-551); (3) the step-0 sender code
//          term hw0 = code_layer(sigmoid(code_bias)) (model.py:199-200) and code_in[0] = sigmoid(code_bias).
#pragma once
#include "mmg_fast.cuh"
#include "mmg_umma.cuh"

namespace mmg {

MMG_DEVICE float packed_src(const float* W, int ld, int o, int r, int K) { return r < K ? ldg(W + (size_t)o * ld + r) : 0.f; }

// value of fwd-image element e (sections produced by dot products are skipped by the caller)
MMG_DEVICE float fwd_image_elem(const Dims& d, const FwdImage& im, const ParamPtrs& P, int e) {
    if (e < im.wb) { int q = e - im.wc; int r = (q / (d.Hi * 4)) * 4 + (q & 3), o = (q >> 2) % d.Hi;
        return packed_src(P.p[MMG_P_SEN_CODE_W], d.M, o, r, d.M); }
    if (e < im.b_code) { int q = e - im.wb; int r = (q / (d.M * 4)) * 4 + (q & 3), o = (q >> 2) % d.M;
        return packed_src(P.p[MMG_P_SEN_BIN_W], d.Ha, o, r, d.Ha); }
    if (e < im.hw0) { int i = e - im.b_code; return i < d.Hi ? ldg(P.p[MMG_P_SEN_CODE_B] + i) : 0.f; }
    if (e < im.b_b) return 0.f;  // hw0: dot role
    if (e < im.sender_end) { int i = e - im.b_b; return i < d.M ? ldg(P.p[MMG_P_SEN_BIN_B] + i) : 0.f; }
    if (e < im.whh) { int q = e - im.wih; int r = (q / (d.G3 * 4)) * 4 + (q & 3), o = (q >> 2) % d.G3;
        return packed_src(P.p[MMG_P_REC_RNN_WIH], d.M, o, r, d.M); }
    if (e < im.whead) { int q = e - im.whh; int r = (q / (d.G3 * 4)) * 4 + (q & 3), o = (q >> 2) % d.G3;
        return packed_src(P.p[MMG_P_REC_RNN_WHH], d.Hr, o, r, d.Hr); }
    if (e < im.ww) { int q = e - im.whead; int r = (q / (d.NH * 4)) * 4 + (q & 3), o = (q >> 2) % d.NH;
        if (o < d.Hr) return packed_src(P.p[MMG_P_REC_Y1_W] + d.y1_hcol, d.Hr + d.WV, o, r, d.Hr);
        if (o < 2 * d.Hr) return packed_src(P.p[MMG_P_REC_WH_W], d.Hr, o - d.Hr, r, d.Hr);
        if (o == 2 * d.Hr) return packed_src(P.p[MMG_P_REC_S_W], d.Hr, 0, r, d.Hr);
        return packed_src(P.p[MMG_P_REC_DH_W], d.Hr, o - 2 * d.Hr - 1, r, d.Hr); }      // d_h.weight (desc_attn)
    if (e < im.b_ih) { int q = e - im.ww; int r = (q / (d.M * 4)) * 4 + (q & 3), o = (q >> 2) % d.M;
        return packed_src(P.p[MMG_P_REC_W_W], d.Hr, o, r, d.Hr); }
    if (e < im.b_hh) { int i = e - im.b_ih; return i < d.G3 ? ldg(P.p[MMG_P_REC_RNN_BIH] + i) : 0.f; }
    if (e < im.b_head) { int i = e - im.b_hh; return i < d.G3 ? ldg(P.p[MMG_P_REC_RNN_BHH] + i) : 0.f; }
    if (e < im.w2) { int i = e - im.b_head;
        if (i < d.Hr) return 0.f;
        if (i < 2 * d.Hr) return ldg(P.p[MMG_P_REC_WH_B] + (i - d.Hr));
        if (i == 2 * d.Hr) return ldg(P.p[MMG_P_REC_S_B]);
        if (i < d.NH) return ldg(P.p[MMG_P_REC_DH_B] + (i - 2 * d.Hr - 1));
        return 0.f; }
    if (e < im.b_w) { int i = e - im.w2; return i < d.Hr ? ldg(P.p[MMG_P_REC_Y2_W] + i) : 0.f; }
    if (e < im.misc) { int i = e - im.b_w; return i < d.M ? ldg(P.p[MMG_P_REC_W_B] + i) : 0.f; }
    if (e < im.y1d) { int i = e - im.misc; return i == 0 ? ldg(P.p[MMG_P_REC_Y2_B]) : 0.f; }
    return 0.f;  // y1d / wdd: dot role
}

MMG_DEVICE float bwd_image_elem(const Dims& d, const BwdImage& im, const ParamPtrs& P, int e) {
    if (e < im.wwT) { int q = e - im.wbT; int r = (q / (d.Ha * 4)) * 4 + (q & 3), o = (q >> 2) % d.Ha;   // W(o=n, r=j) = bin_w[j][n]
        return r < d.M ? ldg(P.p[MMG_P_SEN_BIN_W] + (size_t)r * d.Ha + o) : 0.f; }
    if (e < im.headT) { int q = e - im.wwT; int r = (q / (d.Hr * 4)) * 4 + (q & 3), o = (q >> 2) % d.Hr;  // w_w[j][k]
        return r < d.M ? ldg(P.p[MMG_P_REC_W_W] + (size_t)r * d.Hr + o) : 0.f; }
    if (e < im.whhT) { int q = e - im.headT; int r = (q / (d.Hr * 4)) * 4 + (q & 3), o = (q >> 2) % d.Hr;
        if (r < d.Hr) return ldg(P.p[MMG_P_REC_WH_W] + (size_t)r * d.Hr + o);
        if (r < 2 * d.Hr) return ldg(P.p[MMG_P_REC_Y1_W] + (size_t)(r - d.Hr) * (d.Hr + d.WV) + d.y1_hcol + o);
        if (r < 2 * d.Hr + d.A) return ldg(P.p[MMG_P_REC_DH_W] + (size_t)(r - 2 * d.Hr) * d.Hr + o);   // d_h.weight^T (desc_attn)
        return 0.f; }
    if (e < im.ws) { int q = e - im.whhT; int r = (q / (d.Hr * 4)) * 4 + (q & 3), o = (q >> 2) % d.Hr;
        return r < d.G3 ? ldg(P.p[MMG_P_REC_RNN_WHH] + (size_t)r * d.Hr + o) : 0.f; }
    if (e < im.w2) { int i = e - im.ws; return i < d.Hr ? ldg(P.p[MMG_P_REC_S_W] + i) : 0.f; }
    if (e < im.y1d) { int i = e - im.w2; return i < d.Hr ? ldg(P.p[MMG_P_REC_Y2_W] + i) : 0.f; }
    return 0.f;  // y1d: dot role
}

MMG_GLOBAL void __launch_bounds__(kGemmThreads)
k_pre(Dims d, ParamPtrs P, WsPtrs W, ExchangeInputs in, int n_hx_tiles, int hx_kslice, int fast, int n_cls_tiles, int use_umma,
      int dyn_floats, const MMG_GRID_CONSTANT ImageTmaps tmaps) {
    pdl_wait();                 // PDL: the previous kernel of the stream has completed and flushed
    pdl_launch_dependents();    // let the next kernel's CTAs be scheduled behind this grid
    MMG_SHARED __attribute__((aligned(16))) float gs[kGemmSmemFloats];
    const int tid = threadIdx.x;
    MMG_TRACE_AT(0, 0);
    if (blockIdx.x == 0 && tid == 0) {
        W.tickets[1] = 0;      // "h_x rows ready" counter of the next forward kernel
        // Every exchange that consumes on-device draws (Bernoulli samples without injected uniforms, flipout noise) gets a
        // fresh Philox stream: the iteration counter advances HERE, before the conversation kernel of this launch sequence
        // reads it (this grid has fully completed by then), so forward-only callers (model.exchange, eval with -flipout_dev)
        // never see the same stream twice.
        const bool flips = (d.flip_sen >= 0.f && in.u_flip_sen == nullptr) || (d.flip_rec >= 0.f && in.u_flip_rec == nullptr);
        if ((in.train && in.u_sen == nullptr) || (flips && d.use_binary && (in.train || d.flipout_dev)))
            W.rng_state[1] += 1ull;
    }
#ifndef MMG_CPU_EMU
    if (use_umma && (int)blockIdx.x < n_hx_tiles) {
        // ---- role A on the tensor cores (tcgen05, 3xTF32): 128 hidden units x 64 batch rows x one K-slice per CTA --------------
        MMG_DYN_SMEM(umma_smem);
        const int nbt = cdiv(d.B, umma::kN), nmt = d.Hi / umma::kM;
        int t = blockIdx.x;
        const int bt = t % nbt; t /= nbt;
        const int mt = t % nmt; t /= nmt;
        const int s = t;
        const int k0 = s * hx_kslice, nk = min(hx_kslice, d.F - k0);
        umma::image_layer_tile(P.p[MMG_P_SEN_IMG_W], in.x, d.F, d.Hi, d.B, mt * umma::kM, bt * umma::kN, k0, nk,
                               W.hx_part + (size_t)s * d.B * d.Hi, umma_smem, use_umma == 2 ? &tmaps : nullptr);
        return;
    }
#endif
    if ((int)blockIdx.x < n_hx_tiles) {
        // ---- role A: h_x split-K tile -------------------------------------------------------------------
        const int ntn = cdiv(d.Hi, kTile), ntm = cdiv(d.B, kTile);
        int t = blockIdx.x;
        const int nt = t % ntn; t /= ntn;
        const int mt = t % ntm; t /= ntm;
        const int s = t;
        Operand A = {in.x, nullptr, nullptr, nullptr, d.F, 0, 0, 0, 0, OP_PLAIN};
        Operand Bo = {P.p[MMG_P_SEN_IMG_W], nullptr, nullptr, nullptr, d.F, 0, 0, 0, 0, OP_PLAIN};
        float acc[4][4];
        const int k0 = s * hx_kslice, k1 = min(d.F, k0 + hx_kslice);
        gemm_tile(A, Bo, d.B, d.Hi, mt * kTile, nt * kTile, k0, k1, acc, nullptr, gs);
        const int tx = tid % 16, ty = tid / 16;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int b = mt * kTile + ty * 4 + a;
            if (b >= d.B) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int n = nt * kTile + tx * 4 + c;
                if (n < d.Hi) W.hx_part[((size_t)s * d.B + b) * d.Hi + n] = acc[a][c];
            }
        }
        MMG_TRACE_AT(0, 6);
        return;
    }
    const FwdImage fim = make_fwd_image(d);
    const BwdImage bim = make_bwd_image(d);
    const FastFwdImage ffi = make_fast_fwd_image(d.M, d.D);
    const FastBwdImage fbi = make_fast_bwd_image(d.M, d.D);
    if ((int)blockIdx.x < n_hx_tiles + n_cls_tiles && d.A) {
        // ---- role C, -desc_attn: loop-invariant WORD tables (model.py:352 and the word halves of y1 / w_d) ------------
        //   wtab_y1[n][k] = desc_set[n] . y1.weight[k][:WV]   wtab_wd[n][k] = desc_set[n] . w_d.weight[k]
        //   wtab_dd[n][a] = e^{2 (d_d.bias[a] + desc_set[n] . d_d.weight[a])}: the word factor of tanh(d_d(word) + d_h(h)), see attn_tanh
        const int ntk = cdiv(d.Hr, kTile), nta = cdiv(d.A, kTile), per_m = 2 * ntk + nta;
        int t = (int)blockIdx.x - n_hx_tiles;
        const int mt = t / per_m;
        t %= per_m;
        const int which = t < ntk ? 0 : (t < 2 * ntk ? 1 : 2);
        const int nt = which == 0 ? t : (which == 1 ? t - ntk : t - 2 * ntk);
        const int N = which == 2 ? d.A : d.Hr;
        Operand A = {in.desc_set, nullptr, nullptr, nullptr, d.WV, 0, 0, 0, 0, OP_PLAIN};
        Operand Bo = which == 0 ? Operand{P.p[MMG_P_REC_Y1_W] + d.y1_dcol, nullptr, nullptr, nullptr, d.Hr + d.WV, 0, 0, 0, 0, OP_PLAIN}
                   : which == 1 ? Operand{P.p[MMG_P_REC_WD_W], nullptr, nullptr, nullptr, d.WV, 0, 0, 0, 0, OP_PLAIN}
                                : Operand{P.p[MMG_P_REC_DD_W], nullptr, nullptr, nullptr, d.WV, 0, 0, 0, 0, OP_PLAIN};
        float acc[4][4];
        if (rows_tile_ok(A.p, A.ld, d.WV, nullptr, 0, 0, Bo.p, Bo.ld) && rows_tile_smem_floats(d.WV) <= dyn_floats) {
            MMG_DYN_SMEM(dyn_raw);
            gemm_rows_tile(A.p, A.ld, d.WV, nullptr, 0, 0, Bo.p, Bo.ld, d.NW, N, mt * kTile, nt * kTile, acc, reinterpret_cast<float*>(dyn_raw));
        } else
        gemm_tile_deep(A, Bo, d.NW, N, mt * kTile, nt * kTile, 0, d.WV, acc, gs);
        const int tx = tid % 16, ty = tid / 16;
        float* out = which == 0 ? W.wtab_y1 : (which == 1 ? W.wtab_wd : W.wtab_dd);
        const int NP = align4(N);                   // rows padded to whole float4 groups, padding zero
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int n = mt * kTile + ty * 4 + a;
            if (n >= d.NW) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int k = nt * kTile + tx * 4 + c;
                if (k >= NP) continue;
                out[(size_t)n * NP + k] = k < N ? (which == 2 ? attn_e2(acc[a][c] + ldg(P.p[MMG_P_REC_DD_B] + k)) : acc[a][c]) : 0.f;
            }
        }
        return;
    }
    if ((int)blockIdx.x < n_hx_tiles + n_cls_tiles) {
        // ---- role C: loop-invariant class tables as two small GEMMs over the word-vector dimension -----------------
        //   y1d[dd][k] = y1.bias[k] + sum_v desc[dd][v] * y1.weight[k][Hr + v]     wdd[dd][k] = sum_v desc[dd][v] * w_d.weight[k][v]
        const int ntk = cdiv(d.Hr, kTile), ntd = cdiv(d.D, kTile);
        int t = (int)blockIdx.x - n_hx_tiles;
        const int which = t / (ntd * ntk);
        t %= ntd * ntk;
        const int nt = t % ntk, mt = t / ntk;
        Operand A = {in.desc, nullptr, nullptr, nullptr, d.WV, 0, 0, 0, 0, OP_PLAIN};
        Operand Bo = which == 0 ? Operand{P.p[MMG_P_REC_Y1_W] + d.Hr, nullptr, nullptr, nullptr, d.Hr + d.WV, 0, 0, 0, 0, OP_PLAIN}
                                : Operand{P.p[MMG_P_REC_WD_W], nullptr, nullptr, nullptr, d.WV, 0, 0, 0, 0, OP_PLAIN};
        float acc[4][4];
        MMG_TRACE_AT(0, 3);
        if (rows_tile_ok(A.p, A.ld, d.WV, nullptr, 0, 0, Bo.p, Bo.ld) && rows_tile_smem_floats(d.WV) <= dyn_floats) {
            MMG_DYN_SMEM(dyn_raw);
            gemm_rows_tile(A.p, A.ld, d.WV, nullptr, 0, 0, Bo.p, Bo.ld, d.D, d.Hr, mt * kTile, nt * kTile, acc, reinterpret_cast<float*>(dyn_raw));
            MMG_TRACE_AT(0, 6);
        } else
        gemm_tile_deep(A, Bo, d.D, d.Hr, mt * kTile, nt * kTile, 0, d.WV, acc, gs);
        const int tx = tid % 16, ty = tid / 16;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int dd = mt * kTile + ty * 4 + a;
            if (dd >= d.D) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int k = nt * kTile + tx * 4 + c;
                if (k >= d.Hr) continue;
                const int o = dd * d.Hr + k;
                if (which == 0) {
                    const float v = acc[a][c] + ldg(P.p[MMG_P_REC_Y1_B] + k);
                    if (fast) {
                        W.fwd_image[ffi.y1d + o] = v;
                        W.bwd_image[fbi.y1d + ((dd >> 2) * d.Hr + k) * 4 + (dd & 3)] = v;
                    } else {
                        W.fwd_image[fim.y1d + o] = v;
                        W.bwd_image[bim.y1d + o] = v;
                    }
                } else {
                    if (fast) W.fwd_image[ffi.wdd + ((dd >> 2) * d.Hr + k) * 4 + (dd & 3)] = acc[a][c];
                    else      W.fwd_image[fim.wdd + o] = acc[a][c];
                }
            }
        }
        MMG_TRACE_AT(0, 5);
        return;
    }
    // ---- role B: images + the step-0 code term -------------------------------------------------------------
    const int nblk = gridDim.x - n_hx_tiles - n_cls_tiles, blk = blockIdx.x - n_hx_tiles - n_cls_tiles;
    const int gthreads = nblk * kGemmThreads, gtid = blk * kGemmThreads + tid;
    if (d.A && blk == 0 && tid == 0) {          // word segments of the classes (model.py:372-376)
        int start = 0;
        for (int dd = 0; dd < d.D; ++dd) {
            W.seg[dd] = start;
            const int n = in.desc_set_lens[dd];
            for (int i = 0; i < n && start + i < d.NW; ++i) W.wcls[start + i] = dd;
            start += n;
            if (start > d.NW) start = d.NW;
        }
        W.seg[d.D] = start;
    }
    const bool ff = (fast & 1) != 0, fb = (fast & 2) != 0;      // forward / backward image in the fast-path format
    if (ff) {
        for (int e = gtid; e < ffi.y1d; e += gthreads) {
            if (e >= ffi.hw0 && e < ffi.b_b) continue;
            W.fwd_image[e] = fast_fwd_image_elem(d, ffi, P, e);
        }
    } else {
        for (int e = gtid; e < fim.y1d; e += gthreads) {
            if (e >= fim.hw0 && e < fim.b_b) continue;
            W.fwd_image[e] = fwd_image_elem(d, fim, P, e);
        }
    }
    if (fb) for (int e = gtid; e < fbi.y1d; e += gthreads) W.bwd_image[e] = fast_bwd_image_elem(d, fbi, P, e);
    else    for (int e = gtid; e < bim.y1d; e += gthreads) W.bwd_image[e] = bwd_image_elem(d, bim, P, e);
    // dot role: one warp per output, lanes along the reduction
    const int lane = tid & 31, gwarp = gtid >> 5, nwarps = gthreads >> 5;
    const int n_y1d = d.D * d.Hr;
    const int n_y1d_w = d.A ? 0 : n_y1d;      // -desc_attn: the class tables are replaced by word tables, the sections stay zero
    const int n_out = 2 * n_y1d + d.Hi + d.M;
    for (int o = 2 * n_y1d + gwarp; o < n_out; o += nwarps) {
        if (o < 2 * n_y1d + d.Hi) {     // hw0[n] = code_layer.bias[n] + sum_j sigmoid(code_bias[j]) * code_layer.weight[n][j]
            const int n = o - 2 * n_y1d;
            float s = 0.f;
            for (int j = lane; j < d.M; j += 32)
                s = fmaf(sigmoidf_(ldg(P.p[MMG_P_SEN_CODE_BIAS] + j)), ldg(P.p[MMG_P_SEN_CODE_W] + (size_t)n * d.M + j), s);
            s = warp_sum(s);
            if (lane == 0) W.fwd_image[(ff ? ffi.hw0 : fim.hw0) + n] = s + ldg(P.p[MMG_P_SEN_CODE_B] + n);
        } else {                               // code_in[0][b][j] = sigmoid(code_bias[j]) for every row b
            const int j = o - 2 * n_y1d - d.Hi;
            const float c0 = sigmoidf_(ldg(P.p[MMG_P_SEN_CODE_BIAS] + j));
            for (int b = lane; b < d.B; b += 32) W.code_in[(size_t)b * d.M + j] = c0;
        }
    }
    if (d.mix_mou && d.ignore_code) {
        // -sender_mix mou -ignore_code: after step 0 the code is sigmoid(code_bias_mou) for every example (model.py:201-205):
        // hw0m[n] = code_layer(that code)[n], and the code rows of steps >= 1 (the code_layer weight-gradient operand)
        for (int n = gwarp; n < d.Hi; n += nwarps) {
            float s = 0.f;
            for (int j = lane; j < d.M; j += 32)
                s = fmaf(sigmoidf_(ldg(P.p[MMG_P_SEN_CODE_BIAS_MOU] + j)), ldg(P.p[MMG_P_SEN_CODE_W] + (size_t)n * d.M + j), s);
            s = warp_sum(s);
            if (lane == 0) W.fwd_image[fim.hw0m + n] = s + ldg(P.p[MMG_P_SEN_CODE_B] + n);
        }
        for (long long e = (long long)d.B * d.M + gtid; e < (long long)d.R * d.M; e += gthreads)
            W.code_in[e] = sigmoidf_(ldg(P.p[MMG_P_SEN_CODE_BIAS_MOU] + (int)(e % d.M)));
    }
    // pad tails of dot sections (keep the images fully defined for the bulk copies)
    const int dpad = ((d.D + 3) / 4) * 4;
    if (ff || fb) {
        for (int e = gtid; e < (dpad - d.D) * d.Hr; e += gthreads) {
            const int dd = d.D + e / d.Hr, k = e % d.Hr;
            if (ff) W.fwd_image[ffi.wdd + ((dd >> 2) * d.Hr + k) * 4 + (dd & 3)] = 0.f;
            if (fb) W.bwd_image[fbi.y1d + ((dd >> 2) * d.Hr + k) * 4 + (dd & 3)] = 0.f;
        }
    }
    if (!ff) {
        for (int e = gtid; e < fim.total; e += gthreads) {
            if ((e >= fim.hw0 + d.Hi && e < fim.hw0m) || (e >= fim.hw0m + ((d.mix_mou && d.ignore_code) ? d.Hi : 0) && e < fim.b_b) ||
                (e >= fim.y1d + n_y1d_w && e < fim.wdd) || (e >= fim.wdd + n_y1d_w))
                W.fwd_image[e] = 0.f;
        }
    }
    if (!fb) for (int e = gtid + bim.y1d + n_y1d_w; e < bim.total; e += gthreads) W.bwd_image[e] = 0.f;
    MMG_TRACE_AT(0, 4);
}

}  // namespace mmg
