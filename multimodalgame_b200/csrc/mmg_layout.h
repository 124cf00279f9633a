// Internal layouts shared by host code and kernels: problem dimensions, the flat parameter layout, the
// workspace map and the kernel-layout weight images staged into shared memory by the exchange kernels.
#pragma once
#include "../../include/mmg_b200.h"
#include "mmg_platform.cuh"

namespace mmg {

struct Dims {
    int B, Bg, F, Hi, M, Hr, D, WV, Hb, T;
    int R;        // T * B rows, step-major
    int G3;       // 3 * Hr
    int M4, Hr4, Hi4;   // ceil(x / 4): float4 groups along a reduction dimension
    int Ha, Ha4;        // width of the sender's hidden vector a: Hi, or 4 Hi with sender_mix mou ([h_x ; h_w ; h_x - h_w ; h_x * h_w])
    int NH;       // stacked head rows: [y1.weight[:, h cols] ; w_h.weight ; s.weight ; d_h.weight (desc_attn)] = 2*Hr + 1 + A
    int use_binary, fixed, s_prob_prod, ignore_receiver;
    float first_rec;
    float flip_sen, flip_rec;   // flipout probabilities, < 0 = off (model.py:233-234,467-468)
    int flipout_dev;
    int mix_prod, ignore_code;  // sender hidden: tanh(h_x * h_w) instead of the sum / tanh(h_x) alone (model.py:208-221)
    int mix_mou;                // sender hidden: the 4-block mixture (model.py:211-213, 219-221); with ignore_code the code term after
                                // step 0 is code_layer(sigmoid(code_bias_mou)) (model.py:201-205)
    // -desc_attn (model.py:344-410): A = attention width (0 = off), NW = words of all class descriptions.
    // y1.weight columns are [h_z ; desc] without and [desc ; h_z] with attention (model.py:408-410 vs 548).
    int A, NW, y1_hcol, y1_dcol;
    int row0;     // global index of this rank's first batch row (Philox counters are keyed by the global row)
};

MMG_HOST_DEVICE Dims make_dims(const mmg_config& c) {
    Dims d;
    d.B = c.batch; d.Bg = c.batch_global; d.F = c.img_feat_dim; d.Hi = c.img_h_dim; d.M = c.msg_dim;
    d.Hr = c.rec_hidden; d.D = c.n_classes; d.WV = c.wv_dim; d.Hb = c.baseline_hid; d.T = c.max_exchange;
    d.R = d.T * d.B; d.G3 = 3 * d.Hr;
    d.M4 = cdiv(d.M, 4); d.Hr4 = cdiv(d.Hr, 4); d.Hi4 = cdiv(d.Hi, 4);
    d.use_binary = c.use_binary; d.fixed = c.fixed_exchange; d.s_prob_prod = c.s_prob_prod;
    d.ignore_receiver = c.ignore_receiver;
    d.first_rec = c.first_rec;
    d.flip_sen = c.has_flipout_sen ? c.flipout_sen : -1.f;
    d.flip_rec = c.has_flipout_rec ? c.flipout_rec : -1.f;
    d.flipout_dev = c.flipout_dev;
    d.mix_prod = c.sender_mix == MMG_MIX_PROD; d.ignore_code = c.ignore_code;
    d.mix_mou = c.sender_mix == MMG_MIX_MOU;
    d.Ha = d.mix_mou ? 4 * d.Hi : d.Hi; d.Ha4 = cdiv(d.Ha, 4);
    d.A = c.desc_attn ? c.desc_attn_dim : 0; d.NW = c.desc_attn ? c.n_words : 0;
    d.y1_hcol = d.A ? d.WV : 0; d.y1_dcol = d.A ? 0 : d.Hr;
    d.NH = 2 * d.Hr + 1 + d.A;
    d.row0 = c.batch_offset;
    return d;
}

// ---- weight images ------------------------------------------------------------------------------------
// "Packed" matrix for a mat-vec out[o] = sum_r W(o, r) * v[r]:  P[(r/4) * Nout * 4 + o * 4 + (r % 4)],
// zero padded along r.  Thread `o` then reads one float4 per 4 reduction elements and a warp reads 512
// contiguous bytes: conflict-free shared-memory (or fully coalesced global) access.
// All offsets in floats, every section a multiple of 4 floats (16 bytes) for the TMA bulk copies.
struct FwdImage {
    // sender part
    int wc;      // code_layer.weight   out=Hi  red=M
    int wb;      // binary_layer.weight out=M   red=Hi
    int b_code;  // [Hi]
    int hw0;     // [Hi] code_layer(sigmoid(code_bias)) incl. bias: the step-0 code term (model.py:199-200)
    int hw0m;    // [Hi] code_layer(sigmoid(code_bias_mou)) incl. bias: the code term after step 0 with mou + ignore_code (201-205)
    int b_b;     // [M]
    int sender_end;
    // receiver part
    int wih;     // rnn.weight_ih  out=3Hr red=M
    int whh;     // rnn.weight_hh  out=3Hr red=Hr
    int whead;   // [y1.weight[:, h cols] ; w_h.weight ; s.weight ; d_h.weight]  out=NH red=Hr
    int ww;      // w.weight       out=M   red=Hr
    int b_ih, b_hh;   // [3Hr] each
    int b_head;  // [NH]  (0 for the y1 rows: y1.bias is folded into y1d; w_h.bias; s.bias)
    int w2;      // [Hr] y2.weight
    int b_w;     // [M]
    int misc;    // [4]: y2.bias, -, -, -
    int y1d;     // [D][Hr]  desc_d . y1.weight[:, Hr:]^T + y1.bias   (class half of y1, loop invariant)
    int wdd;     // [D][Hr]  desc_d . w_d.weight^T
    int total;
};
struct BwdImage {
    int wbT;     // binary_layer.weight^T  out=Hi red=M      (sender rows)
    int sender_end;
    int wwT;     // w.weight^T             out=Hr red=M
    int headT;   // [w_h.weight ; y1.weight[:, h cols] ; d_h.weight (desc_attn)]^T  out=Hr red=2Hr+A
    int whhT;    // rnn.weight_hh^T        out=Hr red=3Hr
    int ws;      // [Hr] s.weight
    int w2;      // [Hr] y2.weight
    int y1d;     // [D][Hr]
    int total;
};

MMG_HOST_DEVICE int align4(int x) { return (x + 3) & ~3; }

MMG_HOST_DEVICE FwdImage make_fwd_image(const Dims& d) {
    FwdImage im; int o = 0;
    im.wc = o; o += d.M4 * d.Hi * 4;
    im.wb = o; o += d.Ha4 * d.M * 4;
    im.b_code = o; o += align4(d.Hi);
    im.hw0 = o; o += align4(d.Hi);
    im.hw0m = o; o += align4(d.Hi);
    im.b_b = o; o += align4(d.M);
    im.sender_end = o;
    im.wih = o; o += d.M4 * d.G3 * 4;
    im.whh = o; o += d.Hr4 * d.G3 * 4;
    im.whead = o; o += d.Hr4 * d.NH * 4;
    im.ww = o; o += d.Hr4 * d.M * 4;
    im.b_ih = o; o += align4(d.G3);
    im.b_hh = o; o += align4(d.G3);
    im.b_head = o; o += align4(d.NH);
    im.w2 = o; o += align4(d.Hr);
    im.b_w = o; o += align4(d.M);
    im.misc = o; o += 4;
    im.y1d = o; o += d.A ? 0 : align4(d.D * d.Hr);      // -desc_attn: word tables in global memory replace the class tables
    im.wdd = o; o += d.A ? 0 : align4(d.D * d.Hr);
    im.total = o;
    return im;
}
MMG_HOST_DEVICE BwdImage make_bwd_image(const Dims& d) {
    BwdImage im; int o = 0;
    im.wbT = o; o += d.M4 * d.Ha * 4;
    im.sender_end = o;
    im.wwT = o; o += d.M4 * d.Hr * 4;
    im.headT = o; o += cdiv(2 * d.Hr + d.A, 4) * d.Hr * 4;
    im.whhT = o; o += cdiv(d.G3, 4) * d.Hr * 4;
    im.ws = o; o += align4(d.Hr);
    im.w2 = o; o += align4(d.Hr);
    im.y1d = o; o += align4(d.D * d.Hr);
    im.total = o;
    return im;
}

// ---- fast-path images (img_h_dim = 256, rec_hidden = 64, msg_dim in {32, 64}; mmg_fast.cuh) -------------------
// Every matrix is stored in the order its consumer phase reads it: a warp's float4 loads are 512 contiguous bytes.
enum { kFastThreads = 256, kFastBwdThreads = 256, kFastHi = 256, kFastHr = 64, kFastMaxT = 32 };
struct FastFwdImage {
    int wc;      // code_layer.weight    [k4 < M/4][n < 256][4]
    int wb;      // binary_layer.weight  row-major (M, 256): the state_dict layout
    int b_code, hw0, b_b;
    int sender_end;
    int wih;     // rnn.weight_ih        [(g*(M/16) + q)][k < 64][part < 4][4],  column = part*(M/4) + 4q + c
    int whead;   // rows [y1.weight[:, :64] ; w_h.weight]  [q < 8][o < 128][half < 2][4],  column = half*32 + 4q + c
    int wgh;     // rnn.weight_hh        [(g*4 + q)][k < 64][part < 4][4],  column = part*16 + 4q + c
    int ww;      // w.weight             [q][j < M][part < LPO][4],  LPO = 256/M, column = part*(64/LPO) + 4q + c
    int b_ih;    // [192]
    int b_hh;    // [192]
    int b_wh;    // [64] w_h.bias  (y1.bias lives in y1d)
    int ws;      // [64] s.weight
    int b_w;     // [M]
    int w2;      // [64] y2.weight
    int misc;    // [4]: y2.bias, s.bias
    int y1d;     // [D][64]   desc_d . y1.weight[:, 64:]^T + y1.bias
    int wdd;     // [ceil(D/4)][64][4]   (desc_d . w_d.weight^T)[k] at ((d/4)*64 + k)*4 + d%4
    int total;
};
MMG_HOST_DEVICE FastFwdImage make_fast_fwd_image(int M, int D) {
    FastFwdImage im; int o = 0;
    im.wc = o; o += M * kFastHi;
    im.wb = o; o += M * kFastHi;
    im.b_code = o; o += kFastHi;
    im.hw0 = o; o += kFastHi;
    im.b_b = o; o += M;
    im.sender_end = o;
    im.wih = o; o += 3 * M * kFastHr;
    im.whead = o; o += 2 * kFastHr * kFastHr;
    im.wgh = o; o += 3 * kFastHr * kFastHr;
    im.ww = o; o += M * kFastHr;
    im.b_ih = o; o += 3 * kFastHr;
    im.b_hh = o; o += 3 * kFastHr;
    im.b_wh = o; o += kFastHr;
    im.ws = o; o += kFastHr;
    im.b_w = o; o += M;
    im.w2 = o; o += kFastHr;
    im.misc = o; o += 4;
    im.y1d = o; o += D * kFastHr;
    im.wdd = o; o += ((D + 3) / 4) * 4 * kFastHr;
    im.total = o;
    return im;
}
struct FastBwdImage {
    int wwT;     // w.weight^T            [j4 < M/4][k < 64][4]   element c = w.weight[4 j4 + c][k]
    int whT;     // w_h.weight^T          [k4 < 16][k < 64][4]
    int w1hT;    // y1.weight[:, :64]^T   [k4 < 16][k < 64][4]
    int whhT;    // rnn.weight_hh^T       [q < 12][k < 64][part < 4][4]   element c = weight_hh[part*48 + 4q + c][k]
    int ws, w2;  // [64] each
    int y1d;     // [ceil(D/4)][64][4]
    int total;
};
MMG_HOST_DEVICE FastBwdImage make_fast_bwd_image(int M, int D) {
    FastBwdImage im; int o = 0;
    im.wwT = o; o += M * kFastHr;
    im.whT = o; o += kFastHr * kFastHr;
    im.w1hT = o; o += kFastHr * kFastHr;
    im.whhT = o; o += 3 * kFastHr * kFastHr;
    im.ws = o; o += kFastHr;
    im.w2 = o; o += kFastHr;
    im.y1d = o; o += ((D + 3) / 4) * 4 * kFastHr;
    im.total = o;
    return im;
}
MMG_HOST_DEVICE bool fast_dims(const Dims& d) {
    return d.Hi == kFastHi && d.Hr == kFastHr && (d.M == 32 || d.M == 64) && d.T <= kFastMaxT && d.A == 0 && !d.mix_mou;
}

// ---- workspace ------------------------------------------------------------------------------------------
enum { kHxSplitMax = 32, kWgradSplitMax = 16, kNormCtasMax = 592, kLossCtasMax = 592 };
enum { kStatKinds = 3 };  // 0: sender messages, 1: receiver messages, 2: stop bit
// stats (double): per (kind, t): n, sum w, sum w^2 (w = logs - baseline); then per t: baseline_rec SSE,
// baseline_sen SSE, n_mask; then scalars: nll_sum, topk_correct.
MMG_HOST_DEVICE int stats_count(const Dims& d) { return kStatKinds * d.T * 3 + 3 * d.T + 4; }
MMG_HOST_DEVICE int stat_idx(const Dims& d, int kind, int t, int which) { return (kind * d.T + t) * 3 + which; }
MMG_HOST_DEVICE int stat_bas(const Dims& d, int t, int which) { return kStatKinds * d.T * 3 + t * 3 + which; }
MMG_HOST_DEVICE int stat_scalar(const Dims& d, int which) { return kStatKinds * d.T * 3 + 3 * d.T + which; }

struct Ws {   // byte offsets into the workspace
    mmg_workspace_layout pub;
    // saved activations
    int64_t code_in;   // (T,B,M)  sender code input: sigmoid(code_bias) at t=0, receiver message after
    int64_t a_s;       // (T,B,Hi) sender tanh hidden
    int64_t hw_s;      // (T,B,Hi) sender code term h_w (kept only for sender_mix = prod: d h_x = d pre * h_w)
    int64_t gates;     // (T,B,4,Hr) r, u, n, (W_hn h + b_hn)
    int64_t y1h;       // (T,B,Hr) y1.weight[:, :Hr] . h_z
    int64_t q;         // (T,B,D)  softmax(y)
    int64_t wd;        // (T,B,WV) q . desc
    int64_t h1s, h1r;  // (T,B,Hb) baseline hidden (post relu)
    int64_t bs_part, br_part;  // (T,B,NTb) partial dots with linear2.weight per 64-column tile
    int64_t ubs;       // (B,Hb)  baseline_sen.linear1 applied to h_x only (+ bias): shared by all T steps of an example (fast path)
    // pre-pass
    int64_t hx_part;   // (S,B,Hi)
    int64_t fwd_image, bwd_image;
    // backward deltas
    int64_t d_lz;      // (T,B,M)  dL/d sender logits
    int64_t d_as;      // (T,B,Hi) dL/d sender pre-tanh
    int64_t dhx;       // (B,Hi)
    int64_t dgi, dgh;  // (T,B,3Hr)
    int64_t d_lw;      // (T,B,M)
    int64_t d_hw;      // (T,B,Hr)
    int64_t d_ls;      // (T,B)
    int64_t g_h;       // (B,Hr)   sum_d dy1
    int64_t hsel;      // (B,Hr)   h_z at the prediction step
    int64_t dy1;       // (B,D,Hr)
    int64_t dw2p;      // (B,Hr)
    int64_t dcode_part; // (B,M)   per-example d code_layer-input at t = 0 (fast path), summed into d code_bias
    int64_t slabs;     // (kWgradSplitMax - 1, P) split-K partial gradients (split 0 lands in the gradient buffer itself)
    int64_t norm_part; // (4, kNormCtasMax) per-CTA partial sums of squares (K_grad_norm, K_peer_reduce_scatter)
    int64_t tile_norm; // float (kMaxOutTiles) sum of squares of each finished gradient tile (K_wgrad)
    int64_t tile_tickets; // uint32 (kMaxOutTiles) split-K arrival counters per output tile, zero between launches
    int64_t norm_final; // double[4] per-module sum of squares of the gradient (pre-clip), consumed by K_update
    int64_t loss_part; // double (kLossCtasMax, 8) per-CTA loss partial sums, summed in CTA order by the last CTA
    int64_t tickets;   // uint32[16]: [0] loss reduction, [1] h_x rows ready, [3] finished gradient tiles, [6] norm partials,
                       //             [7] baseline tiles
    int64_t opt_counters; // int64[4]: [0] = number of updates that reached the receiver message head (Adam bias correction)
    int64_t coefs;     // float (6 T + 4) loss coefficients derived once from the batch statistics (single-rank fused iteration)
    int64_t hit;       // (B) 1 when the target is among the top-k classes of the prediction (fused iteration)
    // -desc_attn (all zero-sized otherwise)
    int64_t wtab_dd;   // (NW,A4)  e^{2 d_d(desc_set)}: word factor of the attention tanh (attn_tanh), loop invariant
    int64_t wtab_y1;   // (NW,Hr)  desc_set . y1.weight[:, :WV]^T
    int64_t wtab_wd;   // (NW,Hr)  desc_set . w_d.weight^T
    int64_t seg;       // int32 (D+1) first word of each class;  wcls: int32 (NW) class of each word
    int64_t wcls;
    int64_t qa;        // (T,B,NW) softmax(y)[class of word] * attention weight: rows of the wd = (q a) . desc_set GEMM
    int64_t attn;      // (T,B,NW) attention weights (softmax within each class's words)
    int64_t dh_s;      // (T,B,A)  e^{2 d_h(h_z)}: hidden factor of the attention tanh
    int64_t ddh;       // (T,B,A)  dL/d dh_s
    int64_t dva;       // (T,B,A)  per-row partial of dL/d d_attn.weight;  dba (T,B): of d_attn.bias
    int64_t dba;
    int64_t ddd_part;  // (n_rec_ctas, [NW][A4] ; [NW][Hr4]) per-CTA partials: dL/d wtab_dd summed over the CTA's steps/examples,
                       // and Z[n] = sum_b a[b,n] d y1[b,class(n)] (word factor of the y1 description-column gradient)
    int hx_split, wgrad_split, ntb;
};

}  // namespace mmg
