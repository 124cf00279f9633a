// Host side of the C-ABI (include/mmg_b200.h): layouts, argument validation, launch sequences.
// No allocation, no host synchronisation (except mmg_device_count); every sequence is CUDA-graph capturable.
#include "mmg_pre.cuh"
#include "mmg_exchange_fwd.cuh"
#include "mmg_exchange_bwd.cuh"
#include "mmg_fast.cuh"
#include "mmg_loss.cuh"
#include "mmg_update.cuh"
#include "mmg_single.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>

namespace mmg {
namespace host {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;
void count_launch() { ++g_launches; }

#ifndef MMG_CPU_EMU
// ---- MMG_KTIME=1: in-situ per-kernel device times (diagnostic; mmg_debug_kernel_times reads and clears them) -------------
struct KTimeRec { const char* name; cudaEvent_t a, b; };
static const int kKTimeMax = 1 << 16;
static KTimeRec* g_ktime = nullptr;
static int g_ktime_n = 0;
int ktime_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MMG_KTIME"); v = (e && e[0] == '1') ? 1 : 0; }
    return v;
}
void ktime_begin(const char* name, cudaStream_t st) {
    if (!g_ktime) g_ktime = (KTimeRec*)calloc(kKTimeMax, sizeof(KTimeRec));
    if (g_ktime_n >= kKTimeMax) return;
    KTimeRec& r = g_ktime[g_ktime_n];
    r.name = name;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
}
void ktime_end(cudaStream_t st) {
    if (g_ktime_n >= kKTimeMax) return;
    cudaEventRecord(g_ktime[g_ktime_n].b, st);
    ++g_ktime_n;
}
#endif

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#ifndef MMG_CPU_EMU
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda).  Returns false when the
// driver does not offer it or an encode fails: the caller then stages the operands with asynchronous 16-byte copies.
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmapEncodeFn tmap_encoder() {
    static TmapEncodeFn fn = nullptr;
    static int tried = 0;
    if (!tried) {
        tried = 1;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (TmapEncodeFn)p;
        (void)cudaGetLastError();
    }
    return fn;
}
// K-major fp32 matrix (rows x cols, row stride cols * 4 bytes): boxes of 32 floats (one 128-byte swizzle span) x box_rows rows
static bool encode_kmajor_tmap(CUtensorMap* tm, const float* base, int rows, int cols, int box_rows) {
    TmapEncodeFn enc = tmap_encoder();
    if (!enc) return false;
    // small cache: the weight matrix and the (two) staging slots of the host pipeline come back every step
    struct Slot { const float* base; int rows, cols, box; CUtensorMap tm; };
    static thread_local Slot cache[8];
    static thread_local int next = 0;
    for (int i = 0; i < 8; ++i)
        if (cache[i].base == base && cache[i].rows == rows && cache[i].cols == cols && cache[i].box == box_rows) { *tm = cache[i].tm; return true; }
    Slot& sl = cache[next];
    next = (next + 1) & 7;
    sl.base = nullptr;
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    if (enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    sl.base = base; sl.rows = rows; sl.cols = cols; sl.box = box_rows; sl.tm = *tm;
    return true;
}
static int check_cuda(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MMG_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return MMG_OK;
}
static const int kMaxSmem = 227 * 1024;
#else
static int check_cuda(const char*) { return MMG_OK; }
static const int kMaxSmem = 227 * 1024;
#endif

static int validate(const mmg_config* c) {
    if (c == nullptr) return fail(MMG_ERR_INVALID, "null config");
    if (c->batch < 1 || c->batch_global < c->batch) return fail(MMG_ERR_INVALID, "batch=%d batch_global=%d", c->batch, c->batch_global);
    if (c->batch_offset < 0 || c->batch_offset + c->batch > c->batch_global)
        return fail(MMG_ERR_INVALID, "batch_offset=%d outside the global batch (batch=%d batch_global=%d)", c->batch_offset, c->batch, c->batch_global);
    if (c->img_feat_dim < 1 || c->img_h_dim < 1 || c->msg_dim < 1 || c->rec_hidden < 1 || c->n_classes < 1 ||
        c->wv_dim < 1 || c->baseline_hid < 1 || c->max_exchange < 1)
        return fail(MMG_ERR_INVALID, "all dimensions must be >= 1");
    if (c->rec_hidden > 256) return fail(MMG_ERR_UNSUPPORTED, "rec_hidden=%d > 256 not supported by the fused path", c->rec_hidden);
    if (c->msg_dim > 256) return fail(MMG_ERR_UNSUPPORTED, "msg_dim=%d > 256 not supported by the fused path", c->msg_dim);
    if (c->img_h_dim > (int)kGemmSmemFloats * 64) return fail(MMG_ERR_UNSUPPORTED, "img_h_dim=%d too large", c->img_h_dim);
    if (c->optim_type < 0 || c->optim_type > 2) return fail(MMG_ERR_INVALID, "optim_type=%d", c->optim_type);
    if (c->sender_mix != MMG_MIX_SUM && c->sender_mix != MMG_MIX_PROD && c->sender_mix != MMG_MIX_MOU) return fail(MMG_ERR_INVALID, "sender_mix=%d", c->sender_mix);
    if ((long long)c->max_exchange * c->batch > (1ll << 24)) return fail(MMG_ERR_UNSUPPORTED, "T*B too large");
    if (c->desc_attn) {
        if (c->desc_attn_dim < 1 || c->desc_attn_dim > 256) return fail(MMG_ERR_INVALID, "desc_attn_dim=%d", c->desc_attn_dim);
        if (c->n_words < c->n_classes) return fail(MMG_ERR_INVALID, "n_words=%d < n_classes=%d (every class needs a word)", c->n_words, c->n_classes);
        if (c->n_words > 4096) return fail(MMG_ERR_UNSUPPORTED, "n_words=%d > 4096", c->n_words);
    }
    return MMG_OK;
}

static void param_layout(const Dims& d, mmg_param_layout* L) {
    const int Hr = d.Hr, M = d.M, WV = d.WV, Hi = d.Hi, F = d.F, Hb = d.Hb, A = d.A, on = d.A ? 1 : 0;
    const int mo = (d.mix_mou && d.ignore_code) ? 1 : 0;      // code_bias_mou exists (model.py:73-74)
    struct E { int id, rows, cols, seg; };
    const E tab[MMG_P_COUNT] = {
        {MMG_P_REC_RNN_WIH, 3 * Hr, M, 0}, {MMG_P_REC_RNN_WHH, 3 * Hr, Hr, 0}, {MMG_P_REC_RNN_BIH, 3 * Hr, 1, 0},
        {MMG_P_REC_RNN_BHH, 3 * Hr, 1, 0}, {MMG_P_REC_WH_W, Hr, Hr, 0}, {MMG_P_REC_WH_B, Hr, 1, 0},
        {MMG_P_REC_WD_W, Hr, WV, 0}, {MMG_P_REC_W_W, M, Hr, 0}, {MMG_P_REC_W_B, M, 1, 0},
        {MMG_P_REC_Y1_W, Hr, Hr + WV, 0}, {MMG_P_REC_Y1_B, Hr, 1, 0}, {MMG_P_REC_Y2_W, 1, Hr, 0},
        {MMG_P_REC_Y2_B, 1, 1, 0}, {MMG_P_REC_S_W, 1, Hr, 0}, {MMG_P_REC_S_B, 1, 1, 0},
        {MMG_P_REC_DD_W, A, on * WV, 0}, {MMG_P_REC_DD_B, A, on, 0}, {MMG_P_REC_DH_W, A, on * Hr, 0}, {MMG_P_REC_DH_B, A, on, 0},
        {MMG_P_REC_DA_W, on, A, 0}, {MMG_P_REC_DA_B, on, on, 0},
        {MMG_P_SEN_CODE_BIAS, M, 1, 1}, {MMG_P_SEN_CODE_BIAS_MOU, mo * M, mo, 1}, {MMG_P_SEN_IMG_W, Hi, F, 1}, {MMG_P_SEN_IMG_B, Hi, 1, 1},
        {MMG_P_SEN_CODE_W, Hi, M, 1}, {MMG_P_SEN_CODE_B, Hi, 1, 1}, {MMG_P_SEN_BIN_W, M, d.Ha, 1},
        {MMG_P_SEN_BIN_B, M, 1, 1},
        {MMG_P_BR_L1_W, Hb, M + Hr, 2}, {MMG_P_BR_L1_B, Hb, 1, 2}, {MMG_P_BR_L2_W, 1, Hb, 2}, {MMG_P_BR_L2_B, 1, 1, 2},
        {MMG_P_BS_L1_W, Hb, Hi + M, 3}, {MMG_P_BS_L1_B, Hb, 1, 3}, {MMG_P_BS_L2_W, 1, Hb, 3}, {MMG_P_BS_L2_B, 1, 1, 3}};
    int64_t off = 0;
    int cur_seg = -1;
    for (int i = 0; i < MMG_P_COUNT; ++i) {
        const E& e = tab[i];
        if (e.seg != cur_seg) { cur_seg = e.seg; L->seg_begin[cur_seg] = off; }
        L->offset[e.id] = off;
        L->rows[e.id] = e.rows; L->cols[e.id] = e.cols; L->segment[e.id] = e.seg;
        off += round_up64((int64_t)e.rows * e.cols, 4);
    }
    L->seg_begin[MMG_SEG_COUNT] = off;
    L->total = off;
}

static int64_t take(int64_t& cur, int64_t bytes) {
    const int64_t at = cur;
    cur += round_up64(bytes, 256);
    return at;
}

static void ws_layout(const Dims& d, Ws* w) {
    memset(w, 0, sizeof(*w));
    mmg_param_layout L;
    param_layout(d, &L);
    const int64_t R = d.R, B = d.B, f = sizeof(float);
    int64_t c = 0;
    mmg_workspace_layout& p = w->pub;
    p.sen_feats = take(c, R * d.M * f);
    p.sen_probs = take(c, R * d.M * f);
    p.rec_feats = take(c, (R + B) * d.M * f);
    p.rec_probs = take(c, R * d.M * f);
    p.stop_feat = take(c, R * f);
    p.stop_prob = take(c, R * f);
    p.y = take(c, R * d.D * f);
    p.stop_mask = take(c, R + B);
    p.bs = take(c, R * f);
    p.br = take(c, R * f);
    p.h_x = take(c, B * d.Hi * f);
    p.h_z = take(c, (R + B) * d.Hr * f);
    p.h_w = take(c, R * d.Hr * f);
    p.losses = take(c, MMG_LOSS_COUNT * f);
    p.ystep = take(c, B * 4);
    p.outp = take(c, B * d.D * f);
    p.logs = take(c, B * f);
    p.argmax = take(c, B * 4);
    p.stats_count = stats_count(d);
    p.stats = take(c, p.stats_count * 8);
    p.grad_norms = take(c, 4 * f);
    p.g_sen_probs = take(c, R * d.M * f);
    p.g_rec_probs = take(c, R * d.M * f);
    p.g_stop_prob = take(c, R * f);
    p.g_outp = take(c, B * d.D * f);
    p.g_bs = take(c, R * f);
    p.g_br = take(c, R * f);
    p.rng_state = take(c, 16);
    w->code_in = take(c, R * d.M * f);
    w->a_s = take(c, R * d.Ha * f);
    w->hw_s = take(c, (d.mix_prod || d.mix_mou) ? R * d.Hi * f : 16);
    w->gates = take(c, R * 4 * d.Hr * f);
    w->y1h = take(c, R * d.Hr * f);
    w->q = take(c, R * d.D * f);
    w->wd = take(c, R * d.WV * f);
    w->ntb = cdiv(d.Hb, kTile);
    w->h1s = take(c, R * d.Hb * f);
    w->h1r = take(c, R * d.Hb * f);
    w->bs_part = take(c, R * w->ntb * f);
    w->br_part = take(c, R * w->ntb * f);
    w->ubs = take(c, B * d.Hb * f);
    int hs = d.F / 128;      // image-layer GEMM: K-slices of 128 features per CTA (8 chunks), partials summed by the consumer
    if (hs < 1) hs = 1;
    if (hs > kHxSplitMax) hs = kHxSplitMax;
    w->hx_split = hs;
    w->hx_part = take(c, (int64_t)hs * B * d.Hi * f);
    int64_t fi = make_fwd_image(d).total, bi = make_bwd_image(d).total;
    if (fast_dims(d) || fast_fwd_attn_dims(d)) {   // the fast-path images share the buffers
        const int64_t ffi = make_fast_fwd_image(d.M, d.D).total, fbi = make_fast_bwd_image(d.M, d.D).total;
        if (ffi > fi) fi = ffi;
        if (fbi > bi) bi = fbi;
    }
    w->fwd_image = take(c, fi * f);
    w->bwd_image = take(c, bi * f);
    w->d_lz = take(c, R * d.M * f);
    w->d_as = take(c, R * d.Hi * f);
    w->dhx = take(c, B * d.Hi * f);
    w->dgi = take(c, R * d.G3 * f);
    w->dgh = take(c, R * d.G3 * f);
    w->d_lw = take(c, R * d.M * f);
    w->d_hw = take(c, R * d.Hr * f);
    w->d_ls = take(c, R * f);
    w->g_h = take(c, B * d.Hr * f);
    w->hsel = take(c, B * d.Hr * f);
    w->dy1 = take(c, B * d.D * d.Hr * f);
    w->dw2p = take(c, B * d.Hr * f);
    w->dcode_part = take(c, B * d.M * f);
    w->wgrad_split = kWgradSplitMax;
    w->slabs = take(c, (int64_t)(kWgradSplitMax - 1) * L.total * f);
    w->norm_part = take(c, 4 * kNormCtasMax * f);
    w->tile_norm = take(c, 2 * kMaxOutTiles * f);      // sums of squares | module of each tile
    w->tile_tickets = take(c, kMaxOutTiles * 4);
    w->norm_final = take(c, 4 * 8);
    w->hit = take(c, B * f);
    w->coefs = take(c, (6 * d.T + 4) * f);
    w->loss_part = take(c, (int64_t)(2 * B > kLossCtasMax ? 2 * B : kLossCtasMax) * 8 * 8);   // K_lossgrad CTAs, or 2 CTAs per example (fused backward)
    w->tickets = take(c, 16 * 4);
    w->opt_counters = take(c, 4 * 8);
    p.opt_counters = w->opt_counters;
    if (d.A) {
        const int64_t NW = d.NW, A = d.A;
        const int64_t AP = align4(d.A), HrP = align4(d.Hr);     // table rows padded to whole float4 groups
        w->wtab_dd = take(c, NW * AP * f);
        w->wtab_y1 = take(c, NW * HrP * f);
        w->wtab_wd = take(c, NW * HrP * f);
        w->qa = take(c, R * NW * f);
        w->seg = take(c, (d.D + 1) * 4);
        w->wcls = take(c, NW * 4);
        w->attn = take(c, R * NW * f);
        w->dh_s = take(c, R * A * f);
        w->ddh = take(c, R * A * f);
        w->dva = take(c, R * A * f);
        w->dba = take(c, R * f);
        w->ddd_part = take(c, B * NW * (AP + HrP) * f);     // per receiver CTA (at most one per example): [NW][AP] ; [NW][HrP]
    }
    p.total_bytes = c;
}

static WsPtrs resolve(const Ws& w, void* base) {
    char* b = (char*)base;
    WsPtrs r;
    const mmg_workspace_layout& p = w.pub;
#define F_(name) r.name = (float*)(b + p.name)
    F_(sen_feats); F_(sen_probs); F_(rec_feats); F_(rec_probs); F_(stop_feat); F_(stop_prob); F_(y); F_(bs); F_(br);
    F_(h_x); F_(h_z); F_(h_w); F_(losses); F_(outp); F_(logs); F_(grad_norms);
    F_(g_sen_probs); F_(g_rec_probs); F_(g_stop_prob); F_(g_outp); F_(g_bs); F_(g_br);
#undef F_
    r.stop_mask = (unsigned char*)(b + p.stop_mask);
    r.ystep = (int*)(b + p.ystep);
    r.argmax = (int*)(b + p.argmax);
    r.stats = (double*)(b + p.stats);
    r.rng_state = (unsigned long long*)(b + p.rng_state);
#define G_(name) r.name = (float*)(b + w.name)
    G_(code_in); G_(a_s); G_(hw_s); G_(gates); G_(y1h); G_(q); G_(wd); G_(h1s); G_(h1r); G_(bs_part); G_(br_part); G_(ubs);
    G_(hx_part); G_(fwd_image); G_(bwd_image); G_(d_lz); G_(d_as); G_(dhx); G_(dgi); G_(dgh); G_(d_lw); G_(d_hw);
    G_(d_ls); G_(g_h); G_(hsel); G_(dy1); G_(dw2p); G_(dcode_part); G_(slabs); G_(norm_part);
#undef G_
    r.loss_part = (double*)(b + w.loss_part);
    r.tickets = (unsigned*)(b + w.tickets);
    r.tile_norm = (float*)(b + w.tile_norm);
    r.tile_tickets = (unsigned*)(b + w.tile_tickets);
    r.norm_final = (double*)(b + w.norm_final);
    r.hit = (float*)(b + w.hit);
    r.coefs = (float*)(b + w.coefs);
    r.opt_counters = (long long*)(b + w.opt_counters);
#define G_(name) r.name = (float*)(b + w.name)
    G_(wtab_dd); G_(wtab_y1); G_(wtab_wd); G_(qa); G_(attn); G_(dh_s); G_(ddh); G_(dva); G_(dba); G_(ddd_part);
#undef G_
    r.seg = (int*)(b + w.seg); r.wcls = (int*)(b + w.wcls);
    r.hx_split = w.hx_split; r.wgrad_split = w.wgrad_split; r.ntb = w.ntb;
    return r;
}

static ParamPtrs param_ptrs(const mmg_param_layout& L, const float* base) {
    ParamPtrs P;
    for (int i = 0; i < MMG_P_COUNT; ++i) P.p[i] = base + L.offset[i];
    return P;
}

static ExchangeInputs resolve_inputs(const mmg_inputs* in) {
    ExchangeInputs e;
    e.x = in->d_x; e.desc = in->d_desc; e.target = (const long long*)in->d_target;
    e.u_sen = in->d_u_sen; e.u_stop = in->d_u_stop; e.u_rec = in->d_u_rec;
    e.u_flip_sen = in->d_u_flip_sen; e.u_flip_rec = in->d_u_flip_rec;
    e.corrupt_mask = in->d_corrupt_mask; e.h0 = in->d_h0; e.top_k = in->top_k; e.train = in->train;
    e.desc_set = in->d_desc_set; e.desc_set_lens = in->d_desc_set_lens;
    return e;
}

struct Plan { int BT; int sender_smem; int fwd_smem_bytes; int bwd_smem_bytes; int fast; int attn_acc;
              int fast_fwd; int fast_fwd_smem_bytes; };   // fast_fwd: the register-resident forward kernel (always with `fast`;
                                                          // alone for -desc_attn, whose backward runs on the generic kernel)

static AttnArgs attn_args(const Dims& d, const ParamPtrs& P, const ExchangeInputs& in, const Plan& pl) {
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    if (d.A) {
        a.acc_smem = pl.attn_acc;
        a.dh_w = P.p[MMG_P_REC_DH_W]; a.dh_b = P.p[MMG_P_REC_DH_B]; a.va = P.p[MMG_P_REC_DA_W]; a.ba = P.p[MMG_P_REC_DA_B];
        a.b1 = P.p[MMG_P_REC_Y1_B]; a.desc_set = in.desc_set;
    }
    return a;
}

static int choose_bt(int B) {
    if (B <= 148) return 1;
    if (B <= 2 * 148) return 2;
    if (B <= 4 * 148) return 4;
    return 8;
}

static int g_force_generic = -1;   // MMG_FORCE_GENERIC=1 routes every shape through the generic kernels (tests)

static bool make_fast_plan(const Dims& d, Plan* pl) {
    if (g_force_generic < 0) {
        const char* e = getenv("MMG_FORCE_GENERIC");
        g_force_generic = (e && e[0] == '1') ? 1 : 0;
    }
    if (g_force_generic || !fast_dims(d)) return false;
    const FastFwdImage fi = make_fast_fwd_image(d.M, d.D);
    const FastBwdImage bi = make_fast_bwd_image(d.M, d.D);
    int bw_rec = (bi.total + fast_bwd_rec_state_floats(d.T, d.M, d.D)) * 4;
    int bw_sen = fast_bwd_sen_state_floats(d.T, d.M) * 4;
    pl->bwd_smem_bytes = bw_rec > bw_sen ? bw_rec : bw_sen;
    if (pl->bwd_smem_bytes > kMaxSmem) return false;
    int bt = d.B <= 148 ? 1 : (d.B <= 2 * 148 ? 2 : 4);
    for (; bt >= 1; bt /= 2) {
        int st = fast_fwd_state_floats(bt, d.M, d.D, d.T);
#ifdef MMG_PHASE_TIMING
        st += d.T * 64 + 8;
#endif
        // msg_dim 32: every loop matrix in registers; msg_dim 64: receiver in registers, sender section in shared memory
        const int tail = fi.total - fi.b_ih;
        const int need = ((d.M == 32 ? 0 : fi.sender_end) + tail + st) * 4;
        if (need <= kMaxSmem) {
            pl->sender_smem = d.M == 32 ? 0 : 1;
            pl->fwd_smem_bytes = need > (int)kGemmSmemFloats * 4 ? need : (int)kGemmSmemFloats * 4;   // side-role GEMM tiles
            break;
        }
    }
    if (bt < 1) return false;
    pl->BT = bt;
    pl->fast = 1;
    return true;
}

static int make_plan(const Dims& d, Plan* pl) {
    pl->fast = 0; pl->attn_acc = 0; pl->fast_fwd = 0; pl->fast_fwd_smem_bytes = 0;
    if (make_fast_plan(d, pl)) { pl->fast_fwd = 1; pl->fast_fwd_smem_bytes = pl->fwd_smem_bytes; return MMG_OK; }
    if (!g_force_generic && fast_fwd_attn_dims(d)) {
        // -desc_attn at the fast shapes: forward on the register-resident kernel with the word tables in shared memory
        const FastFwdImage fi = make_fast_fwd_image(d.M, d.D);
        const int need = (fi.total - fi.b_ih + fast_fwd_state_floats(1, d.M, d.D, d.T) + fast_fwd_attn_floats(d.D, d.NW)) * 4;
        const char* env = getenv("MMG_FAST_ATTN");            // =0: generic forward (tests)
        if (need <= kMaxSmem && !(env && env[0] == '0')) { pl->fast_fwd = 1; pl->fast_fwd_smem_bytes = need; }
    }
    const FwdImage fi = make_fwd_image(d);
    const BwdImage bi = make_bwd_image(d);
    pl->BT = choose_bt(d.B);
    for (;;) {
        const int st = fwd_state_floats(d, pl->BT);
        const int full = (fi.total + st) * 4, recv_only = (fi.total - fi.sender_end + st) * 4;
        int bw_rec = (bi.total - bi.sender_end + bwd_rec_state_floats(d, pl->BT)) * 4;
        int bw_sen = (bi.sender_end + bwd_sen_state_floats(d, pl->BT)) * 4;
        int bw = bw_rec > bw_sen ? bw_rec : bw_sen;
        if (recv_only <= kMaxSmem && bw <= kMaxSmem) {
            const int acc_bytes = d.A ? d.NW * align4(d.A) * 4 : 0;      // -desc_attn: d (d_d(word)) sums beside the weights when they fit
            const char* env = getenv("MMG_ATTN_ACC_SMEM");        // =0 forces the global-memory accumulators (tests)
            if (d.A && bw_rec + acc_bytes <= kMaxSmem && !(env && env[0] == '0')) {
                pl->attn_acc = 1;
                if (bw_rec + acc_bytes > bw) bw = bw_rec + acc_bytes;
            }
            pl->sender_smem = full <= kMaxSmem ? 1 : 0;
            pl->fwd_smem_bytes = pl->sender_smem ? full : recv_only;
            pl->bwd_smem_bytes = bw;
            return MMG_OK;
        }
        if (pl->BT == 1) break;
        pl->BT /= 2;
    }
    return fail(MMG_ERR_UNSUPPORTED, "receiver weights (%d floats) do not fit in shared memory", fi.total - fi.sender_end);
}

#ifndef MMG_CPU_EMU
template <typename K>
static int set_smem(K kernel, int bytes) {
    // one driver call per (kernel instantiation, device) and size change, not per launch: the attribute is sticky
    // (kernels of different instantiations can share one function-pointer TYPE, so the cache is keyed by the pointer value)
    struct Ent { const void* fn; int bytes, dev; };
    static thread_local Ent cache[64];
    static thread_local int n_ent = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    Ent* hit = nullptr;
    for (int i = 0; i < n_ent; ++i) if (cache[i].fn == (const void*)kernel && cache[i].dev == dev) { hit = &cache[i]; break; }
    if (hit != nullptr && hit->bytes == bytes) return MMG_OK;
    if (hit == nullptr && n_ent < 64) { hit = &cache[n_ent++]; hit->fn = (const void*)kernel; hit->dev = dev; }
    if (hit != nullptr) hit->bytes = bytes;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return fail(MMG_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", bytes, cudaGetErrorString(e));
    return MMG_OK;
}
#else
template <typename K>
static int set_smem(K, int) { return MMG_OK; }
#endif

template <int BT>
static int launch_fwd(const Dims& d, const WsPtrs& W, const ExchangeInputs& in, const float* b_img, const Plan& pl,
                      cudaStream_t st, const AttnArgs& aa) {
    int rc = set_smem(k_exchange_fwd<BT>, pl.fwd_smem_bytes);
    if (rc) return rc;
    MMG_LAUNCH(k_exchange_fwd<BT>, cdiv(d.B, BT), kLoopThreads, pl.fwd_smem_bytes, st, d, W, in, b_img, pl.sender_smem, d.row0, aa);
    return check_cuda("k_exchange_fwd");
}
template <int BT>
static int launch_bwd(const Dims& d, const WsPtrs& W, const Plan& pl, cudaStream_t st, const AttnArgs& aa) {
    int rc = set_smem(k_exchange_bwd<BT>, pl.bwd_smem_bytes);
    if (rc) return rc;
    const int n_rec = cdiv(d.B, BT), n_sen = d.use_binary ? cdiv(d.B, BT) : 0;
    MMG_LAUNCH(k_exchange_bwd<BT>, n_rec + n_sen, kLoopThreads, pl.bwd_smem_bytes, st, d, W, n_rec, aa);
    return check_cuda("k_exchange_bwd");
}

struct FastFwdArgs { const float* b_img; const float* bs_w1; const float* bs_b1; int epilogue; };
template <int BT, int M, bool SS, bool PERF>
static int launch_fwd_fast_mode(const Dims& d, const WsPtrs& W, const ExchangeInputs& in, const FastFwdArgs& fa, const Plan& pl,
                                cudaStream_t st) {
    auto k_exchange_fwd_fast_ = k_exchange_fwd_fast<BT, M, SS, PERF>;
    int rc = set_smem(k_exchange_fwd_fast_, pl.fwd_smem_bytes);
    if (rc) return rc;
    const int n_conv = cdiv(d.B, BT);
    const int n_side = in.train ? cdiv(d.B, kTile) * cdiv(d.Hb, kTile) : 0;     // baseline pre-activation tiles
    AttnArgs none;
    memset(&none, 0, sizeof(none));
    MMG_LAUNCH(k_exchange_fwd_fast_, n_conv + n_side, kFastThreads, pl.fwd_smem_bytes, st, d, W, in, fa.b_img, d.row0, fa.bs_w1, fa.bs_b1, n_conv, none, fa.epilogue);
    return check_cuda("k_exchange_fwd_fast");
}
template <int BT, int M, bool SS>
static int launch_fwd_fast_one(const Dims& d, const WsPtrs& W, const ExchangeInputs& in, const FastFwdArgs& fa, const Plan& pl,
                               cudaStream_t st) {
    const bool perf = in.train && d.use_binary && in.u_sen == nullptr && in.corrupt_mask == nullptr && !d.ignore_receiver &&
                      d.flip_sen < 0.f && d.flip_rec < 0.f && !d.mix_prod && !d.ignore_code && d.B % BT == 0;
    return perf ? launch_fwd_fast_mode<BT, M, SS, true>(d, W, in, fa, pl, st)
                : launch_fwd_fast_mode<BT, M, SS, false>(d, W, in, fa, pl, st);
}
template <int M, bool SS>
static int launch_fwd_fast_bt(const Dims& d, const WsPtrs& W, const ExchangeInputs& in, const FastFwdArgs& fa, const Plan& pl,
                              cudaStream_t st) {
    switch (pl.BT) {
        case 1: return launch_fwd_fast_one<1, M, SS>(d, W, in, fa, pl, st);
        case 2: return launch_fwd_fast_one<2, M, SS>(d, W, in, fa, pl, st);
        default: return launch_fwd_fast_one<4, M, SS>(d, W, in, fa, pl, st);
    }
}
static int launch_fwd_fast_attn(const Dims& d, const WsPtrs& W, const ExchangeInputs& in, const FastFwdArgs& fa, const Plan& pl,
                                cudaStream_t st, const AttnArgs& aa) {
    // kPerf: the training configuration bench.py measures, every mode flag a compile-time constant
    const bool perf = in.train && d.use_binary && in.u_sen == nullptr && in.corrupt_mask == nullptr && !d.ignore_receiver &&
                      d.flip_sen < 0.f && d.flip_rec < 0.f && !d.mix_prod && !d.ignore_code;
    int rc;
    const int n_side = in.train ? cdiv(d.B, kTile) * cdiv(d.Hb, kTile) : 0;     // baseline pre-activation tiles (U[b])
    if (perf) {
        auto k_exchange_fwd_fast_attn_ = k_exchange_fwd_fast<1, 32, true, true, true>;
        if ((rc = set_smem(k_exchange_fwd_fast_attn_, pl.fast_fwd_smem_bytes))) return rc;
        MMG_LAUNCH(k_exchange_fwd_fast_attn_, d.B + n_side, kFastThreads, pl.fast_fwd_smem_bytes, st, d, W, in, fa.b_img, d.row0, fa.bs_w1, fa.bs_b1, d.B, aa, fa.epilogue);
    } else {
        auto k_exchange_fwd_fast_attn_ = k_exchange_fwd_fast<1, 32, true, false, true>;
        if ((rc = set_smem(k_exchange_fwd_fast_attn_, pl.fast_fwd_smem_bytes))) return rc;
        MMG_LAUNCH(k_exchange_fwd_fast_attn_, d.B + n_side, kFastThreads, pl.fast_fwd_smem_bytes, st, d, W, in, fa.b_img, d.row0, fa.bs_w1, fa.bs_b1, d.B, aa, fa.epilogue);
    }
    return check_cuda("k_exchange_fwd_fast<attn>");
}
static int launch_fwd_fast(const Dims& d, const WsPtrs& W, const ExchangeInputs& in, const FastFwdArgs& fa, const Plan& pl,
                           cudaStream_t st) {
    if (d.M == 32) return launch_fwd_fast_bt<32, true>(d, W, in, fa, pl, st);
    return launch_fwd_fast_bt<64, false>(d, W, in, fa, pl, st);
}
template <int M, bool FUSE>
static int launch_bwd_fast_m(const Dims& d, const WsPtrs& W, const float* bin_w, const float* code_w, const Plan& pl,
                             cudaStream_t st, const mmg_config& cfg, const PeerView& pv, const float* bs_w2) {
    const int loss_off = (pl.bwd_smem_bytes + 15) / 16 * 4;          // floats; the coefficient scratch follows the state
    const int smem = FUSE ? loss_off * 4 + loss_smem_bytes(d, pv.world) : pl.bwd_smem_bytes;
    if (smem > kMaxSmem) return fail(MMG_ERR_UNSUPPORTED, "fused backward: %d bytes of shared memory", smem);
    auto k_exchange_bwd_fast_ = k_exchange_bwd_fast<M, FUSE>;
    int rc = set_smem(k_exchange_bwd_fast_, smem);
    if (rc) return rc;
    const int n_rec = d.B, n_sen = d.use_binary ? d.B : 0;
    MMG_LAUNCH(k_exchange_bwd_fast_, n_rec + n_sen, kFastBwdThreads, smem, st, d, W, bin_w, code_w, n_rec, cfg, pv, loss_off, bs_w2);
    return check_cuda("k_exchange_bwd_fast");
}

static Operand km(const float* p, int ld) { return Operand{p, nullptr, nullptr, nullptr, ld, 0, 1, 0, 0, OP_PLAIN}; }
static Operand ones() { return Operand{nullptr, nullptr, nullptr, nullptr, 0, 0, 1, 0, 0, OP_ONES}; }

struct WgBuilder {
    WgTable* t;
    const mmg_param_layout* L;
    bool covered[MMG_P_COUNT];
    // C (M x N) = A^T B goes to tensor `wid` at column offset `col`; colsum(A) (optional) to tensor `bid`
    void add(const Operand& A, const Operand& B, int M, int N, int K, int wid, int col, int bid, int kind = WG_GEMM,
             const float* sig_rows = nullptr) {
        WgProblem& p = t->p[t->count];
        ++t->count;
        covered[wid] = true;
        if (bid >= 0) covered[bid] = true;
        p.A = A; p.B = B; p.M = M; p.N = N; p.K = K;
        p.c_off = L->offset[wid] + col; p.ldc = N == 1 ? 1 : L->cols[wid];   // N == 1: column sums land contiguously
        p.bias_off = bid >= 0 ? L->offset[bid] : -1;
        p.sig_rows = sig_rows; p.kind = kind;
        p.seg = L->segment[wid];
        int ns = kind != WG_CODEBIAS ? cdiv(K, kWgradKSlice) : 1;     // every problem picks its own split factor
        if (ns < 1) ns = 1;
        if (ns > kWgradSplitMax) ns = kWgradSplitMax;
        p.nsplit = ns;
    }
    void add_zero(int tensor) {
        WgProblem& p = t->p[t->count];
        ++t->count;
        memset(&p, 0, sizeof(p));
        p.kind = WG_ZERO; p.nsplit = 1; p.N = 1; p.K = 1;
        p.M = (int)round_up64((int64_t)L->rows[tensor] * L->cols[tensor], 4);     // floats to clear (the whole 4-float slot)
        p.c_off = L->offset[tensor]; p.bias_off = -1; p.seg = L->segment[tensor];
    }
    int finish() {
        for (int i = 0; i < MMG_P_COUNT; ++i) {     // tensors nobody writes this iteration: cleared (the buffer is fully defined)
            if (covered[i] || (int64_t)L->rows[i] * L->cols[i] == 0) continue;
            if (t->count >= kMaxWgProblems) return -1;
            add_zero(i);
        }
        t->total_tiles = 0; t->total_out = 0;
        for (int i = 0; i < t->count; ++i) {
            WgProblem& p = t->p[i];
            if (p.kind == WG_ZERO) { p.ntm = cdiv(p.M, kZeroChunk); p.ntn = 1; }
            else {
                p.ntm = cdiv(p.M, kTile);
                p.ntn = p.kind == WG_GEMM ? cdiv(p.N, kTile) : (p.kind == WG_ROWVEC ? cdiv(p.N, kRowvecCols) : 1);
            }
            p.tile_begin = t->total_tiles;
            p.out_begin = t->total_out;
            t->tile_begin[i] = p.tile_begin;
            t->total_tiles += p.ntm * p.ntn * p.nsplit;
            t->total_out += p.ntm * p.ntn;
        }
        t->tile_begin[t->count] = t->total_tiles;
        return t->total_out <= kMaxOutTiles ? 0 : -3;
    }
};

static int build_wgrad_table(const Dims& d, const mmg_param_layout& L, const ParamPtrs& P, const WsPtrs& W,
                             const ExchangeInputs& in, int fast, int fast_bs, int n_rec_ctas, WgTable* t) {
    t->count = 0; t->total_tiles = 0; t->slab_stride = L.total;
    WgBuilder b{t, &L, {}};
    for (int i = 0; i < MMG_P_COUNT; ++i) b.covered[i] = false;
    const int R = d.R, B = d.B, Hr = d.Hr, M = d.M, Hi = d.Hi;
    const float* h_after = W.h_z + (size_t)B * Hr;       // h_z after step t, row-aligned with (t, b)
    // receiver
    b.add(km(W.dgi, d.G3), km(W.sen_feats, M), d.G3, M, R, MMG_P_REC_RNN_WIH, 0, MMG_P_REC_RNN_BIH);
    b.add(km(W.dgh, d.G3), km(W.h_z, Hr), d.G3, Hr, R, MMG_P_REC_RNN_WHH, 0, MMG_P_REC_RNN_BHH);
    b.add(km(W.d_lw, M), km(W.h_w, Hr), M, Hr, R, MMG_P_REC_W_W, 0, MMG_P_REC_W_B);
    b.add(km(W.d_hw, Hr), km(h_after, Hr), Hr, Hr, R, MMG_P_REC_WH_W, 0, MMG_P_REC_WH_B);
    b.add(km(W.d_hw, Hr), km(W.wd, d.WV), Hr, d.WV, R, MMG_P_REC_WD_W, 0, -1);
    b.add(km(W.d_ls, 1), km(h_after, Hr), 1, Hr, R, MMG_P_REC_S_W, 0, MMG_P_REC_S_B, WG_ROWVEC);
    b.add(km(W.g_h, Hr), km(W.hsel, Hr), Hr, Hr, B, MMG_P_REC_Y1_W, d.y1_hcol, -1);
    if (d.A) {
        // -desc_attn: the description rows of y1's input are the attended bags of words of the prediction step
        Operand bz = km(in.desc_set, d.WV);
        b.add(km(W.ddd_part + (size_t)d.NW * align4(d.A), align4(Hr)), bz, Hr, d.WV, d.NW, MMG_P_REC_Y1_W, d.y1_dcol, -1);   // Z^T . desc_set
        b.add(km(W.dy1, Hr), ones(), Hr, 1, B * d.D, MMG_P_REC_Y1_B, 0, -1);
        b.add(km(W.ddh, d.A), km(h_after, Hr), d.A, Hr, R, MMG_P_REC_DH_W, 0, MMG_P_REC_DH_B);
        b.add(km(W.dva, d.A), ones(), d.A, 1, R, MMG_P_REC_DA_W, 0, -1);                 // d_attn.weight (1, A)
        b.add(km(W.dba, 1), ones(), 1, 1, R, MMG_P_REC_DA_B, 0, -1);
        Operand bw = km(in.desc_set, d.WV);
        bw.mod = d.NW;                                       // row (cta, n) -> desc_set[n]
        b.add(km(W.ddd_part, align4(d.A)), bw, d.A, d.WV, d.NW, MMG_P_REC_DD_W, 0, MMG_P_REC_DD_B);   // slab 0 = sum over CTAs (K_attn_reduce)
    } else {
        // rows (b, d) of dy1 meet the same description row for every example: sum over the examples while staging (K = D instead
        // of B * D: one K-slice per tile, no split-K round trip for what was the longest reduction of the launch)
        Operand ad = km(W.dy1, Hr);
        ad.kind = OP_BSUM; ad.mod = B; ad.ld2 = d.D;
        b.add(ad, km(in.desc, d.WV), Hr, d.WV, d.D, MMG_P_REC_Y1_W, d.y1_dcol, MMG_P_REC_Y1_B);
    }
    b.add(km(W.dw2p, Hr), ones(), Hr, 1, B, MMG_P_REC_Y2_W, 0, -1);          // y2.weight (1, Hr): sum over examples
    b.add(km(W.g_outp, 1), ones(), 1, 1, B * d.D, MMG_P_REC_Y2_B, 0, -1);
    if (d.use_binary) {
        // sender
        b.add(km(W.d_lz, M), km(W.a_s, d.Ha), M, d.Ha, R, MMG_P_SEN_BIN_W, 0, MMG_P_SEN_BIN_B);
        if (!d.ignore_code || d.mix_mou) {      // -ignore_code without mou: code_layer / code_bias receive no gradient (model.py:208-210); they stay zero
            b.add(km(W.d_as, Hi), km(W.code_in, M), Hi, M, R, MMG_P_SEN_CODE_W, 0, MMG_P_SEN_CODE_B);
            if (fast) b.add(km(W.dcode_part, M), ones(), M, 1, B, MMG_P_SEN_CODE_BIAS, 0, -1, WG_GEMM, P.p[MMG_P_SEN_CODE_BIAS]);
            else {
                // d code_bias from the rows of step 0; with mou + ignore_code, d code_bias_mou from the rows of the later steps
                Operand rows0 = km(W.d_as, Hi);
                rows0.ld2 = 0; rows0.mod = B;
                b.add(rows0, km(W.d_as, Hi), M, 1, 1, MMG_P_SEN_CODE_BIAS, 0, -1, WG_CODEBIAS, P.p[MMG_P_SEN_CODE_BIAS]);
                if (d.mix_mou && d.ignore_code && R > B) {
                    Operand rows1 = km(W.d_as, Hi);
                    rows1.ld2 = B; rows1.mod = R;
                    b.add(rows1, km(W.d_as, Hi), M, 1, 1, MMG_P_SEN_CODE_BIAS_MOU, 0, -1, WG_CODEBIAS, P.p[MMG_P_SEN_CODE_BIAS_MOU]);
                }
            }
        }
        b.add(km(W.dhx, Hi), km(in.x, d.F), Hi, d.F, B, MMG_P_SEN_IMG_W, 0, MMG_P_SEN_IMG_B);
        // baseline_sen: d pre = g_bs * linear2.weight * (hidden > 0); rows [h_x[b] ; z_r[t,b]]
        {
            Operand a = km(W.h1s, d.Hb);
            a.kind = OP_RELUGRAD; a.g = W.g_bs; a.w2 = P.p[MMG_P_BS_L2_W];
            if (fast_bs) {
                // h_x is shared by the T rows of an example: sum the relu-gradient over t first (K = B instead of T*B),
                // the z_r columns keep the full row range.  The fast backward kernel has already left the t-summed matrix S (B, Hb)
                // in the U[b] array (a plain operand); behind the generic backward kernel the sum is formed while staging.
                Operand at = a;
                at.kind = OP_RELUGRAD_TSUM; at.mod = d.T; at.ld2 = B;
                if (fast) at = km(W.ubs, d.Hb);
                b.add(at, km(W.h_x, Hi), d.Hb, Hi, B, MMG_P_BS_L1_W, 0, MMG_P_BS_L1_B);
                b.add(a, km(W.rec_feats, M), d.Hb, M, R, MMG_P_BS_L1_W, Hi, -1);
            } else {
                Operand bb = km(W.h_x, Hi);
                bb.mod = B; bb.p2 = W.rec_feats; bb.ld2 = M; bb.split = Hi;
                b.add(a, bb, d.Hb, Hi + M, R, MMG_P_BS_L1_W, 0, MMG_P_BS_L1_B);
            }
            b.add(km(W.g_bs, 1), km(W.h1s, d.Hb), 1, d.Hb, R, MMG_P_BS_L2_W, 0, MMG_P_BS_L2_B, WG_ROWVEC);
        }
        // baseline_rec: rows [z[t,b] ; h_z after step t]
        {
            Operand a = km(W.h1r, d.Hb);
            a.kind = OP_RELUGRAD; a.g = W.g_br; a.w2 = P.p[MMG_P_BR_L2_W];
            Operand bb = km(W.sen_feats, M);
            bb.p2 = h_after; bb.ld2 = Hr; bb.split = M;
            b.add(a, bb, d.Hb, M + Hr, R, MMG_P_BR_L1_W, 0, MMG_P_BR_L1_B);
            b.add(km(W.g_br, 1), km(W.h1r, d.Hb), 1, d.Hb, R, MMG_P_BR_L2_W, 0, MMG_P_BR_L2_B, WG_ROWVEC);
        }
    }
    return b.finish();
}

// The loss gradients can be evaluated inside the fast backward kernel when both of its roles run (binary messages) and no
// -desc_attn backward (generic kernel) is involved.
static bool fuse_loss_ok(const mmg_config* c) {
    static int off = -1;
    if (off < 0) { const char* e = getenv("MMG_FUSE_LOSS"); off = (e && e[0] == '0') ? 1 : 0; }   // =0: separate K_lossgrad (tests)
    return !off && c->use_binary && !c->desc_attn;
}

static PeerView no_peers() {
    PeerView pv;
    memset(&pv, 0, sizeof(pv));
    return pv;
}
static int resolve_peers(const mmg_peers* p, int64_t step, PeerView* pv) {
    if (!p || p->world < 2 || p->world > MMG_MAX_PEERS || p->rank < 0 || p->rank >= p->world || !p->d_error)
        return fail(MMG_ERR_INVALID, "bad mmg_peers");
    memset(pv, 0, sizeof(*pv));
    pv->world = p->world; pv->rank = p->rank; pv->error = p->d_error; pv->iter = (unsigned long long)step;
    pv->send_mc = p->d_send_mc;
    for (int r = 0; r < p->world; ++r) {
        if (!p->d_send[r] || !p->d_recv[r] || !p->d_stats[r] || !p->d_norms[r] || !p->d_flags[r])
            return fail(MMG_ERR_INVALID, "null peer pointer (rank %d)", r);
        pv->send[r] = p->d_send[r]; pv->recv[r] = p->d_recv[r]; pv->stats[r] = p->d_stats[r]; pv->norms[r] = p->d_norms[r];
        pv->flags[r] = p->d_flags[r];
    }
    return MMG_OK;
}

static int upd_ctas(int64_t total) {
    int64_t n = cdiv64(total / 4, kUpdThreads);
    if (n > kNormCtasMax) n = kNormCtasMax;
    if (n < 1) n = 1;
    return (int)n;
}

// K_update: one float4 per thread (no per-CTA scratch limits its grid): every load of the pass — with peers, every NVLink pull —
// is in flight at once instead of one round trip per loop iteration
static int update_ctas(int64_t total) {
    int64_t n = cdiv64(total / 4, kUpdThreads);
    if (n > 4096) n = 4096;
    if (n < 1) n = 1;
    return (int)n;
}

static SegInfo seg_info(const mmg_param_layout& L, const Dims& d) {
    SegInfo s;
    for (int i = 0; i < 5; ++i) s.begin[i] = L.seg_begin[i];
    s.trained[0] = 1;
    s.trained[1] = s.trained[2] = s.trained[3] = d.use_binary ? 1 : 0;   // model.py:1313
    // tensors that receive no gradient at all are skipped by torch.optim (their .grad stays None):
    //  - the receiver message head (w_h, w_d, w) when no receiver-message loss exists this iteration
    //    (`len(rec_feats[:-1]) == 0`, model.py:1284-1289: one-step conversations);
    //  - the STOP head (s) when the exchange length is fixed (loss_binary_s not built, model.py:1278-1280,1299).
    s.whead_begin = L.offset[MMG_P_REC_WH_W];
    s.whead_end = L.offset[MMG_P_REC_W_B] + round_up64(L.rows[MMG_P_REC_W_B], 4);
    s.shead_begin = L.offset[MMG_P_REC_S_W];
    s.shead_end = L.offset[MMG_P_REC_S_B] + 4;
    s.shead_active = d.fixed ? 0 : 1;
    s.whead_stat = stat_idx(d, 1, 0, 0);
    return s;
}

}  // namespace host
}  // namespace mmg

using namespace mmg;
using namespace mmg::host;

extern "C" {

int mmg_abi_version(void) { return MMG_ABI_VERSION; }
const char* mmg_last_error(void) { return g_err; }
int mmg_launch_count(void) { return g_launches; }
void mmg_launch_count_reset(void) { g_launches = 0; }

int mmg_debug_trace(unsigned long long* out, int32_t count) {
#if defined(MMG_TRACE) && !defined(MMG_CPU_EMU)
    if (!out || count < 7 * 1024 * 8) return fail(MMG_ERR_INVALID, "trace buffer too small");
    if (cudaDeviceSynchronize() != cudaSuccess) return check_cuda("cudaDeviceSynchronize");
    if (cudaMemcpyFromSymbol(out, g_trace, sizeof(g_trace)) != cudaSuccess) return check_cuda("cudaMemcpyFromSymbol");
    static unsigned long long zero[7 * 1024 * 8];
    cudaMemcpyToSymbol(g_trace, zero, sizeof(g_trace));
    return 7 * 1024 * 8;
#else
    (void)out; (void)count;
    return 0;
#endif
}

int mmg_debug_kernel_times(char* out, int32_t cap) {
    if (!out || cap < 1) return fail(MMG_ERR_INVALID, "null output");
    out[0] = 0;
#ifndef MMG_CPU_EMU
    if (!ktime_enabled()) return 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return check_cuda("cudaDeviceSynchronize");
    struct Agg { const char* name; double us; int n; };
    Agg agg[64];
    int na = 0;
    for (int i = 0; i < g_ktime_n; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_ktime[i].a, g_ktime[i].b);
        cudaEventDestroy(g_ktime[i].a);
        cudaEventDestroy(g_ktime[i].b);
        int j = 0;
        while (j < na && agg[j].name != g_ktime[i].name) ++j;
        if (j == na) { if (na == 64) continue; agg[na++] = Agg{g_ktime[i].name, 0.0, 0}; }
        agg[j].us += 1e3 * ms;
        agg[j].n += 1;
    }
    const int n = g_ktime_n;
    g_ktime_n = 0;
    int pos = 0;
    for (int j = 0; j < na && pos < cap - 1; ++j)
        pos += snprintf(out + pos, (size_t)(cap - pos), "%s %d %.3f\n", agg[j].name, agg[j].n, agg[j].us / agg[j].n);
    return n;
#else
    return 0;
#endif
}

int mmg_device_count(void) {
#ifndef MMG_CPU_EMU
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { fail(MMG_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); return MMG_ERR_NO_DEVICE; }
    return n;
#else
    return 1;
#endif
}

int mmg_param_layout_get(const mmg_config* cfg, mmg_param_layout* out) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!out) return fail(MMG_ERR_INVALID, "null output");
    param_layout(make_dims(*cfg), out);
    return MMG_OK;
}

int mmg_workspace_layout_get(const mmg_config* cfg, mmg_workspace_layout* out) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!out) return fail(MMG_ERR_INVALID, "null output");
    Ws w;
    ws_layout(make_dims(*cfg), &w);
    *out = w.pub;
    return MMG_OK;
}

int mmg_workspace_init(const mmg_config* cfg, void* d_workspace, uint64_t seed, void* stream) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!d_workspace) return fail(MMG_ERR_INVALID, "null workspace");
    const Dims d = make_dims(*cfg);
    Ws w;
    ws_layout(d, &w);
    cudaStream_t st = (cudaStream_t)stream;
#ifndef MMG_CPU_EMU
    if (cudaMemsetAsync(d_workspace, 0, (size_t)w.pub.total_bytes, st) != cudaSuccess) return check_cuda("cudaMemsetAsync");
#else
    memset(d_workspace, 0, (size_t)w.pub.total_bytes);
#endif
    WsPtrs W = resolve(w, d_workspace);
    MMG_LAUNCH(k_init_rng, 1, 32, 0, st, W.rng_state, (unsigned long long)seed);
    return check_cuda("k_init_rng");
}

// `fuse`: non-null inside the fused training iteration: the per-example loss inputs are produced by the conversation kernel's
// epilogue and the batch statistics by the last CTA of K_baseline_fwd (with `*fuse` as the peers to publish them to); on
// return `*fused` tells whether that happened (fast kernels only) or K_stats still has to run.
static int exchange_forward_impl(const mmg_config* cfg, const float* d_params, const mmg_inputs* in, void* d_workspace,
                                 void* stream, bool finish_baselines, const PeerView* fuse = nullptr, bool* fused = nullptr) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!d_params || !in || !d_workspace || !in->d_x || !in->d_desc) return fail(MMG_ERR_INVALID, "null pointer argument");
    const Dims d = make_dims(*cfg);
    Plan pl;
    if ((rc = make_plan(d, &pl))) return rc;
    mmg_param_layout L;
    param_layout(d, &L);
    Ws w;
    ws_layout(d, &w);
    const WsPtrs W = resolve(w, d_workspace);
    const ParamPtrs P = param_ptrs(L, d_params);
    const ExchangeInputs ei = resolve_inputs(in);
    cudaStream_t st = (cudaStream_t)stream;
    // K_pre
    const int n_hx = cdiv(d.B, kTile) * cdiv(d.Hi, kTile) * W.hx_split;
    const int hx_kslice = round_up(cdiv(d.F, W.hx_split), 4);
    const int n_pack = 64;
    if (d.A && (!in->d_desc_set || !in->d_desc_set_lens)) return fail(MMG_ERR_INVALID, "desc_attn needs d_desc_set and d_desc_set_lens");
    const int n_cls = d.A ? cdiv(d.NW, kTile) * (2 * cdiv(d.Hr, kTile) + cdiv(d.A, kTile))      // word tables
                          : 2 * cdiv(d.D, kTile) * cdiv(d.Hr, kTile);                            // class tables
    // image-layer GEMM on the tensor cores (tcgen05, 3xTF32) when the shapes allow: 128-row tiles of hidden units, K-slices of
    // whole 32-float swizzle atoms, 16-byte aligned rows (MMG_UMMA=0 keeps the FFMA tiles)
    int use_umma = 0, pre_smem = 0, n_hx_l = n_hx;
    ImageTmaps tmaps;
    memset(&tmaps, 0, sizeof(tmaps));
#ifndef MMG_CPU_EMU
    {
        static int umma_off = -1, tma_off = -1;
        if (umma_off < 0) { const char* e = getenv("MMG_UMMA"); umma_off = (e && e[0] == '0') ? 1 : 0; }
        if (tma_off < 0) { const char* e = getenv("MMG_UMMA_TMA"); tma_off = (e && e[0] == '0') ? 1 : 0; }   // =0: cp.async staging
        if (!umma_off && d.Hi % umma::kM == 0 && d.F % 32 == 0 && hx_kslice % 32 == 0 && hx_kslice <= umma::kMaxSliceK &&
            (((size_t)in->d_x | (size_t)P.p[MMG_P_SEN_IMG_W]) & 15) == 0) {
            use_umma = 1;
            pre_smem = umma::tile_smem_bytes(hx_kslice);
            n_hx_l = (d.Hi / umma::kM) * cdiv(d.B, umma::kN) * W.hx_split;
            if ((rc = set_smem(k_pre, pre_smem))) return rc;
            // operands staged by the TMA unit (tensor maps: box = 32 floats x tile rows, 128-byte swizzle, zero fill past the batch)
            if (!tma_off && encode_kmajor_tmap(&tmaps.w, P.p[MMG_P_SEN_IMG_W], d.Hi, d.F, umma::kM) &&
                encode_kmajor_tmap(&tmaps.x, in->d_x, d.B, d.F, umma::kN))
                use_umma = 2;
        }
    }
#endif
    // image formats: bit 0 = fast forward image, bit 1 = fast backward image
    // the class / word tables run on the small-K row tile when the word-vector width allows (dynamic shared memory)
    const int cls_dyn = d.WV <= kRowsKMax ? rows_tile_smem_floats(d.WV) * 4 : 0;
    if (cls_dyn > pre_smem) { pre_smem = cls_dyn; if ((rc = set_smem(k_pre, pre_smem))) return rc; }
    MMG_LAUNCH(k_pre, n_hx_l + n_cls + n_pack, kGemmThreads, pre_smem, st, d, P, W, ei, n_hx_l, hx_kslice, pl.fast_fwd | (pl.fast << 1), n_cls,
               use_umma, pre_smem / 4, tmaps);
    if ((rc = check_cuda("k_pre"))) return rc;
    // K_exchange_fwd
    const AttnArgs aa = attn_args(d, P, ei, pl);
    const float* b_img = P.p[MMG_P_SEN_IMG_B];
    const int epi = (fuse != nullptr && pl.fast_fwd && in->train && in->d_target != nullptr) ? 1 : 0;
    if (fused != nullptr) *fused = epi != 0;
    if (pl.fast) rc = launch_fwd_fast(d, W, ei, FastFwdArgs{b_img, P.p[MMG_P_BS_L1_W], P.p[MMG_P_BS_L1_B], epi}, pl, st);
    else if (pl.fast_fwd) rc = launch_fwd_fast_attn(d, W, ei, FastFwdArgs{b_img, P.p[MMG_P_BS_L1_W], P.p[MMG_P_BS_L1_B], epi}, pl, st, aa);
    else switch (pl.BT) {
        case 1: rc = launch_fwd<1>(d, W, ei, b_img, pl, st, aa); break;
        case 2: rc = launch_fwd<2>(d, W, ei, b_img, pl, st, aa); break;
        case 4: rc = launch_fwd<4>(d, W, ei, b_img, pl, st, aa); break;
        default: rc = launch_fwd<8>(d, W, ei, b_img, pl, st, aa); break;
    }
    if (rc) return rc;
    if (in->train) {
        const int tiles = 2 * cdiv(d.R, kTile) * W.ntb;
        // wd rows as GEMM tiles (fast path and -desc_attn); otherwise the generic kernel writes wd itself
        const int wd_tiles = (pl.fast || d.A) ? cdiv(d.R, kTile) * cdiv(d.WV, kTile) : 0;
        // the small-K row tile stages [sen_feats ; h_z] x baseline_rec.linear1 rows whole: 2 x 64 rows x (M + Hr, padded) floats
        const int bas_k = d.M + d.Hr;
        const int bas_dyn = bas_k <= kRowsKMax ? rows_tile_smem_floats(bas_k) : 0;
        if ((rc = set_smem(k_baseline_fwd, bas_dyn * 4))) return rc;
        MMG_LAUNCH(k_baseline_fwd, tiles + wd_tiles, kGemmThreads, bas_dyn * 4, st, d, P, W, d.A ? ei.desc_set : ei.desc, tiles, pl.fast_fwd, epi, ei,
                   epi ? *fuse : no_peers(), *cfg, bas_dyn);
        if ((rc = check_cuda("k_baseline_fwd"))) return rc;
        if (finish_baselines) {     // standalone forward: bs / br must be final on return (mmg_loss re-derives them anyway)
            MMG_LAUNCH(k_baseline_finish, cdiv(d.R, 256), 256, 0, st, d, P, W);
            if ((rc = check_cuda("k_baseline_finish"))) return rc;
        }
    }
    return MMG_OK;
}

int mmg_exchange_forward(const mmg_config* cfg, const float* d_params, const mmg_inputs* in, void* d_workspace,
                         void* stream) {
    return exchange_forward_impl(cfg, d_params, in, d_workspace, stream, true);
}

static int loss_impl(const mmg_config* cfg, const float* d_params, const mmg_inputs* in, void* d_workspace, int phase,
                     void* stream, const PeerView& pv, bool stats_done = false) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!d_params || !in || !d_workspace || !in->d_target) return fail(MMG_ERR_INVALID, "null pointer argument");
    const Dims d = make_dims(*cfg);
    mmg_param_layout L;
    param_layout(d, &L);
    Ws w;
    ws_layout(d, &w);
    const WsPtrs W = resolve(w, d_workspace);
    const ParamPtrs P = param_ptrs(L, d_params);
    const ExchangeInputs ei = resolve_inputs(in);
    cudaStream_t st = (cudaStream_t)stream;
    if (phase <= 0 && !stats_done) {
        MMG_LAUNCH(k_stats, 1, kStatsThreadsMax, 0, st, d, P, W, ei, pv);
        if ((rc = check_cuda("k_stats"))) return rc;
    }
    if (phase != 0) {
        int ctas = cdiv(d.R, kLossThreads / 32);
        if (ctas > 4 * 148) ctas = 4 * 148;
        const int smem = loss_smem_bytes(d, pv.world);
        MMG_LAUNCH(k_lossgrad, ctas, kLossThreads, smem, st, d, *cfg, W, pv);
        if ((rc = check_cuda("k_lossgrad"))) return rc;
    }
    return MMG_OK;
}

int mmg_loss(const mmg_config* cfg, const float* d_params, const mmg_inputs* in, void* d_workspace, int phase,
             void* stream) {
    return loss_impl(cfg, d_params, in, d_workspace, phase, stream, no_peers());
}

// `fuse_loss`: the loss gradients (K_lossgrad) are evaluated inside the fast backward kernel; needs the batch statistics in
// the workspace (fused forward sequence) and use_binary (both backward roles run).
// `norm_tiles` (optional): the caller's K_update adds the per-tile sums of squares (and the loss partials) itself, so K_wgrad
// runs without a finishing CTA; receives the number of output tiles.
static int backward_impl(const mmg_config* cfg, const float* d_params, const mmg_inputs* in, void* d_workspace,
                         float* d_grads, void* stream, const PeerView& pv, bool fuse_loss = false, int* norm_tiles = nullptr,
                         int* loss_parts = nullptr) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!d_params || !in || !d_workspace || !d_grads) return fail(MMG_ERR_INVALID, "null pointer argument");
    const Dims d = make_dims(*cfg);
    Plan pl;
    if ((rc = make_plan(d, &pl))) return rc;
    mmg_param_layout L;
    param_layout(d, &L);
    Ws w;
    ws_layout(d, &w);
    const WsPtrs W = resolve(w, d_workspace);
    const ParamPtrs P = param_ptrs(L, d_params);
    const ExchangeInputs ei = resolve_inputs(in);
    cudaStream_t st = (cudaStream_t)stream;
    if (d.A && (!in->d_desc_set || !in->d_desc_set_lens)) return fail(MMG_ERR_INVALID, "desc_attn needs d_desc_set and d_desc_set_lens");
    const AttnArgs aa = attn_args(d, P, ei, pl);
    if (pl.fast) {
        const float *bw = P.p[MMG_P_SEN_BIN_W], *cw = P.p[MMG_P_SEN_CODE_W];
        const float* w2 = P.p[MMG_P_BS_L2_W];
        if (fuse_loss) rc = d.M == 32 ? launch_bwd_fast_m<32, true>(d, W, bw, cw, pl, st, *cfg, pv, w2) : launch_bwd_fast_m<64, true>(d, W, bw, cw, pl, st, *cfg, pv, w2);
        else           rc = d.M == 32 ? launch_bwd_fast_m<32, false>(d, W, bw, cw, pl, st, *cfg, pv, w2) : launch_bwd_fast_m<64, false>(d, W, bw, cw, pl, st, *cfg, pv, w2);
    }
    else switch (pl.BT) {
        case 1: rc = launch_bwd<1>(d, W, pl, st, aa); break;
        case 2: rc = launch_bwd<2>(d, W, pl, st, aa); break;
        case 4: rc = launch_bwd<4>(d, W, pl, st, aa); break;
        default: rc = launch_bwd<8>(d, W, pl, st, aa); break;
    }
    if (rc) return rc;
    if (d.A) {
        const int slab_f4 = d.NW * (align4(d.A) + align4(d.Hr)) / 4;
        int ctas = cdiv(slab_f4, 32);
        if (ctas > 592) ctas = 592;
        MMG_LAUNCH(k_attn_reduce, ctas, 256, 0, st, W.ddd_part, cdiv(d.B, pl.BT), slab_f4);
        if ((rc = check_cuda("k_attn_reduce"))) return rc;
    }
    WgTable tab;
    const int trc = build_wgrad_table(d, L, P, W, ei, pl.fast, pl.fast_fwd, cdiv(d.B, pl.BT), &tab);
    if (trc) return fail(MMG_ERR_UNSUPPORTED, "weight-gradient problem table (%d): too many problems / tiles for these dimensions", trc);
    // tensors of untrained modules keep a zero gradient: their problems are not in the table, so they are cleared there
    const WgSync sy{W.tile_tickets, W.tile_norm, W.tickets + 3, W.norm_final};
    if ((rc = set_smem(k_wgrad, wgrad_smem_bytes()))) return rc;
    const int n_loss_parts = (fuse_loss && pl.fast) ? (d.use_binary ? 2 * d.B : d.B) : 0;
    const int defer = norm_tiles != nullptr && pv.world <= 1 ? 1 : 0;
    // data-parallel: the finishing CTA only raises the "send buffer complete" flags; K_update adds up the loss partials there too
    if (norm_tiles != nullptr) { *norm_tiles = defer ? tab.total_out : 0; *loss_parts = (defer || pv.world > 1) ? n_loss_parts : 0; }
    MMG_LAUNCH(k_wgrad, tab.total_tiles, kGemmThreads, wgrad_smem_bytes(), st, d, tab, d_grads, W.slabs, P.p[MMG_P_SEN_CODE_W],
               P.p[MMG_P_SEN_CODE_BIAS], W.d_as, sy, pv, W, n_loss_parts, defer);
    return check_cuda("k_wgrad");
}

int mmg_backward(const mmg_config* cfg, const float* d_params, const mmg_inputs* in, void* d_workspace, float* d_grads,
                 void* stream) {
    return backward_impl(cfg, d_params, in, d_workspace, d_grads, stream, no_peers());
}

int mmg_grad_norm(const mmg_config* cfg, float* d_grads, void* d_workspace, void* stream) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!d_workspace || !d_grads) return fail(MMG_ERR_INVALID, "null pointer argument");
    const Dims d = make_dims(*cfg);
    mmg_param_layout L;
    param_layout(d, &L);
    Ws w;
    ws_layout(d, &w);
    const WsPtrs W = resolve(w, d_workspace);
    const SegInfo seg = seg_info(L, d);
    MMG_LAUNCH(k_grad_norm, upd_ctas(L.total), kUpdThreads, 0, (cudaStream_t)stream, seg, (const float*)d_grads, W.norm_part,
               W.tickets + 6, W.norm_final);
    return check_cuda("k_grad_norm");
}

static int clip_update_impl(const mmg_config* cfg, float* d_params, float* d_grads, float* d_state1, float* d_state2,
                            int64_t step, float grad_scale, void* d_workspace, void* stream, int norm_tiles, int loss_parts,
                            float* h_losses_out = nullptr) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!d_params || !d_grads || !d_workspace) return fail(MMG_ERR_INVALID, "null pointer argument");
    if (cfg->optim_type != MMG_OPT_SGD && !d_state1) return fail(MMG_ERR_INVALID, "optimizer state required");
    if (cfg->optim_type == MMG_OPT_ADAM && !d_state2) return fail(MMG_ERR_INVALID, "Adam needs d_state2");
    if (grad_scale != 1.0f) return fail(MMG_ERR_UNSUPPORTED, "grad_scale != 1: gradients are already normalised by batch_global");
    const Dims d = make_dims(*cfg);
    mmg_param_layout L;
    param_layout(d, &L);
    Ws w;
    ws_layout(d, &w);
    const WsPtrs W = resolve(w, d_workspace);
    const SegInfo seg = seg_info(L, d);
    OptHyper hp;
    hp.optim = cfg->optim_type; hp.lr = cfg->learning_rate; hp.max_norm = cfg->max_norm; hp.step = step;
    MMG_LAUNCH(k_update, update_ctas(L.total) + (loss_parts > 0 ? 1 : 0), kUpdThreads, 0, (cudaStream_t)stream, seg, hp, d_params, (const float*)d_grads, d_grads,
               d_state1, d_state2, W.norm_final, W.grad_norms, (const double*)W.stats,
               (const long long*)W.opt_counters, no_peers(), (const float*)W.tile_norm, norm_tiles, loss_parts, d, W,
               loss_parts > 0 ? h_losses_out : nullptr);
    return check_cuda("k_update");
}

int mmg_clip_update(const mmg_config* cfg, float* d_params, float* d_grads, float* d_state1, float* d_state2,
                    int64_t step, float grad_scale, void* d_workspace, void* stream) {
    return clip_update_impl(cfg, d_params, d_grads, d_state1, d_state2, step, grad_scale, d_workspace, stream, 0, 0);
}

int mmg_train_step(const mmg_config* cfg, float* d_params, float* d_grads, float* d_state1, float* d_state2,
                   int64_t step, const mmg_inputs* in, void* d_workspace, void* stream) {
    int rc;
    if (!in || !in->train) return fail(MMG_ERR_INVALID, "mmg_train_step needs in->train = 1");
    const PeerView none = no_peers();
    bool fused = false;
    if ((rc = exchange_forward_impl(cfg, d_params, in, d_workspace, stream, false, &none, &fused))) return rc;
    const bool fuse_loss = fused && fuse_loss_ok(cfg);
    if (!fuse_loss && (rc = loss_impl(cfg, d_params, in, d_workspace, -1, stream, none, fused))) return rc;
    int norm_tiles = 0, loss_parts = 0;
    if ((rc = backward_impl(cfg, d_params, in, d_workspace, d_grads, stream, none, fuse_loss, &norm_tiles, &loss_parts))) return rc;
    if ((rc = clip_update_impl(cfg, d_params, d_grads, d_state1, d_state2, step, 1.0f, d_workspace, stream, norm_tiles, loss_parts,
                               in->h_losses_out))) return rc;
#ifndef MMG_CPU_EMU
    if (in->h_losses_out != nullptr && loss_parts == 0) {
        // the loss values were finalised by an earlier kernel of this sequence (no fused loss path for this configuration):
        // deliver them with an ordinary copy
        Ws w;
        ws_layout(make_dims(*cfg), &w);
        if (cudaMemcpyAsync(in->h_losses_out, (char*)d_workspace + w.pub.losses, MMG_LOSS_COUNT * sizeof(float), cudaMemcpyDeviceToHost,
                            (cudaStream_t)stream) != cudaSuccess)
            return check_cuda("D2H losses");
    }
#endif
    return MMG_OK;
}

int mmg_train_step_host(const mmg_config* cfg, float* d_params, float* d_grads, float* d_state1, float* d_state2,
                        int64_t step, const float* h_x, const int64_t* h_target, const float* h_desc,
                        float* d_x_stage, int64_t* d_target_stage, float* d_desc_stage, const mmg_inputs* in,
                        void* d_workspace, float* h_losses, void* stream) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!h_x || !h_target || !d_x_stage || !d_target_stage || !in || !h_losses) return fail(MMG_ERR_INVALID, "null pointer argument");
#ifndef MMG_CPU_EMU
    cudaStream_t st = (cudaStream_t)stream;
    const Dims d = make_dims(*cfg);
    if (cudaMemcpyAsync(d_x_stage, h_x, (size_t)d.B * d.F * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess)
        return check_cuda("H2D x");
    if (cudaMemcpyAsync(d_target_stage, h_target, (size_t)d.B * sizeof(int64_t), cudaMemcpyHostToDevice, st) != cudaSuccess)
        return check_cuda("H2D target");
    if (h_desc != nullptr) {
        if (!d_desc_stage) return fail(MMG_ERR_INVALID, "d_desc_stage required with h_desc");
        if (cudaMemcpyAsync(d_desc_stage, h_desc, (size_t)d.D * d.WV * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess)
            return check_cuda("H2D desc");
    }
    mmg_inputs dev = *in;
    dev.d_x = d_x_stage;
    dev.d_target = d_target_stage;
    if (h_desc != nullptr) dev.d_desc = d_desc_stage;
    if ((rc = mmg_train_step(cfg, d_params, d_grads, d_state1, d_state2, step, &dev, d_workspace, stream))) return rc;
    Ws w;
    ws_layout(d, &w);
    if (cudaMemcpyAsync(h_losses, (char*)d_workspace + w.pub.losses, MMG_LOSS_COUNT * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return check_cuda("D2H losses");
    return MMG_OK;
#else
    (void)d_params; (void)d_grads; (void)d_state1; (void)d_state2; (void)step; (void)h_desc; (void)d_desc_stage; (void)d_workspace; (void)stream;
    return fail(MMG_ERR_UNSUPPORTED, "host-buffer entry point is not part of the emulation build");
#endif
}

static int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

int mmg_peer_buffer_layout(const mmg_config* cfg, int64_t* total_bytes, int64_t* send_off, int64_t* recv_off, int64_t* stats_off,
                           int64_t* norms_off, int64_t* flags_off) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!total_bytes || !send_off || !recv_off || !stats_off || !norms_off || !flags_off) return fail(MMG_ERR_INVALID, "null output");
    const Dims d = make_dims(*cfg);
    mmg_param_layout L;
    param_layout(d, &L);
    *send_off = 0;
    *recv_off = align256(L.total * 4);
    *stats_off = *recv_off + align256(L.total * 4);
    *norms_off = *stats_off + align256((int64_t)MMG_MAX_PEERS * stats_count(d) * 16);    // one slot per pushing rank, 2 packets per value
    *flags_off = *norms_off + align256(4 * MMG_MAX_PEERS * 16);
    *total_bytes = *flags_off + align256(3 * MMG_MAX_PEERS * 8);
    return MMG_OK;
}

int mmg_train_step_peer(const mmg_config* cfg, float* d_params, float* d_grads, float* d_state1, float* d_state2,
                        int64_t step, const mmg_inputs* in, void* d_workspace, const mmg_peers* peers, void* stream) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!in || !in->train) return fail(MMG_ERR_INVALID, "mmg_train_step_peer needs in->train = 1");
    if (step < 1) return fail(MMG_ERR_INVALID, "step must start at 1 and increase by one per call (it is the flag value)");
    if (!d_params || !d_grads) return fail(MMG_ERR_INVALID, "null pointer argument");
    if (cfg->optim_type != MMG_OPT_SGD && !d_state1) return fail(MMG_ERR_INVALID, "optimizer state required");
    if (cfg->optim_type == MMG_OPT_ADAM && !d_state2) return fail(MMG_ERR_INVALID, "Adam needs d_state2");
    PeerView pv;
    if ((rc = resolve_peers(peers, step, &pv))) return rc;
    bool fused = false;
    if ((rc = exchange_forward_impl(cfg, d_params, in, d_workspace, stream, false, &pv, &fused))) return rc;
    const bool fuse_loss = fused && fuse_loss_ok(cfg);
    if (!fuse_loss && (rc = loss_impl(cfg, d_params, in, d_workspace, -1, stream, pv, fused))) return rc;
    // local gradient -> this rank's symmetric send buffer (K_wgrad raises flag row 1 when it is complete)
    int norm_tiles = 0, loss_parts = 0;
    if ((rc = backward_impl(cfg, d_params, in, d_workspace, pv.send[pv.rank], stream, pv, fuse_loss, &norm_tiles, &loss_parts))) return rc;
    const Dims d = make_dims(*cfg);
    mmg_param_layout L;
    param_layout(d, &L);
    Ws w;
    ws_layout(d, &w);
    const WsPtrs W = resolve(w, d_workspace);
    const SegInfo seg = seg_info(L, d);
    cudaStream_t st = (cudaStream_t)stream;
    // two-shot sum: this rank reduces its 1/G slice of all send buffers into its own receive buffer (+ the slice's norms to all) ...
    // one float4 per thread: every load of the slice is in flight at once
    MMG_LAUNCH(k_peer_reduce_scatter, upd_ctas(cdiv64(L.total, pv.world)), kUpdThreads, 0, st, seg, pv, W.norm_part, W.tickets + 6,
               W.norm_final);
    if ((rc = check_cuda("k_peer_reduce_scatter"))) return rc;
    // ... and the update waits for all slices (flag row 2), then clips and steps, pulling every slice from its owner
    OptHyper hp;
    hp.optim = cfg->optim_type; hp.lr = cfg->learning_rate; hp.max_norm = cfg->max_norm; hp.step = step;
    MMG_LAUNCH(k_update, update_ctas(L.total) + (loss_parts > 0 ? 1 : 0), kUpdThreads, 0, st, seg, hp, d_params, (const float*)pv.recv[pv.rank], d_grads, d_state1,
               d_state2, W.norm_final, W.grad_norms, (const double*)W.stats, (const long long*)W.opt_counters, pv,
               (const float*)W.tile_norm, 0, loss_parts, d, W, loss_parts > 0 ? in->h_losses_out : nullptr);
    if ((rc = check_cuda("k_update"))) return rc;
#ifndef MMG_CPU_EMU
    if (in->h_losses_out != nullptr && loss_parts == 0 &&
        cudaMemcpyAsync(in->h_losses_out, (char*)d_workspace + w.pub.losses, MMG_LOSS_COUNT * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return check_cuda("D2H losses");
#endif
    return MMG_OK;
}

int mmg_host_prefetch(const mmg_config* cfg, const float* h_x, const int64_t* h_target, float* d_x_stage,
                      int64_t* d_target_stage, void* copy_stream, void* ev_free, void* ev_ready) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!h_x || !h_target || !d_x_stage || !d_target_stage || !ev_ready) return fail(MMG_ERR_INVALID, "null pointer argument");
#ifndef MMG_CPU_EMU
    cudaStream_t cs = (cudaStream_t)copy_stream;
    const Dims d = make_dims(*cfg);
    if (ev_free != nullptr && cudaStreamWaitEvent(cs, (cudaEvent_t)ev_free, 0) != cudaSuccess) return check_cuda("wait ev_free");
    if (cudaMemcpyAsync(d_x_stage, h_x, (size_t)d.B * d.F * sizeof(float), cudaMemcpyHostToDevice, cs) != cudaSuccess)
        return check_cuda("H2D x");
    if (cudaMemcpyAsync(d_target_stage, h_target, (size_t)d.B * sizeof(int64_t), cudaMemcpyHostToDevice, cs) != cudaSuccess)
        return check_cuda("H2D target");
    if (cudaEventRecord((cudaEvent_t)ev_ready, cs) != cudaSuccess) return check_cuda("record ev_ready");
    return MMG_OK;
#else
    (void)copy_stream; (void)ev_free;
    return fail(MMG_ERR_UNSUPPORTED, "host-buffer entry point is not part of the emulation build");
#endif
}

int mmg_train_step_staged(const mmg_config* cfg, float* d_params, float* d_grads, float* d_state1, float* d_state2,
                          int64_t step, const mmg_inputs* in, void* d_workspace, float* h_losses, void* stream,
                          void* ev_ready, void* ev_free) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!in || (!h_losses && !in->h_losses_out) || !ev_ready || !ev_free) return fail(MMG_ERR_INVALID, "null pointer argument");
#ifndef MMG_CPU_EMU
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaStreamWaitEvent(st, (cudaEvent_t)ev_ready, 0) != cudaSuccess) return check_cuda("wait ev_ready");
    if ((rc = mmg_train_step(cfg, d_params, d_grads, d_state1, d_state2, step, in, d_workspace, stream))) return rc;
    if (cudaEventRecord((cudaEvent_t)ev_free, st) != cudaSuccess) return check_cuda("record ev_free");
    Ws w;
    ws_layout(make_dims(*cfg), &w);
    if (h_losses != nullptr &&
        cudaMemcpyAsync(h_losses, (char*)d_workspace + w.pub.losses, MMG_LOSS_COUNT * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return check_cuda("D2H losses");
    return MMG_OK;
#else
    (void)d_params; (void)d_grads; (void)d_state1; (void)d_state2; (void)step; (void)d_workspace; (void)stream;
    return fail(MMG_ERR_UNSUPPORTED, "host-buffer entry point is not part of the emulation build");
#endif
}

int mmg_sender_forward(const mmg_config* cfg, const float* d_params, int32_t rows, const float* d_x, const float* d_w,
                       int32_t t, int32_t train, const double* d_u, const double* d_u_flip, uint64_t seed, uint64_t counter,
                       float* d_msg, float* d_probs, float* d_h_x, void* stream) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!d_params || !d_x || !d_msg || !d_h_x || rows < 1) return fail(MMG_ERR_INVALID, "null pointer argument");
    if (t != 0 && !d_w) return fail(MMG_ERR_INVALID, "Sender.forward at t > 0 needs the receiver's message w");
    if (cfg->use_binary && !d_probs) return fail(MMG_ERR_INVALID, "d_probs required with use_binary");
    const Dims d = make_dims(*cfg);
    mmg_param_layout L;
    param_layout(d, &L);
    const ParamPtrs P = param_ptrs(L, d_params);
    const int smem = sender_step_smem_floats(d) * 4;
    if (smem > kMaxSmem) return fail(MMG_ERR_UNSUPPORTED, "img_feat_dim=%d too large for the single-turn sender kernel", d.F);
    if ((rc = set_smem(k_sender_step, smem))) return rc;
    MMG_LAUNCH(k_sender_step, rows, kSingleThreads, smem, (cudaStream_t)stream, d, P, d_x, d_w, (int)t, (int)train, d_u, d_u_flip,
               (unsigned long long)seed, (unsigned long long)counter, d_msg, d_probs, d_h_x);
    return check_cuda("k_sender_step");
}

int mmg_receiver_forward(const mmg_config* cfg, const float* d_params, int32_t rows, const float* d_z, const float* d_desc,
                         const float* d_desc_set, const int32_t* d_desc_set_lens, float* d_h_z, float* d_s_prob_prod,
                         int32_t first, int32_t train, const double* d_u_stop, const double* d_u_rec, const double* d_u_flip,
                         uint64_t seed, uint64_t counter, float* d_s, float* d_s_prob, float* d_w, float* d_w_probs,
                         float* d_y, float* d_h_w, void* stream) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!d_params || !d_z || !d_desc || !d_h_z || !d_s || !d_s_prob || !d_w || !d_y || !d_h_w || rows < 1)
        return fail(MMG_ERR_INVALID, "null pointer argument");
    if (!train && !d_s_prob_prod) return fail(MMG_ERR_INVALID, "eval mode needs d_s_prob_prod (model.py:423-427)");
    if (cfg->use_binary && !d_w_probs) return fail(MMG_ERR_INVALID, "d_w_probs required with use_binary");
    if (cfg->desc_attn && (!d_desc_set || !d_desc_set_lens)) return fail(MMG_ERR_INVALID, "desc_attn needs d_desc_set and d_desc_set_lens");
    const Dims d = make_dims(*cfg);
    mmg_param_layout L;
    param_layout(d, &L);
    const ParamPtrs P = param_ptrs(L, d_params);
    ReceiverStepIO io;
    io.z = d_z; io.desc = d_desc; io.desc_set = d_desc_set; io.desc_set_lens = d_desc_set_lens; io.h_z = d_h_z;
    io.s_prob_prod = d_s_prob_prod; io.u_stop = d_u_stop; io.u_rec = d_u_rec; io.u_flip = d_u_flip;
    io.s = d_s; io.s_prob = d_s_prob; io.w = d_w; io.w_probs = d_w_probs; io.y = d_y; io.h_w = d_h_w;
    io.first = first; io.train = train; io.seed = seed; io.counter = counter;
    const int smem = receiver_step_smem_floats(d) * 4;
    if (smem > kMaxSmem) return fail(MMG_ERR_UNSUPPORTED, "single-turn receiver kernel: %d bytes of shared memory", smem);
    if ((rc = set_smem(k_receiver_step, smem))) return rc;
    MMG_LAUNCH(k_receiver_step, rows, kSingleThreads, smem, (cudaStream_t)stream, d, P, io);
    return check_cuda("k_receiver_step");
}

int mmg_baseline_forward(const mmg_config* cfg, const float* d_params, int32_t which, int32_t rows, const float* d_x,
                         int32_t x_dim, const float* d_binary, int32_t binary_dim, const float* d_inp, int32_t inp_dim,
                         float* d_out, void* stream) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (!d_params || !d_out || rows < 1) return fail(MMG_ERR_INVALID, "null pointer argument");
    if (which != MMG_SEG_BASELINE_SEN && which != MMG_SEG_BASELINE_REC) return fail(MMG_ERR_INVALID, "which=%d", which);
    if ((x_dim > 0 && !d_x) || (binary_dim > 0 && !d_binary) || (inp_dim > 0 && !d_inp) || x_dim < 0 || binary_dim < 0 || inp_dim < 0)
        return fail(MMG_ERR_INVALID, "piece pointer / width mismatch");
    const Dims d = make_dims(*cfg);
    mmg_param_layout L;
    param_layout(d, &L);
    const ParamPtrs P = param_ptrs(L, d_params);
    const int w1 = which == MMG_SEG_BASELINE_SEN ? MMG_P_BS_L1_W : MMG_P_BR_L1_W;
    if (x_dim + binary_dim + inp_dim != L.cols[w1])
        return fail(MMG_ERR_INVALID, "input width %d != linear1 width %d", x_dim + binary_dim + inp_dim, L.cols[w1]);
    const int smem = align4(L.cols[w1]) * 4;
    MMG_LAUNCH(k_baseline_step, rows, kSingleThreads, smem, (cudaStream_t)stream, P.p[w1], P.p[w1 + 1], P.p[w1 + 2], P.p[w1 + 3], d.Hb,
               d_x, (int)x_dim, d_binary, (int)binary_dim, d_inp, (int)inp_dim, d_out);
    return check_cuda("k_baseline_step");
}

}  // extern "C"
