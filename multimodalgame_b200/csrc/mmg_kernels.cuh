// Kernel-side shared declarations: resolved pointer tables and launch helpers.
#pragma once
#include "mmg_layout.h"
#include "mmg_gemm.cuh"

namespace mmg {

struct ParamPtrs {            // one pointer per state_dict tensor (params, or grads / slabs with the same layout)
    const float* p[MMG_P_COUNT];
};

struct WsPtrs {               // resolved workspace arrays (device pointers)
    float *sen_feats, *sen_probs, *rec_feats, *rec_probs, *stop_feat, *stop_prob, *y, *bs, *br, *h_x, *h_z, *h_w;
    unsigned char* stop_mask;
    float *losses, *outp, *logs, *grad_norms;
    int *ystep, *argmax;
    double* stats;
    float *g_sen_probs, *g_rec_probs, *g_stop_prob, *g_outp, *g_bs, *g_br;
    unsigned long long* rng_state;
    float *code_in, *a_s, *hw_s, *gates, *y1h, *q, *wd, *h1s, *h1r, *bs_part, *br_part, *ubs;
    float *hx_part, *fwd_image, *bwd_image;
    float *d_lz, *d_as, *dhx, *dgi, *dgh, *d_lw, *d_hw, *d_ls, *g_h, *hsel, *dy1, *dw2p, *dcode_part, *slabs, *norm_part, *tile_norm, *hit, *coefs;
    unsigned* tile_tickets;
    double* norm_final;
    double* loss_part;
    unsigned* tickets;
    long long* opt_counters;
    // -desc_attn
    float *wtab_dd, *wtab_y1, *wtab_wd, *qa, *attn, *dh_s, *ddh, *dva, *dba, *ddd_part;
    int *seg, *wcls;
    int hx_split, wgrad_split, ntb;
};

struct PeerView {             // mmg_peers resolved for the kernels; world <= 1: single-rank run, nothing is touched
    int world, rank;
    float* send[MMG_MAX_PEERS];     // local gradients (read by the peers)
    float* send_mc;                 // multicast address of the send buffers (in-switch reduction), or nullptr
    float* recv[MMG_MAX_PEERS];     // global gradient, written slice by slice by the rank that owns the slice
    double* stats[MMG_MAX_PEERS];
    double* norms[MMG_MAX_PEERS];   // [world][4] per-slice sums of squares of the global gradient
    unsigned long long* flags[MMG_MAX_PEERS];   // [3][MMG_MAX_PEERS]: statistics / send buffer / slice published
    int* error;
    unsigned long long iter;
};

struct ExchangeInputs {       // device pointers of one exchange (mmg_inputs resolved)
    const float* x;
    const float* desc;
    const long long* target;
    const double *u_sen, *u_stop, *u_rec, *u_flip_sen, *u_flip_rec;
    const float* corrupt_mask;
    const float* h0;
    int top_k, train;
    const float* desc_set;     // (NW,WV) -desc_attn
    const int* desc_set_lens;  // (D)
};

struct AttnArgs {             // -desc_attn parameters read straight from the flat parameter buffer (all L2 resident)
    const float* dh_w;         // d_h.weight (A,Hr)
    const float* dh_b;         // d_h.bias (A)
    const float* va;           // d_attn.weight (A)
    const float* ba;           // d_attn.bias (1)
    const float* b1;           // y1.bias (Hr)
    const float* desc_set;     // (NW,WV)
    int acc_smem;              // backward: the running d (d_d(word)) sums live in shared memory (they fit beside the weights)
};

}  // namespace mmg

// ---- launch macro --------------------------------------------------------------------------------------
#ifndef MMG_CPU_EMU
#include <utility>
#include <stdlib.h>
namespace mmg {
namespace host {
// Every launch carries the programmatic-stream-serialization attribute (PDL): the next kernel's CTAs are scheduled while
// the previous grid drains and block in `griddepcontrol.wait` (first statement of every kernel) until it has completed
// and flushed, which hides launch + CTA-scheduling latency at each of the 8 kernel boundaries of an iteration.
// MMG_PDL=0 disables the attribute (plain stream order).
inline int pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MMG_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}
// MMG_KTIME=1 (diagnostic, see mmg_debug_kernel_times): every launch is bracketed by two CUDA events on its stream, without
// PDL, so that per-kernel device times can be read IN SITU (inside the caller's own cache / flush protocol) instead of
// from a profiler's serialised cold-cache replays.
int ktime_enabled();
void ktime_begin(const char* name, cudaStream_t st);
void ktime_end(cudaStream_t st);
template <typename... KArgs, typename... Args>
inline void launch(const char* name, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    const int kt = ktime_enabled();
    cfg.numAttrs = (pdl_enabled() && !kt) ? 1 : 0;
    if (kt) ktime_begin(name, st);
    cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
    if (kt) ktime_end(st);
}
}  // namespace host
}  // namespace mmg
#define MMG_LAUNCH(kernel, grid, block, smem, stream, ...)                                          \
    do {                                                                                             \
        mmg::host::launch(#kernel, kernel, dim3(grid), dim3(block), (size_t)(smem), (stream), __VA_ARGS__);  \
        mmg::host::count_launch();                                                                   \
    } while (0)
#else
#define MMG_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                      \
        mmg::emu::launch(dim3(grid), dim3(block), (smem), [&]() { kernel(__VA_ARGS__); });    \
        mmg::host::count_launch();                                                            \
    } while (0)
#endif

namespace mmg {
namespace host {
void count_launch();
}
}  // namespace mmg
