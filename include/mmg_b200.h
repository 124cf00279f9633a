/*
 * mmg_b200.h — C-ABI of the B200-native referential-game hot path (libmmg_b200.so).
 *
 * Drop-in boundary: the reference (nyu-dl/MultimodalGame) has no FFI layer; its operator
 * API for this path is the Python surface of model.py.  The host-side mirror of that surface lives in
 * multimodalgame_b200/model.py (Sender / Receiver / Baseline / exchange / loss functions, same names and
 * argument meaning); underneath it, every tensor operation of the training iteration is executed by the
 * entry points declared here.  Each entry point cites the reference code it replaces (file:line into
 * /root/reference).
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types in any signature;
 *   - all `void* d_*` / `float* d_*` arguments are DEVICE pointers into caller-owned storage (the caller
 *     keeps them alive until the stream has drained); `h_*` arguments are HOST pointers (pinned for
 *     asynchronous copies);
 *   - nothing is allocated inside the library; `mmg_workspace_layout` tells the caller how much scratch
 *     to provide and where each named array lives in it;
 *   - every function returns 0 on success or a negative mmg_status; nothing throws across the ABI;
 *     `mmg_last_error()` returns a static string describing the last failure on the calling thread;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), stream-ordered, no host sync
 *     unless stated; the sequences are CUDA-graph capturable.
 */
#ifndef MMG_B200_H_
#define MMG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMG_ABI_VERSION 6

typedef enum mmg_status {
    MMG_OK = 0,
    MMG_ERR_INVALID = -1,      /* bad argument / unsupported dimension */
    MMG_ERR_CUDA = -2,         /* a CUDA runtime call failed (see mmg_last_error) */
    MMG_ERR_UNSUPPORTED = -3,  /* flag combination outside the fused path */
    MMG_ERR_NO_DEVICE = -4     /* no CUDA device: the library never falls back to the CPU */
} mmg_status;

enum { MMG_OPT_RMSPROP = 0, MMG_OPT_ADAM = 1, MMG_OPT_SGD = 2 }; /* model.py:1111-1137 */
enum { MMG_MIX_SUM = 0, MMG_MIX_PROD = 1, MMG_MIX_MOU = 2 };

/* Flags that shape the path.  Field names follow the reference's gflags (model.py:1641-1741). */
typedef struct mmg_config {
    int32_t batch;          /* rows handled by THIS rank (FLAGS.batch_size / world size)                  */
    int32_t batch_global;   /* global batch: denominators of every mean (== batch on a single GPU)         */
    int32_t img_feat_dim;   /* F   model.py:1693 */
    int32_t img_h_dim;      /* Hi  model.py:1694 */
    int32_t msg_dim;        /* M = rec_w_dim = sender_out_dim (asserted equal, model.py:1756) */
    int32_t rec_hidden;     /* Hr  model.py:1697 */
    int32_t n_classes;      /* D   rows of `desc` */
    int32_t wv_dim;         /* WV  model.py:1674 */
    int32_t baseline_hid;   /* Hb  model.py:1695 */
    int32_t max_exchange;   /* T   model.py:1736 */
    int32_t use_binary;     /* model.py:1701 */
    int32_t fixed_exchange; /* model.py:1737 */
    int32_t s_prob_prod;    /* model.py:1713 (eval only) */
    int32_t optim_type;     /* MMG_OPT_* */
    int32_t has_entropy_s, has_entropy_sen, has_entropy_rec; /* 0 when the flag is None (model.py:925) */
    float entropy_s, entropy_sen, entropy_rec;                /* model.py:1730-1732 */
    float first_rec;        /* model.py:1709 */
    float learning_rate;    /* model.py:1728 */
    float max_norm;         /* clip_grad_norm(..., max_norm=1.) model.py:1310 */
    int32_t ignore_receiver; /* model.py:1703,470-472 */
    int32_t has_flipout_sen, has_flipout_rec; /* 0 when the flag is None (model.py:1710-1711) */
    int32_t flipout_dev;     /* model.py:1712: flip in eval mode too */
    float flipout_sen, flipout_rec;           /* bit-flip probabilities (model.py:233-234,467-468,554-568) */
    int32_t sender_mix;      /* MMG_MIX_SUM: tanh(h_x + h_w); MMG_MIX_PROD: tanh(h_x * h_w); MMG_MIX_MOU: tanh([h_x ; h_w ; h_x - h_w ; h_x * h_w]), binary_layer 4 Hi wide (model.py:1692,71-76,208-221) */
    int32_t ignore_code;     /* model.py:1704,208-213: the sender ignores the receiver's message, hidden = tanh(h_x) */
    int32_t desc_attn;       /* model.py:1719,344-410: the receiver attends over the words of each class description */
    int32_t desc_attn_dim;   /* A   model.py:1720 */
    int32_t n_words;         /* NW  rows of `desc_set` (sum of desc_set_lens); only read when desc_attn != 0 */
    int32_t batch_offset;    /* index of this rank's first row inside the global batch (rank * batch in a data-parallel run, else 0):
                                the on-device sampler keys its Philox counters by the GLOBAL row, so the shards of one global batch
                                draw independent noise and a G-way run consumes the same stream as a single-GPU run of the global batch */
} mmg_config;

/* ---- parameter layout -------------------------------------------------------------------------------
 * One flat fp32 buffer holds the four modules' parameters, in optimizer order (model.py:1308-1330):
 * receiver | sender | baseline_rec | baseline_sen.  Every tensor keeps the reference's state_dict name,
 * shape and row-major layout; offsets are multiples of 4 floats, padding stays zero.
 * Gradients and optimizer state use the same layout. */
enum {
    MMG_P_REC_RNN_WIH = 0,  /* receiver.rnn.weight_ih   (3Hr, M)   gate order r,z,n  model.py:256 */
    MMG_P_REC_RNN_WHH,      /* receiver.rnn.weight_hh   (3Hr, Hr) */
    MMG_P_REC_RNN_BIH,      /* receiver.rnn.bias_ih     (3Hr)     */
    MMG_P_REC_RNN_BHH,      /* receiver.rnn.bias_hh     (3Hr)     */
    MMG_P_REC_WH_W,         /* receiver.w_h.weight      (Hr, Hr)   model.py:258 */
    MMG_P_REC_WH_B,         /* receiver.w_h.bias        (Hr)      */
    MMG_P_REC_WD_W,         /* receiver.w_d.weight      (Hr, WV)   model.py:259 (no bias) */
    MMG_P_REC_W_W,          /* receiver.w.weight        (M, Hr)    model.py:260 */
    MMG_P_REC_W_B,          /* receiver.w.bias          (M)       */
    MMG_P_REC_Y1_W,         /* receiver.y1.weight       (Hr, Hr+WV) columns [h_z ; desc]  model.py:262,548;
                               with desc_attn the columns are [desc ; h_z]  model.py:408-410 */
    MMG_P_REC_Y1_B,         /* receiver.y1.bias         (Hr)      */
    MMG_P_REC_Y2_W,         /* receiver.y2.weight       (1, Hr)    model.py:263 */
    MMG_P_REC_Y2_B,         /* receiver.y2.bias         (1)       */
    MMG_P_REC_S_W,          /* receiver.s.weight        (1, Hr)    model.py:265 */
    MMG_P_REC_S_B,          /* receiver.s.bias          (1)       */
    /* desc_attn only (zero-sized otherwise), model.py:267-271 */
    MMG_P_REC_DD_W,         /* receiver.d_d.weight      (A, WV)   */
    MMG_P_REC_DD_B,         /* receiver.d_d.bias        (A)       */
    MMG_P_REC_DH_W,         /* receiver.d_h.weight      (A, Hr)   */
    MMG_P_REC_DH_B,         /* receiver.d_h.bias        (A)       */
    MMG_P_REC_DA_W,         /* receiver.d_attn.weight   (1, A)    */
    MMG_P_REC_DA_B,         /* receiver.d_attn.bias     (1)       */
    MMG_P_SEN_CODE_BIAS,    /* sender.code_bias         (M)        model.py:69 */
    MMG_P_SEN_CODE_BIAS_MOU,/* sender.code_bias_mou     (M)        model.py:73-74; only with sender_mix mou AND ignore_code, zero-sized otherwise */
    MMG_P_SEN_IMG_W,        /* sender.image_layer.weight (Hi, F)   model.py:67 */
    MMG_P_SEN_IMG_B,        /* sender.image_layer.bias  (Hi)      */
    MMG_P_SEN_CODE_W,       /* sender.code_layer.weight (Hi, M)    model.py:68 */
    MMG_P_SEN_CODE_B,       /* sender.code_layer.bias   (Hi)      */
    MMG_P_SEN_BIN_W,        /* sender.binary_layer.weight (M, Hi), (M, 4 Hi) with sender_mix mou  model.py:72,76 */
    MMG_P_SEN_BIN_B,        /* sender.binary_layer.bias (M)       */
    MMG_P_BR_L1_W,          /* baseline_rec.linear1.weight (Hb, M+Hr) columns [z ; h_z]  model.py:492,842 */
    MMG_P_BR_L1_B,          /* baseline_rec.linear1.bias   (Hb)   */
    MMG_P_BR_L2_W,          /* baseline_rec.linear2.weight (1, Hb) */
    MMG_P_BR_L2_B,          /* baseline_rec.linear2.bias   (1)    */
    MMG_P_BS_L1_W,          /* baseline_sen.linear1.weight (Hb, Hi+M) columns [h_x ; z_r]  model.py:492,835 */
    MMG_P_BS_L1_B,          /* baseline_sen.linear1.bias   (Hb)   */
    MMG_P_BS_L2_W,          /* baseline_sen.linear2.weight (1, Hb) */
    MMG_P_BS_L2_B,          /* baseline_sen.linear2.bias   (1)    */
    MMG_P_COUNT
};
enum { MMG_SEG_RECEIVER = 0, MMG_SEG_SENDER = 1, MMG_SEG_BASELINE_REC = 2, MMG_SEG_BASELINE_SEN = 3, MMG_SEG_COUNT = 4 };

typedef struct mmg_param_layout {
    int64_t offset[MMG_P_COUNT];     /* in floats from the start of the flat buffer */
    int32_t rows[MMG_P_COUNT];
    int32_t cols[MMG_P_COUNT];       /* 1 for vectors */
    int32_t segment[MMG_P_COUNT];    /* MMG_SEG_* */
    int64_t seg_begin[MMG_SEG_COUNT + 1];
    int64_t total;                   /* floats, multiple of 4 */
} mmg_param_layout;

/* ---- workspace layout -------------------------------------------------------------------------------
 * One caller-provided device buffer holds the per-iteration outputs (what `exchange()` returns,
 * model.py:872-876), the activations saved for the backward pass and all scratch.  Offsets in BYTES.
 * R = max_exchange * batch rows, ordered step-major: row = t * batch + b. */
typedef struct mmg_workspace_layout {
    int64_t total_bytes;
    /* outputs of exchange(): fp32 unless noted */
    int64_t sen_feats;   /* (T,B,M)   z: sender messages {0,1} or raw scores (continuous)   model.py:855 */
    int64_t sen_probs;   /* (T,B,M)   model.py:856 (unused when !use_binary) */
    int64_t rec_feats;   /* (T+1,B,M) slot 0 = first_rec fill (model.py:786); rec_feats[t] = slot t+1  model.py:857 */
    int64_t rec_probs;   /* (T,B,M)   model.py:858 */
    int64_t stop_feat;   /* (T,B)     model.py:853 */
    int64_t stop_prob;   /* (T,B)     model.py:854 */
    int64_t y;           /* (T,B,D)   model.py:859 */
    int64_t stop_mask;   /* uint8 (T+1,B) raw chain min(prev, s) with row 0 = 1 (model.py:775,852)        */
    int64_t bs;          /* (T,B)     baseline_sen scores  model.py:863 */
    int64_t br;          /* (T,B)     baseline_rec scores  model.py:862 */
    int64_t h_x;         /* (B,Hi)    sender.h_x           model.py:195 */
    int64_t h_z;         /* (T+1,B,Hr) slot 0 = initial state (zeros or caller state); h_z after step t = slot t+1 */
    int64_t h_w;         /* (T,B,Hr)  receiver.h_w         model.py:452 */
    /* per-iteration results */
    int64_t losses;      /* float[MMG_LOSS_COUNT] */
    int64_t ystep;       /* int32 (B) step at which each example's prediction is read (model.py:893-896) */
    int64_t outp;        /* (B,D) selected prediction scores (get_rec_outp) */
    int64_t logs;        /* (B)   per-example log-likelihood (model.py:1274) */
    int64_t argmax;      /* int32 (B) model.py:1268 */
    int64_t stats;       /* double[stats_count]: batch statistics that a data-parallel run all-reduces (sum) */
    int64_t stats_count;
    int64_t grad_norms;  /* float[4] pre-clip global L2 norm per module (clip_grad_norm's return value) */
    /* upstream gradients dLoss/d(output) consumed by mmg_backward (written by mmg_loss or by the caller) */
    int64_t g_sen_probs; /* (T,B,M) */
    int64_t g_rec_probs; /* (T,B,M) */
    int64_t g_stop_prob; /* (T,B)   */
    int64_t g_outp;      /* (B,D)   gradient w.r.t. y[ystep[b], b, :] */
    int64_t g_bs;        /* (T,B)   */
    int64_t g_br;        /* (T,B)   */
    int64_t rng_state;   /* uint64[2]: {seed, iteration counter} of the on-device Philox sampler */
    int64_t opt_counters; /* int64[4]: [0] = updates that reached the receiver message head (w_h, w_d, w): torch.optim keeps
                             one step count per parameter and skips parameters without a gradient (one-step conversations),
                             so Adam's bias correction for these tensors uses this count.  Part of a checkpoint. */
} mmg_workspace_layout;

enum {
    MMG_LOSS_NLL = 0,        /* nll_loss         model.py:1271 */
    MMG_LOSS_REC,            /* loss_rec         model.py:1296-1300 */
    MMG_LOSS_SEN,            /* loss_sen         model.py:1301 */
    MMG_LOSS_BAS_REC,        /* loss_bas_rec     model.py:1293 */
    MMG_LOSS_BAS_SEN,        /* loss_bas_sen     model.py:1294 */
    MMG_LOSS_BINARY_S,       /* loss_binary_s    model.py:1279 */
    MMG_LOSS_BINARY_REC,     /* loss_binary_rec  model.py:1285 */
    MMG_LOSS_BINARY_SEN,     /* loss_binary_sen  model.py:1291 */
    MMG_LOSS_TOPK_CORRECT,   /* number of rows (this rank) whose target is in the top-k (model.py:1333-1338) */
    MMG_LOSS_ACTIVE_STEPS,   /* T' = number of exchange steps the reference would have executed (model.py:866) */
    MMG_LOSS_COUNT = 16
};

/* Inputs of one exchange.  `d_u_*` are optional float64 uniforms in the reference's draw order
 * (sender (T,B,M) model.py:227, stop (T,B) model.py:420, receiver (T,B,M) model.py:460);
 * when NULL, training draws come from the on-device Philox stream seeded through `rng_state`. */
typedef struct mmg_inputs {
    const float* d_x;          /* (B,F)  image features       exchange_args["data"]   model.py:761 */
    const float* d_desc;       /* (D,WV) class descriptions   exchange_args["desc"]   model.py:764 */
    const int64_t* d_target;   /* (B)    class labels         exchange_args["target"] model.py:763 (may be NULL for forward) */
    const double* d_u_sen;
    const double* d_u_stop;
    const double* d_u_rec;
    const float* d_corrupt_mask; /* (M) 0/1, eval-time bit flips (model.py:814-820) or NULL */
    const float* d_h0;           /* (B,Hr) initial receiver state or NULL (zeros, model.py:336-337) */
    int32_t top_k;               /* FLAGS.top_k_train */
    int32_t train;               /* exchange_args["train"] model.py:767 */
    /* optional float64 uniforms of the flipout draws, (T,B,M) each, consumed right after the message draw of the same
     * agent (model.py:233-234 after 227; 467-468 after 460); NULL: on-device Philox stream */
    const double* d_u_flip_sen;
    const double* d_u_flip_rec;
    /* desc_attn only (model.py:765-766): the words of all class descriptions, class after class, and the number of
     * words of each class; every class has at least one word */
    const float* d_desc_set;        /* (NW,WV) exchange_args["desc_set"] */
    const int32_t* d_desc_set_lens; /* (D)     exchange_args["desc_set_lens"] */
    /* optional, fused training entry points only: MMG_LOSS_COUNT floats of MAPPED PINNED HOST memory (cudaHostAlloc; with unified
     * addressing the host pointer is the device pointer).  The update kernel writes the iteration's loss values there itself, so no
     * device-to-host copy has to sit between two iterations; they are valid once the iteration has completed on the stream. */
    float* h_losses_out;
} mmg_inputs;

int mmg_abi_version(void);
const char* mmg_last_error(void);
/* Number of CUDA devices visible; <0 on error.  The library refuses to run without one. */
int mmg_device_count(void);

/* Validate `cfg` and fill the parameter layout.  Replaces the shape bookkeeping of Sender/Receiver/Baseline
 * .__init__ (model.py:53-88, 245-273, 484-494). */
int mmg_param_layout_get(const mmg_config* cfg, mmg_param_layout* out);
int mmg_workspace_layout_get(const mmg_config* cfg, mmg_workspace_layout* out);

/* One-time initialisation of a workspace (zeroes it, seeds the sampler, writes the launch plans). */
int mmg_workspace_init(const mmg_config* cfg, void* d_workspace, uint64_t seed, void* stream);

/* Forward conversation: replaces exchange() (model.py:725-876) including Sender.forward (195-238),
 * Receiver.forward (333-342,412-477), both Baseline.forward calls (496-516, train only), build_inp (519-551)
 * and the stop-mask chain.  Runs all `max_exchange` steps; rows that the reference would have skipped after an
 * early break (model.py:866) are masked.  Outputs land in the workspace. */
int mmg_exchange_forward(const mmg_config* cfg, const float* d_params, const mmg_inputs* in,
                         void* d_workspace, void* stream);

/* Losses and their gradients w.r.t. the exchange outputs: replaces get_rec_outp (879-904), NLL +
 * loglikelihood (1264-1275, 571-577), multistep_loss_binary x3 / calculate_loss_binary (907-968),
 * multistep_loss_bas x2 (971-988), the mask wiring (1248-1262), loss assembly (1296-1305) and top-k accuracy
 * (1333-1338).  phase 0: per-rank batch statistics -> workspace.stats (all-reduce them across ranks when
 * batch_global > batch); phase 1: loss values + upstream gradients.  phase -1 runs both. */
int mmg_loss(const mmg_config* cfg, const float* d_params, const mmg_inputs* in, void* d_workspace, int phase,
             void* stream);

/* Backward through both agents and both baselines given the upstream gradients in the workspace: replaces
 * the four loss.backward() calls (model.py:1309,1316,1322,1328).  Writes the flat gradient buffer `d_grads`
 * (same layout as the parameters; fully overwritten). */
int mmg_backward(const mmg_config* cfg, const float* d_params, const mmg_inputs* in, void* d_workspace,
                 float* d_grads, void* stream);

/* Data-parallel helper: recompute the per-module sums of squares of `d_grads` after the ranks have all-reduced it
 * (mmg_backward already leaves them in the workspace for the single-GPU case). */
int mmg_grad_norm(const mmg_config* cfg, float* d_grads, void* d_workspace, void* stream);

/* Per-module global-norm clip + optimizer step: replaces 4x clip_grad_norm(params, 1.) + optimizer.step()
 * (model.py:1310-1311,1317-1318,1323-1324,1329-1330).  `d_state1` = RMSprop square_avg / Adam exp_avg_sq,
 * `d_state2` = Adam exp_avg (may be NULL otherwise), `step` = 1-based update count (Adam bias correction).
 * `grad_scale` multiplies gradients first (1/world for averaged all-reduce; 1 otherwise).
 * Modules not trained by the flags (sender + baselines when !use_binary, model.py:1313) are left untouched. */
int mmg_clip_update(const mmg_config* cfg, float* d_params, float* d_grads, float* d_state1, float* d_state2,
                    int64_t step, float grad_scale, void* d_workspace, void* stream);

/* Whole training iteration on device-resident inputs = forward + loss + backward + clip_update
 * (model.py:1240-1339).  Single-GPU convenience used by the fused `train_step`. */
int mmg_train_step(const mmg_config* cfg, float* d_params, float* d_grads, float* d_state1, float* d_state2,
                   int64_t step, const mmg_inputs* in, void* d_workspace, void* stream);

/* End-to-end variant with HOST buffers: copies x (B,F) / target (B) [/ desc (D,WV) when h_desc != NULL] from
 * pinned host memory into the caller's device staging buffers, runs mmg_train_step and copies the
 * MMG_LOSS_COUNT loss floats back to `h_losses`.  All asynchronous on `stream`; the caller synchronises. */
int mmg_train_step_host(const mmg_config* cfg, float* d_params, float* d_grads, float* d_state1, float* d_state2,
                        int64_t step, const float* h_x, const int64_t* h_target, const float* h_desc,
                        float* d_x_stage, int64_t* d_target_stage, float* d_desc_stage, const mmg_inputs* in,
                        void* d_workspace, float* h_losses, void* stream);

/* Pipelined host-buffer variant: the H2D copy of batch i+1 runs on `copy_stream` while batch i trains on `stream`.
 * mmg_host_prefetch: waits (on copy_stream) for `ev_free` (the staging slot's previous consumer), copies x (B,F) and
 *   target (B) from pinned host memory into the slot and records `ev_ready`.
 * mmg_train_step_staged: `stream` waits for `ev_ready`, runs mmg_train_step on the slot (in->d_x / in->d_target point
 *   at it), records `ev_free` and copies the MMG_LOSS_COUNT loss floats to `h_losses` (pinned); `h_losses` may be NULL
 *   when in->h_losses_out is set (the update kernel then delivers them, no copy is enqueued).  Streams and events are
 *   caller-owned handles (cudaStream_t / cudaEvent_t passed as void*); nothing synchronises the host. */
int mmg_host_prefetch(const mmg_config* cfg, const float* h_x, const int64_t* h_target, float* d_x_stage,
                      int64_t* d_target_stage, void* copy_stream, void* ev_free, void* ev_ready);
int mmg_train_step_staged(const mmg_config* cfg, float* d_params, float* d_grads, float* d_state1, float* d_state2,
                          int64_t step, const mmg_inputs* in, void* d_workspace, float* h_losses, void* stream,
                          void* ev_ready, void* ev_free);

/* ---- data-parallel iteration over NVLink peer memory ----------------------------------------------------------------
 * New capability (the reference is single-process).  Each rank owns one SYMMETRIC buffer (peer-mapped on every other
 * rank, e.g. torch.distributed._symmetric_memory) holding, in this order:
 *   send   : float[param_layout.total]      this rank's local gradient, read by every peer
 *   recv   : float[param_layout.total]      slice `rank` of the GLOBAL gradient, reduced HERE by this rank and PULLED by the peers
 *   stats  : packets[MMG_MAX_PEERS][2 * workspace.stats_count]   slot r = rank r's batch statistics, PUSHED by rank r
 *   norms  : packets[MMG_MAX_PEERS][2 * 4]  slot r = per-module sums of squares of slice r of the global gradient, PUSHED by rank r
 *   flags  : uint64[3][MMG_MAX_PEERS]       arrival counters written REMOTELY by the peers (row 1: send buffer complete,
 *                                           row 2: reduced slice stored; row 0 unused); value = training step, monotonic
 * A packet is one 8-byte word {float bits | step << 32}; a double travels as two (head, remainder) and is valid once both carry
 * the current step, so pushed values need neither a flag nor a fence.  Bulk data is only written to the writer's OWN memory and
 * pulled by the peers behind a device-scope fence + remote flag store (the peers' loads are served by the owner's L2).
 * No collective library call sits on the path; every sum runs in rank order, so every rank computes bit-identical numbers.
 * Gradient: rank r sums ITS 1/G slice over all send buffers into its recv buffer; the update kernel of every rank pulls each slice
 * from its owner — (G-1)/G of a gradient over NVLink on each leg.  Waits are bounded (~15 s: ranks must run in
 * lockstep within that window); a timeout sets *d_error, which is STICKY: later waits give up at once and the update kernel
 * leaves parameters and optimizer state untouched, so the caller must poll it (GameEngine.peer_error()) and abort.
 * The buffer must start zeroed. */
#define MMG_MAX_PEERS 8
typedef struct mmg_peers {
    int32_t world, rank;
    float* d_send[MMG_MAX_PEERS];
    float* d_recv[MMG_MAX_PEERS];
    double* d_stats[MMG_MAX_PEERS];
    double* d_norms[MMG_MAX_PEERS];
    unsigned long long* d_flags[MMG_MAX_PEERS];
    int32_t* d_error;      /* local device int, set non-zero when a peer wait timed out */
    /* optional: MULTICAST address of the `send` region (one address that the NVSwitch resolves to every rank's send buffer,
     * e.g. torch symmetric memory's multicast_ptr + send_off), or NULL.  With it the owner of a slice sums it with one
     * multimem.ld_reduce per float4 (the switch reduces in flight) instead of one peer load per rank. */
    float* d_send_mc;
} mmg_peers;
/* Bytes of the symmetric buffer and the offsets of its five sections. */
int mmg_peer_buffer_layout(const mmg_config* cfg, int64_t* total_bytes, int64_t* send_off, int64_t* recv_off, int64_t* stats_off,
                           int64_t* norms_off, int64_t* flags_off);
/* One training iteration on this rank's batch shard (cfg->batch rows of a global batch of cfg->batch_global rows):
 * mmg_train_step with the statistics and the gradient summed across `peers` in-kernel.  `d_grads` (local, not
 * symmetric) receives the clipped global gradient; `step` must be identical on all ranks and increase by one per call. */
int mmg_train_step_peer(const mmg_config* cfg, float* d_params, float* d_grads, float* d_state1, float* d_state2,
                        int64_t step, const mmg_inputs* in, void* d_workspace, const mmg_peers* peers, void* stream);

/* ---- module-level single-turn forwards ---------------------------------------------------------------------------------
 * The reference's nn.Module.forward entry points, for callers that drive the agents turn by turn instead of through
 * exchange().  `rows` examples (independent of cfg->batch); parameters are read from the same flat buffer / layout as
 * everywhere else; no activations are saved (no backward).  Training draws use the injected float64 uniforms when given,
 * else the Philox stream (seed, counter) — the caller advances `counter` per call.
 *
 * mmg_sender_forward: Sender.forward(x, w, g, t) default path, model.py:144-238 (sender_mix sum/prod/mou, ignore_code,
 *   flipout_sen).  d_w (rows,M) may be NULL when t == 0 (the code is sigmoid(code_bias), model.py:199-200).
 *   Outputs: d_msg (rows,M) message {0,1} or raw scores (continuous), d_probs (rows,M) (untouched when !use_binary),
 *   d_h_x (rows,Hi) = sender.h_x (model.py:195). */
int mmg_sender_forward(const mmg_config* cfg, const float* d_params, int32_t rows, const float* d_x, const float* d_w,
                       int32_t t, int32_t train, const double* d_u, const double* d_u_flip, uint64_t seed, uint64_t counter,
                       float* d_msg, float* d_probs, float* d_h_x, void* stream);
/* mmg_receiver_forward: Receiver.forward(z, desc, desc_set, desc_set_lens), model.py:303-477 incl. the -desc_attn branch
 *   (344-410).  State: d_h_z (rows,Hr) in/out (`first` != 0: the state was reset, h_z starts at zero, model.py:336-337);
 *   d_s_prob_prod (rows) in/out, eval only (model.py:423-427).  Outputs: d_s, d_s_prob (rows), d_w, d_w_probs (rows,M),
 *   d_y (rows,D), d_h_w (rows,Hr). */
int mmg_receiver_forward(const mmg_config* cfg, const float* d_params, int32_t rows, const float* d_z, const float* d_desc,
                         const float* d_desc_set, const int32_t* d_desc_set_lens, float* d_h_z, float* d_s_prob_prod,
                         int32_t first, int32_t train, const double* d_u_stop, const double* d_u_rec, const double* d_u_flip,
                         uint64_t seed, uint64_t counter, float* d_s, float* d_s_prob, float* d_w, float* d_w_probs,
                         float* d_y, float* d_h_w, void* stream);
/* mmg_baseline_forward: Baseline.forward(x, binary, inp), model.py:496-516.  `which`: MMG_SEG_BASELINE_SEN or
 *   MMG_SEG_BASELINE_REC; absent pieces are NULL with width 0; widths must add up to linear1's input width.  d_out (rows). */
int mmg_baseline_forward(const mmg_config* cfg, const float* d_params, int32_t which, int32_t rows, const float* d_x,
                         int32_t x_dim, const float* d_binary, int32_t binary_dim, const float* d_inp, int32_t inp_dim,
                         float* d_out, void* stream);

/* Number of kernels the last call of each entry point enqueued (for launch accounting). */
int mmg_launch_count(void);
void mmg_launch_count_reset(void);
/* Diagnostic.  With MMG_KTIME=1 in the environment every kernel launch is bracketed by CUDA events on its own stream (and
 * launched without programmatic dependent launch); this call synchronises the device, writes one line per kernel
 * ("name launches mean_us") into `out` (NUL-terminated, at most `cap` bytes), clears the records and returns how many
 * launches were recorded.  Without MMG_KTIME it writes an empty string and returns 0. */
int mmg_debug_kernel_times(char* out, int32_t cap);
/* Diagnostic, debug builds only (-DMMG_TRACE): copies the per-CTA time stamps [7 kernels][1024 CTAs][8 slots] (global timer,
 * ns) of the launches since the last call into `out` (count >= 7*1024*8) and clears them.  Release builds return 0. */
int mmg_debug_trace(unsigned long long* out, int32_t count);

#ifdef __cplusplus
}
#endif
#endif /* MMG_B200_H_ */
