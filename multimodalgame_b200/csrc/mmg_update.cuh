// K_wgrad, K_grad_norm, K_peer_reduce_scatter, K_update.
//
// K_wgrad: every weight / bias gradient of the four modules as ONE grouped launch of 64x64 fp32 tiles C = A^T . B reduced
//   over the T*B (or B, or B*D) rows.  Each CTA stages its WHOLE K-slice (<= 128 rows of both operands, 70 KB) in shared
//   memory with asynchronous 16-byte copies — every load of the CTA is in flight at once, one barrier — then multiplies.
//   Each problem picks its own split-K factor; a split CTA writes its partial tile (split 0 into the gradient buffer,
//   the others into arena slabs) and takes a ticket on the OUTPUT tile: the last CTA to arrive re-reads all partials in
//   split order (bit-reproducible whatever the arrival order, no float atomics), writes the final tile and its sum of
//   squares.  The CTA that finishes the last output tile adds the per-tile sums in tile order into the four per-module
//   squared gradient norms, so clip + optimizer can follow directly: no separate reduction pass over the gradient.
//   Column sums (bias-like gradients) run through the same tiles with a constant-one operand.
// K_grad_norm: per-module sums of squares of a finished gradient buffer (after an NCCL all-reduce).
// K_peer_reduce_scatter: data-parallel gradient sum over NVLink peer memory, two-shot: every rank sums ITS 1/G slice of
//   all ranks' send buffers and stores the result (+ the slice's sums of squares) into every rank's receive buffer.
// K_update: torch.nn.utils.clip_grad_norm(params, 1.) per module (model.py:1310,1317,1323,1329) fused with the
//   optimizer step (RMSprop default, model.py:1725; Adam / SGD, 1111-1137).
#pragma once
#include "mmg_kernels.cuh"
#include "mmg_loss.cuh"

namespace mmg {

enum { WG_GEMM = 0, WG_ROWVEC = 1, WG_CODEBIAS = 2, WG_ZERO = 3 };
enum { kRowvecCols = 256, kZeroChunk = 4096 };
enum { kMaxWgProblems = 48, kWgradKSlice = 128, kWgLd = kTile + 4, kMaxOutTiles = 8192 };
MMG_HOST_DEVICE int wgrad_smem_bytes() { return (2 * kWgradKSlice * kWgLd + kWgradKSlice) * 4; }

struct WgProblem {
    Operand A, B;
    int M, N, K;
    long long c_off;      // float offset of C[0][0] inside the flat layout (includes any column offset)
    int ldc;
    long long bias_off;   // float offset of colsum(A) output, or -1
    const float* sig_rows; // non-null: row i of C is scaled by s (1 - s), s = sigmoid(sig_rows[i])  (d sigmoid(code_bias))
    int kind;
    int nsplit;
    int tile_begin, ntm, ntn;
    int out_begin;        // index of this problem's first OUTPUT tile (tickets, per-tile sums of squares)
    int seg;              // module (MMG_SEG_*) the output belongs to
};
struct WgTable {
    WgProblem p[kMaxWgProblems];
    int tile_begin[kMaxWgProblems + 1];   // compact copy for the per-CTA problem lookup
    int count, total_tiles, total_out;
    long long slab_stride;   // floats between consecutive arena slabs (= flat layout total)
};
struct WgSync {               // workspace pieces of the split-K / norm protocol
    unsigned* tile_tickets;   // [kMaxOutTiles], zero between launches
    float* tile_norm;         // [kMaxOutTiles] sum of squares of each finished output tile
    unsigned* done;           // finished output tiles of this launch
    double* norm_final;       // [4] per-module sum of squares of the gradient
};

// Stages rows [k0, k1) x columns [i0, i0 + 64) of a k-major operand into S[kk][kWgLd]; rows up to `rows_pad` are zeroed.
// PLAIN / RELUGRAD rows travel as asynchronous 16-byte copies (RELUGRAD: the raw hidden values, transformed in place later).
MMG_DEVICE void wg_stage(const Operand& op, int k0, int k1, int i0, int ilim, float* S, int rows_pad, int tid) {
    const bool vec = op.kind != OP_ONES && (op.ld & 3) == 0 && aligned16(op.p) &&
                     (op.kind == OP_RELUGRAD_TSUM ? aligned16(op.w2)
                                                  : (op.p2 == nullptr || op.split == 0 ||
                                                     ((op.split & 3) == 0 && (op.ld2 & 3) == 0 && aligned16(op.p2))));
    for (int idx = tid; idx < rows_pad * (kTile / 4); idx += kGemmThreads) {
        const int kk = idx >> 4, g = idx & 15;
        const int k = k0 + kk, i = i0 + 4 * g;
        float* dst = S + kk * kWgLd + 4 * g;
        if (k >= k1 || i >= ilim) { *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
        if (vec && i + 3 < ilim) {
            if (op.kind == OP_BSUM) {
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int b0 = 0; b0 < op.mod; b0 += 8) {        // 8 examples' rows in flight at once, added in example order
                    float4 h[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        h[u] = b0 + u < op.mod ? ldg4(reinterpret_cast<const float4*>(op.p + ((size_t)(b0 + u) * op.ld2 + k) * op.ld + i))
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < 8; ++u) { r.x += h[u].x; r.y += h[u].y; r.z += h[u].z; r.w += h[u].w; }
                }
                *reinterpret_cast<float4*>(dst) = r;
            } else if (op.kind == OP_RELUGRAD_TSUM) {
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int t0 = 0; t0 < op.mod; t0 += 8) {        // 8 steps' loads in flight at once, summed in step order
                    float4 h[8];
                    float gg[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int t = t0 + u;
                        h[u] = make_float4(0.f, 0.f, 0.f, 0.f); gg[u] = 0.f;
                        if (t < op.mod) {
                            h[u] = ldg4(reinterpret_cast<const float4*>(op.p + ((size_t)t * op.ld2 + k) * op.ld + i));
                            gg[u] = ldg(op.g + (size_t)t * op.ld2 + k);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        r.x += h[u].x > 0.f ? gg[u] : 0.f; r.y += h[u].y > 0.f ? gg[u] : 0.f;
                        r.z += h[u].z > 0.f ? gg[u] : 0.f; r.w += h[u].w > 0.f ? gg[u] : 0.f;
                    }
                }
                const float4 w = ldg4(reinterpret_cast<const float4*>(op.w2 + i));
                *reinterpret_cast<float4*>(dst) = make_float4(r.x * w.x, r.y * w.y, r.z * w.z, r.w * w.w);
            } else if (op.split > 0 && i >= op.split) {
                cp_async16(dst, op.p2 + (size_t)k * op.ld2 + (i - op.split));
            } else {
                const int row = op.mod > 0 ? k % op.mod : k;
                cp_async16(dst, op.p + (size_t)row * op.ld + i);
            }
            continue;
        }
        float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (i + c >= ilim) continue;
            if (op.kind == OP_RELUGRAD) v[c] = ldg(op.p + (size_t)k * op.ld + i + c);      // raw, see wg_relu_pass
            else v[c] = operand_load(op, k, i + c);
        }
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    }
}
// RELUGRAD operand in place: hidden value h -> (h > 0) ? g[k] : 0; the per-column factor w2[i] is applied to the rows of C.
MMG_DEVICE void wg_relu_pass(const Operand& op, int k0, int k1, float* S, float* gk, int tid) {
    const int nk = k1 - k0;
    for (int kk = tid; kk < nk; kk += kGemmThreads) gk[kk] = ldg(op.g + k0 + kk);
    MMG_SYNCTHREADS();
    for (int idx = tid; idx < nk * (kTile / 4); idx += kGemmThreads) {
        const int kk = idx >> 4, g = idx & 15;
        float4* q = reinterpret_cast<float4*>(S + kk * kWgLd + 4 * g);
        float4 h = *q;
        const float gg = gk[kk];
        h.x = h.x > 0.f ? gg : 0.f; h.y = h.y > 0.f ? gg : 0.f; h.z = h.z > 0.f ? gg : 0.f; h.w = h.w > 0.f ? gg : 0.f;
        *q = h;
    }
}

MMG_DEVICE int wg_seg_of_out(const WgTable& tab, int o) {
    int pi = 0;
    while (pi + 1 < tab.count && o >= tab.p[pi + 1].out_begin) ++pi;
    return tab.p[pi].seg;
}

// Block-wide sum of one float per thread (256 threads), result valid in thread 0.
MMG_DEVICE float block_sum_256(float v, float* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    MMG_SYNCTHREADS();
    float s = 0.f;
    if (tid == 0) for (int w = 0; w < kGemmThreads / 32; ++w) s += red[w];
    return s;
}

MMG_GLOBAL void __launch_bounds__(kGemmThreads, 3)
k_wgrad(Dims d, WgTable tab, float* grads, float* arena, const float* code_w, const float* code_bias, const float* d_as,
        WgSync sy, PeerView pv, WsPtrs W, int n_loss_parts, int defer_tail) {
    pdl_wait();                 // PDL: the previous kernel of the stream has completed and flushed
    pdl_launch_dependents();    // let the next kernel's CTAs be scheduled behind this grid
    MMG_TRACE_AT(4, 0);
    // fused iteration: the loss values are added up beside K_update (its extra CTA); the receiver message head's update counter is
    // advanced HERE, one kernel earlier, so that K_update reads a stable value (the statistics are final since the backward kernel)
    if (n_loss_parts > 0 && defer_tail + (pv.world > 1) > 0 && blockIdx.x == 0 && threadIdx.x == 0 &&
        W.stats[stat_idx(d, 1, 0, 0)] > 0.0)
        W.opt_counters[0] += 1;
    MMG_DYN_SMEM(smem_raw);
    float* As = reinterpret_cast<float*>(smem_raw);
    float* Bs = As + kWgradKSlice * kWgLd;
    float* gk = Bs + kWgradKSlice * kWgLd;
    MMG_SHARED float red[kGemmThreads / 32];
    MMG_SHARED int s_flag;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    int tile = blockIdx.x, pi = 0;
    while (pi + 1 < tab.count && tile >= tab.tile_begin[pi + 1]) ++pi;
    const WgProblem& pr = tab.p[pi];
    tile -= pr.tile_begin;
    const int per_split = pr.ntm * pr.ntn;
    const int s = tile / per_split;
    tile %= per_split;
    const int out_tile = pr.out_begin + tile;
    const int nt = tile % pr.ntn, mt = tile / pr.ntn;
    float* slab = s == 0 ? grads : arena + (size_t)(s - 1) * tab.slab_stride;
    const int ks = round_up(cdiv(pr.K, pr.nsplit), 4);
    const int k0 = s * ks, k1 = min(pr.K, k0 + ks);
    float ss = 0.f;             // this thread's share of the finished tile's sum of squares
    bool fin = pr.nsplit == 1;  // this CTA holds the final values of its output tile

    if (pr.kind == WG_GEMM) {
        const int m0 = mt * kTile, n0 = nt * kTile;
        float2 acc2[4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a) { acc2[a][0] = make_float2(0.f, 0.f); acc2[a][1] = make_float2(0.f, 0.f); }
        const bool want_bias = pr.bias_off >= 0 && nt == 0;
        float csa[4] = {0.f, 0.f, 0.f, 0.f};       // column sums of A (bias gradients) of rows ty*4.., held by the tx == 0 threads
        MMG_SHARED float csum_sm[kTile];
        if (want_bias && tid < kTile) csum_sm[tid] = 0.f;
        // the K-slice in passes of at most kWgradKSlice rows (one pass at the BASELINE.json shapes)
        for (int kb = k0; kb < k1; kb += kWgradKSlice) {
        const int ke = min(k1, kb + kWgradKSlice);
        const int rows_pad = round_up(ke - kb, 8);
        if (kb > k0) MMG_SYNCTHREADS();            // the previous pass has been consumed
        wg_stage(pr.A, kb, ke, m0, pr.M, As, rows_pad, tid);
        wg_stage(pr.B, kb, ke, n0, pr.N, Bs, rows_pad, tid);
        cp_async_wait_all();
        MMG_SYNCTHREADS();
        MMG_TRACE_AT(4, 1);
        if (pr.A.kind == OP_RELUGRAD) { wg_relu_pass(pr.A, kb, ke, As, gk, tid); MMG_SYNCTHREADS(); }
#pragma unroll 8
        for (int kk = 0; kk < rows_pad; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(As + kk * kWgLd + ty * 4);
            const float4 b4 = *reinterpret_cast<const float4*>(Bs + kk * kWgLd + tx * 4);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const float2 b01 = make_float2(b4.x, b4.y), b23 = make_float2(b4.z, b4.w);
#pragma unroll
            for (int a = 0; a < 4; ++a) {      // packed fp32x2 FMA (FFMA2): 8 instructions for the 4x4 update
                const float2 aa = make_float2(av[a], av[a]);
                acc2[a][0] = ffma2(aa, b01, acc2[a][0]);
                acc2[a][1] = ffma2(aa, b23, acc2[a][1]);
            }
        }
        if (want_bias) {
            // column sums of A (bias gradients): thread (column c, quarter q) adds rows q, q + 4, ...; the quarters meet in the
            // idle tail of gk (128 floats: 64 columns x ... two halves at a time)
            const int c = tid & 63, q = tid >> 6;
            float s0 = 0.f, s1 = 0.f;
            for (int kk = q; kk < rows_pad; kk += 8) { s0 += As[kk * kWgLd + c]; s1 += As[(kk + 4) * kWgLd + c]; }
            MMG_SYNCTHREADS();                      // gk (relu pass) is no longer read
            if (q >= 2) gk[(q - 2) * 64 + c] = s0 + s1;
            MMG_SYNCTHREADS();
            float part = s0 + s1;
            if (q < 2) part += gk[q * 64 + c];
            MMG_SYNCTHREADS();
            if (q == 1) gk[c] = part;
            MMG_SYNCTHREADS();
            if (q == 0) csum_sm[c] += part + gk[c];
        }
        }
        MMG_TRACE_AT(4, 2);
        // row factors: w2[i] of a relu-gradient operand, d sigmoid(code_bias)
        auto row_scale = [&](int i) {
            float r = 1.f;
            if (pr.A.kind == OP_RELUGRAD) r = ldg(pr.A.w2 + i);
            if (pr.sig_rows != nullptr) { const float c0 = sigmoidf_(ldg(pr.sig_rows + i)); r *= c0 * (1.f - c0); }
            return r;
        };
        const bool scaled = pr.A.kind == OP_RELUGRAD || pr.sig_rows != nullptr;
        const int j0 = n0 + tx * 4;
        const bool vec = j0 + 3 < pr.N && (pr.ldc & 3) == 0 && (pr.c_off & 3) == 0;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = m0 + ty * 4 + a;
            const float rs = (scaled && i < pr.M) ? row_scale(i) : 1.f;
            acc[a][0] = acc2[a][0].x * rs; acc[a][1] = acc2[a][0].y * rs; acc[a][2] = acc2[a][1].x * rs; acc[a][3] = acc2[a][1].y * rs;
        }
        const bool bias_thread = want_bias && tx == 0;
        if (want_bias) {
            MMG_SYNCTHREADS();
            if (tx == 0) { csa[0] = csum_sm[ty * 4]; csa[1] = csum_sm[ty * 4 + 1]; csa[2] = csum_sm[ty * 4 + 2]; csa[3] = csum_sm[ty * 4 + 3]; }
        }
        if (bias_thread && scaled) {
#pragma unroll
            for (int a = 0; a < 4; ++a) if (m0 + ty * 4 + a < pr.M) csa[a] *= row_scale(m0 + ty * 4 + a);
        }
        // this CTA's (partial or final) tile
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = m0 + ty * 4 + a;
            if (i >= pr.M) continue;
            float* row = slab + pr.c_off + (size_t)i * pr.ldc;
            if (vec) *reinterpret_cast<float4*>(row + j0) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
            else {
#pragma unroll
                for (int c = 0; c < 4; ++c) if (j0 + c < pr.N) row[j0 + c] = acc[a][c];
            }
        }
        if (bias_thread) {
#pragma unroll
            for (int a = 0; a < 4; ++a) if (m0 + ty * 4 + a < pr.M) slab[pr.bias_off + m0 + ty * 4 + a] = csa[a];
        }
        if (pr.nsplit > 1) {
            // the CTA's partial tile is published by ONE fence: the barrier orders every thread's stores before thread 0's
            // device-scope fence + ticket (the grid-barrier pattern); the finishing CTA reads with L1-bypassing loads
            MMG_SYNCTHREADS();
            if (tid == 0) {
                fence_acquire();
                s_flag = (ticket_take(sy.tile_tickets + out_tile) == (unsigned)pr.nsplit - 1) ? 1 : 0;
                fence_acquire();
            }
            MMG_SYNCTHREADS();
            fin = s_flag != 0;
            MMG_TRACE_AT(4, 3);
            if (fin) {
                // last split to arrive: sum all partials in split order (this CTA's own share from its registers would change the
                // order with the arrival order, so every share is re-read) and write the final tile
                if (vec) {
                    // ROLLED over the slabs (this block runs once per output tile: straight-line code would be fetched at L2 speed),
                    // the next slab's four rows in flight while the current ones are added; split order => reproducible sums
                    const size_t off0 = (size_t)pr.c_off + (size_t)(m0 + ty * 4) * pr.ldc + j0;
                    const int nrow = min(4, pr.M - (m0 + ty * 4));
                    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 v[4], nx[2][4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        v[a] = a < nrow ? ld_cg4(grads + off0 + (size_t)a * pr.ldc) : z4;
                        nx[0][a] = a < nrow ? ld_cg4(arena + off0 + (size_t)a * pr.ldc) : z4;          // slab 1 (nsplit > 1 here)
                        nx[1][a] = (a < nrow && pr.nsplit > 2) ? ld_cg4(arena + tab.slab_stride + off0 + (size_t)a * pr.ldc) : z4;
                    }
#pragma unroll 1
                    for (int q = 1; q < pr.nsplit; q += 2) {           // slabs q, q + 1 are in nx; q + 2, q + 3 go in flight
                        float4 cur[2][4];
#pragma unroll
                        for (int u = 0; u < 2; ++u)
#pragma unroll
                            for (int a = 0; a < 4; ++a) cur[u][a] = nx[u][a];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const bool more = q + 2 + u < pr.nsplit;
                            const float* nb = arena + (size_t)(q + 1 + u) * tab.slab_stride + off0;
#pragma unroll
                            for (int a = 0; a < 4; ++a) nx[u][a] = (more && a < nrow) ? ld_cg4(nb + (size_t)a * pr.ldc) : z4;
                        }
#pragma unroll
                        for (int u = 0; u < 2; ++u)
#pragma unroll
                            for (int a = 0; a < 4; ++a) {
                                v[a].x += cur[u][a].x; v[a].y += cur[u][a].y; v[a].z += cur[u][a].z; v[a].w += cur[u][a].w;
                            }
                    }
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        if (a < nrow) *reinterpret_cast<float4*>(grads + off0 + (size_t)a * pr.ldc) = v[a];
                        acc[a][0] = v[a].x; acc[a][1] = v[a].y; acc[a][2] = v[a].z; acc[a][3] = v[a].w;
                    }
                }
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int i = m0 + ty * 4 + a;
                    if (i >= pr.M || vec) continue;
                    const size_t off = (size_t)pr.c_off + (size_t)i * pr.ldc + j0;
                    {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (j0 + c >= pr.N) continue;
                            float v = ld_cg(grads + off + c);
                            for (int q = 1; q < pr.nsplit; ++q) v += ld_cg(arena + (size_t)(q - 1) * tab.slab_stride + off + c);
                            grads[off + c] = v;
                            acc[a][c] = v;
                        }
                    }
                }
                if (bias_thread) {
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        if (m0 + ty * 4 + a >= pr.M) continue;
                        const size_t off = (size_t)pr.bias_off + m0 + ty * 4 + a;
                        float v = ld_cg(grads + off);
                        for (int q = 1; q < pr.nsplit; ++q) v += ld_cg(arena + (size_t)(q - 1) * tab.slab_stride + off);
                        grads[off] = v;
                        csa[a] = v;
                    }
                }
            }
        }
        if (fin) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                if (m0 + ty * 4 + a >= pr.M) continue;
#pragma unroll
                for (int c = 0; c < 4; ++c) if (j0 + c < pr.N) ss = fmaf(acc[a][c], acc[a][c], ss);
            }
            if (bias_thread) {
#pragma unroll
                for (int a = 0; a < 4; ++a) if (m0 + ty * 4 + a < pr.M) ss = fmaf(csa[a], csa[a], ss);
            }
        }
    } else if (pr.kind == WG_ROWVEC) {
        // single-row products C[0][j] = sum_k a[k] B[k][j] (STOP head, linear2 of both baselines): one column per thread
        // instead of a 64x64 tile with 63 idle rows.  A: plain (K, 1) vector; B: plain k-major rows.
        // thread (column quad cq, row group kg): rows k0 + kg, + 4, ... of 4 adjacent columns, 8 independent loads in flight;
        // the 4 row groups meet in shared memory, thread tid then owns column nt * 256 + tid
        const int j = nt * kRowvecCols + tid;
        const int cq = tid & 63, kg = tid >> 6, jq = nt * kRowvecCols + 4 * cq;
        const float* a = pr.A.p;
        const bool vecb = (pr.B.ld & 3) == 0 && aligned16(pr.B.p) && jq + 3 < pr.N;
        float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float asum_t = 0.f;
        for (int kb = k0 + kg; kb < k1; kb += 32) {
            float av[8];
            float4 bv4[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int k = kb + 4 * u;
                av[u] = k < k1 ? ldg(a + (size_t)k * pr.A.ld) : 0.f;
                bv4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < k1) {
                    const float* bp = pr.B.p + (size_t)k * pr.B.ld + jq;
                    if (vecb) bv4[u] = ldg4(reinterpret_cast<const float4*>(bp));
                    else {
                        if (jq < pr.N) bv4[u].x = ldg(bp);
                        if (jq + 1 < pr.N) bv4[u].y = ldg(bp + 1);
                        if (jq + 2 < pr.N) bv4[u].z = ldg(bp + 2);
                        if (jq + 3 < pr.N) bv4[u].w = ldg(bp + 3);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                asum_t += av[u];
                acc4.x = fmaf(av[u], bv4[u].x, acc4.x); acc4.y = fmaf(av[u], bv4[u].y, acc4.y);
                acc4.z = fmaf(av[u], bv4[u].z, acc4.z); acc4.w = fmaf(av[u], bv4[u].w, acc4.w);
            }
        }
        float* rv = As;                                   // [4 row groups][256 columns] + [4] sums of a
        *reinterpret_cast<float4*>(rv + kg * kRowvecCols + 4 * cq) = acc4;
        if (cq == 0) rv[4 * kRowvecCols + kg] = asum_t;
        MMG_SYNCTHREADS();
        float v = (rv[tid] + rv[kRowvecCols + tid]) + (rv[2 * kRowvecCols + tid] + rv[3 * kRowvecCols + tid]);
        float bv = (rv[4 * kRowvecCols] + rv[4 * kRowvecCols + 1]) + (rv[4 * kRowvecCols + 2] + rv[4 * kRowvecCols + 3]);
        const bool want_bias = pr.bias_off >= 0 && nt == 0 && tid == 0;
        if (j < pr.N) slab[pr.c_off + j] = v;
        if (want_bias) slab[pr.bias_off] = bv;
        if (pr.nsplit > 1) {
            MMG_SYNCTHREADS();
            if (tid == 0) {
                fence_acquire();
                s_flag = (ticket_take(sy.tile_tickets + out_tile) == (unsigned)pr.nsplit - 1) ? 1 : 0;
                fence_acquire();
            }
            MMG_SYNCTHREADS();
            fin = s_flag != 0;
            if (fin) {
                if (j < pr.N) {
                    v = ld_cg(grads + pr.c_off + j);
                    for (int q = 1; q < pr.nsplit; ++q) v += ld_cg(arena + (size_t)(q - 1) * tab.slab_stride + pr.c_off + j);
                    grads[pr.c_off + j] = v;
                }
                if (want_bias) {
                    bv = ld_cg(grads + pr.bias_off);
                    for (int q = 1; q < pr.nsplit; ++q) bv += ld_cg(arena + (size_t)(q - 1) * tab.slab_stride + pr.bias_off);
                    grads[pr.bias_off] = bv;
                }
            }
        }
        if (fin) {
            if (j < pr.N) ss = v * v;
            if (want_bias) ss = fmaf(bv, bv, ss);
        }
    } else if (pr.kind == WG_ZERO) {
        // tensors no gradient problem covers this iteration (untrained modules, the code path under -ignore_code ...):
        // the gradient buffer is fully defined after every backward pass
        const long long lo = (long long)tile * kZeroChunk, hi = min((long long)pr.M, lo + kZeroChunk);
        for (long long i = lo + 4 * tid; i < hi; i += 4 * kGemmThreads)
            *reinterpret_cast<float4*>(grads + pr.c_off + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        // generic-path only (the fast backward kernel emits per-example partials instead):
        // d code_bias[j] = c0 (1 - c0) sum_n code_layer.weight[n][j] * (sum_b d_as[t=0][b][n])   (model.py:199-200)
        // rows [A.ld2, A.mod) of d_as: step 0 for code_bias, the later steps for code_bias_mou (model.py:201-205); sig_rows = the bias
        float* v = As;
        const int vcap = 2 * kWgradKSlice * kWgLd;
        for (int j0 = 0; j0 < d.M; j0 += kGemmThreads) {
            const int j = j0 + tid;
            float accj = 0.f;
            for (int nb = 0; nb < d.Hi; nb += vcap) {
                const int nlim = min(d.Hi - nb, vcap);
                MMG_SYNCTHREADS();
                for (int n = tid; n < nlim; n += kGemmThreads) {
                    float sv = 0.f;
                    for (int b = pr.A.ld2; b < pr.A.mod; ++b) sv += d_as[(size_t)b * d.Hi + nb + n];
                    v[n] = sv;
                }
                MMG_SYNCTHREADS();
                if (j < d.M) for (int n = 0; n < nlim; ++n) accj = fmaf(ldg(code_w + (size_t)(nb + n) * d.M + j), v[n], accj);
            }
            if (j < d.M) {
                const float c0 = sigmoidf_(ldg((pr.sig_rows != nullptr ? pr.sig_rows : code_bias) + j));
                const float r = accj * c0 * (1.f - c0);
                grads[pr.c_off + j] = r;
                ss = fmaf(r, r, ss);
            }
        }
    }
    MMG_TRACE_AT(4, 4);
    if (!fin) return;
    // ---- finished output tile: its sum of squares; the CTA that finishes the LAST tile adds them up per module --------------
    const float tile_ss = block_sum_256(ss, red);
    if (tid == 0) {
        sy.tile_norm[out_tile] = tile_ss;
        sy.tile_norm[kMaxOutTiles + out_tile] = (float)pr.seg;          // the module the tile belongs to (K_update's per-module sums)
        if (pr.nsplit > 1) sy.tile_tickets[out_tile] = 0;                // ready for the next launch
        if (!defer_tail) s_flag = (ticket_take(sy.done) == (unsigned)tab.total_out - 1) ? 1 : 0;
    }
    if (defer_tail) return;      // single-rank fused iteration: K_update adds the tile sums itself, no finishing CTA
    MMG_SYNCTHREADS();
    MMG_TRACE_AT(4, 5);
    if (!s_flag) return;
    fence_acquire();
    if (pv.world > 1) {
        // data-parallel: `grads` is this rank's symmetric send buffer and it is complete: raise flag row 1 on every peer.  The
        // per-module norms come out of the cross-rank sum, the loss values are added up by K_update: nothing else to do here.
        if (tid == 0) *sy.done = 0;
        MMG_SYNCTHREADS();
        if (tid < pv.world) peer_signal(pv.flags[tid] + MMG_MAX_PEERS + pv.rank, pv.iter);      // device-scope fence inside: the data is local
        MMG_TRACE_AT(4, 6);
        return;
    }
    {
        MMG_SHARED double dred[4][kGemmThreads / 32];
        double s4[4] = {0.0, 0.0, 0.0, 0.0};
        for (int o = tid; o < tab.total_out; o += kGemmThreads) s4[wg_seg_of_out(tab, o)] += (double)sy.tile_norm[o];
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double v = warp_sum_d(s4[k]);
            if (lane == 0) dred[k][warp] = v;
        }
        MMG_SYNCTHREADS();
        if (tid < 4) {
            double v = 0.0;
            for (int w = 0; w < kGemmThreads / 32; ++w) v += dred[tid][w];
            sy.norm_final[tid] = v;
        }
        if (tid == 0) *sy.done = 0;
        // fused iteration: the backward kernel left per-CTA partials of the five loss values; they are added up here, off
        // every critical path (the values are only reported)
        if (n_loss_parts > 0) { MMG_SYNCTHREADS(); loss_finalize(d, W, n_loss_parts); }
    }
    MMG_TRACE_AT(4, 6);
}

enum { kUpdThreads = 256 };

struct SegInfo {
    long long begin[5];
    int trained[4];
    long long whead_begin, whead_end, shead_begin, shead_end;
    int shead_active, whead_stat;
};

MMG_DEVICE int seg_of(const long long* seg_begin, long long i) {
    return i < seg_begin[1] ? 0 : (i < seg_begin[2] ? 1 : (i < seg_begin[3] ? 2 : 3));
}

// Per-CTA partial sums of squares per module -> norm_part[module * gridDim.x + cta]; the last CTA to finish adds the
// partials in CTA order into norm_final[4] (deterministic).  Returns true in the finishing CTA (all its threads).
MMG_DEVICE bool finish_norms(const float (&ss)[4], float* norm_part, unsigned* ticket, double* norm_final) {
    MMG_SHARED float red[4][kUpdThreads / 32];
    MMG_SHARED int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float v = warp_sum(ss[k]);
        if (lane == 0) red[k][warp] = v;
    }
    MMG_SYNCTHREADS();
    if (tid < 4) {
        float v = 0.f;
        for (int w = 0; w < kUpdThreads / 32; ++w) v += red[tid][w];
        norm_part[tid * gridDim.x + blockIdx.x] = v;
    }
    fence_acquire();
    MMG_SYNCTHREADS();
    if (tid == 0) s_last = (ticket_take(ticket) == gridDim.x - 1) ? 1 : 0;
    MMG_SYNCTHREADS();
    if (!s_last) return false;
    fence_acquire();
    if (warp < 4) {
        double v = 0.0;
        for (int c0 = 0; c0 < (int)gridDim.x; c0 += 32 * 8) {       // 8 loads in flight per lane, added in a fixed order
            float pv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = c0 + u * 32 + lane;
                pv[u] = c < (int)gridDim.x ? norm_part[warp * gridDim.x + c] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) v += (double)pv[u];
        }
        v = warp_sum_d(v);
        if (lane == 0) norm_final[warp] = v;
    }
    if (tid == 0) *ticket = 0;
    MMG_SYNCTHREADS();
    return true;
}

// Per-module sums of squares of a finished gradient buffer (data-parallel NCCL variant: after the all-reduce).
MMG_GLOBAL void __launch_bounds__(kUpdThreads)
k_grad_norm(SegInfo seg, const float* grads, float* norm_part, unsigned* ticket, double* norm_final) {
    pdl_wait();                 // PDL: the previous kernel of the stream has completed and flushed
    pdl_launch_dependents();    // let the next kernel's CTAs be scheduled behind this grid
    const int tid = threadIdx.x;
    const long long total = seg.begin[4];
    float ss[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long i = 4 * ((long long)blockIdx.x * kUpdThreads + tid); i < total; i += 4ll * gridDim.x * kUpdThreads) {
        const int sg = seg_of(seg.begin, i);
        const float4 g = *reinterpret_cast<const float4*>(grads + i);
        ss[sg] = fmaf(g.x, g.x, ss[sg]); ss[sg] = fmaf(g.y, g.y, ss[sg]);
        ss[sg] = fmaf(g.z, g.z, ss[sg]); ss[sg] = fmaf(g.w, g.w, ss[sg]);
    }
    finish_norms(ss, norm_part, ticket, norm_final);
}

// Data-parallel gradient sum over NVLink peer memory, two-shot, PULL on both legs.  Rank r owns the r-th 1/G slice of the flat
// gradient: it waits until every rank has published its send buffer, sums the slice over all send buffers (peer loads, rank
// order => the result is the same number whoever computes it) into its OWN receive buffer, pushes only the slice's four
// per-module sums of squares to every rank and raises flag row 2 everywhere; K_update then pulls each slice from its owner.
// Per rank and iteration: (G-1)/G of the gradient in over NVLink on each leg, no bulk remote stores (a system-scope fence
// behind bulk remote stores cost 5-12 us per CTA in the push variant, profiles/r02_trace_2gpu_push.txt).
MMG_GLOBAL void __launch_bounds__(kUpdThreads)
k_peer_reduce_scatter(SegInfo seg, PeerView pv, float* norm_part, unsigned* ticket, double* norm_scratch) {
    pdl_wait();                 // PDL: the previous kernel of the stream has completed and flushed
    pdl_launch_dependents();    // let the next kernel's CTAs be scheduled behind this grid
    const int tid = threadIdx.x;
    MMG_TRACE_AT(6, 0);
    if (tid < pv.world && !peer_wait(pv.flags[pv.rank] + MMG_MAX_PEERS + tid, pv.iter, pv.error)) *pv.error = 2;
    MMG_SYNCTHREADS();
    MMG_TRACE_AT(6, 1);
    const long long n4 = seg.begin[4] / 4;
    const long long per = cdiv64(n4, pv.world);
    const long long lo = per * pv.rank, hi = min(n4, lo + per);
    float ss[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long q = lo + (long long)blockIdx.x * kUpdThreads + tid; q < hi; q += (long long)gridDim.x * kUpdThreads) {
        const long long i = 4 * q;
        const int sg = seg_of(seg.begin, i);
        if (pv.send_mc != nullptr) {
            // in-switch reduction: ONE multimem load returns the sum over all ranks (1/G of the inbound bytes); the slice has a
            // single reducer, so every replica still receives the same numbers
            const float4 g = multimem_sum4(pv.send_mc + i);
            *reinterpret_cast<float4*>(pv.recv[pv.rank] + i) = g;
            ss[sg] = fmaf(g.x, g.x, ss[sg]); ss[sg] = fmaf(g.y, g.y, ss[sg]);
            ss[sg] = fmaf(g.z, g.z, ss[sg]); ss[sg] = fmaf(g.w, g.w, ss[sg]);
            continue;
        }
        // every rank's share of this float4 in flight at once (one NVLink round trip, not one per rank), added in rank order
        float4 sh[MMG_MAX_PEERS];
#pragma unroll
        for (int r = 0; r < MMG_MAX_PEERS; ++r) sh[r] = r < pv.world ? peer_load4(pv.send[r] + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 g = sh[0];
#pragma unroll
        for (int r = 1; r < MMG_MAX_PEERS; ++r) if (r < pv.world) { g.x += sh[r].x; g.y += sh[r].y; g.z += sh[r].z; g.w += sh[r].w; }
        *reinterpret_cast<float4*>(pv.recv[pv.rank] + i) = g;       // the owner keeps the slice: the peers PULL it in K_update
        ss[sg] = fmaf(g.x, g.x, ss[sg]); ss[sg] = fmaf(g.y, g.y, ss[sg]);
        ss[sg] = fmaf(g.z, g.z, ss[sg]); ss[sg] = fmaf(g.w, g.w, ss[sg]);
    }
    MMG_TRACE_AT(6, 2);
    if (finish_norms(ss, norm_part, ticket, norm_scratch)) {
        // the whole slice is reduced and stored: publish its sums of squares and raise flag row 2 everywhere
        if (tid < 4 * pv.world) ll_store(pv.norms[tid >> 2], 4 * pv.rank + (tid & 3), norm_scratch[tid & 3], pv.iter);
        MMG_SYNCTHREADS();
        if (tid < pv.world) peer_signal(pv.flags[tid] + 2 * MMG_MAX_PEERS + pv.rank, pv.iter);
        MMG_TRACE_AT(6, 4);
    }
}

struct OptHyper { int optim; float lr, max_norm; long long step; };

// `norm_final`: per-module sum of squares of the gradient (K_wgrad / K_grad_norm), or with peers: the per-rank slice sums
// that K_peer_reduce_scatter published (added in rank order after the slices have arrived).
// `norm_tiles` > 0 (single-rank fused iteration): K_wgrad left one sum of squares per finished output tile (+ the tile's module)
// and no finishing CTA; every CTA adds them up itself, in tile order (the same numbers in every CTA).
MMG_DEVICE void update_body(const SegInfo& seg, const OptHyper& hp, float* params, const float* grads_in, float* grads_out,
                            float* state1, float* state2, double* norm_final, float* grad_norms, const double* stats,
                            const long long* opt_counters, const PeerView& pv, const float* tile_norm = nullptr, int norm_tiles = 0,
                            int nb = 0) {
    if (nb <= 0) nb = (int)gridDim.x;          // CTAs that share the update (the grid may carry one extra CTA for the loss values)
    MMG_SHARED float coef[4];
    MMG_SHARED int s_err;
    MMG_SHARED double tsum[4][kUpdThreads / 32];
    const int tid = threadIdx.x;
    if (tid == 0) s_err = 0;
    if (norm_tiles > 0) {
        double s4[4] = {0.0, 0.0, 0.0, 0.0};
        for (int o0 = 0; o0 < norm_tiles; o0 += 4 * kUpdThreads) {      // 8 loads in flight per thread
            float nv[4], sv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int o = o0 + u * kUpdThreads + tid;
                nv[u] = o < norm_tiles ? tile_norm[o] : 0.f;
                sv[u] = o < norm_tiles ? tile_norm[kMaxOutTiles + o] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int sg = (int)sv[u];
#pragma unroll
                for (int k = 0; k < 4; ++k) s4[k] += sg == k ? (double)nv[u] : 0.0;
            }
        }
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double v = warp_sum_d(s4[k]);
            if (lane == 0) tsum[k][warp] = v;
        }
        MMG_SYNCTHREADS();
    }
    if (pv.world > 1) {
        if (tid < pv.world && !peer_wait(pv.flags[pv.rank] + 2 * MMG_MAX_PEERS + tid, pv.iter, pv.error)) *pv.error = 3;
        MMG_SYNCTHREADS();
        MMG_TRACE_AT(5, 1);
        if (tid == 0) s_err = *reinterpret_cast<volatile int*>(pv.error);
    }
    if (tid < 4) {   // global L2 norm per module
        double v = 0.0;
        if (pv.world > 1) for (int r = 0; r < pv.world; ++r) v += ll_load(pv.norms[pv.rank], 4 * r + tid, pv.iter, pv.error);
        else if (norm_tiles > 0) {
            for (int w = 0; w < kUpdThreads / 32; ++w) v += tsum[tid][w];
            if (blockIdx.x == 0) norm_final[tid] = v;
        }
        else v = norm_final[tid];
        const float total = (float)sqrt(v);
        const float cc = hp.max_norm / (total + 1e-6f);       // clip_grad_norm: scale only when coef < 1
        coef[tid] = cc < 1.f ? cc : 1.f;
        if (blockIdx.x == 0) grad_norms[tid] = total;
    }
    MMG_SYNCTHREADS();
    if (s_err != 0) return;      // a peer wait timed out (sticky): no update from stale or partial sums; the host raises
    const long long total = seg.begin[4];
    float bc1 = 1.f, bc2s = 1.f, bc1w = 1.f, bc2sw = 1.f;
    const bool whead_active = stats[seg.whead_stat] > 0.0;
    if (hp.optim == MMG_OPT_ADAM) {
        bc1 = 1.f - powf(0.9f, (float)hp.step);
        bc2s = sqrtf(1.f - powf(0.999f, (float)hp.step));
        const float ws = (float)opt_counters[0];     // torch.optim.Adam keeps one step count per parameter
        bc1w = 1.f - powf(0.9f, ws);
        bc2sw = sqrtf(1.f - powf(0.999f, ws));
    }
    // all range boundaries are multiples of 4 floats, so a float4 group never straddles a module or a head
    for (long long i = 4 * ((long long)blockIdx.x * kUpdThreads + tid); i < total; i += 4ll * nb * kUpdThreads) {
        const int sg = seg_of(seg.begin, i);
        if (!seg.trained[sg]) continue;
        const bool in_whead = i >= seg.whead_begin && i < seg.whead_end;
        if (in_whead && !whead_active) continue;
        if (i >= seg.shead_begin && i < seg.shead_end && !seg.shead_active) continue;
        float4 g4;
        if (pv.world > 1) {       // the reduced slice lives in its owner's receive buffer (NVLink peer load, L1 bypass)
            const long long per4 = cdiv64(total / 4, pv.world) * 4;
            g4 = peer_load4(pv.recv[(int)(i / per4)] + i);
        } else g4 = *reinterpret_cast<const float4*>(grads_in + i);
        float4 p4 = *reinterpret_cast<const float4*>(params + i);
        const float cf = coef[sg];
        float g[4] = {g4.x * cf, g4.y * cf, g4.z * cf, g4.w * cf};
        float p[4] = {p4.x, p4.y, p4.z, p4.w};
        if (hp.optim == MMG_OPT_RMSPROP) {            // alpha 0.99, eps 1e-8, no momentum
            float4 v4 = *reinterpret_cast<const float4*>(state1 + i);
            float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                v[c] = 0.99f * v[c] + 0.01f * g[c] * g[c];
                p[c] -= hp.lr * g[c] / (sqrtf(v[c]) + 1e-8f);
            }
            *reinterpret_cast<float4*>(state1 + i) = make_float4(v[0], v[1], v[2], v[3]);
        } else if (hp.optim == MMG_OPT_ADAM) {        // betas (0.9, 0.999), eps 1e-8
            float4 v4 = *reinterpret_cast<const float4*>(state1 + i);
            float4 m4 = *reinterpret_cast<const float4*>(state2 + i);
            float v[4] = {v4.x, v4.y, v4.z, v4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                m[c] = 0.9f * m[c] + 0.1f * g[c];
                v[c] = 0.999f * v[c] + 0.001f * g[c] * g[c];
                p[c] -= (hp.lr / (in_whead ? bc1w : bc1)) * m[c] / (sqrtf(v[c]) / (in_whead ? bc2sw : bc2s) + 1e-8f);
            }
            *reinterpret_cast<float4*>(state1 + i) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(state2 + i) = make_float4(m[0], m[1], m[2], m[3]);
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) p[c] -= hp.lr * g[c];
        }
        *reinterpret_cast<float4*>(grads_out + i) = make_float4(g[0], g[1], g[2], g[3]);
        *reinterpret_cast<float4*>(params + i) = make_float4(p[0], p[1], p[2], p[3]);
    }
}

MMG_GLOBAL void __launch_bounds__(kUpdThreads)
k_update(SegInfo seg, OptHyper hp, float* params, const float* grads_in, float* grads_out, float* state1, float* state2,
         double* norm_final, float* grad_norms, const double* stats, const long long* opt_counters, PeerView pv,
         const float* tile_norm, int norm_tiles, int n_loss_parts, Dims d, WsPtrs W, float* h_losses_out) {
    pdl_wait();                 // PDL: the previous kernel of the stream has completed and flushed
    pdl_launch_dependents();    // let the next kernel's CTAs be scheduled behind this grid
    MMG_TRACE_AT(5, 0);
    // fused iteration: the backward kernel left per-CTA partials of the five loss values; ONE EXTRA CTA adds them up beside the
    // update (the values are only reported; as a tail of an update CTA this once-executed code lengthened the kernel by 3 us)
    if (n_loss_parts > 0 && blockIdx.x == gridDim.x - 1) { loss_finalize(d, W, n_loss_parts, false, h_losses_out); MMG_TRACE_AT(5, 6); return; }
    update_body(seg, hp, params, grads_in, grads_out, state1, state2, norm_final, grad_norms, stats, opt_counters, pv, tile_norm,
                norm_tiles, n_loss_parts > 0 ? (int)gridDim.x - 1 : (int)gridDim.x);
    MMG_TRACE_AT(5, 7);
}

MMG_GLOBAL void k_init_rng(unsigned long long* st, unsigned long long seed) {
    pdl_wait();
    if (threadIdx.x == 0 && blockIdx.x == 0) { st[0] = seed; st[1] = 0ull; }
}

}  // namespace mmg
