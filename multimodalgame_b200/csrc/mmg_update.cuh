// K_wgrad, K_reduce_norm, K_update.
//
// K_wgrad: every weight / bias gradient of the four modules as ONE grouped launch of 64x64 fp32 tiles
//   C = A^T . B reduced over the T*B (or B, or B*D) rows, split-K into `nsplit` slabs for parallelism; slabs are
//   summed in a fixed order by K_reduce_norm, so gradients are bit-reproducible run to run (no float atomics).
// K_reduce_norm: slabs -> flat gradient, plus per-CTA partial sums of squares per module.
// K_update: torch.nn.utils.clip_grad_norm(params, 1.) per module (model.py:1310,1317,1323,1329) fused with the
//   optimizer step (RMSprop default, model.py:1725; Adam / SGD, 1111-1137).
#pragma once
#include "mmg_kernels.cuh"

namespace mmg {

enum { WG_GEMM = 0, WG_COLSUM = 1, WG_CODEBIAS = 2 };
enum { kMaxWgProblems = 20 };

struct WgProblem {
    Operand A, B;
    int M, N, K;
    long long c_off;      // float offset of C[0][0] inside the flat layout (includes any column offset)
    int ldc;
    long long bias_off;   // float offset of colsum(A) output, or -1
    int kind;
    int tile_begin, ntm, ntn;
};
struct WgTable {
    WgProblem p[kMaxWgProblems];
    int count, total_tiles, nsplit;
    long long slab_stride;   // floats between consecutive slabs (= flat layout total)
};

MMG_GLOBAL void __launch_bounds__(kGemmThreads)
k_wgrad(Dims d, WgTable tab, float* slabs, const float* code_w, const float* code_bias, const float* d_as) {
    MMG_SHARED __attribute__((aligned(16))) float As[kChunk * kLd];
    MMG_SHARED __attribute__((aligned(16))) float Bs[kChunk * kLd];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    int tile = blockIdx.x, pi = 0;
    while (pi + 1 < tab.count && tile >= tab.p[pi + 1].tile_begin) ++pi;
    const WgProblem& pr = tab.p[pi];
    tile -= pr.tile_begin;
    const int per_split = pr.ntm * pr.ntn;
    const int s = tile / per_split;
    tile %= per_split;
    const int nt = tile % pr.ntn, mt = tile / pr.ntn;
    float* slab = slabs + (size_t)s * tab.slab_stride;
    const int ks = cdiv(pr.K, tab.nsplit);
    const int k0 = s * ks, k1 = min(pr.K, k0 + ks);

    if (pr.kind == WG_GEMM) {
        float acc[4][4];
        float cs = 0.f;
        const bool want_bias = pr.bias_off >= 0 && nt == 0;
        gemm_tile(pr.A, pr.B, pr.M, pr.N, mt * kTile, nt * kTile, k0, k1, acc, want_bias ? &cs : nullptr, As, Bs);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = mt * kTile + ty * 4 + a;
            if (i >= pr.M) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = nt * kTile + tx * 4 + c;
                if (j < pr.N) slab[pr.c_off + (size_t)i * pr.ldc + j] = acc[a][c];
            }
        }
        if (want_bias && tid < kTile && mt * kTile + tid < pr.M) slab[pr.bias_off + mt * kTile + tid] = cs;
    } else if (pr.kind == WG_COLSUM) {
        // C[i] = sum_k A(k, i): 4 row-interleaved partial sums per column, combined through shared memory
        const int c = tid % kTile, qd = tid / kTile, i = mt * kTile + c;
        float sacc = 0.f;
        if (i < pr.M) for (int k = k0 + qd; k < k1; k += kGemmThreads / kTile) sacc += operand_load(pr.A, k, i);
        As[qd * kTile + c] = sacc;
        MMG_SYNCTHREADS();
        if (tid < kTile && i < pr.M) slab[pr.c_off + i] = As[c] + As[kTile + c] + As[2 * kTile + c] + As[3 * kTile + c];
    } else {
        // d code_bias[j] = c0 (1 - c0) sum_n code_layer.weight[n][j] * (sum_b d_as[t=0][b][n])   (model.py:199-200)
        float* v = As;   // Hi <= kChunk * kLd * 2 checked on the host (As and Bs are contiguous only by luck: use As + loop)
        for (int j0 = 0; j0 < d.M; j0 += kGemmThreads) {
            const int j = j0 + tid;
            float accj = 0.f;
            for (int nb = 0; nb < d.Hi; nb += kChunk * kLd) {
                const int nlim = min(d.Hi - nb, kChunk * kLd);
                MMG_SYNCTHREADS();
                for (int n = tid; n < nlim; n += kGemmThreads) {
                    float sv = 0.f;
                    if (s == 0) for (int b = 0; b < d.B; ++b) sv += d_as[(size_t)b * d.Hi + nb + n];
                    v[n] = sv;
                }
                MMG_SYNCTHREADS();
                if (j < d.M) for (int n = 0; n < nlim; ++n) accj = fmaf(ldg(code_w + (size_t)(nb + n) * d.M + j), v[n], accj);
            }
            if (j < d.M) {
                const float c0 = sigmoidf_(ldg(code_bias + j));
                slab[pr.c_off + j] = accj * c0 * (1.f - c0);
            }
        }
    }
}

enum { kUpdThreads = 256 };

MMG_DEVICE int seg_of(const long long* seg_begin, long long i) {
    return i < seg_begin[1] ? 0 : (i < seg_begin[2] ? 1 : (i < seg_begin[3] ? 2 : 3));
}

struct SegInfo {
    long long begin[5];
    int trained[4];
    long long whead_begin, whead_end, shead_begin, shead_end;
    int shead_active, whead_stat;
};

MMG_GLOBAL void __launch_bounds__(kUpdThreads)
k_reduce_norm(SegInfo seg, const float* slabs, long long slab_stride, int nsplit, float* grads, float scale,
              int do_reduce, float* norm_part) {
    MMG_SHARED float red[4][kUpdThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long total = seg.begin[4];
    const long long chunk = round_up64(cdiv64(total, (long long)gridDim.x), 4);
    const long long lo = (long long)blockIdx.x * chunk, hi = min(total, lo + chunk);
    float ss[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long i = lo + tid; i < hi; i += kUpdThreads) {
        const int sg = seg_of(seg.begin, i);
        float g;
        if (do_reduce) {
            g = 0.f;
            if (seg.trained[sg]) for (int s = 0; s < nsplit; ++s) g += slabs[(size_t)s * slab_stride + i];
            g *= scale;
            grads[i] = g;
        } else {
            g = grads[i];
        }
        ss[sg] = fmaf(g, g, ss[sg]);
    }
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
}

struct OptHyper { int optim; float lr, max_norm; long long step; };

MMG_GLOBAL void __launch_bounds__(kUpdThreads)
k_update(SegInfo seg, OptHyper hp, float* params, float* grads, float* state1, float* state2, const float* norm_part,
         int n_norm_ctas, float* grad_norms, const double* stats, const long long* opt_counters) {
    MMG_SHARED float coef[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp < 4) {   // global L2 norm per module, partials summed in a fixed order
        double v = 0.0;
        for (int c = lane; c < n_norm_ctas; c += 32) v += (double)norm_part[warp * n_norm_ctas + c];
        v = warp_sum_d(v);
        if (lane == 0) {
            const float total = (float)sqrt(v);
            const float cc = hp.max_norm / (total + 1e-6f);       // clip_grad_norm: scale only when coef < 1
            coef[warp] = cc < 1.f ? cc : 1.f;
            if (blockIdx.x == 0) grad_norms[warp] = total;
        }
    }
    MMG_SYNCTHREADS();
    const long long total = seg.begin[4];
    const long long chunk = round_up64(cdiv64(total, (long long)gridDim.x), 4);
    const long long lo = (long long)blockIdx.x * chunk, hi = min(total, lo + chunk);
    float bc1 = 1.f, bc2s = 1.f, bc1w = 1.f, bc2sw = 1.f;
    const bool whead_active = stats[seg.whead_stat] > 0.0;
    if (hp.optim == MMG_OPT_ADAM) {
        bc1 = 1.f - powf(0.9f, (float)hp.step);
        bc2s = sqrtf(1.f - powf(0.999f, (float)hp.step));
        const float ws = (float)opt_counters[0];     // torch.optim.Adam keeps one step count per parameter
        bc1w = 1.f - powf(0.9f, ws);
        bc2sw = sqrtf(1.f - powf(0.999f, ws));
    }
    for (long long i = lo + tid; i < hi; i += kUpdThreads) {
        const int sg = seg_of(seg.begin, i);
        if (!seg.trained[sg]) continue;
        const bool in_whead = i >= seg.whead_begin && i < seg.whead_end;
        if (in_whead && !whead_active) continue;
        if (i >= seg.shead_begin && i < seg.shead_end && !seg.shead_active) continue;
        const float g = grads[i] * coef[sg];
        grads[i] = g;
        float p = params[i];
        if (hp.optim == MMG_OPT_RMSPROP) {            // alpha 0.99, eps 1e-8, no momentum
            const float v = 0.99f * state1[i] + 0.01f * g * g;
            state1[i] = v;
            p -= hp.lr * g / (sqrtf(v) + 1e-8f);
        } else if (hp.optim == MMG_OPT_ADAM) {        // betas (0.9, 0.999), eps 1e-8
            const float m = 0.9f * state2[i] + 0.1f * g;
            const float v = 0.999f * state1[i] + 0.001f * g * g;
            state2[i] = m; state1[i] = v;
            p -= (hp.lr / (in_whead ? bc1w : bc1)) * m / (sqrtf(v) / (in_whead ? bc2sw : bc2s) + 1e-8f);
        } else {
            p -= hp.lr * g;
        }
        params[i] = p;
    }
}

MMG_GLOBAL void k_init_rng(unsigned long long* st, unsigned long long seed) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { st[0] = seed; st[1] = 0ull; }
}

}  // namespace mmg
