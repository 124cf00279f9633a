// K_wgrad, K_reduce_norm, K_update.
//
// K_wgrad: every weight / bias gradient of the four modules as ONE grouped launch of 64x64 fp32 tiles
//   C = A^T . B reduced over the T*B (or B, or B*D) rows.  Each tensor picks its own split-K factor so that every CTA
//   multiplies a K-slice of ~128 rows (balanced single wave); split 0 writes straight into the flat gradient buffer,
//   splits 1.. into arena slabs that K_reduce_norm adds in a fixed order, so gradients are bit-reproducible run to run
//   (no float atomics).  Column sums (bias-like gradients) run through the same tiles with a constant-one operand.
// K_reduce_norm: gradient += slabs, plus per-CTA partial sums of squares per module (float4 streams).
// K_update: torch.nn.utils.clip_grad_norm(params, 1.) per module (model.py:1310,1317,1323,1329) fused with the
//   optimizer step (RMSprop default, model.py:1725; Adam / SGD, 1111-1137).
#pragma once
#include "mmg_kernels.cuh"

namespace mmg {

enum { WG_GEMM = 0, WG_ROWVEC = 1, WG_CODEBIAS = 2 };
enum { kRowvecCols = 256 };
enum { kMaxWgProblems = 26, kWgradKSlice = 128 };

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
};
struct WgTable {
    WgProblem p[kMaxWgProblems];
    int tile_begin[kMaxWgProblems + 1];   // compact copy for the per-CTA problem lookup
    int count, total_tiles;
    long long slab_stride;   // floats between consecutive arena slabs (= flat layout total)
};

MMG_GLOBAL void __launch_bounds__(kGemmThreads, 4)
k_wgrad(Dims d, WgTable tab, float* grads, float* arena, const float* code_w, const float* code_bias, const float* d_as) {
    pdl_wait();                 // PDL: the previous kernel of the stream has completed and flushed
    pdl_launch_dependents();    // let the next kernel's CTAs be scheduled behind this grid
    MMG_SHARED __attribute__((aligned(16))) float gs[kGemmSmemFloats];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    int tile = blockIdx.x, pi = 0;
    while (pi + 1 < tab.count && tile >= tab.tile_begin[pi + 1]) ++pi;
    const WgProblem& pr = tab.p[pi];
    tile -= pr.tile_begin;
    const int per_split = pr.ntm * pr.ntn;
    const int s = tile / per_split;
    tile %= per_split;
    const int nt = tile % pr.ntn, mt = tile / pr.ntn;
    float* slab = s == 0 ? grads : arena + (size_t)(s - 1) * tab.slab_stride;
    const int ks = round_up(cdiv(pr.K, pr.nsplit), 4);
    const int k0 = s * ks, k1 = min(pr.K, k0 + ks);

    if (pr.kind == WG_GEMM) {
        float acc[4][4];
        float cs = 0.f;
        const bool want_bias = pr.bias_off >= 0 && nt == 0;
        gemm_tile(pr.A, pr.B, pr.M, pr.N, mt * kTile, nt * kTile, k0, k1, acc, want_bias ? &cs : nullptr, gs);
        const int j0 = nt * kTile + tx * 4;
        const bool vec = j0 + 3 < pr.N && (pr.ldc & 3) == 0 && (pr.c_off & 3) == 0 && pr.sig_rows == nullptr;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = mt * kTile + ty * 4 + a;
            if (i >= pr.M) continue;
            if (vec) {
                *reinterpret_cast<float4*>(slab + pr.c_off + (size_t)i * pr.ldc + j0) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
                continue;
            }
            float rs = 1.f;
            if (pr.sig_rows != nullptr) { const float c0 = sigmoidf_(ldg(pr.sig_rows + i)); rs = c0 * (1.f - c0); }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = j0 + c;
                if (j < pr.N) slab[pr.c_off + (size_t)i * pr.ldc + j] = acc[a][c] * rs;
            }
        }
        if (want_bias && tid < kTile && mt * kTile + tid < pr.M) slab[pr.bias_off + mt * kTile + tid] = cs;
    } else if (pr.kind == WG_ROWVEC) {
        // single-row products C[0][j] = sum_k a[k] B[k][j] (STOP head, linear2 of both baselines): one column per thread
        // instead of a 64x64 tile with 63 idle rows.  A: plain (K, 1) vector; B: plain k-major rows.
        const int j = nt * kRowvecCols + tid;
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, asum[4] = {0.f, 0.f, 0.f, 0.f};
        const float* a = pr.A.p;
        const float* bp = pr.B.p + (j < pr.N ? j : 0);
        int k = k0;
        for (; k + 3 < k1; k += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float av = ldg(a + (size_t)(k + u) * pr.A.ld);
                asum[u] += av;
                acc[u] = fmaf(av, ldg(bp + (size_t)(k + u) * pr.B.ld), acc[u]);
            }
        }
        for (; k < k1; ++k) {
            const float av = ldg(a + (size_t)k * pr.A.ld);
            asum[0] += av;
            acc[0] = fmaf(av, ldg(bp + (size_t)k * pr.B.ld), acc[0]);
        }
        if (j < pr.N) slab[pr.c_off + j] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        if (pr.bias_off >= 0 && nt == 0 && tid == 0) slab[pr.bias_off] = (asum[0] + asum[1]) + (asum[2] + asum[3]);
    } else {
        // generic-path only (the fast backward kernel emits per-example partials instead):
        // d code_bias[j] = c0 (1 - c0) sum_n code_layer.weight[n][j] * (sum_b d_as[t=0][b][n])   (model.py:199-200)
        float* v = gs;
        for (int j0 = 0; j0 < d.M; j0 += kGemmThreads) {
            const int j = j0 + tid;
            float accj = 0.f;
            for (int nb = 0; nb < d.Hi; nb += kGemmSmemFloats) {
                const int nlim = min(d.Hi - nb, (int)kGemmSmemFloats);
                MMG_SYNCTHREADS();
                for (int n = tid; n < nlim; n += kGemmThreads) {
                    float sv = 0.f;
                    for (int b = 0; b < d.B; ++b) sv += d_as[(size_t)b * d.Hi + nb + n];
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

struct SegInfo {
    long long begin[5];
    int trained[4];
    long long whead_begin, whead_end, shead_begin, shead_end;
    int shead_active, whead_stat;
};
// Per state_dict tensor: where it lives, how many floats are real (the rest of its 4-float slot is padding) and how
// many split-K partials K_wgrad produced for it.
struct SplitTable {
    long long begin[MMG_P_COUNT + 1];
    int numel[MMG_P_COUNT];
    int nsplit[MMG_P_COUNT];
};

MMG_DEVICE int seg_of(const long long* seg_begin, long long i) {
    return i < seg_begin[1] ? 0 : (i < seg_begin[2] ? 1 : (i < seg_begin[3] ? 2 : 3));
}
MMG_DEVICE int tensor_of(const SplitTable& st, long long i) {
    int lo = 0, hi = MMG_P_COUNT - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (st.begin[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// grads (+)= slabs, scaled; per-CTA partial sums of squares per module -> norm_part[module * gridDim.x + cta].
// do_reduce = 0: gradients are final already (after the data-parallel all-reduce), only the norms are computed.
MMG_DEVICE void reduce_norm_body(const SegInfo& seg, const SplitTable& st, const float* arena, long long slab_stride,
                                 float* grads, float scale, int do_reduce, float* norm_part) {
    MMG_SHARED float red[4][kUpdThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long total = seg.begin[4];
    float ss[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long i = 4 * ((long long)blockIdx.x * kUpdThreads + tid); i < total; i += 4ll * gridDim.x * kUpdThreads) {
        const int sg = seg_of(seg.begin, i);
        float4 g = *reinterpret_cast<const float4*>(grads + i);
        if (do_reduce) {
            const int tn = tensor_of(st, i);
            const long long valid = st.begin[tn] + st.numel[tn] - i;   // real floats in this group (>= 1)
            if (!seg.trained[sg] || st.nsplit[tn] == 0) {      // untrained module, or a tensor no gradient problem covers
                g = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                for (int s = 1; s < st.nsplit[tn]; ++s) {
                    const float4 a = *reinterpret_cast<const float4*>(arena + (size_t)(s - 1) * slab_stride + i);
                    g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w;
                }
                if (valid < 4) { if (valid < 2) g.y = 0.f; if (valid < 3) g.z = 0.f; g.w = 0.f; }
            }
            g.x *= scale; g.y *= scale; g.z *= scale; g.w *= scale;
            *reinterpret_cast<float4*>(grads + i) = g;
        }
        ss[sg] = fmaf(g.x, g.x, ss[sg]); ss[sg] = fmaf(g.y, g.y, ss[sg]);
        ss[sg] = fmaf(g.z, g.z, ss[sg]); ss[sg] = fmaf(g.w, g.w, ss[sg]);
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

MMG_GLOBAL void __launch_bounds__(kUpdThreads)
k_reduce_norm(SegInfo seg, SplitTable st, const float* arena, long long slab_stride, float* grads, float scale,
              int do_reduce, float* norm_part, PeerView pv, unsigned* ticket) {
    pdl_wait();                 // PDL: the previous kernel of the stream has completed and flushed
    pdl_launch_dependents();    // let the next kernel's CTAs be scheduled behind this grid
    const int tid = threadIdx.x;
    reduce_norm_body(seg, st, arena, slab_stride, grads, scale, do_reduce, norm_part);
    if (pv.world > 1) {
        // `grads` is this rank's symmetric send buffer: once every CTA has written its part, raise flag row 1 on all peers
        MMG_SHARED int s_last;
        fence_system();
        MMG_SYNCTHREADS();
        if (tid == 0) s_last = (ticket_take(ticket) == gridDim.x - 1) ? 1 : 0;
        MMG_SYNCTHREADS();
        if (s_last) {
            if (tid < pv.world) peer_signal(pv.flags[tid] + MMG_MAX_PEERS + pv.rank, pv.iter);
            if (tid == 0) *ticket = 0;
        }
    }
}

// Data-parallel gradient sum over NVLink peer memory, fused with the per-module sum of squares: every rank reads all
// send buffers (one-shot, rank order => bit-identical results everywhere) and writes the global gradient locally.
MMG_GLOBAL void __launch_bounds__(kUpdThreads)
k_peer_allreduce_norm(SegInfo seg, PeerView pv, float* grads_out, float* norm_part) {
    pdl_wait();                 // PDL: the previous kernel of the stream has completed and flushed
    pdl_launch_dependents();    // let the next kernel's CTAs be scheduled behind this grid
    MMG_SHARED float red[4][kUpdThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < pv.world && !peer_wait(pv.flags[pv.rank] + MMG_MAX_PEERS + tid, pv.iter)) *pv.error = 2;
    MMG_SYNCTHREADS();
    const long long total = seg.begin[4];
    float ss[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long i = 4 * ((long long)blockIdx.x * kUpdThreads + tid); i < total; i += 4ll * gridDim.x * kUpdThreads) {
        const int sg = seg_of(seg.begin, i);
        float4 g = peer_load4(pv.send[0] + i);
        for (int r = 1; r < pv.world; ++r) {
            const float4 a = peer_load4(pv.send[r] + i);
            g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w;
        }
        *reinterpret_cast<float4*>(grads_out + i) = g;
        ss[sg] = fmaf(g.x, g.x, ss[sg]); ss[sg] = fmaf(g.y, g.y, ss[sg]);
        ss[sg] = fmaf(g.z, g.z, ss[sg]); ss[sg] = fmaf(g.w, g.w, ss[sg]);
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

MMG_DEVICE void update_body(const SegInfo& seg, const OptHyper& hp, float* params, float* grads, float* state1, float* state2,
                            const float* norm_part, int n_norm_ctas, float* grad_norms, const double* stats,
                            const long long* opt_counters) {
    MMG_SHARED float coef[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp < 4) {   // global L2 norm per module, partials summed in a fixed order
        // per-CTA partials of this module: all loads in flight together (a plain accumulate loop would pay one L2 round
        // trip per element), added in a fixed order
        double v = 0.0;
        for (int c0 = 0; c0 < n_norm_ctas; c0 += 32 * 8) {
            float pv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = c0 + u * 32 + lane;
                pv[u] = c < n_norm_ctas ? norm_part[warp * n_norm_ctas + c] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) v += (double)pv[u];
        }
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
    for (long long i = 4 * ((long long)blockIdx.x * kUpdThreads + tid); i < total; i += 4ll * gridDim.x * kUpdThreads) {
        const int sg = seg_of(seg.begin, i);
        if (!seg.trained[sg]) continue;
        const bool in_whead = i >= seg.whead_begin && i < seg.whead_end;
        if (in_whead && !whead_active) continue;
        if (i >= seg.shead_begin && i < seg.shead_end && !seg.shead_active) continue;
        float4 g4 = *reinterpret_cast<const float4*>(grads + i);
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
        *reinterpret_cast<float4*>(grads + i) = make_float4(g[0], g[1], g[2], g[3]);
        *reinterpret_cast<float4*>(params + i) = make_float4(p[0], p[1], p[2], p[3]);
    }
}

MMG_GLOBAL void __launch_bounds__(kUpdThreads)
k_update(SegInfo seg, OptHyper hp, float* params, float* grads, float* state1, float* state2, const float* norm_part,
         int n_norm_ctas, float* grad_norms, const double* stats, const long long* opt_counters) {
    pdl_wait();                 // PDL: the previous kernel of the stream has completed and flushed
    pdl_launch_dependents();    // let the next kernel's CTAs be scheduled behind this grid
    update_body(seg, hp, params, grads, state1, state2, norm_part, n_norm_ctas, grad_norms, stats, opt_counters);
}

#ifndef MMG_CPU_EMU
// Single-GPU fusion of K_reduce_norm and K_update: the global gradient norm is a grid-wide dependency, resolved by a
// software grid barrier (arrival counter + spin).  Legal because the host launches at most as many CTAs as are
// co-resident (occupancy x SM count), so every CTA is running when the first one starts to wait.
MMG_DEVICE void grid_barrier(unsigned* ctr, unsigned n) {     // ctr[0] arrivals, ctr[1] departures; self-resetting
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        for (int spin = 0; spin < (1 << 26) && *reinterpret_cast<volatile unsigned*>(ctr) < n; ++spin) { }   // bounded
        __threadfence();
        if (atomicAdd(ctr + 1, 1u) == n - 1) { ctr[0] = 0; ctr[1] = 0; __threadfence(); }
    }
    __syncthreads();
}
MMG_GLOBAL void __launch_bounds__(kUpdThreads)
k_reduce_update(SegInfo seg, SplitTable st, const float* arena, long long slab_stride, OptHyper hp, float* params,
                float* grads, float* state1, float* state2, float* norm_part, float* grad_norms, const double* stats,
                const long long* opt_counters, unsigned* barrier_ctr) {
    pdl_wait();
    pdl_launch_dependents();
    reduce_norm_body(seg, st, arena, slab_stride, grads, 1.0f, 1, norm_part);
    grid_barrier(barrier_ctr, gridDim.x);
    update_body(seg, hp, params, grads, state1, state2, norm_part, (int)gridDim.x, grad_norms, stats, opt_counters);
}
#endif

MMG_GLOBAL void k_init_rng(unsigned long long* st, unsigned long long seed) {
    pdl_wait();
    if (threadIdx.x == 0 && blockIdx.x == 0) { st[0] = seed; st[1] = 0ull; }
}

}  // namespace mmg
