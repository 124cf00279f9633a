// Generic fp32 FFMA tile GEMM used for every non-recurrent product of the path: the image-layer GEMM, the two
// baseline MLPs and all weight gradients.  C[i][j] = sum_k A(k, i) * B(k, j) over k in [k0, k1).
// 64x64 output tile per 256-thread CTA, 4x4 register micro-tile per thread, 16-deep chunks, shared memory double
// buffered through a register prefetch: the global loads of chunk c+1 are in flight while chunk c is multiplied, one
// barrier per chunk.  Operand rows are fetched as float4 (LDG.128 -> STS.128) whenever the layout allows it.
// fp32 on the CUDA cores on purpose: parity with the reference is 1e-4 on logits and bit-exact on sampled
// bits, which TF32 tensor-core products (10-bit mantissa) do not meet at K=2048.
#pragma once
#include "mmg_platform.cuh"

namespace mmg {

enum { kTile = 64, kChunk = 16, kLd = kTile + 4, kGemmThreads = 256 };
enum { kGemmSmemFloats = 4 * kChunk * kLd };     // [buffer 0/1][A, B][kChunk][kLd]
enum { OP_PLAIN = 0, OP_RELUGRAD = 1, OP_ONES = 2, OP_RELUGRAD_TSUM = 4, OP_BSUM = 8 };

// Operand descriptor.  Element (k, i): k = reduction index, i = output index (row of C for A, column for B).
struct Operand {
    const float* p;     // primary source
    const float* p2;    // secondary source (concatenation), or nullptr
    const float* g;     // RELUGRAD: per-k scale
    const float* w2;    // RELUGRAD: per-i scale
    int ld, ld2;        // leading dimensions
    int kmajor;         // 1: element at p[row(k) * ld + i]   (rows are the reduction index)
                        // 0: element at p[row(i) * ld + k]   (reduction index contiguous)
    int mod;            // > 0: row index taken modulo `mod` (applies to the primary source only)
    int split;          // concat boundary: kmajor ? (i >= split -> p2[k * ld2 + i - split])
                        //                         : (k >= split -> p2[i * ld2 + k - split]); 0 = none
    int kind;           // OP_PLAIN; OP_RELUGRAD: (p > 0) ? g[k] * w2[i] : 0; OP_ONES: 1 (column sums as a GEMM);
                        // OP_RELUGRAD_TSUM (k-major, k = example): w2[i] * sum_{t < mod} g[t * ld2 + k] * (p[(t * ld2 + k) * ld + i] > 0)
                        // OP_BSUM (k-major, k = class): sum_{b < mod} p[(b * ld2 + k) * ld + i]   (rows (example, class) summed over the examples)
};

MMG_DEVICE float operand_load(const Operand& op, int k, int i) {
    if (op.kind == OP_ONES) return 1.f;
    if (op.kind == OP_BSUM) {
        float v = 0.f;
        for (int b = 0; b < op.mod; ++b) v += ldg(op.p + ((size_t)b * op.ld2 + k) * op.ld + i);
        return v;
    }
    if (op.kind == OP_RELUGRAD_TSUM) {
        float v = 0.f;
        for (int t = 0; t < op.mod; ++t)
            if (ldg(op.p + ((size_t)t * op.ld2 + k) * op.ld + i) > 0.f) v += ldg(op.g + (size_t)t * op.ld2 + k);
        return v * ldg(op.w2 + i);
    }
    float v;
    if (op.kmajor) {
        if (op.split > 0 && i >= op.split) {
            v = ldg(op.p2 + (size_t)k * op.ld2 + (i - op.split));
        } else {
            int r = op.mod > 0 ? k % op.mod : k;
            v = ldg(op.p + (size_t)r * op.ld + i);
        }
        if (op.kind == OP_RELUGRAD) v = v > 0.f ? ldg(op.g + k) * ldg(op.w2 + i) : 0.f;
    } else {
        if (op.split > 0 && k >= op.split) {
            v = ldg(op.p2 + (size_t)i * op.ld2 + (k - op.split));
        } else {
            int r = op.mod > 0 ? i % op.mod : i;
            v = ldg(op.p + (size_t)r * op.ld + k);
        }
    }
    return v;
}

MMG_DEVICE bool aligned16(const void* p) { return (((size_t)p) & 15) == 0; }

// Per-thread fetch pattern of one 16 x 64 chunk (4 elements per thread):
//   mode 0 (scalar, k-major rows): (kk, ii) = (tid/64 + 4 l, tid % 64)
//   mode 1 (scalar, k contiguous): (kk, ii) = (tid % 16, tid/16 + 16 l)
//   mode 2 (float4 along i):       (kk, ii) = (tid/16, 4 (tid % 16) + c)
//   mode 3 (float4 along k):       (kk, ii) = (4 (tid % 4) + c, tid/4)
MMG_DEVICE int operand_mode(const Operand& op, int k0) {
    if (op.kind == OP_ONES) return 0;
    if (op.kind == OP_RELUGRAD_TSUM) return ((op.ld & 3) == 0 && aligned16(op.p) && aligned16(op.w2)) ? 2 : 0;
    if (op.kind == OP_BSUM) return 0;
    const bool ok2 = op.p2 == nullptr || op.split == 0 || ((op.split & 3) == 0 && (op.ld2 & 3) == 0 && aligned16(op.p2));
    if ((op.ld & 3) != 0 || !aligned16(op.p) || !ok2) return op.kmajor ? 0 : 1;
    if (op.kmajor) {
        if (op.kind == OP_RELUGRAD && !aligned16(op.w2)) return 0;
        return 2;
    }
    return (k0 & 3) == 0 ? 3 : 1;
}

MMG_DEVICE float4 chunk_fetch(const Operand& op, int mode, int kbase, int kend, int ibase, int ilim, int tid) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (mode == 2) {
        const int k = kbase + (tid >> 4), i = ibase + 4 * (tid & 15);
        if (k < kend && i + 3 < ilim && op.kind == OP_RELUGRAD_TSUM) {
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int t = 0; t < op.mod; ++t) {
                const float4 h = ldg4(reinterpret_cast<const float4*>(op.p + ((size_t)t * op.ld2 + k) * op.ld + i));
                const float g = ldg(op.g + (size_t)t * op.ld2 + k);
                r.x += h.x > 0.f ? g : 0.f; r.y += h.y > 0.f ? g : 0.f; r.z += h.z > 0.f ? g : 0.f; r.w += h.w > 0.f ? g : 0.f;
            }
            const float4 w = ldg4(reinterpret_cast<const float4*>(op.w2 + i));
            r.x *= w.x; r.y *= w.y; r.z *= w.z; r.w *= w.w;
            return r;
        }
        if (k < kend && i + 3 < ilim) {
            float4 r;
            if (op.split > 0 && i >= op.split) {
                r = ldg4(reinterpret_cast<const float4*>(op.p2 + (size_t)k * op.ld2 + (i - op.split)));
            } else {
                const int row = op.mod > 0 ? k % op.mod : k;
                r = ldg4(reinterpret_cast<const float4*>(op.p + (size_t)row * op.ld + i));
            }
            if (op.kind == OP_RELUGRAD) {
                const float g = ldg(op.g + k);
                const float4 w = ldg4(reinterpret_cast<const float4*>(op.w2 + i));
                r.x = r.x > 0.f ? g * w.x : 0.f; r.y = r.y > 0.f ? g * w.y : 0.f;
                r.z = r.z > 0.f ? g * w.z : 0.f; r.w = r.w > 0.f ? g * w.w : 0.f;
            }
            return r;
        }
        if (k < kend) {
#pragma unroll
            for (int c = 0; c < 4; ++c) if (i + c < ilim) v[c] = operand_load(op, k, i + c);
        }
    } else if (mode == 3) {
        const int k = kbase + 4 * (tid & 3), i = ibase + (tid >> 2);
        if (i < ilim) {
            if (k + 3 < kend && !(op.split > 0 && k + 3 >= op.split && k < op.split)) {
                if (op.split > 0 && k >= op.split)
                    return ldg4(reinterpret_cast<const float4*>(op.p2 + (size_t)i * op.ld2 + (k - op.split)));
                const int row = op.mod > 0 ? i % op.mod : i;
                return ldg4(reinterpret_cast<const float4*>(op.p + (size_t)row * op.ld + k));
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) if (k + c < kend) v[c] = operand_load(op, k + c, i);
        }
    } else if (mode == 0) {
        const int i = ibase + (tid & 63);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int k = kbase + (tid >> 6) + 4 * l;
            if (k < kend && i < ilim) v[l] = operand_load(op, k, i);
        }
    } else {
        const int k = kbase + (tid & 15);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int i = ibase + (tid >> 4) + 16 * l;
            if (k < kend && i < ilim) v[l] = operand_load(op, k, i);
        }
    }
    return make_float4(v[0], v[1], v[2], v[3]);
}

MMG_DEVICE void chunk_store(float* S, int mode, const float4& r, int tid) {
    if (mode == 2) {
        *reinterpret_cast<float4*>(S + (tid >> 4) * kLd + 4 * (tid & 15)) = r;
    } else if (mode == 3) {
        float* q = S + (4 * (tid & 3)) * kLd + (tid >> 2);
        q[0] = r.x; q[kLd] = r.y; q[2 * kLd] = r.z; q[3 * kLd] = r.w;
    } else if (mode == 0) {
        float* q = S + (tid >> 6) * kLd + (tid & 63);
        q[0] = r.x; q[4 * kLd] = r.y; q[8 * kLd] = r.z; q[12 * kLd] = r.w;
    } else {
        float* q = S + (tid & 15) * kLd + (tid >> 4);
        q[0] = r.x; q[16] = r.y; q[32] = r.z; q[48] = r.w;
    }
}

// Accumulates the 4x4 micro-tile of thread (ty = tid / 16, tx = tid % 16): rows i = ty*4.., cols j = tx*4..
// `colsum` (optional, threads < kTile): sum_k A(k, m0 + tid) as a by-product (bias gradients).
// `smem`: kGemmSmemFloats floats, 16-byte aligned.
MMG_DEVICE void gemm_tile(const Operand& A, const Operand& Bm, int Mdim, int Ndim, int m0, int n0, int k0, int k1,
                          float (&acc)[4][4], float* colsum, float* smem) {
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    float2 acc2[4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a) { acc2[a][0] = make_float2(0.f, 0.f); acc2[a][1] = make_float2(0.f, 0.f); }
    float cs = 0.f;
    if (k0 < k1) {
        const int modeA = operand_mode(A, k0), modeB = operand_mode(Bm, k0);
        float4 ra = chunk_fetch(A, modeA, k0, k1, m0, Mdim, tid);
        float4 rb = chunk_fetch(Bm, modeB, k0, k1, n0, Ndim, tid);
        chunk_store(smem, modeA, ra, tid);
        chunk_store(smem + kChunk * kLd, modeB, rb, tid);
        MMG_SYNCTHREADS();
        int buf = 0;
        for (int kb = k0; kb < k1; kb += kChunk) {
            const bool more = kb + kChunk < k1;
            if (more) {
                ra = chunk_fetch(A, modeA, kb + kChunk, k1, m0, Mdim, tid);
                rb = chunk_fetch(Bm, modeB, kb + kChunk, k1, n0, Ndim, tid);
            }
            const float* As = smem + buf * 2 * kChunk * kLd;
            const float* Bs = As + kChunk * kLd;
#pragma unroll
            for (int kk = 0; kk < kChunk; ++kk) {
                const float4 a4 = *reinterpret_cast<const float4*>(As + kk * kLd + ty * 4);
                const float4 b4 = *reinterpret_cast<const float4*>(Bs + kk * kLd + tx * 4);
                const float av[4] = {a4.x, a4.y, a4.z, a4.w};
                const float2 b01 = make_float2(b4.x, b4.y), b23 = make_float2(b4.z, b4.w);
#pragma unroll
                for (int a = 0; a < 4; ++a) {      // packed fp32x2 FMA (FFMA2): 8 instructions for the 4x4 update
                    const float2 aa = make_float2(av[a], av[a]);
                    acc2[a][0] = ffma2(aa, b01, acc2[a][0]);
                    acc2[a][1] = ffma2(aa, b23, acc2[a][1]);
                }
            }
            if (colsum != nullptr && tid < kTile) {
#pragma unroll
                for (int kk = 0; kk < kChunk; ++kk) cs += As[kk * kLd + tid];
            }
            if (more) {
                float* Sn = smem + (buf ^ 1) * 2 * kChunk * kLd;
                chunk_store(Sn, modeA, ra, tid);
                chunk_store(Sn + kChunk * kLd, modeB, rb, tid);
            }
            MMG_SYNCTHREADS();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        acc[a][0] = acc2[a][0].x; acc[a][1] = acc2[a][0].y; acc[a][2] = acc2[a][1].x; acc[a][3] = acc2[a][1].y;
    }
    if (colsum != nullptr && tid < kTile) *colsum = cs;
}

// Same tile for SHORT reductions (k1 - k0 <= kDeep * kChunk): every chunk of both operands is fetched into registers up front
// (all global loads of the CTA in flight at once: one memory latency instead of one per chunk), then stored / multiplied chunk
// by chunk through the double-buffered shared memory.  Longer reductions fall back to gemm_tile.
enum { kDeep = 8 };
MMG_DEVICE void gemm_tile_deep(const Operand& A, const Operand& Bm, int Mdim, int Ndim, int m0, int n0, int k0, int k1,
                               float (&acc)[4][4], float* smem) {
    if (k1 - k0 > kDeep * kChunk) { gemm_tile(A, Bm, Mdim, Ndim, m0, n0, k0, k1, acc, nullptr, smem); return; }
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    float2 acc2[4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a) { acc2[a][0] = make_float2(0.f, 0.f); acc2[a][1] = make_float2(0.f, 0.f); }
    const int modeA = operand_mode(A, k0), modeB = operand_mode(Bm, k0);
    float4 ra[kDeep], rb[kDeep];
#pragma unroll
    for (int c = 0; c < kDeep; ++c) {
        const int kb = k0 + c * kChunk;
        if (kb < k1) {
            ra[c] = chunk_fetch(A, modeA, kb, k1, m0, Mdim, tid);
            rb[c] = chunk_fetch(Bm, modeB, kb, k1, n0, Ndim, tid);
        }
    }
#pragma unroll
    for (int c = 0; c < kDeep; ++c) {
        const int kb = k0 + c * kChunk;
        if (kb >= k1) break;
        float* As = smem + (c & 1) * 2 * kChunk * kLd;
        float* Bs = As + kChunk * kLd;
        chunk_store(As, modeA, ra[c], tid);
        chunk_store(Bs, modeB, rb[c], tid);
        MMG_SYNCTHREADS();          // one barrier per chunk: buffer (c & 1) is only rewritten two chunks later, after the next barrier
#pragma unroll
        for (int kk = 0; kk < kChunk; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(As + kk * kLd + ty * 4);
            const float4 b4 = *reinterpret_cast<const float4*>(Bs + kk * kLd + tx * 4);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const float2 b01 = make_float2(b4.x, b4.y), b23 = make_float2(b4.z, b4.w);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float2 aa = make_float2(av[a], av[a]);
                acc2[a][0] = ffma2(aa, b01, acc2[a][0]);
                acc2[a][1] = ffma2(aa, b23, acc2[a][1]);
            }
        }
    }
    MMG_SYNCTHREADS();              // the caller may reuse `smem`
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        acc[a][0] = acc2[a][0].x; acc[a][1] = acc2[a][0].y; acc[a][2] = acc2[a][1].x; acc[a][3] = acc2[a][1].y;
    }
}

// ---- small-K tile for operands whose reduction index is contiguous in memory (rows of activations x rows of a Linear weight) ----
// C[i][j] = sum_k A_i[k] * B_j[k], k < ka1 + ka2 <= kRowsKMax, where row i of A is the concatenation [a1[i * lda1 + 0..ka1) ;
// a2[i * lda2 + 0..ka2)] and row j of B is b[j * ldb + 0..K).  The whole K range of both operands is staged with asynchronous
// 16-byte copies (every load of the CTA in flight at once, no register round trip), rows kept k-contiguous in shared memory
// with the 16-byte group index XOR-ed by (row / 4) % 8 so that the float4 reads of 8 consecutive thread columns hit 32 distinct
// banks; then ONE rolled loop of 8 LDS.128 + 32 packed FMAs per 4 k.  The code is a few hundred instructions: these tiles run
// once per CTA, and the fully unrolled generic tile spent 65 % of its stall samples waiting for instructions (ncu, r02).
enum { kRowsKMax = 128 };
MMG_HOST_DEVICE int rows_tile_kp(int K) { return (K + 31) & ~31; }                        // padded row length (floats)
MMG_HOST_DEVICE int rows_tile_smem_floats(int K) { return 2 * kTile * rows_tile_kp(K); }
MMG_HOST_DEVICE bool rows_tile_ok(const float* a1, int lda1, int ka1, const float* a2, int lda2, int ka2, const float* b, int ldb) {
    const size_t bits = (size_t)a1 | (size_t)b | (ka2 > 0 ? (size_t)a2 : 0);
    return (bits & 15) == 0 && ((lda1 | ka1 | ldb | (ka2 > 0 ? (lda2 | ka2) : 0)) & 3) == 0 && ka1 + ka2 <= kRowsKMax && ka1 > 0;
}
MMG_DEVICE void gemm_rows_tile(const float* a1, int lda1, int ka1, const float* a2, int lda2, int ka2, const float* b, int ldb,
                               int Mdim, int Ndim, int m0, int n0, float (&acc)[4][4], float* smem) {
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int K = ka1 + ka2, K4 = K >> 2, KP = rows_tile_kp(K);
    float* As = smem;
    float* Bs = smem + kTile * KP;
    // staging: thread -> (row = tid / 4 + 64 p?, groups tid % 4, + 4, ...): rows advance by kGemmThreads / 4 = 64 -> one pass per operand
    {
        const int row = tid >> 2;
        const int sw = (row >> 2) & 7;
        const bool oka = m0 + row < Mdim, okb = n0 + row < Ndim;
        const float* ar1 = a1 + (size_t)(m0 + row) * lda1;
        const float* ar2 = ka2 > 0 ? a2 + (size_t)(m0 + row) * lda2 - ka1 : ar1;
        const float* br = b + (size_t)(n0 + row) * ldb;
        for (int g = tid & 3; g < K4; g += 4) {
            float* da = As + row * KP + 4 * (g ^ sw);
            float* db = Bs + row * KP + 4 * (g ^ sw);
            if (oka) cp_async16(da, 4 * g < ka1 ? ar1 + 4 * g : ar2 + 4 * g);
            else *reinterpret_cast<float4*>(da) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (okb) cp_async16(db, br + 4 * g);
            else *reinterpret_cast<float4*>(db) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    cp_async_wait_all();
    MMG_SYNCTHREADS();
    float2 s2[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) s2[a][c] = make_float2(0.f, 0.f);
    const float* arow = As + (ty * 4) * KP;
    const float* brow = Bs + (tx * 4) * KP;
    const int swa = ty & 7, swb = tx & 7;            // (row / 4) % 8 of this thread's four A rows / four B rows
#pragma unroll 2
    for (int g = 0; g < K4; ++g) {
        float4 av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = *reinterpret_cast<const float4*>(arow + a * KP + 4 * (g ^ swa));
#pragma unroll
        for (int c = 0; c < 4; ++c) bv[c] = *reinterpret_cast<const float4*>(brow + c * KP + 4 * (g ^ swb));
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) {       // (even k, odd k) partial sums ride in one packed register pair
                s2[a][c] = ffma2(make_float2(av[a].x, av[a].y), make_float2(bv[c].x, bv[c].y), s2[a][c]);
                s2[a][c] = ffma2(make_float2(av[a].z, av[a].w), make_float2(bv[c].z, bv[c].w), s2[a][c]);
            }
    }
    MMG_SYNCTHREADS();              // the caller may reuse `smem`
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = s2[a][c].x + s2[a][c].y;
}

}  // namespace mmg
