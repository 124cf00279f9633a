// Generic fp32 FFMA tile GEMM used for every non-recurrent product of the path: the image-layer GEMM, the two
// baseline MLPs and all weight gradients.  C[i][j] = sum_k A(k, i) * B(k, j) over k in [k0, k1).
// 64x64 output tile per 256-thread CTA, 4x4 register micro-tile per thread, 16-deep shared-memory chunks.
// fp32 on the CUDA cores on purpose: parity with the reference is 1e-4 on logits and bit-exact on sampled
// bits, which TF32 tensor-core products (10-bit mantissa) do not meet at K=2048.
#pragma once
#include "mmg_platform.cuh"

namespace mmg {

enum { kTile = 64, kChunk = 16, kLd = kTile + 4, kGemmThreads = 256 };
enum { OP_PLAIN = 0, OP_RELUGRAD = 1 };

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
    int kind;
};

MMG_DEVICE float operand_load(const Operand& op, int k, int i) {
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

// Load a kChunk x kTile chunk of an operand into shared memory S[kk][ii] (row stride kLd), zero filled outside
// [0,K) x [0,N).  Thread->element mapping follows the operand's contiguous direction for coalescing.
MMG_DEVICE void load_chunk(const Operand& op, float* S, int kbase, int kend, int ibase, int ilim, int tid) {
#pragma unroll
    for (int l = 0; l < (kChunk * kTile) / kGemmThreads; ++l) {
        int idx = tid + l * kGemmThreads;
        int kk, ii;
        if (op.kmajor) { kk = idx / kTile; ii = idx % kTile; }
        else           { kk = idx % kChunk; ii = idx / kChunk; }
        int k = kbase + kk, i = ibase + ii;
        float v = (k < kend && i < ilim) ? operand_load(op, k, i) : 0.f;
        S[kk * kLd + ii] = v;
    }
}

// Accumulates the 4x4 micro-tile of thread (ty = tid / 16, tx = tid % 16): rows i = ty*4.., cols j = tx*4..
// `colsum` (optional, threads < kTile): sum_k A(k, m0 + tid) as a by-product (bias gradients).
MMG_DEVICE void gemm_tile(const Operand& A, const Operand& Bm, int Mdim, int Ndim, int m0, int n0, int k0, int k1,
                          float (&acc)[4][4], float* colsum, float* As, float* Bs) {
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    float cs = 0.f;
    for (int kb = k0; kb < k1; kb += kChunk) {
        load_chunk(A, As, kb, k1, m0, Mdim, tid);
        load_chunk(Bm, Bs, kb, k1, n0, Ndim, tid);
        MMG_SYNCTHREADS();
#pragma unroll
        for (int kk = 0; kk < kChunk; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(As + kk * kLd + ty * 4);
            const float4 b4 = *reinterpret_cast<const float4*>(Bs + kk * kLd + tx * 4);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }
        if (colsum != nullptr && tid < kTile) {
#pragma unroll
            for (int kk = 0; kk < kChunk; ++kk) cs += As[kk * kLd + tid];
        }
        MMG_SYNCTHREADS();
    }
    if (colsum != nullptr && tid < kTile) *colsum = cs;
}

}  // namespace mmg
