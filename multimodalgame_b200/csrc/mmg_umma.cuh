// 5th-generation tensor cores (tcgen05) for the one contraction of the path that is GEMM-shaped enough to pay for them:
// the sender's image layer  h_x = x . W_img^T  (model.py:195; (B x F) x (F x Hi), F = 2048..4096).
//
//   * operands: both are K-major in HBM (the feature dimension is contiguous in x and in image_layer.weight), so a tile is a set
//     of rows x 128-byte segments.  They are staged into shared memory in the canonical K-major SWIZZLE_128B layout of the
//     UMMA shared-memory descriptor (8-row x 128-byte atoms, the 16-byte chunk index XOR-ed with the row index) by
//     asynchronous 16-byte copies — the swizzle granule is exactly the copy size, so every copy lands at its final address;
//   * precision: fp32 parity with the reference is 1e-4 on logits and bit-exact on sampled bits, which a plain TF32 product
//     (10-bit mantissa) does not meet at K = 2048.  Every operand is therefore split in shared memory into hi = the TF32-
//     representable head (low 13 mantissa bits cleared) and lo = x - hi, and the product is accumulated as
//     hi.hi + lo.hi + hi.lo in the fp32 TMEM accumulator (3xTF32: relative error ~2^-21, the dropped lo.lo term);
//   * issue: one elected thread issues the tcgen05.mma instructions (kind::tf32, cta_group::1, M = 128, N = 64, K = 8 per
//     instruction) and commits them to an mbarrier; the accumulator (128 lanes x 64 columns of TMEM) is read back with
//     tcgen05.ld (32 lanes x 16 columns per warp-instruction) for the split-K partial store.
// The emulated CPU build (tests/emu) has no tensor cores: it keeps the FFMA tiles (mmg_gemm.cuh) for this product.
#pragma once
#include "mmg_platform.cuh"

#ifndef MMG_CPU_EMU
#include <cuda.h>        // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)
namespace mmg {
// TMA descriptors of the two operands of the image layer (K-major, fp32, box = 32 floats x tile rows, SWIZZLE_128B): the
// tensor-map unit writes a tile straight into the canonical UMMA shared-memory layout, out-of-range rows arrive as zeros.
struct ImageTmaps { CUtensorMap w, x; };
#define MMG_GRID_CONSTANT __grid_constant__
namespace umma {

enum { kM = 128, kN = 64, kAtomK = 32, kMaxSliceK = 128, kTmemCols = 64 };
// shared memory of one tile: A hi/lo (kM rows) + B hi/lo (kN rows), each rows x slice_k floats; + 1 KB alignment slack
MMG_HOST_DEVICE int tile_smem_bytes(int slice_k) { return 2 * (kM + kN) * slice_k * 4 + 1024; }

MMG_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
MMG_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
MMG_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14), leading
// byte offset (unused by swizzled K-major layouts, canonical value 1) in [16,30), stride byte offset = 1024 B between 8-row
// groups in [32,46), descriptor version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
MMG_DEVICE uint64_t smem_desc(const void* p) {
    const uint32_t a = smem_u32(p);
    return (uint64_t)((a >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits [4,6) = 1), A = B = TF32 (bits [7,10), [10,13) = 2), both
// K-major (bits 15, 16 = 0), N >> 3 in [17,23), M >> 4 in [24,29).
MMG_DEVICE uint32_t instr_desc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
MMG_DEVICE void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
// All MMAs issued so far by this thread arrive on `bar` when they have completed (implies tcgen05.fence::before_thread_sync).
MMG_DEVICE void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One warp allocates `kTmemCols` TMEM columns; the base address lands in *slot (shared memory).
MMG_DEVICE void tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
MMG_DEVICE void tmem_free(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"((uint32_t)kTmemCols) : "memory");
}
// 32 lanes x 16 consecutive columns of the accumulator -> 16 registers per thread (thread l of the warp reads lane base + l)
MMG_DEVICE void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"        // same asm statement: the registers are only read after the wait
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Stage `rows` rows x `nk` floats (nk % 32 == 0) of a K-major matrix (row r at src + r * ld, 16-byte aligned) into the
// SWIZZLE_128B layout: atom a (32 floats of K) occupies rows * 128 bytes, row r at + r * 128, 16-byte chunk c at ((c ^ (r & 7)) * 16).
// Rows >= row_limit are zero.  Then split in place: hi keeps the TF32 head, lo = x - hi goes to the same offset of `lo`.
// Every thread splits exactly the chunks it copied itself, so its own cp.async.wait_all is all the ordering it needs.
template <int NT>
MMG_DEVICE void stage_split(unsigned char* hi, unsigned char* lo, const float* src, size_t ld, int rows, int row_limit, int nk, int tid) {
    const int chunks_per_row = nk >> 2, cpa = kAtomK >> 2;       // 16-byte chunks per row / per atom row
    const int total = rows * chunks_per_row;
    for (int idx = tid; idx < total; idx += NT) {
        const int r = idx / chunks_per_row, cc = idx % chunks_per_row;
        const int a = cc / cpa, c = cc % cpa;
        unsigned char* d = hi + (size_t)a * rows * 128 + r * 128 + ((c ^ (r & 7)) << 4);
        if (r < row_limit) cp_async16(d, src + (size_t)r * ld + 4 * cc);
        else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_wait_all();
    for (int idx = tid; idx < total; idx += NT) {
        const int r = idx / chunks_per_row, cc = idx % chunks_per_row;
        const int a = cc / cpa, c = cc % cpa;
        const size_t off = (size_t)a * rows * 128 + r * 128 + ((c ^ (r & 7)) << 4);
        const float4 v = *reinterpret_cast<const float4*>(hi + off);
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
        *reinterpret_cast<float4*>(hi + off) = h;
        *reinterpret_cast<float4*>(lo + off) = l;
    }
}

// 2-D tensor-map load of one box (32 floats of K x `rows` rows) into shared memory; completion is counted in bytes on `bar`.
MMG_DEVICE void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int k, int row, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tm), "r"(k), "r"(row), "r"(smem_u32(bar))
        : "memory");
}

// One split-K tile of the image layer, transposed:  D^T[n][b] = sum_{f in slice} W_img[n0 + n][f] * x[b0 + b][f]
// (M = 128 hidden units on the TMEM lanes, N = 64 batch rows on the columns).  256 threads.  Partial sums go to
// part[(b0 + b) * Hi + n0 + n] for b0 + b < B.
// `tm` != nullptr: the operands are staged by the TMA unit (cp.async.bulk.tensor, one box per 32-float atom and operand) and
// only split (hi / lo) by the threads; nullptr: asynchronous 16-byte copies issued by the threads.
MMG_DEVICE void image_layer_tile(const float* w_img, const float* x, int F, int Hi, int B, int n0, int b0, int k0, int nk,
                                 float* part, unsigned char* smem_raw, const ImageTmaps* tm) {
    constexpr int NT = 256;
    MMG_SHARED uint64_t s_bar;
    MMG_SHARED uint64_t s_ld;
    MMG_SHARED uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // SWIZZLE_128B atoms: 1024-byte aligned
    unsigned char* a_hi = base;
    unsigned char* a_lo = a_hi + (size_t)kM * nk * 4;
    unsigned char* b_hi = a_lo + (size_t)kM * nk * 4;
    unsigned char* b_lo = b_hi + (size_t)kN * nk * 4;
    if (tid == 0) { mbar_init(&s_bar, 1); mbar_init(&s_ld, 1); mbar_fence_init(); }
    if (tm != nullptr) {
        __syncthreads();                 // the barrier is initialised before the TMA unit may signal it
        if (tid == 0) {
            const int atoms = nk / kAtomK;
            mbar_expect_tx(&s_ld, (uint32_t)((kM + kN) * nk * 4));
            for (int a = 0; a < atoms; ++a) {
                tma_load_2d(a_hi + (size_t)a * kM * 128, &tm->w, k0 + a * kAtomK, n0, &s_ld);
                tma_load_2d(b_hi + (size_t)a * kN * 128, &tm->x, k0 + a * kAtomK, b0, &s_ld);
            }
        }
        if (warp == 0) tmem_alloc(&s_tmem);
        mbar_wait(&s_ld, 0);
        // split in place: hi keeps the TF32 head, lo = x - hi at the same (swizzled) offset of the lo buffer; A | lo A | B | lo B
        // are contiguous, every thread walks 16-byte chunks of the raw tiles
        const int chunks_a = kM * nk / 4, chunks_b = kN * nk / 4;
        for (int idx = tid; idx < chunks_a + chunks_b; idx += NT) {
            unsigned char* h = idx < chunks_a ? a_hi + (size_t)idx * 16 : b_hi + (size_t)(idx - chunks_a) * 16;
            unsigned char* l = idx < chunks_a ? a_lo + (size_t)idx * 16 : b_lo + (size_t)(idx - chunks_a) * 16;
            const float4 v = *reinterpret_cast<const float4*>(h);
            float4 hh, ll;
            hh.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); ll.x = v.x - hh.x;
            hh.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); ll.y = v.y - hh.y;
            hh.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); ll.z = v.z - hh.z;
            hh.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); ll.w = v.w - hh.w;
            *reinterpret_cast<float4*>(h) = hh;
            *reinterpret_cast<float4*>(l) = ll;
        }
    } else {
        if (warp == 0) tmem_alloc(&s_tmem);
        stage_split<NT>(a_hi, a_lo, w_img + (size_t)n0 * F + k0, (size_t)F, kM, Hi - n0, nk, tid);
        stage_split<NT>(b_hi, b_lo, x + (size_t)b0 * F + k0, (size_t)F, kN, B - b0, nk, tid);
    }
    MMG_TRACE_AT(0, 1);
    fence_proxy_async();            // the staged operands (generic-proxy writes) become visible to the tensor-core (async) proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t idesc = instr_desc(kM, kN);
        uint32_t acc = 0;
        for (int a = 0; a < nk / kAtomK; ++a) {
            const uint64_t dah = smem_desc(a_hi + (size_t)a * kM * 128), dal = smem_desc(a_lo + (size_t)a * kM * 128);
            const uint64_t dbh = smem_desc(b_hi + (size_t)a * kN * 128), dbl = smem_desc(b_lo + (size_t)a * kN * 128);
#pragma unroll
            for (int k = 0; k < kAtomK / 8; ++k) {            // K = 8 TF32 (32 bytes) per instruction: + 2 in the address field
                mma_tf32(tmem, dal + 2 * k, dbh + 2 * k, idesc, acc);
                acc = 1;
                mma_tf32(tmem, dah + 2 * k, dbl + 2 * k, idesc, 1);
                mma_tf32(tmem, dah + 2 * k, dbh + 2 * k, idesc, 1);
            }
        }
        mma_commit(&s_bar);
    }
    mbar_wait(&s_bar, 0);
    tc_fence_after();
    MMG_TRACE_AT(0, 2);
    // epilogue: warp w reads TMEM lanes 32 (w % 4) .. + 31 (hidden units), columns 16-wide groups (batch rows); warps 0-3 take
    // the even groups, warps 4-7 the odd ones.  For a fixed batch row the 32 lanes store 128 contiguous bytes.
    const int n = n0 + 32 * (warp & 3) + lane;
    for (int c0 = 16 * (warp >> 2); c0 < kN; c0 += 32) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0, v);
        if (n < Hi) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int b = b0 + c0 + i;
                if (b < B) part[(size_t)b * Hi + n] = v[i];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem);
    MMG_TRACE_AT(0, 7);
}

}  // namespace umma
}  // namespace mmg
#else
namespace mmg {
struct ImageTmaps { int unused; };
}
#define MMG_GRID_CONSTANT
#endif
