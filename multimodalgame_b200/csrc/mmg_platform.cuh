// Platform layer.  Product build: nvcc, sm_100a.  When MMG_CPU_EMU is defined (tests/emu only) the very same
// kernel sources are compiled by g++ and every CUDA thread becomes an OS thread, so kernel logic (indexing,
// barrier structure, shared-memory hazards under ThreadSanitizer) can be checked in the GPU-less build
// container.  The emulated library is test infrastructure: the product package never loads it.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <string.h>

#ifndef MMG_CPU_EMU
// =====================================================================================================
//                                              CUDA
// =====================================================================================================
#include <cuda_runtime.h>

#define MMG_DEVICE __device__ __forceinline__
#define MMG_HOST_DEVICE __host__ __device__ __forceinline__
#define MMG_GLOBAL __global__
#define MMG_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#define MMG_SHARED __shared__

namespace mmg {

MMG_DEVICE float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
MMG_DEVICE float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
MMG_DEVICE double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sum over aligned groups of 16 lanes
MMG_DEVICE float half_warp_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

MMG_DEVICE float shfl_xor_f(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
MMG_DEVICE double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
// Blackwell packed fp32: one FFMA2 instruction = two IEEE-rn fused multiply-adds (bitwise equal to two fmaf)
MMG_DEVICE float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
// Short-latency transcendentals for the recurrent fast path (MUFU.EX2 / MUFU.RCP; absolute error ~2e-7, far inside
// the 1e-4 parity bar).  The generic kernels keep the libm-accurate versions.
MMG_DEVICE float fast_exp(float x) { return __expf(x); }
MMG_DEVICE float fast_rcp(float x) { return __fdividef(1.0f, x); }
MMG_DEVICE float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
MMG_DEVICE float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }
// Additive attention scores tanh(w + h) with the two halves kept as e2(w) = e^{2w}, e2(h) = e^{2h} (arguments clamped so
// that neither factor is 0 or inf): tanh(w + h) = 1 - 2 / (e^{2w} e^{2h} + 1), one reciprocal per element.
MMG_DEVICE float attn_e2(float x) { return __expf(2.0f * fminf(fmaxf(x, -40.0f), 40.0f)); }
MMG_DEVICE float attn_tanh(float ew, float eh) { return 1.0f - 2.0f * __fdividef(1.0f, fmaf(ew, eh, 1.0f)); }
// sum over aligned groups of N lanes (N power of two <= 32); every lane of the group receives the sum
template <int N>
MMG_DEVICE float group_sum(float v) {
#pragma unroll
    for (int o = N / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int N>
MMG_DEVICE float group_max(float v) {
#pragma unroll
    for (int o = N / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// "last CTA done" ticket: returns the number of CTAs that arrived before this one (release/acquire around it)
MMG_DEVICE unsigned ticket_take(unsigned* counter) {
    __threadfence();
    return atomicAdd(counter, 1u);
}
MMG_DEVICE void fence_acquire() { __threadfence(); }
// Producer side: publish this CTA's global writes, then count it.  Consumer side: spin until `target` CTAs have arrived.
// Only legal when the producers cannot be starved by the spinning CTAs (producers have lower block indices and never wait).
MMG_DEVICE void flag_arrive(unsigned* counter) { __threadfence(); atomicAdd(counter, 1u); }
MMG_DEVICE void flag_wait(const unsigned* counter, unsigned target) {
    while (*reinterpret_cast<const volatile unsigned*>(counter) < target) __nanosleep(64);
    __threadfence();
}

// ---- cross-GPU flags over peer-mapped memory (NVLink) ----------------------------------------------------------------
// Protocol rule: everything a flag announces lives in the SIGNALLING rank's own memory (send buffer, reduced slice) and is
// PULLED by the peers over NVLink, whose loads are served by the owner's L2 (the owner's point of coherence; peer addresses are
// never cached in the reader's L2 and the readers bypass L1).  A device-scope fence therefore orders the data (at the owner's
// L2) before the remote flag store leaves; the system-scope fence the first version used here cost 5-12 us per signal because it
// also waits for every outstanding write to leave the GPU.  Small values that are PUSHED (statistics, slice norms) travel as
// self-validating packets instead (ll_store / ll_load below): no flag, no fence.
MMG_DEVICE void peer_signal(unsigned long long* remote_flag, unsigned long long value) {
    __threadfence();
    *reinterpret_cast<volatile unsigned long long*>(remote_flag) = value;
}
// Bounded wait (~15 s: ranks must run in lockstep within that window; a dead peer must not hang the GPU for good).  `err`
// is the rank's sticky error word: once a wait has timed out every later wait gives up at once, the update kernels skip
// their work and the host raises (GameEngine.peer_error()).
MMG_DEVICE bool peer_wait(const unsigned long long* local_flag, unsigned long long target, const int* err) {
    if (*reinterpret_cast<const volatile unsigned long long*>(local_flag) >= target) { __threadfence(); return true; }
    if (*reinterpret_cast<const volatile int*>(err) != 0) return false;
    for (int spin = 0; spin < (1 << 26); ++spin) {
        if (*reinterpret_cast<const volatile unsigned long long*>(local_flag) >= target) { __threadfence(); return true; }
        __nanosleep(spin < 4096 ? 32 : 256);
    }
    return false;
}
MMG_DEVICE void fence_system() { __threadfence_system(); }
// Self-validating packets for small pushed values (the LL idea of NCCL): a double travels as two 8-byte words
// {float bits | tag << 32} (head and remainder, ~48 bits of mantissa), each written by ONE 8-byte store, which NVLink delivers
// whole; the reader spins until both words carry the expected tag (the step number: monotonic, never 0, buffers start zeroed).
MMG_DEVICE void ll_store(double* remote_area, int idx, double v, unsigned long long tag) {
    const float hi = (float)v, lo = (float)(v - (double)hi);
    volatile unsigned long long* q = reinterpret_cast<volatile unsigned long long*>(remote_area) + 2 * (size_t)idx;
    q[0] = (unsigned long long)__float_as_uint(hi) | (tag << 32);
    q[1] = (unsigned long long)__float_as_uint(lo) | (tag << 32);
}
MMG_DEVICE double ll_load(const double* local_area, int idx, unsigned long long tag, int* err) {
    const volatile unsigned long long* q = reinterpret_cast<const volatile unsigned long long*>(local_area) + 2 * (size_t)idx;
    const unsigned t32 = (unsigned)tag;
    unsigned long long a = q[0], b = q[1];
    for (int spin = 0; ((unsigned)(a >> 32) != t32 || (unsigned)(b >> 32) != t32) && spin < (1 << 26); ++spin) {
        if (*reinterpret_cast<const volatile int*>(err) != 0) break;
        __nanosleep(spin < 4096 ? 32 : 256);
        a = q[0]; b = q[1];
    }
    if ((unsigned)(a >> 32) != t32 || (unsigned)(b >> 32) != t32) { *err = 4; return 0.0; }
    return (double)__uint_as_float((unsigned)a) + (double)__uint_as_float((unsigned)b);
}
MMG_DEVICE float4 peer_load4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }   // L1 bypass
// One float4 summed over every rank's copy by the NVSwitch (NVLS): `mc` is a multicast address.
MMG_DEVICE float4 multimem_sum4(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
MMG_DEVICE double peer_load_d(const double* p) { return __ldcg(p); }
// L1-bypassing loads for data another CTA of the SAME grid has just written (split-K partial tiles)
MMG_DEVICE float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
MMG_DEVICE float ld_cg(const float* p) { return __ldcg(p); }
MMG_DEVICE double ld_cg_d(const double* p) { return __ldcg(p); }

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) ------------------------------------------
MMG_DEVICE uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

MMG_DEVICE void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
MMG_DEVICE void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
MMG_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
MMG_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
MMG_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy executed by the TMA unit; bytes % 16 == 0, both addresses 16-byte aligned
MMG_DEVICE void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// One thread stages `bytes` (multiple of 16) in <=32 KB pieces and arms the barrier with the total.
MMG_DEVICE void tma_stage(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    mbar_expect_tx(bar, bytes);
    const uint32_t kPiece = 32768;
    for (uint32_t off = 0; off < bytes; off += kPiece) {
        uint32_t n = bytes - off < kPiece ? bytes - off : kPiece;
        tma_bulk_g2s((char*)smem_dst + off, (const char*)gmem_src + off, n, bar);
    }
}
// Two segments, one barrier phase (a second arrive.expect_tx would count as a second arrival).
MMG_DEVICE void tma_stage2(void* dst1, const void* src1, uint32_t bytes1, void* dst2, const void* src2, uint32_t bytes2,
                           uint64_t* bar) {
    mbar_expect_tx(bar, bytes1 + bytes2);
    const uint32_t kPiece = 32768;
    for (uint32_t off = 0; off < bytes1; off += kPiece)
        tma_bulk_g2s((char*)dst1 + off, (const char*)src1 + off, bytes1 - off < kPiece ? bytes1 - off : kPiece, bar);
    for (uint32_t off = 0; off < bytes2; off += kPiece)
        tma_bulk_g2s((char*)dst2 + off, (const char*)src2 + off, bytes2 - off < kPiece ? bytes2 - off : kPiece, bar);
}
// Ampere-style asynchronous 16-byte global -> shared copies (LDGSTS): fire-and-forget, one wait for all of them.
MMG_DEVICE void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
MMG_DEVICE void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// Programmatic dependent launch: wait for the producer grid's memory to be visible / let dependents start.
// -DMMG_TRACE (debug build only, scripts/trace.py): thread 0 of every CTA stamps the global nanosecond timer at a few points;
// mmg_debug_trace copies the table out.  kid: 0 pre, 1 fwd, 2 baseline, 3 bwd, 4 wgrad, 5 update, 6 peer reduce-scatter.
#ifdef MMG_TRACE
__device__ unsigned long long g_trace[7][1024][8];
#define MMG_TRACE_AT(kid, slot)                                                                  \
    do {                                                                                         \
        if (threadIdx.x == 0 && blockIdx.x < 1024) {                                             \
            unsigned long long t_;                                                               \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                               \
            g_trace[kid][blockIdx.x][slot] = t_;                                                 \
        }                                                                                        \
    } while (0)
#else
#define MMG_TRACE_AT(kid, slot) do { } while (0)
#endif
MMG_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
MMG_DEVICE void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

MMG_DEVICE float ldg(const float* p) { return __ldg(p); }
MMG_DEVICE float4 ldg4(const float4* p) { return __ldg(p); }

}  // namespace mmg

#define MMG_SYNCTHREADS() __syncthreads()
#define MMG_SYNCWARP() __syncwarp()

#else
// =====================================================================================================
//                                   CPU emulation (tests only)
// =====================================================================================================
#include <algorithm>
#include <atomic>
#include <barrier>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct float4 {
    float x, y, z, w;
};
struct float2 {
    float x, y;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0

#define MMG_DEVICE static inline
#define MMG_HOST_DEVICE static inline
#define MMG_GLOBAL static
#define __launch_bounds__(...)
#define __restrict__
#define MMG_DYN_SMEM(name) unsigned char* name = mmg::emu::dyn_smem()
#define MMG_SHARED static

namespace mmg {
namespace emu {
struct BlockCtx {
    std::barrier<>* block_bar;
    std::vector<std::unique_ptr<std::barrier<>>>* warp_bars;
    unsigned char* dyn;
    double* shfl;  // [nthreads]
};
extern thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
extern thread_local BlockCtx* t_ctx;
unsigned char* dyn_smem();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void syncthreads();
void syncwarp();
double shfl_xor(double v, int lane_mask);
}  // namespace emu
}  // namespace mmg

#define threadIdx (mmg::emu::t_threadIdx)
#define blockIdx (mmg::emu::t_blockIdx)
#define blockDim (mmg::emu::t_blockDim)
#define gridDim (mmg::emu::t_gridDim)
#define MMG_SYNCTHREADS() mmg::emu::syncthreads()
#define MMG_SYNCWARP() mmg::emu::syncwarp()

using std::max;
using std::min;

namespace mmg {
MMG_DEVICE float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += (float)emu::shfl_xor(v, o);
    return v;
}
MMG_DEVICE float warp_max(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, (float)emu::shfl_xor(v, o));
    return v;
}
MMG_DEVICE double warp_sum_d(double v) {
    for (int o = 16; o > 0; o >>= 1) v += emu::shfl_xor(v, o);
    return v;
}
MMG_DEVICE float half_warp_sum(float v) {
    for (int o = 8; o > 0; o >>= 1) v += (float)emu::shfl_xor(v, o);
    return v;
}
MMG_DEVICE float shfl_xor_f(float v, int m) { return (float)emu::shfl_xor(v, m); }
MMG_DEVICE double shfl_xor_d(double v, int m) { return emu::shfl_xor(v, m); }
MMG_DEVICE float2 ffma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
MMG_DEVICE float fast_exp(float x) { return expf(x); }
MMG_DEVICE float fast_rcp(float x) { return 1.0f / x; }
MMG_DEVICE float fast_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
MMG_DEVICE float fast_tanh(float x) { return 1.0f - 2.0f / (expf(2.0f * x) + 1.0f); }
MMG_DEVICE float attn_e2(float x) { return expf(2.0f * fminf(fmaxf(x, -40.0f), 40.0f)); }
MMG_DEVICE float attn_tanh(float ew, float eh) { return 1.0f - 2.0f / (fmaf(ew, eh, 1.0f)); }
template <int N>
MMG_DEVICE float group_sum(float v) {
    for (int o = N / 2; o > 0; o >>= 1) v += (float)emu::shfl_xor(v, o);
    return v;
}
template <int N>
MMG_DEVICE float group_max(float v) {
    for (int o = N / 2; o > 0; o >>= 1) v = fmaxf(v, (float)emu::shfl_xor(v, o));
    return v;
}
MMG_DEVICE unsigned ticket_take(unsigned* counter) { return __atomic_fetch_add(counter, 1u, __ATOMIC_SEQ_CST); }
MMG_DEVICE void fence_acquire() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
MMG_DEVICE void flag_arrive(unsigned* counter) { __atomic_fetch_add(counter, 1u, __ATOMIC_SEQ_CST); }
MMG_DEVICE void flag_wait(const unsigned* counter, unsigned target) {   // blocks run in index order: already satisfied
    while (__atomic_load_n(counter, __ATOMIC_SEQ_CST) < target) {}
}
MMG_DEVICE void peer_signal(unsigned long long* f, unsigned long long v) { __atomic_store_n(f, v, __ATOMIC_SEQ_CST); }
MMG_DEVICE bool peer_wait(const unsigned long long* f, unsigned long long target, const int*) { return __atomic_load_n(f, __ATOMIC_SEQ_CST) >= target; }
MMG_DEVICE void fence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
MMG_DEVICE void ll_store(double* area, int idx, double v, unsigned long long) { area[2 * (size_t)idx] = v; }
MMG_DEVICE double ll_load(const double* area, int idx, unsigned long long, int*) { return area[2 * (size_t)idx]; }
MMG_DEVICE float4 peer_load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
MMG_DEVICE float4 multimem_sum4(const float* p) { return *reinterpret_cast<const float4*>(p); }
MMG_DEVICE double peer_load_d(const double* p) { return *p; }
MMG_DEVICE float4 ld_cg4(const float* p) { return *reinterpret_cast<const float4*>(p); }
MMG_DEVICE float ld_cg(const float* p) { return *p; }
MMG_DEVICE double ld_cg_d(const double* p) { return *p; }
MMG_DEVICE void mbar_init(uint64_t* bar, int) { *bar = 0; }
MMG_DEVICE void mbar_fence_init() {}
MMG_DEVICE void mbar_wait(uint64_t*, uint32_t) {}
MMG_DEVICE void tma_stage(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t*) {
    memcpy(smem_dst, gmem_src, bytes);
}
MMG_DEVICE void tma_stage2(void* dst1, const void* src1, uint32_t bytes1, void* dst2, const void* src2, uint32_t bytes2,
                           uint64_t*) {
    memcpy(dst1, src1, bytes1);
    memcpy(dst2, src2, bytes2);
}
MMG_DEVICE void cp_async16(void* smem_dst, const void* gmem_src) { memcpy(smem_dst, gmem_src, 16); }
MMG_DEVICE void cp_async_wait_all() {}
MMG_DEVICE void pdl_wait() {}
MMG_DEVICE void pdl_launch_dependents() {}
#define MMG_TRACE_AT(kid, slot) do { } while (0)
MMG_DEVICE float ldg(const float* p) { return *p; }
MMG_DEVICE float4 ldg4(const float4* p) { return *p; }
}  // namespace mmg
#endif

namespace mmg {
MMG_HOST_DEVICE int cdiv(int a, int b) { return (a + b - 1) / b; }
MMG_HOST_DEVICE int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
MMG_HOST_DEVICE int round_up(int a, int b) { return cdiv(a, b) * b; }
MMG_HOST_DEVICE int64_t round_up64(int64_t a, int64_t b) { return cdiv64(a, b) * b; }

MMG_DEVICE float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// Philox4x32-10 counter-based generator (Salmon et al. 2011) for the on-device sampler.
MMG_DEVICE void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}
// 4 uniforms in (0,1) for counter (iter, a, b) under key `seed`
MMG_DEVICE void philox_uniform4(uint64_t seed, uint64_t iter, uint32_t a, uint32_t b, float (&u)[4]) {
    uint32_t c[4] = {(uint32_t)iter, (uint32_t)(iter >> 32), a, b};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int i = 0; i < 10; ++i) philox_round(c, k);
    for (int i = 0; i < 4; ++i) u[i] = ((float)(c[i] >> 8) + 0.5f) * (1.0f / 16777216.0f);
}
}  // namespace mmg
