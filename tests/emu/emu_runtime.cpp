// TEST INFRASTRUCTURE: CPU emulation runtime for the kernels in multimodalgame_b200/csrc (see mmg_platform.cuh).
// Every CUDA thread of a block is an OS thread; blocks run one after another.  Built only by tests/emu/build_emu.sh
// into tests/emu/libmmg_emu.so, loaded only by tests — never by the product package.
#define MMG_CPU_EMU 1
#include "../../multimodalgame_b200/csrc/mmg_platform.cuh"

namespace mmg {
namespace emu {
thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
thread_local BlockCtx* t_ctx = nullptr;

unsigned char* dyn_smem() { return t_ctx->dyn; }
void syncthreads() { t_ctx->block_bar->arrive_and_wait(); }
void syncwarp() { (*t_ctx->warp_bars)[t_threadIdx.x / 32]->arrive_and_wait(); }
double shfl_xor(double v, int lane_mask) {
    const int tid = t_threadIdx.x;
    t_ctx->shfl[tid] = v;
    syncwarp();
    const double r = t_ctx->shfl[tid ^ lane_mask];
    syncwarp();
    return r;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const int nthreads = (int)block.x;
    std::barrier<> block_bar(nthreads);
    std::vector<std::unique_ptr<std::barrier<>>> warp_bars;
    for (int w = 0; w < (nthreads + 31) / 32; ++w) {
        int cnt = std::min(32, nthreads - w * 32);
        warp_bars.emplace_back(new std::barrier<>(cnt));
    }
    std::vector<unsigned char> dyn(smem + 256);
    unsigned char* dyn_aligned = (unsigned char*)(((uintptr_t)dyn.data() + 127) & ~(uintptr_t)127);
    std::vector<double> shfl(nthreads);
    BlockCtx ctx{&block_bar, &warp_bars, dyn_aligned, shfl.data()};
    auto worker = [&](int tid) {
        t_ctx = &ctx;
        t_threadIdx = dim3(tid);
        t_blockDim = block;
        t_gridDim = grid;
        for (unsigned b = 0; b < grid.x; ++b) {
            t_blockIdx = dim3(b);
            body();
            block_bar.arrive_and_wait();
        }
    };
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
}
}  // namespace emu
}  // namespace mmg
