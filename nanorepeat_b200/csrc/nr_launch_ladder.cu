// ladder_kernel<true, FLAG>: round-3 ladders on 32-bit words, flag words (mode 2, redo launches) or span words (mode 1)
#include "nr_launch.h"
#include <mutex>
namespace nrl {
cudaError_t launch_ladder(bool flag, int blocks, int threads, size_t smem, cudaStream_t st, const nr::LadderTask* tasks,
                          const nr::RestArgs& ra, const int* n_order_dev, const uint32_t* qpool, const uint32_t* pool,
                          const nr::LadderRegion* regs, const nr::ScoreW& k, int* counter, int stride, int4* out, int4* sel) {
    static std::mutex mu;
    static bool done[2] = {false, false};
    {
        std::lock_guard<std::mutex> lk(mu);
        cudaError_t e = flag ? prepare(nr::ladder_kernel<true, true>, done[1]) : prepare(nr::ladder_kernel<true, false>, done[0]);
        if (e != cudaSuccess) return e;
    }
    if (flag) nr::ladder_kernel<true, true><<<blocks, threads, smem, st>>>(tasks, ra, n_order_dev, qpool, pool, regs, k, counter, stride, out, sel);
    else nr::ladder_kernel<true, false><<<blocks, threads, smem, st>>>(tasks, ra, n_order_dev, qpool, pool, regs, k, counter, stride, out, sel);
    return cudaGetLastError();
}
}  // namespace nrl
