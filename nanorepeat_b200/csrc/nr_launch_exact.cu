// exact_kernel<FIXED>: (score, tstart, tend) records on 32-bit words (nr_kernels.cuh)
#include "nr_launch.h"
#include <mutex>
namespace nrl {
cudaError_t launch_exact(bool fixed, int blocks, int threads, size_t smem, cudaStream_t st, const nr::Task* tasks,
                         const nr::RestArgs& ra, const uint32_t* pool, const nr::ScoreW& k, int* counter, int stride, int4* out) {
    static std::mutex mu;
    static bool done[2] = {false, false};
    {
        std::lock_guard<std::mutex> lk(mu);
        cudaError_t e = fixed ? prepare(nr::exact_kernel<true>, done[1]) : prepare(nr::exact_kernel<false>, done[0]);
        if (e != cudaSuccess) return e;
    }
    if (fixed) nr::exact_kernel<true><<<blocks, threads, smem, st>>>(tasks, ra, pool, k, counter, stride, out);
    else nr::exact_kernel<false><<<blocks, threads, smem, st>>>(tasks, ra, pool, k, counter, stride, out);
    return cudaGetLastError();
}
}  // namespace nrl
