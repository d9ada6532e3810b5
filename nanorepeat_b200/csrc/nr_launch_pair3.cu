// pair_ladder_kernel: round 3 on u16x2 words, two reads per warp, plus the batch's 32-bit entries (nr_pair_kernels.cuh)
#define NR_DEFINE_PAIR_LADDER_KERNEL
#include "nr_launch.h"
#include <mutex>
namespace nrl {
cudaError_t launch_pair_ladder(int blocks, int threads, size_t smem, cudaStream_t st, const nr::pr::Pair3* pairs,
                               const nr::pr::Deal& deal, const nr::LadderTask* tasks, const nr::RestArgs& ra,
                               const uint32_t* qpool, const uint32_t* pool, const nr::LadderRegion* regs, const nr::ScoreW& k,
                               int* counter, int stride, uint2* prung, int4* out, int4* sel, int* redo_count, int32_t* redo,
                               const uint32_t* qstate) {
    static std::mutex mu;
    static bool done = false;
    {
        std::lock_guard<std::mutex> lk(mu);
        cudaError_t e = prepare(nr::pr::pair_ladder_kernel, done);
        if (e != cudaSuccess) return e;
    }
    nr::pr::pair_ladder_kernel<<<blocks, threads, smem, st>>>(pairs, deal, tasks, ra, qpool, pool, regs, k, counter, stride, prung,
                                                             out, sel, redo_count, redo, qstate);
    return cudaGetLastError();
}
}  // namespace nrl
