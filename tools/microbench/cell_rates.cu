// Microbenchmark: sustained rate of the W32 DP cell (6 DPX + 5 adds) as a function of where its constants live
// (uniform registers / ordinary registers / immediates) and of how the adds are issued (IMAD on the FMA pipe vs IADD).
// NCH independent cells per thread, so latency is hidden and the number is the pipe / operand-delivery bound.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cell_rates cell_rates.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
constexpr int NCH = 6, ITERS = 2048;
struct SC { int ho1, he1, ho2, he2, vo1, ve1, vo2, ve2, one; };
__device__ __forceinline__ int madd(int a, int m, int b) { int d; asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(m), "r"(b)); return d; }
__device__ __forceinline__ int pin(int v) { int r; asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v)); return r; }   // keep in an R register
enum { V_UR = 0, V_REG = 1, V_IMM = 2, V_IADD = 3, V_IMM_IADD = 4, V_COUNT };
const char* vnames[] = {"constants uniform (UR / c[]), adds as IMAD", "constants in registers, adds as IMAD",
                        "constants immediate, adds as IMAD", "constants uniform, adds as IADD (ALU pipe)",
                        "constants immediate, adds as IADD"};
template <int V>
__global__ void __launch_bounds__(128) cellk(int* out, SC p, long long* cycles) {
    SC sc = p;
    if (V == V_REG) { sc.ho1 = pin(p.ho1); sc.he1 = pin(p.he1); sc.ho2 = pin(p.ho2); sc.he2 = pin(p.he2);
                      sc.vo1 = pin(p.vo1); sc.ve1 = pin(p.ve1); sc.vo2 = pin(p.vo2); sc.ve2 = pin(p.ve2); }
    if (V == V_IMM || V == V_IMM_IADD) { sc.ho1 = -393217; sc.he1 = -131073; sc.ho2 = -1638401; sc.he2 = -65537;
                      sc.vo1 = -393216; sc.ve1 = -131072; sc.vo2 = -1638400; sc.ve2 = -65536; }
    int hd[NCH], e1[NCH], e2[NCH], f1[NCH], f2[NCH], s[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { hd[i] = threadIdx.x * (i + 1); e1[i] = sc.ho1 - i; e2[i] = sc.ho2 - i; f1[i] = sc.vo1 - i; f2[i] = sc.vo2 - i; s[i] = (threadIdx.x & 1) ? 131071 : -262145; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const bool iadd = (V == V_IADD || V == V_IMM_IADD);
            const int d = iadd ? hd[i] + s[i] : madd(hd[i], sc.one, s[i]);
            const int t = __vimax3_s32(d, e1[i], e2[i]);
            const int h = __vimax3_s32_relu(t, f1[i], f2[i]);
            e1[i] = __viaddmax_s32(h, sc.ho1, iadd ? e1[i] + sc.he1 : madd(e1[i], sc.one, sc.he1));
            e2[i] = __viaddmax_s32(h, sc.ho2, iadd ? e2[i] + sc.he2 : madd(e2[i], sc.one, sc.he2));
            f1[i] = __viaddmax_s32(h, sc.vo1, iadd ? f1[i] + sc.ve1 : madd(f1[i], sc.one, sc.ve1));
            f2[i] = __viaddmax_s32(h, sc.vo2, iadd ? f2[i] + sc.ve2 : madd(f2[i], sc.one, sc.ve2));
            hd[i] = h;
        }
    }
    long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc ^= hd[i] ^ e1[i] ^ e2[i] ^ f1[i] ^ f2[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int V> void run(int nsm, int blocks_per_sm, int* dout, long long* dcyc) {
    SC p = {-393217, -131073, -1638401, -65537, -393216, -131072, -1638400, -65536, 1};
    const int nb = nsm * blocks_per_sm;
    cellk<V><<<nb, 128>>>(dout, p, dcyc); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    cellk<V><<<nb, 128>>>(dout, p, dcyc);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long* h = (long long*)malloc(nb * sizeof(long long));
    CK(cudaMemcpy(h, dcyc, nb * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < nb; ++i) avg += h[i]; avg /= nb; free(h);
    // per SMSP: blocks_per_sm warps, each ITERS * NCH cells
    const double clk_per_cell = avg / ((double)ITERS * NCH * blocks_per_sm);
    const double gcups = (double)nb * 128.0 * ITERS * NCH / (ms * 1e-3) / 1e9;      // from the event time of the whole launch
    printf("{\"variant\":\"%s\",\"warps_per_smsp\":%d,\"clk_per_warp_cell_per_smsp\":%.2f,\"launch_ms\":%.4f,\"gcups_from_event_time\":%.0f}\n",
           vnames[V], blocks_per_sm, clk_per_cell, ms, gcups);
}
int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int nsm = pr.multiProcessorCount;
    int* dout; long long* dcyc; CK(cudaMalloc(&dout, nsm * 16 * 128 * sizeof(int))); CK(cudaMalloc(&dcyc, nsm * 16 * sizeof(long long)));
    for (int w : {1, 2, 4, 8}) { run<V_UR>(nsm, w, dout, dcyc); run<V_REG>(nsm, w, dout, dcyc); run<V_IMM>(nsm, w, dout, dcyc); run<V_IADD>(nsm, w, dout, dcyc); run<V_IMM_IADD>(nsm, w, dout, dcyc); }
    return 0;
}
