// Microbenchmark: dependent-issue latency (cycles) of the instructions on the DP kernels' critical path,
// one warp per SM, one dependent chain.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
constexpr int ITERS = 8192;
enum Op { L_IADD, L_VIADDMNMX, L_VIMNMX3, L_VIMNMX3_RELU, L_CELLCHAIN, L_SHFL, L_LDS, L_DPX_THEN_IADD, L_IADD_THEN_DPX, L_COUNT };
const char* names[] = {"IADD -> IADD", "VIADDMNMX -> VIADDMNMX", "VIMNMX3 -> VIMNMX3", "VIMNMX3.RELU -> VIMNMX3.RELU",
                       "cell chain: VIMNMX3.RELU(h) -> VIADDMNMX(f) per row (2 instr)", "SHFL.UP -> SHFL.UP", "LDS -> LDS (dependent address)",
                       "VIADDMNMX -> IADD (2 instr)", "IADD -> VIADDMNMX (2 instr)"};
template<int OP>
__global__ void lat(int* out, int c1, int c2, long long* cycles) {
  __shared__ int sm[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = (i + 1) & 1023;
  __syncwarp();
  int x = threadIdx.x + c1, y = threadIdx.x ^ c2, f = c2;
  long long t0 = clock64();
  #pragma unroll 16
  for (int it = 0; it < ITERS; ++it) {
    if (OP == L_IADD) { asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(y)); }
    else if (OP == L_VIADDMNMX) x = __viaddmax_s32(x, c1, y);
    else if (OP == L_VIMNMX3) x = __vimax3_s32(x, y, c2);
    else if (OP == L_VIMNMX3_RELU) x = __vimax3_s32_relu(x, y, c2);
    else if (OP == L_CELLCHAIN) { x = __vimax3_s32_relu(y, f, c2); f = __viaddmax_s32(x, c1, f); }
    else if (OP == L_SHFL) x = __shfl_up_sync(0xffffffffu, x, 1);
    else if (OP == L_LDS) x = sm[x & 1023];
    else if (OP == L_DPX_THEN_IADD) { x = __viaddmax_s32(x, c1, y); asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(c2)); }
    else if (OP == L_IADD_THEN_DPX) { asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(c2)); x = __viaddmax_s32(x, c1, y); }
  }
  long long t1 = clock64();
  out[blockIdx.x * 32 + threadIdx.x] = x ^ f;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template<int OP> void run(int* dout, long long* dcyc) {
  lat<OP><<<1, 32>>>(dout, -393216, -131072, dcyc); CK(cudaDeviceSynchronize());
  lat<OP><<<1, 32>>>(dout, -393216, -131072, dcyc); CK(cudaDeviceSynchronize());
  long long h; CK(cudaMemcpy(&h, dcyc, sizeof h, cudaMemcpyDeviceToHost));
  printf("{\"chain\":\"%s\",\"cycles_per_iteration\":%.2f}\n", names[OP], (double)h / ITERS);
}
int main() {
  int* dout; long long* dcyc; CK(cudaMalloc(&dout, 4096)); CK(cudaMalloc(&dcyc, 64));
  run<L_IADD>(dout, dcyc); run<L_VIADDMNMX>(dout, dcyc); run<L_VIMNMX3>(dout, dcyc); run<L_VIMNMX3_RELU>(dout, dcyc);
  run<L_CELLCHAIN>(dout, dcyc); run<L_SHFL>(dout, dcyc); run<L_LDS>(dout, dcyc); run<L_DPX_THEN_IADD>(dout, dcyc); run<L_IADD_THEN_DPX>(dout, dcyc);
  return 0;
}
