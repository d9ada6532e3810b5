// Microbenchmark: issue rates of the integer / DPX instructions the DP kernels are built from.
// Measures warp-instructions per clock per SM for long independent chains, so the DPX roofline
// (SMs x clock x lanes/clk x cells/instr / instr/cell) uses measured rates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

constexpr int NCHAIN = 8;
constexpr int ITERS = 4096;
constexpr int UNROLL = 8;

enum Op { OP_IADD, OP_VIADDMNMX, OP_VIMNMX3, OP_VIADDMNMX_S16, OP_VIMNMX3_S16, OP_VIADD16, OP_IMAD,
          OP_MIX_2DPX_1IMAD, OP_MIX_1DPX_1IMAD, OP_MIX_1DPX_1IADD, OP_SHFL, OP_LDS128, OP_MIX_DPX_SHFL,
          OP_MIX_DPX_LDS, OP_VIMNMX_S16, OP_LOP3, OP_ISETP_SEL, OP_MIX_3DPX_2IMAD, OP_COUNT };
const char* names[] = {"IADD3","VIADDMNMX.s32","VIMNMX3.s32","VIADDMNMX.s16x2","VIMNMX3.s16x2","VIADD.16x2","IMAD",
  "mix 2 VIADDMNMX + 1 IMAD","mix 1 VIADDMNMX + 1 IMAD","mix 1 VIADDMNMX + 1 IADD3","SHFL.UP","LDS.128",
  "mix 4 VIADDMNMX + 1 SHFL","mix 4 VIADDMNMX + 1 LDS.128","VIMNMX.s16x2","LOP3","ISETP+SEL","mix 3 VIADDMNMX + 2 IMAD"};
// warp-instructions executed per inner body per chain
const int instr_per_body[] = {1,1,1,1,1,1,1,3,2,2,1,1,5,5,1,1,2,5};

template<int OP>
__global__ void __launch_bounds__(1024,1) bench(int* out, int one, int c1, int c2, long long* cycles) {
  __shared__ int4 sm[1024];
  int x[NCHAIN], y[NCHAIN];
  #pragma unroll
  for (int i=0;i<NCHAIN;i++){ x[i]=threadIdx.x*(i+1)+c1; y[i]=threadIdx.x^(i*77)+c2; }
  sm[threadIdx.x]=make_int4(x[0],x[1],x[2],x[3]);
  __syncthreads();
  long long t0=clock64();
  for (int it=0; it<ITERS/UNROLL; ++it) {
    #pragma unroll
    for (int u=0;u<UNROLL;u++) {
      #pragma unroll
      for (int i=0;i<NCHAIN;i++) {
        if (OP==OP_IADD) { asm volatile("add.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i])); }
        else if (OP==OP_VIADDMNMX) x[i]=__viaddmax_s32(x[i],c1,y[i]);
        else if (OP==OP_VIMNMX3) x[i]=__vimax3_s32(x[i],y[i],c2+u);
        else if (OP==OP_VIADDMNMX_S16) x[i]=__viaddmax_s16x2(x[i],c1,y[i]);
        else if (OP==OP_VIMNMX3_S16) x[i]=__vimax3_s16x2(x[i],y[i],c2+u);
        else if (OP==OP_VIADD16) x[i]=__vadd2(x[i],y[i]);
        else if (OP==OP_VIMNMX_S16) x[i]=__vmaxs2(x[i]^u,y[i]);
        else if (OP==OP_IMAD) x[i]=x[i]*one+y[i];
        else if (OP==OP_LOP3) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(c1)); }
        else if (OP==OP_ISETP_SEL) x[i]= (x[i]>y[i]) ? c1 : (x[i]+u);
        else if (OP==OP_MIX_2DPX_1IMAD) { x[i]=__viaddmax_s32(x[i],c1,y[i]); y[i]=y[i]*one+c2; x[i]=__viaddmax_s32(x[i],c2,y[i]); }
        else if (OP==OP_MIX_3DPX_2IMAD) { x[i]=__viaddmax_s32(x[i],c1,y[i]); y[i]=y[i]*one+c2; x[i]=__viaddmax_s32(x[i],c2,y[i]); y[i]=y[i]*one+c1; x[i]=__vimax3_s32(x[i],c2,y[i]); }
        else if (OP==OP_MIX_1DPX_1IMAD) { x[i]=__viaddmax_s32(x[i],c1,y[i]); y[i]=y[i]*one+c2; }
        else if (OP==OP_MIX_1DPX_1IADD) { x[i]=__viaddmax_s32(x[i],c1,y[i]); asm volatile("add.s32 %0, %0, %1;" : "+r"(y[i]) : "r"(c2)); }
        else if (OP==OP_SHFL) x[i]=__shfl_up_sync(0xffffffffu,x[i],1);
        else if (OP==OP_LDS128) { int4 v=sm[(threadIdx.x+x[i])&1023]; x[i]=v.x^v.y^v.z^v.w; }
        else if (OP==OP_MIX_DPX_SHFL) { x[i]=__viaddmax_s32(x[i],c1,y[i]); y[i]=__viaddmax_s32(y[i],c2,x[i]); x[i]=__viaddmax_s32(x[i],c2,y[i]); y[i]=__viaddmax_s32(y[i],c1,x[i]); x[i]=__shfl_up_sync(0xffffffffu,x[i],1); }
        else if (OP==OP_MIX_DPX_LDS) { x[i]=__viaddmax_s32(x[i],c1,y[i]); y[i]=__viaddmax_s32(y[i],c2,x[i]); x[i]=__viaddmax_s32(x[i],c2,y[i]); y[i]=__viaddmax_s32(y[i],c1,x[i]); int4 v=sm[(threadIdx.x+it+i)&1023]; x[i]=__vimax3_s32(x[i],v.x,v.w); }
      }
    }
  }
  long long t1=clock64();
  int acc=0;
  #pragma unroll
  for (int i=0;i<NCHAIN;i++) acc^=x[i]^y[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
  if (threadIdx.x==0) cycles[blockIdx.x]=t1-t0;
}

template<int OP> void run(int nsm, int* dout, long long* dcyc) {
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  bench<OP><<<nsm,1024>>>(dout,1,-393216,-131072,dcyc); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  bench<OP><<<nsm,1024>>>(dout,1,-393216,-131072,dcyc);
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms,e0,e1));
  long long* h=(long long*)malloc(nsm*sizeof(long long));
  CK(cudaMemcpy(h,dcyc,nsm*sizeof(long long),cudaMemcpyDeviceToHost));
  double avg=0; for(int i=0;i<nsm;i++) avg+=h[i]; avg/=nsm; free(h);
  double winstr = (double)ITERS*NCHAIN*instr_per_body[OP]*32.0; // warp instrs per SM (32 warps)
  printf("{\"op\":\"%s\",\"warp_instr_per_clk_per_sm\":%.3f,\"lane_ops_per_clk_per_sm\":%.1f,\"ms\":%.3f,\"eff_mhz\":%.0f}\n",
         names[OP], winstr/avg, winstr*32/avg, ms, avg/ms/1e3);
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int nsm=p.multiProcessorCount;
  printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n",p.name,nsm,p.clockRate);
  int* dout; long long* dcyc; CK(cudaMalloc(&dout,nsm*1024*sizeof(int))); CK(cudaMalloc(&dcyc,nsm*sizeof(long long)));
  run<OP_IADD>(nsm,dout,dcyc); run<OP_VIADDMNMX>(nsm,dout,dcyc); run<OP_VIMNMX3>(nsm,dout,dcyc);
  run<OP_VIADDMNMX_S16>(nsm,dout,dcyc); run<OP_VIMNMX3_S16>(nsm,dout,dcyc); run<OP_VIADD16>(nsm,dout,dcyc);
  run<OP_VIMNMX_S16>(nsm,dout,dcyc); run<OP_IMAD>(nsm,dout,dcyc); run<OP_LOP3>(nsm,dout,dcyc); run<OP_ISETP_SEL>(nsm,dout,dcyc);
  run<OP_MIX_2DPX_1IMAD>(nsm,dout,dcyc); run<OP_MIX_3DPX_2IMAD>(nsm,dout,dcyc); run<OP_MIX_1DPX_1IMAD>(nsm,dout,dcyc); run<OP_MIX_1DPX_1IADD>(nsm,dout,dcyc);
  run<OP_SHFL>(nsm,dout,dcyc); run<OP_LDS128>(nsm,dout,dcyc); run<OP_MIX_DPX_SHFL>(nsm,dout,dcyc); run<OP_MIX_DPX_LDS>(nsm,dout,dcyc);
  return 0;
}
