// Microbenchmark: the paired (two reads per 32-bit word, u16x2 halves) DP column step with its real dependency pattern.
// Words are biased unsigned halves so that the plain adds stay IMADs on the FMA pipe (no borrow can cross the halves
// while every half stays inside [0, 65535]); the local-alignment floor is the third operand of the diagonal's VIADDMNMX.
//   variant 0: 7 DPX (VIADDMNMX.U16x2 / VIMNMX3.U16x2) + 4 IMAD per cell PAIR, floor folded into the diagonal add
//   variant 1: 7 DPX + 5 IMAD, floor as an extra 2-input max
//   variant 2: signed s16x2 with .RELU, all adds as VIADDMNMX (11 DPX per pair)
// Levels as in cell_chain.cu (0 cells only ... 4 + running-maximum tracking).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pair_chain pair_chain.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
constexpr int STEPS = 4096;
typedef unsigned u32;
__device__ __forceinline__ u32 madd(u32 a, u32 m, u32 b) { u32 d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(m), "r"(b)); return d; }
__host__ __device__ constexpr u32 pk(int v) { return ((u32)(v & 0xffff) << 16) | (u32)(v & 0xffff); }
__host__ __device__ constexpr u32 nk(int v) { return (u32)(-(int)(((u32)v << 16) | (u32)v)); }   // minus (v, v) as a 32-bit addend
constexpr int B = 64;
template <int R, int LEVEL, int VAR>
__global__ void __launch_bounds__(128) chaink(u32* out, u32 one, u32 four, u32 tw0) {
    extern __shared__ uint4 prof[];
    constexpr int CH = (R + 3) / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* wprof = prof + warp * (4 * CH * 32);
    const u32 sm = VAR == 2 ? pk(2) : (2u << 16 | 2u), sx = VAR == 2 ? pk(-4) : nk(4);
    for (int i = lane; i < 4 * CH * 32; i += 32) wprof[i] = make_uint4(sm, sx, sx, sm);
    __syncwarp();
    u32 H[R], E1[R], E2[R];
    const u32 base = VAR == 2 ? 0u : pk(B);
#pragma unroll
    for (int r = 0; r < R; ++r) { H[r] = base; E1[r] = base; E2[r] = base; }
    u32 hup_prev = base, h_out = base, f1_out = base, f2_out = base, best = 0;
    const u32 nz = lane != 0, bz = lane != 0 ? 0u : base;
    u32 twl = tw0 * (lane + 1);
    const char* prof_lane = reinterpret_cast<const char*>(wprof + lane);
#pragma unroll 1
    for (int st0 = 0; st0 < STEPS; st0 += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int st = st0 + u;
            u32 hup = h_out, f1 = f1_out, f2 = f2_out;
            if (LEVEL >= 1) { hup = __shfl_up_sync(~0u, h_out, 1); f1 = __shfl_up_sync(~0u, f1_out, 1); f2 = __shfl_up_sync(~0u, f2_out, 1); }
            u32 mul0 = one;
            u32 tb = (st & 3);
            if (LEVEL >= 3) {
                f1 = madd(f1, nz, bz); f2 = madd(f2, nz, bz); mul0 = nz;
                u32 lo, hi;
                asm("{\n\t.reg .u64 pp;\n\tmul.wide.u32 pp, %2, %3;\n\tmov.b64 {%0, %1}, pp;\n\t}" : "=r"(lo), "=r"(hi) : "r"(twl), "r"(four));
                twl = lo + 0x9e3779b9u * (hi == 3); tb = hi;
            }
            const uint4* pp = reinterpret_cast<const uint4*>(prof_lane + tb * (unsigned)(CH * 512));
            u32 hd = hup_prev; hup_prev = hup;
            u32 cm = best;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                uint4 sv = make_uint4(sm, sx, sx, sm);
                if (LEVEL >= 2) sv = pp[c * 32];
                const u32 s4[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = 4 * c + q;
                    if (r < R) {
                        const u32 hleft = H[r];
                        u32 h;
                        if (VAR == 0) {
                            // floor folded into the diagonal: max(hd + s, B)
                            const u32 d = __viaddmax_u16x2(r == 0 ? madd(hd, mul0, bz) : hd, s4[q], pk(B));
                            const u32 t = __vimax3_u16x2(d, E1[r], E2[r]);
                            h = __vimax3_u16x2(t, f1, f2);
                            E1[r] = __viaddmax_u16x2(h, nk(6), madd(E1[r], one, nk(2)));
                            E2[r] = __viaddmax_u16x2(h, nk(25), madd(E2[r], one, nk(1)));
                            f1 = __viaddmax_u16x2(h, nk(6), madd(f1, one, nk(2)));
                            f2 = __viaddmax_u16x2(h, nk(25), madd(f2, one, nk(1)));
                        } else if (VAR == 1) {
                            const u32 t = __vimax3_u16x2(madd(hd, r == 0 ? mul0 : one, s4[q]), E1[r], E2[r]);
                            const u32 t2 = __vimax3_u16x2(t, f1, f2);
                            h = __vmaxu2(t2, pk(B));
                            E1[r] = __viaddmax_u16x2(h, nk(6), madd(E1[r], one, nk(2)));
                            E2[r] = __viaddmax_u16x2(h, nk(25), madd(E2[r], one, nk(1)));
                            f1 = __viaddmax_u16x2(h, nk(6), madd(f1, one, nk(2)));
                            f2 = __viaddmax_u16x2(h, nk(25), madd(f2, one, nk(1)));
                        } else {
                            const u32 ninf = pk(-32768);
                            const u32 d = __viaddmax_s16x2(hd, s4[q], ninf);
                            const u32 t = __vimax3_s16x2(d, E1[r], E2[r]);
                            h = __vimax3_s16x2_relu(t, f1, f2);
                            E1[r] = __viaddmax_s16x2(h, pk(-6), __viaddmax_s16x2(E1[r], pk(-2), ninf));
                            E2[r] = __viaddmax_s16x2(h, pk(-25), __viaddmax_s16x2(E2[r], pk(-1), ninf));
                            f1 = __viaddmax_s16x2(h, pk(-6), __viaddmax_s16x2(f1, pk(-2), ninf));
                            f2 = __viaddmax_s16x2(h, pk(-25), __viaddmax_s16x2(f2, pk(-1), ninf));
                        }
                        hd = hleft; H[r] = h;
                        if (LEVEL >= 4) {
                            if (r & 1) cm = VAR == 2 ? __vimax3_s16x2(cm, h, H[r - 1]) : __vimax3_u16x2(cm, h, H[r - 1]);
                            else if (r == R - 1) cm = VAR == 2 ? __vmaxs2(cm, h) : __vmaxu2(cm, h);
                        }
                    }
                }
            }
            h_out = H[R - 1]; f1_out = f1; f2_out = f2;
            best = cm;
        }
    }
    u32 acc = best ^ h_out ^ f1_out ^ f2_out;
#pragma unroll
    for (int r = 0; r < R; ++r) acc ^= H[r] ^ E1[r] ^ E2[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int R, int LEVEL, int VAR> void run(int nsm, u32* dout) {
    const size_t smem = 4 * (4 * ((R + 3) / 4) * 32) * sizeof(uint4);
    CK(cudaFuncSetAttribute((const void*)chaink<R, LEVEL, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int w : {1, 2, 4}) {
        const int nb = nsm * w;
        chaink<R, LEVEL, VAR><<<nb, 128, smem>>>(dout, 1u, 4u, 0x1b2e4d93u); CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        chaink<R, LEVEL, VAR><<<nb, 128, smem>>>(dout, 1u, 4u, 0x1b2e4d93u);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double gcups = 2.0 * (double)nb * 128.0 * STEPS * R / (ms * 1e-3) / 1e9;
        printf("{\"R\":%d,\"level\":%d,\"variant\":%d,\"warps_per_smsp\":%d,\"ms\":%.4f,\"gcups\":%.0f,\"clk_per_step_per_warp\":%.0f}\n",
               R, LEVEL, VAR, w, ms, gcups, ms * 1e-3 * 1.965e9 / STEPS);
    }
}
int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int nsm = pr.multiProcessorCount;
    u32* dout; CK(cudaMalloc(&dout, nsm * 8 * 128 * sizeof(u32)));
    run<8, 0, 0>(nsm, dout); run<8, 4, 0>(nsm, dout);
    run<8, 0, 1>(nsm, dout); run<8, 4, 1>(nsm, dout);
    run<8, 0, 2>(nsm, dout); run<8, 4, 2>(nsm, dout);
    run<12, 0, 0>(nsm, dout); run<12, 4, 0>(nsm, dout);
    run<16, 4, 0>(nsm, dout);
    return 0;
}
