// nr_kernels.cuh -- sm_100a DP kernels for NanoRepeat's repeat-size hot path.
//
// The arithmetic replaces what the reference delegates to pyminimap2.main() at
// src/NanoRepeat/nanoRepeat_bam.py:362 (round 2) and :497 (round 3): an exact local alignment with
// minimap2's map-ont two-piece affine gap model.  Contract = oracle/nr_oracle.c (score, tstart, tend).
//
// Exact kernel ("P32"): every DP value is ONE 32-bit integer holding (score << 16) | start_column.
// Integer max on that word is the lexicographic max of (score, start), adding d << 16 adds d to the score and
// keeps the start, so the whole recurrence is max/plus on packed words -- exactly the shape of Blackwell's DPX
// instructions (VIADDMNMX = max(a + b, c), VIMNMX3 = max(a, b, c)); start tracking costs no instruction.
//
// Mapping: one warp per task.  The query is cut into stripes of 32 * R rows; inside a stripe lane l owns R
// consecutive rows whose H / E1 / E2 state lives in registers.  The warp sweeps the target column by column as
// a skewed wavefront (lane l works on column step - l); the bottom row's H, F1, F2 move to lane l + 1 by warp
// shuffle.  Substitution scores come from a per-warp query profile in shared memory (one LDS.128 per four rows,
// off the integer pipe).  Between stripes the bottom row goes through an L2-resident scratch row.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nr {

struct Task {          // 16 bytes, one per (query, target) pair
    uint32_t q_word;   // first 32-bit word of the 2-bit packed query in the sequence pool
    int32_t  q_len;
    uint32_t t_word;   // first 32-bit word of the 2-bit packed target
    int32_t  t_len;
};

struct ScoreP32 {      // scoring constants already shifted into the score field (<< 16)
    int match, mismatch_neg;   // +a << 16, -b << 16
    int qe1_neg, e1_neg;       // -(q + e) << 16, -e << 16
    int qe2_neg, e2_neg;
};

constexpr int kPadScore = -(16384 << 16);   // substitution score of rows below the query's end
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int vmax3(int a, int b, int c) { return __vimax3_s32(a, b, c); }
__device__ __forceinline__ int vaddmax(int a, int b, int c) { return __viaddmax_s32(a, b, c); }

// One DP cell of the exact kernel.  hd: H(i-1, j-1); s: substitution score; e1/e2: E(i, j); f1/f2: F(i, j).
// On return h = H(i, j), e1/e2 = E(i, j+1), f1/f2 = F(i+1, j).
__device__ __forceinline__ void cell_p32(int hd, int s, int fresh, const ScoreP32& sc,
                                         int& h, int& e1, int& e2, int& f1, int& f2) {
    int t = vaddmax(hd, s, e1);          // max(hd + s, E1)
    t = vmax3(t, e2, fresh);             // ... E2, fresh start (0, j)
    h = vmax3(t, f1, f2);                // vertical gaps last: they carry the row-to-row dependency
    e1 = vaddmax(h, sc.qe1_neg, e1 + sc.e1_neg);
    e2 = vaddmax(h, sc.qe2_neg, e2 + sc.e2_neg);
    f1 = vaddmax(h, sc.qe1_neg, f1 + sc.e1_neg);
    f2 = vaddmax(h, sc.qe2_neg, f2 + sc.e2_neg);
}

// Per-task result before the final warp reduction: (score<<16|start) and the column it was found in.
struct Best { int v; int j; };

// Lexicographic order of the contract: score desc, tend asc, tstart desc -> one 64-bit key, larger is better.
__device__ __forceinline__ unsigned long long best_key(int v, int j) {
    unsigned score = (unsigned)(v >> 16) & 0xffffu;       // scores here are >= 0
    unsigned start = (unsigned)v & 0xffffu;
    return ((unsigned long long)score << 32) | ((unsigned long long)(0xffffu - (unsigned)j) << 16) | start;
}

template <int R>
struct StripeCfg {
    static constexpr int CH = (R + 3) / 4;             // LDS.128 per column step
    static constexpr int PROF_INT4 = 4 * CH * 32;      // int4 entries per warp
};

// Build the stripe's query profile: prof[(c * CH + chunk) * 32 + lane].{x,y,z,w} = score of rows 4*chunk..+3
// of this lane against target code c.
template <int R>
__device__ __forceinline__ void build_profile(int4* prof, const uint32_t* __restrict__ qwords, int q_len,
                                              int row0, int lane, const ScoreP32& sc) {
    constexpr int CH = StripeCfg<R>::CH;
    int* p = reinterpret_cast<int*>(prof);
#pragma unroll
    for (int r = 0; r < 4 * CH; ++r) {
        int i = row0 + r;
        int code = 4;
        if (r < R && i < q_len) code = (qwords[i >> 4] >> (2 * (i & 15))) & 3;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int v = (code == 4) ? kPadScore : (code == c ? sc.match : sc.mismatch_neg);
            p[((c * CH + (r >> 2)) * 32 + lane) * 4 + (r & 3)] = v;
        }
    }
}

// One stripe of one task.  MULTI = false: single-stripe task (no boundary traffic at all).
// MULTI = true: `top` says the stripe has a predecessor (read bnd_in), `bot` says it has a successor
// (lane 31 writes bnd_out).  Boundary entry for column j: (H(last row, j), F1(next row, j), F2(next row, j)).
template <int R, bool MULTI>
__device__ __forceinline__ Best run_stripe(const int4* prof, const uint32_t* __restrict__ twords, int t_len,
                                           int lane, const ScoreP32& sc, bool top, bool bot,
                                           const int4* bnd_in, int4* bnd_out) {
    constexpr int CH = StripeCfg<R>::CH;
    int H[R], E1[R], E2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { H[r] = 0; E1[r] = sc.qe1_neg; E2[r] = sc.qe2_neg; }
    int hup_prev = 0;                       // H(row0 - 1, j - 1); column 0 is (0, start 0)
    int h_out = 0, f1_out = 0, f2_out = 0;  // bottom-row outputs of the previous step
    int best = 0, bestcap = 0xffff, bestj = 0;
    uint32_t tw = 0, tw_next = twords[0];
    int4 bcur = make_int4(0, 0, 0, 0), bnxt = make_int4(0, 0, 0, 0);
    if (MULTI && top) bnxt = __ldcg(&bnd_in[lane < t_len ? lane : t_len - 1]);
    const int nsteps = t_len + 31;
    for (int step = 0; step < nsteps; ++step) {
        const int jj = step - lane;         // 0-based target column of this lane
        int hup = __shfl_up_sync(kFull, h_out, 1);
        int f1 = __shfl_up_sync(kFull, f1_out, 1);
        int f2 = __shfl_up_sync(kFull, f2_out, 1);
        int bh = 0, bf1 = 0, bf2 = 0;
        if (MULTI && top) {                 // uniform branch
            if ((step & 31) == 0) {
                bcur = bnxt;
                int nj = step + 32 + lane;
                bnxt = __ldcg(&bnd_in[nj < t_len ? nj : t_len - 1]);
            }
            bh = __shfl_sync(kFull, bcur.x, step & 31);
            bf1 = __shfl_sync(kFull, bcur.y, step & 31);
            bf2 = __shfl_sync(kFull, bcur.z, step & 31);
        }
        if (jj >= 0 && jj < t_len) {
            const int fresh = jj + 1;       // (score 0, start = j): an alignment that starts after column j
            if (lane == 0) {
                if (MULTI && top) { hup = bh; f1 = bf1; f2 = bf2; }
                else { hup = fresh; f1 = fresh + sc.qe1_neg; f2 = fresh + sc.qe2_neg; }
            }
            if ((jj & 15) == 0) { tw = tw_next; tw_next = twords[(jj >> 4) + 1]; }
            const int tb = tw & 3;
            tw >>= 2;
            const int4* pp = prof + tb * (CH * 32) + lane;
            int hd = hup_prev;
            hup_prev = hup;
            int cm = 0;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const int4 sv = pp[c * 32];
                const int s4[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = 4 * c + u;
                    if (r < R) {
                        int h;
                        const int hleft = H[r];
                        cell_p32(hd, s4[u], fresh, sc, h, E1[r], E2[r], f1, f2);
                        hd = hleft;
                        H[r] = h;
                        if (r & 1) cm = vmax3(cm, h, H[r - 1]);
                        else if (r == R - 1) cm = max(cm, h);
                    }
                }
            }
            h_out = H[R - 1]; f1_out = f1; f2_out = f2;
            if (cm > bestcap) { best = cm; bestcap = cm | 0xffff; bestj = fresh; }
            if (MULTI && bot && lane == 31) __stcg(&bnd_out[jj], make_int4(h_out, f1_out, f2_out, 0));
        }
    }
    Best b; b.v = best; b.j = bestj;
    return b;
}

// Exact (score, tstart, tend) kernel.  Persistent: every warp pulls task indices from *counter.
// order[] lists the tasks of this launch (host sorts them by decreasing cost); out[] is indexed by task id.
// scratch: MULTI only; per warp 2 rows of scratch_stride int4.
template <int R, bool MULTI>
__global__ void __launch_bounds__(128, 4)
exact_kernel(const Task* __restrict__ tasks, const int32_t* __restrict__ order, int n_order,
             const uint32_t* __restrict__ pool, ScoreP32 sc, int* counter,
             int4* scratch, long long scratch_stride, int4* out) {
    extern __shared__ int4 smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int4* prof = smem + warp * StripeCfg<R>::PROF_INT4;
    const long long gwarp = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    int4* bnd_a = MULTI ? scratch + gwarp * 2 * scratch_stride : nullptr;
    int4* bnd_b = MULTI ? bnd_a + scratch_stride : nullptr;
    for (;;) {
        int oi = 0;
        if (lane == 0) oi = atomicAdd(counter, 1);
        oi = __shfl_sync(kFull, oi, 0);
        if (oi >= n_order) break;
        const int tid = order[oi];
        const Task tk = tasks[tid];
        const uint32_t* qwords = pool + tk.q_word;
        const uint32_t* twords = pool + tk.t_word;
        unsigned long long key = 0;
        const int rows_per_stripe = 32 * R;
        const int n_stripes = MULTI ? (tk.q_len + rows_per_stripe - 1) / rows_per_stripe : 1;
        for (int s = 0; s < n_stripes; ++s) {
            __syncwarp();
            build_profile<R>(prof, qwords, tk.q_len, s * rows_per_stripe + lane * R, lane, sc);
            __syncwarp();
            const bool top = s > 0, bot = s + 1 < n_stripes;
            Best b = run_stripe<R, MULTI>(prof, twords, tk.t_len, lane, sc, top, bot,
                                          (s & 1) ? bnd_a : bnd_b, (s & 1) ? bnd_b : bnd_a);
            unsigned long long k = (b.v >> 16) > 0 ? best_key(b.v, b.j) : 0ull;
            key = k > key ? k : key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(kFull, key, o);
            key = other > key ? other : key;
        }
        if (lane == 0) {
            int score = (int)(key >> 32);
            int4 r = make_int4(0, 0, 0, 0);
            if (score > 0) { r.x = score; r.y = (int)(key & 0xffffu); r.z = 0xffff - (int)((key >> 16) & 0xffffu); }
            out[tid] = r;
        }
    }
}

}  // namespace nr
