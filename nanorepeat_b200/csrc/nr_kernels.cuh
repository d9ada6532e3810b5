// nr_kernels.cuh -- sm_100a DP kernels for NanoRepeat's repeat-size hot path.
//
// The arithmetic replaces what the reference delegates to pyminimap2.main() at
// src/NanoRepeat/nanoRepeat_bam.py:362 (round 2) and :497 (round 3): an exact local alignment with
// minimap2's map-ont two-piece affine gap model.  Contract = oracle/nr_oracle.c (score, tstart, tend).
//
// DP word ("W32"): every DP value is ONE 32-bit integer  w = score * 65536 - span,  span = number of target
// columns the alignment has consumed so far (0 <= span < 65536), i.e. tstart = column - span.  Integer max on
// that word is the lexicographic max of (score, -span) = (score, tstart), so the recurrence is plain max/plus on
// packed words -- the shape of Blackwell's DPX instructions (VIADDMNMX = max(a + b, c), VIMNMX3 = max(a, b, c)).
// A move that consumes a target column (diagonal, horizontal gap) subtracts 1 on top of its score; a fresh start
// is the word 0, so the local-alignment floor is the free .RELU of VIMNMX3.  Per cell: 6 DPX-class instructions
// (2 for H, 1 each for E1, E2, F1, F2) + 5 adds that issue on the other integer pipe.
//
// Mapping: one warp per task.  The query is cut into stripes of 32 * R rows; inside a stripe lane l owns R
// consecutive rows whose H / E1 / E2 state lives in registers.  The warp sweeps the target column by column as
// a skewed wavefront (lane l works on column step - l); the bottom row's H, F1, F2 move to lane l + 1 by warp
// shuffle.  Substitution scores come from a per-warp query profile in shared memory (one LDS.128 per four rows,
// off the integer pipes).  A query longer than one stripe is cut into stripes that run on DIFFERENT warps at the same
// time, each a few dozen columns behind the one above it; the bottom row travels through an L2-resident scratch row
// as tagged 16-byte entries (CoopInfo, load_bnd below).  These 32-bit kernels are the second family beside the paired
// u16x2 kernels of nr_pair_kernels.cuh: they take what needs coordinates, other scorings, or more than one stripe.
//
// Ladder kernel (round 3): all rungs k of one read share their prefix L + motif^k and their suffix R, so the warp
// does ONE backward sweep (reversed read x reversed R, final column kept in shared memory) and ONE forward sweep
// over L + motif^kmax.  Whenever a lane finishes a junction column c_k = |L| + k*|motif| it joins its forward
// state with the backward state row by row (three junction states: H, E1, E2 -- a gap may span the junction), and
// a token (prefix best, junction best) travels down the lanes with the wavefront; the last lane combines it with
// the R-only optimum and emits the rung's exact (score, tstart, tend).  Identical, bit for bit, to scoring every
// rung as its own rectangle (tests/test_gpu_parity.py checks that against the oracle).
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

struct LadderTask {    // 32 bytes, one per read of a round-3 region
    uint32_t q_word;
    int32_t  q_len;
    int32_t  kmin, kmax;   // rungs of this read (kmax >= kmin >= 0)
    int32_t  out_off;      // index of rung kmin in out[]
    int32_t  region;       // index into the launch's LadderRegion table
    int32_t  read;         // index of the read in the batch (row of the selection output)
    int32_t  pad;
};

struct LadderRegion {  // per-region constants of a ladder launch (20 bytes)
    uint32_t fwd_word;     // L + motif^K, K >= every kmax of the launch
    uint32_t rev_word;     // reverse(R)
    int32_t  n_left, n_right, m;
};

struct ScoreW {        // scoring constants as W32 increments
    int sub_match, sub_mismatch;             // (a << 16) - 1, -(b << 16) - 1      (diagonal: one column consumed)
    int sub_amb;                             // -(sc_ambi << 16) - 1: a read base other than ACGT against any template base
    int h_open1, h_ext1, h_open2, h_ext2;    // horizontal gap: -((q + e) << 16) - 1, -(e << 16) - 1
    int v_open1, v_ext1, v_open2, v_ext2;    // vertical gap:   -((q + e) << 16),     -(e << 16)
    int refund1, refund2;                    // q << 16, q2 << 16: a gap that spans a junction pays its opening once
    int one, mone;                           // 1, -1 and 4, passed as kernel parameters so that ptxas cannot fold them:
    unsigned four;                           // adds written as a * one + b issue as IMAD on the FMA pipe, beside the DPX pipe
    int min_score;                           // minimap2 -s: rungs scoring below it do not exist for the selection (>= 1)
};

// Device-side view of the scoring constants.  FIXED = the reference's only scoring (tk.py:502-517: all five data types
// map to minimap2's map-ont: +2 / -4 / 4,2 / 24,1): the constants become immediates of the DPX instructions, which
// frees the register-operand slots (measured +7 %, tools/microbench/cell_chain.cu).  Otherwise they come from the
// kernel parameter.
template <bool FIXED> struct ScoreView;
template <> struct ScoreView<false> : ScoreW {
    __device__ __forceinline__ explicit ScoreView(const ScoreW& w) : ScoreW(w) {}
};
template <> struct ScoreView<true> {
    static constexpr int sub_match = (2 << 16) - 1, sub_mismatch = -(4 << 16) - 1, sub_amb = -(1 << 16) - 1;
    static constexpr int h_open1 = -(6 << 16) - 1, h_ext1 = -(2 << 16) - 1, h_open2 = -(25 << 16) - 1, h_ext2 = -(1 << 16) - 1;
    static constexpr int v_open1 = -(6 << 16), v_ext1 = -(2 << 16), v_open2 = -(25 << 16), v_ext2 = -(1 << 16);
    static constexpr int refund1 = 4 << 16, refund2 = 24 << 16;
    int one, mone;
    unsigned four;
    int min_score;
    __device__ __forceinline__ explicit ScoreView(const ScoreW& w) : one(w.one), mone(w.mone), four(w.four), min_score(w.min_score) {}
};

constexpr int kPadScore = -(16384 << 16);   // substitution score of rows below the query's end / void junction state
constexpr unsigned kFull = 0xffffffffu;
constexpr int kExact = 0, kBwd = 1, kFwd = 2;
typedef unsigned long long u64;

__device__ __forceinline__ int w_cap(int w) { return (w + 0xffff) & (int)0xffff0000; }   // score(w) << 16, any sign

// a * mul + b as one IMAD (FMA pipe).  The DPX instructions (VIMNMX3, VIADDMNMX) issue on the ALU pipe at 64 lanes
// per clock per SM and so would plain integer adds; with the adds on the other pipe the cell is 6 ALU + 5 FMA slots.
__device__ __forceinline__ int madd(int a, int mul, int b) {
    int d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(mul), "r"(b));
    return d;
}

// One DP cell.  hd: H(i-1, j-1), taken times dmul (1, or 0 for the first row of a lane whose upper neighbour is the
// matrix border); s: substitution increment; e1/e2: E(i, j); f1/f2: F(i, j).
// On return h = H(i, j), e1/e2 = E(i, j+1), f1/f2 = F(i+1, j).
template <class SC>
__device__ __forceinline__ void cell_w32(int hd, int dmul, int s, const SC& sc, int& h, int& e1, int& e2, int& f1,
                                         int& f2) {
    const int t = __vimax3_s32(madd(hd, dmul, s), e1, e2);
    h = __vimax3_s32_relu(t, f1, f2);        // vertical gaps last: they carry the row-to-row dependency
    e1 = __viaddmax_s32(h, sc.h_open1, madd(e1, sc.one, sc.h_ext1));
    e2 = __viaddmax_s32(h, sc.h_open2, madd(e2, sc.one, sc.h_ext2));
    f1 = __viaddmax_s32(h, sc.v_open1, madd(f1, sc.one, sc.v_ext1));
    f2 = __viaddmax_s32(h, sc.v_open2, madd(f2, sc.one, sc.v_ext2));
}

// Junction candidate: forward word wf (score_f, span) joined with backward word wb (score_b, ext).
// (bhi, blo) keeps the lexicographic best of (score_f + score_b, -ext, -span), both biased by a constant
// (bhi by -1, blo by +1: the cap is taken as (wf - 1) | 0xffff = cap - 1, one LOP3); junction_unbias() undoes it.
template <class SC>
__device__ __forceinline__ void jcand(int wf, int wb, const SC& sc, int& bhi, int& blo) {
    const int cf = madd(wf, sc.one, -1) | 0xffff;       // (score_f << 16) - 1
    const int khi = madd(cf, sc.one, wb);                // ((score_f + score_b) << 16) - ext - 1
    const int klo = madd(cf, sc.mone, wf);               // -span + 1
    if (khi > bhi) { bhi = khi; blo = klo; }
    else if (khi == bhi) blo = max(blo, klo);
}
constexpr int kJuncNone = -(3 << 29);                   // "no candidate yet": below every real candidate, room to subtract
__device__ __forceinline__ void junction_unbias(int& bhi, int& blo) {
    if (bhi == kJuncNone) { bhi = 0; blo = 0; } else { bhi += 1; blo -= 1; }
}

// (best word, its column) -> 64-bit key ordered like the contract: score desc, tend asc, tstart desc.
__device__ __forceinline__ u64 key_of_best(int w, int j) {
    if (w <= 0) return 0ull;
    const unsigned cap = (unsigned)w_cap(w);
    const unsigned span = cap - (unsigned)w;
    return ((u64)(cap >> 16) << 32) | ((u64)(0xffffu - (unsigned)j) << 16) | (u64)(0xffffu - span);
}

// junction best (khi, klo) -> key (score, 0xffff - ext, 0xffff - span)
__device__ __forceinline__ u64 key_of_junction(int khi, int klo) {
    if (khi <= 0) return 0ull;
    const unsigned cap = (unsigned)w_cap(khi);
    const unsigned ext = cap - (unsigned)khi;
    return ((u64)(cap >> 16) << 32) | ((u64)(0xffffu - ext) << 16) | (u64)(0xffffu - (unsigned)(-klo));
}

// Stripes of one long task run on different warps (any SM) at the same time, each a few dozen columns behind the one
// above it.  No flags and no fences on the column path: a boundary entry (H, F1, F2 of one column) is one 16-byte store
// and one 16-byte load of TWO 64-bit elements, each carrying 48 bits of the payload and its own 16-bit tag
// (tag = (run epoch, stripe)).  PTX treats a vector access as independent accesses of its elements, and a 64-bit
// element access is single-copy atomic: an element whose tag matches holds the payload bits that were stored with that
// tag, so an entry is taken only when BOTH elements carry the expected tag; an entry that is not there yet (or is left
// over from the stripe that used the row before, or from an earlier run) is recognised and simply read again.
//   lo = H | (F1 & 0xffff) << 32 | tag << 48          hi = (F1 >> 16) | F2 << 16 | tag << 48
// What keeps stale tags apart (host side, nr_api.cu): the scratch comes from a pool of its own that is zeroed when
// allocated; every launch draws a fresh epoch from 1..1023; a buffer is zeroed again before its first launch in a new
// "era" (every 1023 launches), so a tag left in it can never equal a tag of the running launch.
__device__ __forceinline__ ulonglong2 load_bnd(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void store_bnd(ulonglong2* p, int h, int f1, int f2, int tag) {
    const unsigned long long t = (unsigned long long)(unsigned)tag << 48;
    const unsigned long long lo = (unsigned long long)(unsigned)h | ((unsigned long long)((unsigned)f1 & 0xffffu) << 32) | t;
    const unsigned long long hi = (unsigned long long)((unsigned)f1 >> 16) | ((unsigned long long)(unsigned)f2 << 16) | t;
    asm volatile("st.global.cg.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ bool bnd_valid(const ulonglong2& v, int tag) {
    return (int)(v.x >> 48) == tag && (int)(v.y >> 48) == tag;
}
__device__ __forceinline__ int4 unpack_bnd(const ulonglong2& v) {
    return make_int4((int)(unsigned)v.x, (int)((unsigned)(v.x >> 32) & 0xffffu) | (int)((unsigned)v.y << 16), (int)(unsigned)(v.y >> 16), 0);
}
__device__ __forceinline__ ulonglong2 load_tok(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int peek_cols(const int* flag) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    return v;
}
// Safety valve of every wait between stripes: what a stripe waits for is running or done (entries are dealt in
// dependency order to resident blocks), so a wait ends within the producer's run time.  Should that ever not hold, the
// warp gives up after a few seconds' worth of retries, raises this flag and runs on with what it has; the host turns
// the flag into NR_ERR_CUDA instead of a hung GPU.  The flag is a word of the batch's own counters (RestArgs::spin).
constexpr int kSpinLimit = 1 << 23;
__device__ __forceinline__ void wait_cols(const int* flag, int need, int lane, int* spin) {
    if (lane == 0)
        for (int tries = 0; peek_cols(flag) < need; ++tries) {
            if (tries > kSpinLimit) { *spin = 1; break; }
            __nanosleep(100);
        }
    __syncwarp();
}

__device__ __forceinline__ u64 shfl_up64(u64 v) {
    unsigned lo = __shfl_up_sync(kFull, (unsigned)v, 1);
    unsigned hi = __shfl_up_sync(kFull, (unsigned)(v >> 32), 1);
    return ((u64)hi << 32) | lo;
}

__device__ __forceinline__ u64 warp_max64(u64 key) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u64 other = __shfl_xor_sync(kFull, key, o);
        key = other > key ? other : key;
    }
    return key;
}

// Rung k's record from its three alignment classes: P = ends at or before the junction column c, J = crosses it,
// R-only = lies inside the right flank (same for every rung of the read).
__device__ __forceinline__ int4 finalize_rung(u64 P, u64 J, int c, int r_score, int r_end, int r_start) {
    int bs = 0, be = 0, bst = 0;
    if (P) {
        const int s = (int)(P >> 32), j = 0xffff - (int)((P >> 16) & 0xffffu), span = 0xffff - (int)(P & 0xffffu);
        bs = s; be = j; bst = j - span;
    }
    if (J) {
        const int s = (int)(J >> 32), e = c + 0xffff - (int)((J >> 16) & 0xffffu), st = c - (0xffff - (int)(J & 0xffffu));
        if (s > bs || (s == bs && (e < be || (e == be && st > bst)))) { bs = s; be = e; bst = st; }
    }
    if (r_score > 0) {
        const int e = c + r_end, st = c + r_start;
        if (r_score > bs || (r_score == bs && (e < be || (e == be && st > bst)))) { bs = r_score; be = e; bst = st; }
    }
    return make_int4(bs, bst, be, 0);
}

// Flag ladder: rung record from the prefix-class best word P, the junction-class best key J and the R-only candidate key
// (keys: (score << 16) - 2 * ext - mark).  Output: (score, starts_in_left && ends_in_right, ends_in_right, 0).
// An alignment that ends at or before the junction column has the smaller tend, so the prefix class wins ties.
__device__ __forceinline__ int4 finalize_flag_rung(int P, int J, int rcand) {
    const int np = max(J, rcand);
    const int s_np = np > 0 ? w_cap(np) : 0;
    const int s_p = w_cap(P);
    const int in_right = s_np > s_p;
    return make_int4(max(s_p, s_np) >> 16, in_right & np & 1, in_right, 0);
}

template <int R>
struct StripeCfg {
    static constexpr int CH = (R + 3) / 4;             // LDS.128 per column step
    static constexpr int PROF_INT4 = 4 * CH * 32;      // int4 entries per warp: query profile
    static constexpr int BVEC_INT4 = R * 32;           // int4 entries per warp: backward junction vectors
};

// A read in the sequence pool: ceil(len / 16) words of 2-bit codes, then one word that is 0 for a read of ACGT only;
// otherwise it is the distance (in words, from the read's first word) to a bit plane: bit i & 31 of word i >> 5 says
// that base i is not ACGT -- minimap2's code 4, scored -sc_ambi against every template base (oracle/nr_oracle.c).
// The plane costs nothing where there is no such base.
__device__ __forceinline__ uint32_t read_ambiguity_plane(const uint32_t* __restrict__ qwords, int q_len) {
    return qwords[(q_len + 15) >> 4];
}
__device__ __forceinline__ bool read_base_ambiguous(const uint32_t* __restrict__ qwords, uint32_t plane, int qi) {
    return (qwords[plane + (qi >> 5)] >> (qi & 31)) & 1u;
}

// Build the stripe's query profile: prof[(c * CH + chunk) * 32 + lane].{x,y,z,w} = substitution increment of rows
// 4*chunk..+3 of this lane against target code c.  reverse: the stripe's rows index the reversed query.
template <int R, class SC>
__device__ __forceinline__ void build_profile(int4* prof, const uint32_t* __restrict__ qwords, int q_len,
                                              int row0, int lane, const SC& sc, bool reverse) {
    constexpr int CH = StripeCfg<R>::CH;
    int* p = reinterpret_cast<int*>(prof);
    const uint32_t amb = q_len > 0 ? read_ambiguity_plane(qwords, q_len) : 0u;      // warp-uniform
#pragma unroll
    for (int r = 0; r < 4 * CH; ++r) {
        const int i = row0 + r;
        int code = 4;
        if (r < R && i < q_len) {
            const int qi = reverse ? q_len - 1 - i : i;
            code = (qwords[qi >> 4] >> (30 - 2 * (qi & 15))) & 3;
            if (amb && read_base_ambiguous(qwords, amb, qi)) code = 5;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int v = (code == 4) ? kPadScore : code == 5 ? sc.sub_amb : (code == c ? sc.sub_match : sc.sub_mismatch);
            p[((c * CH + (r >> 2)) * 32 + lane) * 4 + (r & 3)] = v;
        }
    }
}

// Position (int4 index inside the forward-layout backward-vector array) of forward cell row idx0.
template <int R>
__device__ __forceinline__ int bvec_pos(int idx0) {
    const int stripe = idx0 / (32 * R), in = idx0 - stripe * (32 * R);
    return stripe * (32 * R) + (in % R) * 32 + in / R;
}

// Per-sweep view of the scoring words.  DEC = what a move that consumes a target column subtracts from the word's
// low half: 1 for span words (tstart = column - span), 2 for the backward sweep of the flag ladder (2 * ext, leaving
// bit 0 for the forward mark), 0 for the forward sweep of the flag ladder (the low half holds only the mark).
template <class SC, int DEC>
struct ModeScore {
    int sub_match, sub_mismatch, sub_amb, h_open1, h_ext1, h_open2, h_ext2, v_open1, v_ext1, v_open2, v_ext2, one;
    __device__ __forceinline__ explicit ModeScore(const SC& sc)
        : sub_match(sc.sub_match + 1 - DEC), sub_mismatch(sc.sub_mismatch + 1 - DEC), sub_amb(sc.sub_amb + 1 - DEC),
          h_open1(sc.h_open1 + 1 - DEC),
          h_ext1(sc.h_ext1 + 1 - DEC), h_open2(sc.h_open2 + 1 - DEC), h_ext2(sc.h_ext2 + 1 - DEC), v_open1(sc.v_open1),
          v_ext1(sc.v_ext1), v_open2(sc.v_open2), v_ext2(sc.v_ext2), one(sc.one) {}
};

constexpr int kBwdF = 3, kFwdF = 4;     // flag ladder (see ladder_task)

// One stripe of one sweep.
//   kExact: plain task, running best per the contract.
//   kBwd:   reversed read x reversed right flank; best per the R-only ordering; the final column's junction
//           state (H, E1 + refund1, E2 + refund2) is written to bdst in the forward layout.
//   kFwd:   read x L + motif^kmax with junction tokens (see the file header).
//   kBwdF / kFwdF: the same two sweeps on flag words (ladder_task, FLAG).
// MULTI: the task has several stripes; `top` = this stripe has a predecessor (read bnd_in), `bot` = it has a
// successor (lane 31 writes bnd_out).  Boundary entry for column j: (H(last row, j), F1(next row, j), F2(next row, j)).
//
// Step kinds.  Lane l works on column st - l, so for 31 <= st < t_len every lane is inside the matrix: those steps
// run the FAST body (no guards, no junction logic), in blocks of 16 with one uniform refill of the target stream per
// block.  The first 32 steps, the tail, (kFwd) the junction zone and (kFwdF) the 32 steps in which the lanes pass
// the end of the left flank run the guarded body.
// Target stream: every lane reads the target through its own 32-bit window, MSB first, pre-shifted by the lane's
// skew, so that all lanes refill at the same step; taking the next base is one IMAD.WIDE (window * 4: the base
// falls out of the top), and the profile row address one more IMAD -- both off the DPX pipe.
template <int R, int MODE, bool MULTI>
struct Sweep {
    static constexpr bool kIsFwd = MODE == kFwd || MODE == kFwdF;
    static constexpr bool kIsBwd = MODE == kBwd || MODE == kBwdF;
    static constexpr int DEC = MODE == kFwdF ? 0 : MODE == kBwdF ? 2 : 1;
    // inputs
    const int4* prof;
    const uint32_t* twords;
    int t_len, lane;
    bool top, bot;
    const ulonglong2* bnd_in;
    ulonglong2* bnd_out;
    int tag_in, tag_out;           // MULTI: what marks an entry written by the stripe above / by this stripe in this run
    int* spin;                     // MULTI: the batch's give-up flag (RestArgs::spin)
    // backward sweeps
    int4* bdst;
    int q_len, brow0;
    // forward sweeps
    const int4* bsm;
    // MULTI forward sweeps: the junction vectors (and the R-only optimum) are fetched when the sweep reaches its first
    // junction column, so that the stripes of the forward sweep run beside those of the backward sweep
    int4* bsm_w;
    const int4* bglob_rows;        // this stripe's rows of the backward vectors (forward layout), null: no backward sweep
    const int* bdone;              // backward stripes that are done
    int bdone_need, n_right;
    bool late_pending;
    const ulonglong2* tok_in;
    ulonglong2* tok_out;
    int4* out;
    int jnext, m, kcnt, zone_start;
    int r_score, r_end, r_start;   // kFwd: the R-only optimum
    int rcand, mark_col;           // kFwdF: the R-only candidate key; last column of the left flank (-1: none)
    int k0, min_score;             // kFwdF: first rung of the read; selection threshold
    int sel_top, sel_n;            // kFwdF, lane 31: round-3 selection over the rungs finalised so far
    long long sel_sum;             //   (nanoRepeat_bam.py:423-431: top score, tied rungs that span both flanks)
    // state
    int H[R], E1[R], E2[R];
    int hup_prev, h_out, f1_out, f2_out;
    int best, bestor, bestst;      // kExact / kFwd: best word, best | 0xffff (its score class), step it was found at
                                   // backward: bestor = best key; kFwdF: best = running maximum word
    int nz;                        // 0 on lane 0, else 1: multiplier that blanks what lane 0 "receives" from SHFL.UP
    uint32_t twl, w0, w1;          // target window (MSB first) and the two words the next window is cut from
    int wi, wmax, wsh;
    const char* prof_lane;
    int4 bcur;                     // MULTI: boundary entries of the current 32 columns (lane l: column block + l), unpacked
    ulonglong2 bnxt;               //        and of the next 32, as loaded (validated when they are taken over)
    u64 tokP, tokJ;                // kFwd: 64-bit keys; kFwdF: 32-bit words in the low halves

    __device__ __forceinline__ uint32_t tword(int i) const { return __ldg(&twords[min(max(i, 0), wmax)]); }

    template <class MS>
    __device__ __forceinline__ void init(const MS& sc) {
#pragma unroll
        for (int r = 0; r < R; ++r) { H[r] = 0; E1[r] = sc.h_open1; E2[r] = sc.h_open2; }
        hup_prev = 0;                       // H(row0 - 1, j - 1); column 0 is the word 0
        h_out = 0; f1_out = 0; f2_out = 0;  // bottom-row outputs of the previous step
        best = 0; bestor = kIsBwd ? 0 : 0xffff; bestst = 0;
        nz = lane != 0;
        wmax = (t_len + 15) >> 4;           // the zero slack word behind the sequence
        wi = (-lane) >> 4;                  // word of column -lane (floor)
        wsh = 2 * ((-lane) & 15);
        w0 = tword(wi); w1 = tword(wi + 1);
        twl = 0;
        prof_lane = reinterpret_cast<const char*>(prof + lane);
        bcur = make_int4(0, 0, 0, 0); bnxt = make_ulonglong2(0ull, 0ull);
        if (MULTI && top) bnxt = load_bnd(&bnd_in[lane < t_len ? lane : t_len - 1]);   // checked when it is taken over
        tokP = 0; tokJ = 0;
    }

    // every 16 steps, all lanes together: next 16 columns of this lane's skewed view of the target
    __device__ __forceinline__ void refill() {
        twl = __funnelshift_l(w1, w0, wsh);
        w0 = w1;
        ++wi;
        w1 = tword(wi + 1);
    }

    __device__ __forceinline__ unsigned next_base(unsigned four) {
        unsigned lo, hi;
        asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(lo), "=r"(hi) : "r"(twl), "r"(four));
        twl = lo;
        return hi;
    }

    __device__ __forceinline__ int best_col() const { return bestst - lane + 1; }   // 1-based column of `best`

    // JUNC: also evaluate the junction candidates of this column.  kFwd: lexicographic (jhi, jlo); kFwdF: one key jhi.
    template <bool JUNC, class MS, class SC>
    __device__ __forceinline__ int cells(const int4* pp, int hd, int mul0, int cm, int& f1, int& f2, const MS& ms,
                                         const SC& sc, int& jhi, int& jlo) {
        constexpr int CH = StripeCfg<R>::CH;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int4 sv = pp[c * 32];
            const int s4[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = 4 * c + u;
                if (r < R) {
                    const int hleft = H[r];
                    const int e1pre = E1[r], e2pre = E2[r];
                    int h;
                    cell_w32(hd, r == 0 ? mul0 : ms.one, s4[u], ms, h, E1[r], E2[r], f1, f2);
                    if (JUNC) {
                        const int4 b = bsm[r * 32 + lane];
                        if (MODE == kFwdF) {       // the sums on the FMA pipe, two three-input maxima on the DPX pipe
                            jhi = __vimax3_s32(jhi, madd(h, ms.one, b.x), madd(e1pre, ms.one, b.y));
                            jhi = max(jhi, madd(e2pre, ms.one, b.z));
                        } else {
                            jcand(h, b.x, sc, jhi, jlo);
                            jcand(e1pre, b.y, sc, jhi, jlo);
                            jcand(e2pre, b.z, sc, jhi, jlo);
                        }
                    }
                    hd = hleft;
                    H[r] = h;
                    if (r & 1) cm = __vimax3_s32(cm, h, H[r - 1]);
                    else if (r == R - 1) cm = max(cm, h);
                }
            }
        }
        return cm;
    }

    // kFwdF: the reference's per-read selection, rung by rung in increasing k (only the lane that finalises rungs)
    __device__ __forceinline__ void select_rung(const int4& rec, int k) {
        if (rec.x >= min_score) {
            if (rec.x > sel_top) { sel_top = rec.x; sel_n = 0; sel_sum = 0; }
            if (rec.x == sel_top && rec.y) { ++sel_n; sel_sum += k; }
        }
    }

    // kFwdF, once per lane: everything alive after the last column of the left flank started inside the flank
    __device__ __forceinline__ void mark_started_in_left() {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            H[r] = __viaddmax_s32(H[r], -1, 0);
            E1[r] -= E1[r] > 0;
            E2[r] -= E2[r] > 0;
        }
        hup_prev = __viaddmax_s32(hup_prev, -1, 0);
    }

    // FAST: every lane is inside the matrix and (forward sweeps) no lane is at a junction or mark column.
    // Otherwise the guarded body; forward sweeps from zone_start on: some lane may be at a junction column, tokens
    // are moving.  A step in which any lane is at a junction evaluates the junction candidates on every lane (one
    // code path for the warp; the lanes that are not at a junction drop theirs).
    // ZONE (guarded body of forward sweeps): the caller knows on which side of zone_start the step lies, so each of the
    // two guarded loops holds one copy of the cell code (with or without the junction candidates), not both.
    template <bool FAST, bool ZONE, class MS, class SC>
    __device__ __forceinline__ void step(int st, const MS& ms, const SC& sc) {
        constexpr bool zone = !FAST && kIsFwd && ZONE;
        constexpr int CH = StripeCfg<R>::CH;
        int hup = __shfl_up_sync(kFull, h_out, 1);
        int f1 = __shfl_up_sync(kFull, f1_out, 1);
        int f2 = __shfl_up_sync(kFull, f2_out, 1);
        u64 tP = 0, tJ = 0;
        if (zone) {
            if (MODE == kFwdF) {
                tP = (unsigned)__shfl_up_sync(kFull, (int)tokP, 1);
                tJ = (unsigned)__shfl_up_sync(kFull, (int)tokJ, 1);
            } else {
                tP = shfl_up64(tokP); tJ = shfl_up64(tokJ);
            }
        }
        int mul0;
        if (MULTI && top) {                 // uniform branch
            if ((st & 31) == 0) {
                ulonglong2 raw = bnxt;
                const int cj = st + lane;       // the column this lane's entry stands for
                // a stripe right behind its producer finds the entries it asked for 32 steps ago not written yet: it reads
                // them again (one L2 round trip), falls back a little, and from then on its prefetches arrive valid
                for (int tries = 0; !__all_sync(kFull, bnd_valid(raw, tag_in) || cj >= t_len); ++tries) {
                    if (tries > kSpinLimit) { *spin = 1; break; }
                    if (tries > 3) __nanosleep(32);
                    raw = load_bnd(&bnd_in[cj < t_len ? cj : t_len - 1]);
                }
                bcur = unpack_bnd(raw);
                const int nj = st + 32 + lane;
                bnxt = load_bnd(&bnd_in[nj < t_len ? nj : t_len - 1]);      // in flight for the next 32 steps
            }
            const int bh = __shfl_sync(kFull, bcur.x, st & 31);
            const int bf1 = __shfl_sync(kFull, bcur.y, st & 31);
            const int bf2 = __shfl_sync(kFull, bcur.z, st & 31);
            if (lane == 0) { hup = bh; f1 = bf1; f2 = bf2; }
            mul0 = ms.one;
        } else {
            // matrix border above lane 0: H = 0 and any F <= 0 (the floor of H makes every non-positive F equivalent);
            // the diagonal is blanked where it is used
            f1 = madd(f1, nz, 0);
            f2 = madd(f2, nz, 0);
            mul0 = nz;
        }
        const unsigned tb = next_base(sc.four);  // unconditional: the window advances one column per step
        const int jj = st - lane;           // 0-based target column of this lane
        if (FAST || (jj >= 0 && jj < t_len)) {
            const int4* pp = reinterpret_cast<const int4*>(prof_lane + tb * (unsigned)(CH * 512));
            const int hd = hup_prev;
            hup_prev = hup;
            const bool last_col = !FAST && kIsBwd && jj == t_len - 1;
            if (kIsBwd && !FAST) {
                if (last_col) {             // E(i', n_right): the state entering the last column
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int idx0 = q_len - 2 - (brow0 + r);
                        if (idx0 >= 0) {
                            int* d = reinterpret_cast<int*>(&bdst[bvec_pos<R>(idx0)]);
                            d[1] = E1[r] + sc.refund1;
                            d[2] = E2[r] + sc.refund2;
                        }
                    }
                }
            }
            int cm, jhi = kJuncNone, jlo = 0;
            const bool junc = zone && (jj + 1 == jnext);
            const int cm0 = MODE == kFwdF ? best : 0;      // flag ladder: only the running maximum matters
            // in the zone the junction candidates are evaluated on every step (with 32 lanes a motif apart or less some
            // lane is at a junction column on almost every one)
            cm = cells<zone>(pp, hd, mul0, cm0, f1, f2, ms, sc, jhi, jlo);
            h_out = H[R - 1]; f1_out = f1; f2_out = f2;
            if (kIsBwd) {
                // R-only ordering in forward coordinates: score desc, end asc (= reversed start desc), start desc
                // (= reversed end asc: keep the earlier column on a full tie).  cm + DEC * column = (score << 16) +
                // DEC * reversed start of the column's best cell; within a lane column = st + const, so compare
                // cm + DEC * st.
                const int key = madd(cm, ms.one, DEC * st);
                if (key > bestor) { bestor = key; bestst = st; }
                if (last_col) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int idx0 = q_len - 2 - (brow0 + r);
                        if (idx0 >= 0) reinterpret_cast<int*>(&bdst[bvec_pos<R>(idx0)])[0] = H[r];
                    }
                }
            } else if (MODE == kFwdF) {
                best = cm;
                if (!FAST && jj == mark_col) mark_started_in_left();
            } else {
                // a strictly higher score (span is below 65536, so cm > best | 0xffff compares the score fields)
                if (cm > bestor) { best = cm; bestor = cm | 0xffff; bestst = st; }
            }
            if (junc) {
                if (MODE == kFwdF) {
                    if (lane == 0) {
                        if (MULTI && top) {     // both 64-bit elements of the token carry the tag in their upper half
                            ulonglong2 t = load_tok(&tok_in[kcnt]);
                            for (int tries = 0; (int)(t.x >> 32) != tag_in || (int)(t.y >> 32) != tag_in; ++tries) {
                                if (tries > kSpinLimit) { *spin = 1; break; }
                                if (tries > 3) __nanosleep(32);
                                t = load_tok(&tok_in[kcnt]);
                            }
                            tP = (unsigned)t.x; tJ = (unsigned)t.y;
                        } else { tP = 0; tJ = (unsigned)kJuncNone; }
                    }
                    // rung 0's junction column is the left flank's last column: every forward part that scores has
                    // started inside the flank but is marked only after this step (the candidates with an empty
                    // forward part are the R-only class again, so the blanket -1 cannot lose anything)
                    const int myP = max((int)tP, best), myJ = max((int)tJ, jj == mark_col ? jhi - 1 : jhi);
                    tokP = (unsigned)myP; tokJ = (unsigned)myJ;
                    if (lane == 31) {
                        if (MULTI && bot) __stcg(&tok_out[kcnt], make_ulonglong2(tokP | ((u64)(unsigned)tag_out << 32), tokJ | ((u64)(unsigned)tag_out << 32)));
                        else {
                            const int4 rec = finalize_flag_rung(myP, myJ, rcand);
                            out[kcnt] = rec;
                            select_rung(rec, k0 + kcnt);
                        }
                    }
                } else {
                    junction_unbias(jhi, jlo);
                    if (lane == 0) {
                        // (its store was fenced before the boundary entry of this column, which this stripe has seen; the
                        // fence on this side orders the token load behind that observation)
                        if (MULTI && top) { __threadfence(); const ulonglong2 t = load_tok(&tok_in[kcnt]); tP = t.x; tJ = t.y; }
                        else { tP = 0; tJ = 0; }
                    }
                    const u64 myP = key_of_best(best, best_col()), myJ = key_of_junction(jhi, jlo);
                    tokP = myP > tP ? myP : tP;
                    tokJ = myJ > tJ ? myJ : tJ;
                    if (lane == 31) {
                        if (MULTI && bot) { __stcg(&tok_out[kcnt], make_ulonglong2(tokP, tokJ)); __threadfence(); }
                        else out[kcnt] = finalize_rung(tokP, tokJ, jj + 1, r_score, r_end, r_start);
                    }
                }
                jnext += m;
                ++kcnt;
            }
            if (MULTI && bot && lane == 31) store_bnd(&bnd_out[jj], h_out, f1_out, f2_out, tag_out);
        }
    }

    __device__ __forceinline__ void late_load() {
        u64 rkey = 0;
        if (bglob_rows) {
            wait_cols(bdone, bdone_need, lane, spin);
            rkey = __ldcg(reinterpret_cast<const u64*>(bdone + 2));
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            bsm_w[r * 32 + lane] = (bglob_rows && brow0 + r < q_len - 1) ? __ldcg(&bglob_rows[r * 32 + lane])
                                   : (brow0 + r < q_len ? make_int4(0, kPadScore, kPadScore, 0) : make_int4(kPadScore, kPadScore, kPadScore, 0));
        r_score = 0; r_end = 0; r_start = 0; rcand = kJuncNone;
        if (rkey) {
            if (MODE == kFwdF) {
                rcand = (int)(unsigned)rkey - 2 * n_right;
            } else {
                r_score = (int)(rkey >> 32);
                r_end = n_right - (int)((rkey >> 16) & 0xffffu);
                r_start = n_right - (0xffff - (int)(rkey & 0xffffu));
            }
        }
        __syncwarp();
        late_pending = false;
    }

    template <class MS, class SC>
    __device__ __forceinline__ void slow_until(int& st, int end, const MS& ms, const SC& sc) {
        const int pre = kIsFwd ? min(end, zone_start) : end;      // guarded steps before any lane can be at a junction
#pragma unroll 1
        for (; st < pre; ++st) {
            if ((st & 15) == 0) refill();
            step<false, false>(st, ms, sc);
        }
        if (kIsFwd) {
#pragma unroll 1
            for (; st < end; ++st) {
                if (MULTI && late_pending) late_load();
                if ((st & 15) == 0) refill();
                step<false, true>(st, ms, sc);
            }
        }
    }

    // Steps per trip of the fast loop.  A single-stripe task runs four (loop overhead off the DPX pipe); the stripes of a
    // long read run one: their warps are latency-bound (a stripe waits on the one above it) and sit on the same SM in
    // different sweeps and phases, so what limits them is instruction fetch (`no_instruction` was the top stall of
    // config 5's launches, 2.5 cycles per issue) and a loop body a quarter of the size is worth more than the saved branch.
    static constexpr int kUnroll = (MULTI && MODE != kExact) ? 1 : 4;
    template <class MS, class SC>
    __device__ __forceinline__ void fast_until(int& st, int end, const MS& ms, const SC& sc) {
        for (; st + 16 <= end; st += 16) {       // st is a multiple of 16 here
            refill();
#pragma unroll 1
            for (int b = 0; b < 16; b += kUnroll) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) step<true, false>(st + b + u, ms, sc);
            }
        }
    }

    // zone_start_: forward sweeps, first step at which a lane can be at a junction column
    template <class SC>
    __device__ __forceinline__ void run(const SC& sc, int zone_start_) {
        const ModeScore<SC, DEC> ms(sc);
        init(ms);
        zone_start = zone_start_;
        const int nsteps = t_len + 31;
        int fast_end = ((t_len - 1) >> 4) << 4;       // steps below it: every lane's column is < t_len - 1
        if (kIsFwd) fast_end = min(fast_end, (zone_start >> 4) << 4);
        int st = 0;
        slow_until(st, min(32, nsteps), ms, sc);
        if (MODE == kFwdF && mark_col >= 0) {
            // the lanes pass the end of the left flank at steps mark_col .. mark_col + 31: guarded steps
            fast_until(st, min(fast_end, (mark_col >> 4) << 4), ms, sc);
            slow_until(st, min(nsteps, ((mark_col + 32 + 15) >> 4) << 4), ms, sc);
        }
        fast_until(st, fast_end, ms, sc);
        slow_until(st, nsteps, ms, sc);
    }

    // backward sweeps: (score << 16) + DEC * reversed start, and the 1-based reversed end column
    __device__ __forceinline__ void bwd_best(int& key, int& col) const {
        key = bestor - DEC * (lane - 1);              // cm + DEC * column
        col = best_col();
    }
};

// The stripe heights the multi-stripe code is built for (the host uses one per batch, plan_batch): every height is a
// few hundred KB of unrolled code, and the ones in between would never run.
__host__ __device__ constexpr bool is_coop_height(int R) { return R == 4 || R == 6 || R == 8 || R == 12 || R == 16; }
__host__ __device__ constexpr int coop_height_at_least(int R) { return R <= 4 ? 4 : R <= 6 ? 6 : R <= 8 ? 8 : R <= 12 ? 12 : 16; }

constexpr int kMinR = 4;
constexpr int kMaxRExact = 16;    // single-stripe tasks up to 512 rows
constexpr int kMaxRLadder = 12;   // 384 rows: profile + junction vectors stay within 12 KB of shared memory per warp
// ladder_kernel launched on its own (the paired kernel's redo list, mode 2 / 1 batches) also takes single-stripe reads of
// up to 512 rows -- the reads the paired kernels now pair -- at 16 KB per warp, i.e. with fewer warps per block
constexpr int kMaxRLadderOwn = 16;

// Stripe height for a query: single stripe when it fits max_r rows per lane, else the fewest equal stripes.
__host__ __device__ __forceinline__ void stripe_shape(int q_len, int max_r, int& R, int& n_stripes) {
    if (q_len <= 32 * max_r) {
        n_stripes = 1;
        R = (q_len + 31) / 32;
        if (R < kMinR) R = kMinR;
    } else {
        n_stripes = (q_len + 32 * max_r - 1) / (32 * max_r);
        R = (q_len + 32 * n_stripes - 1) / (32 * n_stripes);
    }
}

// Scratch of one multi-stripe task.  flags (ints, zeroed before every run) at flag_off (even):
//   [0, 2S) unused        [2S] stripes of the first sweep (exact: the only sweep; ladder: backward) that are done
//   [2S + 2, 2S + 4) 64-bit key: exact: the task's best; ladder: the R-only optimum of the backward sweep
struct CoopInfo {
    long long data_off;     // int4 index into scratch: [boundary rows a, b of the first sweep | (ladder) boundary rows a, b of the
                            // forward sweep | backward vectors | token row a | token row b]: the ladder's two sweeps overlap
    int flag_off;
    int bnd_stride, b_stride, tok_stride;   // int4, int4, ulonglong2
    int n_stripes;
    int rows;               // rows per lane of every stripe (the host picks from a short list, see plan_batch)
};
constexpr int kCoopFlagInts(int S) { return 2 * S + 4; }
// 16-bit tag of the entries stripe s writes in the launch with this epoch (1..1023, drawn by the host; never 0)
__device__ __forceinline__ int stripe_tag(int epoch, int s) { return ((epoch & 0x3ff) << 6) | (s & 63); }

// what the 32-bit kernels need beside the tasks (also handed to the fused kernels of nr_pair_kernels.cuh)
struct RestArgs {
    const int32_t* order;       // entries, (task << 7) | code
    int n_order;
    int4* scratch;
    const CoopInfo* coop;
    const int32_t* coop_idx;    // exact tasks only (ladder tasks carry theirs)
    int* flags;
    int* spin;                  // set to 1 by a warp that gave up waiting for another stripe (see wait_cols)
    int epoch;
    int32_t* redo_long;         // paired round 3: long reads whose selection the 16-bit words cannot decide (task indices;
    int* redo_long_count;       // the host rescoring them on 32-bit words, nr_api.cu)
};

template <int R, class SC>
__device__ __noinline__ u64 exact_task(const Task& tk, const uint32_t* __restrict__ pool, const SC& sc, int4* prof, int lane) {
    __syncwarp();
    build_profile<R>(prof, pool + tk.q_word, tk.q_len, lane * R, lane, ModeScore<SC, 1>(sc), false);
    __syncwarp();
    Sweep<R, kExact, false> sw;
    sw.prof = prof; sw.twords = pool + tk.t_word; sw.t_len = tk.t_len; sw.lane = lane;
    sw.top = false; sw.bot = false;
    sw.run(sc, 0);
    return key_of_best(sw.best, sw.best_col());
}

template <int R, class SC>
__device__ __forceinline__ u64 exact_dispatch(int r, const Task& tk, const uint32_t* __restrict__ pool, const SC& sc,
                                              int4* prof, int lane) {
    if (r == R) return exact_task<R>(tk, pool, sc, prof, lane);
    if constexpr (R < kMaxRExact) return exact_dispatch<R + 1>(r, tk, pool, sc, prof, lane);
    return 0ull;
}

// One stripe of a multi-stripe task; the stripe that finishes last writes the record.
template <int R, class SC>
__device__ __noinline__ void exact_stripe(const Task& tk, int tid, int s, const CoopInfo& ci, const RestArgs& ra,
                                             const uint32_t* __restrict__ pool, const SC& sc, int4* prof, int lane, int4* out) {
    const int S = ci.n_stripes;
    const int epoch = ra.epoch;
    int* F = ra.flags + ci.flag_off;
    ulonglong2* bnd_a = reinterpret_cast<ulonglong2*>(ra.scratch + ci.data_off);
    ulonglong2* bnd_b = bnd_a + ci.bnd_stride;
    __syncwarp();
    build_profile<R>(prof, pool + tk.q_word, tk.q_len, s * 32 * R + lane * R, lane, ModeScore<SC, 1>(sc), false);
    __syncwarp();
    Sweep<R, kExact, true> sw;
    sw.prof = prof; sw.twords = pool + tk.t_word; sw.t_len = tk.t_len; sw.lane = lane;
    sw.top = s > 0; sw.bot = s + 1 < S;
    sw.bnd_in = (s & 1) ? bnd_a : bnd_b; sw.bnd_out = (s & 1) ? bnd_b : bnd_a;
    sw.tag_in = stripe_tag(epoch, s - 1); sw.tag_out = stripe_tag(epoch, s); sw.spin = ra.spin;
    sw.run(sc, 0);
    const u64 key = warp_max64(key_of_best(sw.best, sw.best_col()));
    if (lane == 0) {
        u64* K = reinterpret_cast<u64*>(F + 2 * S + 2);
        atomicMax(K, key);
        __threadfence();
        if (atomicAdd(F + 2 * S, 1) == S - 1) {
            __threadfence();
            out[tid] = finalize_rung(atomicMax(K, 0ull), 0ull, 0, 0, 0, 0);
        }
    }
}

template <int R, class SC>
__device__ __forceinline__ void exact_stripe_dispatch(int r, const Task& tk, int tid, int s, const CoopInfo& ci, const RestArgs& ra,
                                                      const uint32_t* __restrict__ pool, const SC& sc, int4* prof,
                                                      int lane, int4* out) {
    if constexpr (is_coop_height(R))
        if (r == R) { exact_stripe<R>(tk, tid, s, ci, ra, pool, sc, prof, lane, out); return; }
    if constexpr (R < kMaxRExact) exact_stripe_dispatch<R + 1>(r, tk, tid, s, ci, ra, pool, sc, prof, lane, out);
}

// Rows per lane of a long task that the host cut into n_stripes stripes (few long tasks: short stripes, so that more
// warps work on each; many: tall ones, which cost less per cell).
__host__ __device__ __forceinline__ int coop_rows(int q_len, int n_stripes) {
    const int R = (q_len + 32 * n_stripes - 1) / (32 * n_stripes);
    return R < kMinR ? kMinR : R;
}

// The task-level functions of the 32-bit kernels above (one per stripe height and sweep kind) are NOT inlined into the
// persistent kernels: each is compiled as a function of its own, so that ptxas allocates registers and decides on spills
// per variant.  Inlined, the kernels were one 2 MB body whose allocation flipped between builds (with or without a
// 1.5-3 KB stack frame, 5-8 % apart in speed) on edits to unrelated variants; a call per task costs nothing.
constexpr int kWarpsPerBlock = 16;     // one persistent 512-thread block per SM: 4 warps per scheduler (the host may launch fewer)

// order[] entries: (task << 7) | code.  code 0: the whole (single-stripe) task; 1 + s: stripe s of the first sweep of a
// multi-stripe task (exact: the only sweep; ladder: backward); 64 + s: stripe s of the ladder's forward sweep.
// Entries are listed so that whatever an entry waits for comes before it; warps take them in that order (one static
// round, slot-major, so that a small batch spreads over all SMs; then dynamic pulls from *counter), every block of the
// grid is resident (host: at most one block per SM), so whatever an entry waits for is running or done.
constexpr int kCodeBits = 7, kCodeFwd = 64;
struct TaskCursor {
    int n_order, lane, warp, wpb;
    int* counter;
    bool first;
    __device__ __forceinline__ TaskCursor(int n_order_, int* counter_)
        : n_order(n_order_), lane(threadIdx.x & 31), warp(threadIdx.x >> 5), wpb(blockDim.x >> 5), counter(counter_), first(true) {}
    // next index into order[], -1: done
    __device__ __forceinline__ int next() {
        if (first) {
            first = false;
            const int oi = warp * (int)gridDim.x + (int)blockIdx.x;
            return oi < n_order ? oi : -1;
        }
        int oi = 0;
        if (lane == 0) oi = atomicAdd(counter, 1);
        oi = __shfl_sync(kFull, oi, 0) + wpb * (int)gridDim.x;
        return oi < n_order ? oi : -1;
    }
};

// Exact (score, tstart, tend) kernel.  Persistent: the stripe height and the single- / multi-stripe path are picked
// per entry (warp-uniform dispatch), so one launch covers a whole batch, long expanded alleles included.
// out[] is indexed by task id.  smem_stride: int4 of shared memory per warp.

template <class SC>
__device__ __forceinline__ void exact_entry(int e, const Task* __restrict__ tasks, const uint32_t* __restrict__ pool, const SC& sc,
                                            const RestArgs& ra, int4* prof, int lane, int4* out) {
    const int tid = e >> kCodeBits, code = e & ((1 << kCodeBits) - 1);
    const Task tk = tasks[tid];
    if (code == 0) {
        int R, n_stripes;
        stripe_shape(tk.q_len, kMaxRExact, R, n_stripes);
        const u64 key = warp_max64(exact_dispatch<kMinR>(R, tk, pool, sc, prof, lane));
        if (lane == 0) out[tid] = finalize_rung(key, 0ull, 0, 0, 0, 0);
    } else {
        const CoopInfo ci = ra.coop[ra.coop_idx[tid]];
        exact_stripe_dispatch<kMinR>(ci.rows, tk, tid, code - 1, ci, ra, pool, sc, prof, lane, out);
    }
}

template <bool FIXED>
__global__ void __launch_bounds__(32 * kWarpsPerBlock, 1)
exact_kernel(const Task* __restrict__ tasks, RestArgs ra, const uint32_t* __restrict__ pool, ScoreW scw, int* counter,
             int smem_stride, int4* out) {
    extern __shared__ int4 smem[];
    const ScoreView<FIXED> sc(scw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int4* prof = smem + warp * smem_stride;
    TaskCursor cur(ra.n_order, counter);
    for (;;) {
        const int oi = cur.next();
        if (oi < 0) break;
        exact_entry(ra.order[oi], tasks, pool, sc, ra, prof, lane, out);
    }
}

// Round 3 for one read: all rungs kmin..kmax from one backward and one forward sweep (file header).
//
// FLAG = false: span words everywhere; every rung gets its exact (score, tstart, tend).
// FLAG = true ("flag ladder"): what the reference looks at per rung is the score and the two span
// predicates (nanoRepeat_bam.py:423-427), so the words carry just that.  Forward words are (score << 16) - mark, mark =
// the alignment started inside the left flank (set once, when a lane passes the flank's last column; a later fresh
// start is unmarked and, being the larger word, wins ties like the larger tstart does).  Backward words are
// (score << 16) - 2 * ext.  A junction candidate is then ONE VIADDMNMX: forward + backward = (total << 16) - 2 * ext -
// mark, whose integer order is the contract's (score, smallest tend, largest tstart); the prefix class needs only its
// best score, so the forward sweep has no per-step best tracking at all.  Output per rung: (score, spans both flanks,
// ends in right flank).  Needs 2 * |R| < 65536.
//
// ladder_task: single-stripe read, both sweeps on one warp, junction vectors in shared memory.
// ladder_bwd_stripe / ladder_fwd_stripe: one stripe of a long read; the stripes of a sweep run on different warps at the
// same time (CoopInfo), the forward stripes start when the last backward stripe is done.
struct LadderCtx {
    const uint32_t* qwords;
    const uint32_t* pool;
    LadderRegion reg;
    int q_len;
};

template <bool FLAG>
__device__ __forceinline__ u64 bwd_key(int key, int col) {
    if (key < 65536) return 0ull;     // no positive score
    if (FLAG) return (u64)(unsigned)key;
    return ((u64)((unsigned)key >> 16) << 32) | ((u64)((unsigned)key & 0xffffu) << 16) | (u64)(0xffffu - (unsigned)col);
}

// R-only class from the backward sweep's best key
template <bool FLAG>
__device__ __forceinline__ void r_only(u64 rkey, int n_right, int& r_score, int& r_end, int& r_start, int& rcand) {
    r_score = 0; r_end = 0; r_start = 0; rcand = kJuncNone;
    if (!rkey) return;
    if (FLAG) {
        rcand = (int)(unsigned)rkey - 2 * n_right;                       // (score << 16) - 2 * (forward end inside R)
    } else {
        r_score = (int)(rkey >> 32);
        r_end = n_right - (int)((rkey >> 16) & 0xffffu);                 // forward end inside R (exclusive)
        r_start = n_right - (0xffff - (int)(rkey & 0xffffu));            // forward start inside R
    }
}

// junction vector of forward row idx0 when the backward sweep did not write it: right part empty (H = 0, no gap state)
// for rows of the read, void below it
__device__ __forceinline__ int4 bvec_default(int idx0, int q_len) {
    return idx0 < q_len ? make_int4(0, kPadScore, kPadScore, 0) : make_int4(kPadScore, kPadScore, kPadScore, 0);
}

template <int R, bool MULTI, bool FLAG, class SW, class SC>
__device__ __forceinline__ void setup_fwd(SW& sw, const LadderTask& tk, const LadderCtx& cx, const SC& sc, int4* prof,
                                          const int4* bsm, int lane, int4* out, int r_score, int r_end, int r_start, int rcand) {
    const int c_first = cx.reg.n_left + cx.reg.m * tk.kmin;
    sw.prof = prof; sw.twords = cx.pool + cx.reg.fwd_word; sw.t_len = cx.reg.n_left + cx.reg.m * tk.kmax; sw.lane = lane;
    sw.bsm = bsm;
    sw.out = out + tk.out_off;
    sw.m = cx.reg.m;
    sw.jnext = c_first > 0 ? c_first : cx.reg.m;     // a junction at column 0 has no forward part
    sw.kcnt = c_first > 0 ? 0 : 1;
    sw.r_score = r_score; sw.r_end = r_end; sw.r_start = r_start;
    sw.rcand = rcand; sw.mark_col = cx.reg.n_left - 1;
    sw.k0 = tk.kmin; sw.min_score = sc.min_score;
    sw.sel_top = 0; sw.sel_n = 0; sw.sel_sum = 0;
    if (FLAG && c_first == 0) {        // rung 0 of a region without a left flank: the R-only class alone
        const int4 rec = finalize_flag_rung(0, kJuncNone, rcand);
        if (rec.x >= sc.min_score) { sw.sel_top = rec.x; sw.sel_n = rec.y ? 1 : 0; }
    }
}

// what the warp that holds the last rows writes when the forward sweep is over
template <bool FLAG, class SW>
__device__ __forceinline__ void finish_read(const SW& sw, const LadderTask& tk, const LadderCtx& cx, int lane, int4* out, int4* sel,
                                            int r_score, int r_end, int r_start, int rcand) {
    if (FLAG) {
        // rungs are finalised by lane 31; without a sweep every lane holds rung 0's selection
        if (lane == 31) sel[tk.read] = make_int4(sw.sel_top, sw.sel_n, (int)(unsigned)sw.sel_sum, (int)(sw.sel_sum >> 32));
    }
    if (cx.reg.n_left + cx.reg.m * tk.kmin == 0 && lane == 0)
        out[tk.out_off] = FLAG ? finalize_flag_rung(0, kJuncNone, rcand) : finalize_rung(0ull, 0ull, 0, r_score, r_end, r_start);
}

template <int R, bool FLAG, class SC>
__device__ __noinline__ void ladder_task(const LadderTask& tk, const LadderCtx& cx, const SC& sc, int4* prof, int lane,
                                            int4* out, int4* sel) {
    constexpr int BWD = FLAG ? kBwdF : kBwd, FWD = FLAG ? kFwdF : kFwd;
    int4* bsm = prof + StripeCfg<R>::PROF_INT4;
    const int q_len = cx.q_len;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) bsm[r * 32 + lane] = bvec_default(lane * R + r, q_len);
    __syncwarp();
    // ---- backward sweep: reversed read x reversed right flank ----
    u64 rkey = 0;
    if (cx.reg.n_right > 0) {
        build_profile<R>(prof, cx.qwords, q_len, lane * R, lane, ModeScore<SC, Sweep<R, BWD, false>::DEC>(sc), true);
        __syncwarp();
        Sweep<R, BWD, false> sw;
        sw.prof = prof; sw.twords = cx.pool + cx.reg.rev_word; sw.t_len = cx.reg.n_right; sw.lane = lane;
        sw.top = false; sw.bot = false;
        sw.bdst = bsm; sw.q_len = q_len; sw.brow0 = lane * R;
        sw.run(sc, 0);
        int key, col;
        sw.bwd_best(key, col);
        rkey = warp_max64(bwd_key<FLAG>(key, col));
    }
    int r_score, r_end, r_start, rcand;
    r_only<FLAG>(rkey, cx.reg.n_right, r_score, r_end, r_start, rcand);
    // ---- forward sweep over L + motif^kmax with junction tokens ----
    Sweep<R, FWD, false> sw;
    setup_fwd<R, false, FLAG>(sw, tk, cx, sc, prof, bsm, lane, out, r_score, r_end, r_start, rcand);
    sw.top = false; sw.bot = false;
    if (sw.t_len > 0) {
        __syncwarp();
        build_profile<R>(prof, cx.qwords, q_len, lane * R, lane, ModeScore<SC, Sweep<R, FWD, false>::DEC>(sc), false);
        __syncwarp();
        sw.run(sc, sw.jnext - 1);
    }
    finish_read<FLAG>(sw, tk, cx, lane, out, sel, r_score, r_end, r_start, rcand);
}

template <int R, bool FLAG, class SC>
__device__ __noinline__ void ladder_bwd_stripe(const LadderTask& tk, const LadderCtx& cx, int s, const CoopInfo& ci,
                                                  const RestArgs& ra, const SC& sc, int4* prof, int lane) {
    constexpr int BWD = FLAG ? kBwdF : kBwd;
    const int S = ci.n_stripes;
    const int epoch = ra.epoch;
    int* F = ra.flags + ci.flag_off;
    ulonglong2* bnd_a = reinterpret_cast<ulonglong2*>(ra.scratch + ci.data_off);
    ulonglong2* bnd_b = bnd_a + ci.bnd_stride;
    int4* bglob = ra.scratch + ci.data_off + 4 * ci.bnd_stride;
    __syncwarp();
    build_profile<R>(prof, cx.qwords, cx.q_len, s * 32 * R + lane * R, lane, ModeScore<SC, Sweep<R, BWD, true>::DEC>(sc), true);
    __syncwarp();
    Sweep<R, BWD, true> sw;
    sw.prof = prof; sw.twords = cx.pool + cx.reg.rev_word; sw.t_len = cx.reg.n_right; sw.lane = lane;
    sw.top = s > 0; sw.bot = s + 1 < S;
    sw.bnd_in = (s & 1) ? bnd_a : bnd_b; sw.bnd_out = (s & 1) ? bnd_b : bnd_a;
    sw.tag_in = stripe_tag(epoch, s - 1); sw.tag_out = stripe_tag(epoch, s); sw.spin = ra.spin;
    sw.bdst = bglob; sw.q_len = cx.q_len; sw.brow0 = s * 32 * R + lane * R;
    sw.run(sc, 0);
    int key, col;
    sw.bwd_best(key, col);
    const u64 rkey = warp_max64(bwd_key<FLAG>(key, col));
    __syncwarp();                       // every lane's junction vectors are stored
    if (lane == 0) {
        atomicMax(reinterpret_cast<u64*>(F + 2 * S + 2), rkey);
        __threadfence();
        atomicAdd(F + 2 * S, 1);
    }
}

template <int R, bool FLAG, class SC>
__device__ __noinline__ void ladder_fwd_stripe(const LadderTask& tk, const LadderCtx& cx, int s, const CoopInfo& ci,
                                                  const RestArgs& ra, const SC& sc, int4* prof, int lane, int4* out,
                                                  int4* sel) {
    constexpr int FWD = FLAG ? kFwdF : kFwd;
    const int S = ci.n_stripes;
    const int epoch = ra.epoch;
    int* F = ra.flags + ci.flag_off;
    ulonglong2* bnd_a = reinterpret_cast<ulonglong2*>(ra.scratch + ci.data_off + 2 * ci.bnd_stride);
    ulonglong2* bnd_b = bnd_a + ci.bnd_stride;
    int4* bglob = ra.scratch + ci.data_off + 4 * ci.bnd_stride;
    ulonglong2* tok_a = reinterpret_cast<ulonglong2*>(bglob + ci.b_stride);
    ulonglong2* tok_b = tok_a + ci.tok_stride;
    int4* bsm = prof + StripeCfg<R>::PROF_INT4;
    Sweep<R, FWD, true> sw;
    setup_fwd<R, true, FLAG>(sw, tk, cx, sc, prof, bsm, lane, out, 0, 0, 0, kJuncNone);
    sw.top = s > 0; sw.bot = s + 1 < S;
    sw.bsm_w = bsm;
    sw.bglob_rows = cx.reg.n_right > 0 ? bglob + s * 32 * R : nullptr;
    sw.bdone = F + 2 * S; sw.bdone_need = S; sw.n_right = cx.reg.n_right;
    sw.q_len = cx.q_len; sw.brow0 = s * 32 * R + lane * R;
    sw.late_pending = true;
    sw.spin = ra.spin;
    __syncwarp();
    if (sw.t_len == 0 || cx.reg.n_left + cx.reg.m * tk.kmin == 0) {
        // no sweep, or rung 0 of a region without a left flank (its record is the R-only class alone): needs the
        // backward sweep's result up front
        sw.late_load();
        if (FLAG && cx.reg.n_left + cx.reg.m * tk.kmin == 0) {
            const int4 rec = finalize_flag_rung(0, kJuncNone, sw.rcand);
            sw.sel_top = 0; sw.sel_n = 0;
            if (rec.x >= sc.min_score) { sw.sel_top = rec.x; sw.sel_n = rec.y ? 1 : 0; }
        }
    }
    if (sw.t_len > 0) {
        build_profile<R>(prof, cx.qwords, cx.q_len, s * 32 * R + lane * R, lane, ModeScore<SC, Sweep<R, FWD, true>::DEC>(sc), false);
        __syncwarp();
        sw.bnd_in = (s & 1) ? bnd_a : bnd_b; sw.bnd_out = (s & 1) ? bnd_b : bnd_a;
        sw.tok_in = (s & 1) ? tok_a : tok_b; sw.tok_out = (s & 1) ? tok_b : tok_a;
        sw.tag_in = stripe_tag(epoch, s - 1); sw.tag_out = stripe_tag(epoch, s);
        sw.run(sc, sw.jnext - 1);
    }
    if (s == S - 1) finish_read<FLAG>(sw, tk, cx, lane, out, sel, sw.r_score, sw.r_end, sw.r_start, sw.rcand);
}

template <int R, bool FLAG, int MAXR, class SC>
__device__ __forceinline__ void ladder_dispatch(int r, const LadderTask& tk, const LadderCtx& cx, const SC& sc, int4* prof,
                                                int lane, int4* out, int4* sel) {
    if (r == R) { ladder_task<R, FLAG>(tk, cx, sc, prof, lane, out, sel); return; }
    if constexpr (R < MAXR) ladder_dispatch<R + 1, FLAG, MAXR>(r, tk, cx, sc, prof, lane, out, sel);
}

template <int R, bool FLAG, class SC>
__device__ __forceinline__ void ladder_stripe_dispatch(int r, const LadderTask& tk, const LadderCtx& cx, int code, const CoopInfo& ci,
                                                       const RestArgs& ra, const SC& sc, int4* prof, int lane, int4* out,
                                                       int4* sel) {
    if constexpr (is_coop_height(R))
        if (r == R) {
            if (code < kCodeFwd) ladder_bwd_stripe<R, FLAG>(tk, cx, code - 1, ci, ra, sc, prof, lane);
            else ladder_fwd_stripe<R, FLAG>(tk, cx, code - kCodeFwd, ci, ra, sc, prof, lane, out, sel);
            return;
        }
    if constexpr (R < kMaxRLadder) ladder_stripe_dispatch<R + 1, FLAG>(r, tk, cx, code, ci, ra, sc, prof, lane, out, sel);
}

// MAXR: tallest single stripe of a code-0 entry (the host cuts longer reads into cooperative stripes, except for the
// redo list of the paired kernel, whose reads of up to 512 rows run as one stripe)
template <bool FLAG, int MAXR, class SC>
__device__ __forceinline__ void ladder_entry(int e, const LadderTask* __restrict__ tasks, const uint32_t* __restrict__ qpool,
                                             const uint32_t* __restrict__ pool, const LadderRegion* __restrict__ regs, const SC& sc,
                                             const RestArgs& ra, int4* prof, int lane, int4* out, int4* sel) {
    const int tid = e >> kCodeBits, code = e & ((1 << kCodeBits) - 1);
    const LadderTask tk = tasks[tid];
    LadderCtx cx;
    cx.qwords = qpool + tk.q_word;      // reads may live in the round-2 batch's pool
    cx.pool = pool;
    cx.reg = regs[tk.region];
    cx.q_len = tk.q_len;
    if (code == 0) {
        int R, n_stripes;
        stripe_shape(tk.q_len, MAXR, R, n_stripes);
        ladder_dispatch<kMinR, FLAG, MAXR>(R, tk, cx, sc, prof, lane, out, sel);
    } else {
        const CoopInfo ci = ra.coop[tk.pad];
        ladder_stripe_dispatch<kMinR, FLAG>(ci.rows, tk, cx, code, ci, ra, sc, prof, lane, out, sel);
    }
}

// Round-3 ladder kernel: one warp per read (single stripe) or per stripe of a long read.
// n_order_dev non-null: order[] was filled on the device (redo list), its length is there.
template <bool FIXED, bool FLAG>
__global__ void __launch_bounds__(32 * kWarpsPerBlock, 1)
ladder_kernel(const LadderTask* __restrict__ tasks, RestArgs ra, const int* __restrict__ n_order_dev,
              const uint32_t* __restrict__ qpool, const uint32_t* __restrict__ pool, const LadderRegion* __restrict__ regs, ScoreW scw, int* counter,
              int smem_stride, int4* out, int4* sel) {
    extern __shared__ int4 smem[];
    const ScoreView<FIXED> sc(scw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int4* prof = smem + warp * smem_stride;
    TaskCursor cur(n_order_dev ? *n_order_dev : ra.n_order, counter);
    for (;;) {
        const int oi = cur.next();
        if (oi < 0) break;
        ladder_entry<FLAG, kMaxRLadderOwn>(ra.order[oi], tasks, qpool, pool, regs, sc, ra, prof, lane, out, sel);
    }
}

}  // namespace nr
