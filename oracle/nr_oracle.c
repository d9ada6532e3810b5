/*
 * nr_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY) for the NanoRepeat repeat-size hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this file's shared object.  The product (nanorepeat_b200/) never links, imports or calls it.
 *
 * PARITY UNPINNED: the arithmetic of this path is not in the reference tree.  The reference shells
 * every alignment out to `pyminimap2.main(cmd)` (pyminimap2>=2.30.0, reference setup.py:24,
 * pyproject.toml:20 -- lower bound only, not vendored, not installable here), at the call sites
 * src/NanoRepeat/nanoRepeat_bam.py:362 (round 2) and :497 (round 3).  The reference ships no tests,
 * fixtures or golden vectors, so there is nothing to pin this restatement against except hand-derived
 * known answers (tests/golden/micro_cases.json).  What is restated here is minimap2's published
 * scoring model for `-x map-ont` (match +2, mismatch -4, two-piece affine gap: a gap of length l costs
 * min(4 + 2l, 24 + l), ambiguous base -1) as an EXACT full-rectangle local alignment -- the optimum that
 * minimap2's seed-chain-extend heuristic approximates (band, z-drop, minimizer seeding are not modelled).
 *
 * Result contract per (query, target) task -- this is the tie-break the CUDA kernels must reproduce:
 *   score  = max over all local alignments (0 if nothing scores above 0)
 *   tend   = the smallest target end (0-based, exclusive) over all alignments reaching `score`
 *   tstart = the largest target start (0-based) over all alignments reaching `score` and ending at `tend`
 * (PAF coordinate convention of reference src/NanoRepeat/paf.py:39-52: tstart/tend 0-based half-open.)
 * score == 0 reports tstart = tend = 0.
 *
 * The DP carries (score, start) pairs compared lexicographically; adding a constant to the score keeps
 * the order, so max/plus over pairs is still a semiring and the DP yields, per cell, the best score and
 * the largest start among the paths that reach it.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t match;      /* +a            (minimap2 -A) */
    int32_t mismatch;   /* b, charged -b (minimap2 -B) */
    int32_t gap_open1;  /* q             (minimap2 -O first value) */
    int32_t gap_ext1;   /* e             (minimap2 -E first value) */
    int32_t gap_open2;  /* q2 */
    int32_t gap_ext2;   /* e2 */
    int32_t ambiguous;  /* sc_ambi, charged -sc_ambi when either base is not ACGT */
    int32_t min_dp_score; /* minimap2 -s: alignments below it are not reported (used by selection only) */
} nro_scoring_t;

typedef struct { int32_t score, tstart, tend; } nro_aln_t;

typedef struct { int32_t s, k; } cell_t;   /* score, start key (larger start wins ties) */

static inline cell_t cmax(cell_t a, cell_t b) {
    if (a.s != b.s) return a.s > b.s ? a : b;
    return a.k >= b.k ? a : b;
}
static inline cell_t cadd(cell_t a, int32_t d) { a.s += d; return a; }

static inline int code_of(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

#define NEG_INF (-(1 << 28))

/* One task, full rectangle, explicit (score, start) structs: the DEFINITIONAL version.
   Returns 0, or -1 on allocation failure. */
int nro_align_tuple(const nro_scoring_t* sc, const char* query, int32_t qlen,
              const char* target, int32_t tlen, nro_aln_t* out)
{
    out->score = 0; out->tstart = 0; out->tend = 0;
    if (qlen <= 0 || tlen <= 0) return 0;
    /* column-major sweep: arrays indexed by query row i (0..qlen) */
    cell_t* H  = (cell_t*)malloc(sizeof(cell_t) * (size_t)(qlen + 1));
    cell_t* E1 = (cell_t*)malloc(sizeof(cell_t) * (size_t)(qlen + 1));
    cell_t* E2 = (cell_t*)malloc(sizeof(cell_t) * (size_t)(qlen + 1));
    uint8_t* qc = (uint8_t*)malloc((size_t)qlen);
    if (!H || !E1 || !E2 || !qc) { free(H); free(E1); free(E2); free(qc); return -1; }
    for (int32_t i = 0; i < qlen; ++i) qc[i] = (uint8_t)code_of(query[i]);
    const int32_t qe1 = sc->gap_open1 + sc->gap_ext1, qe2 = sc->gap_open2 + sc->gap_ext2;
    /* column 0: H(i,0) = (0, start 0); E(i,1) = H(i,0) - q - e */
    for (int32_t i = 0; i <= qlen; ++i) {
        H[i].s = 0; H[i].k = 0;
        E1[i].s = -qe1; E1[i].k = 0;
        E2[i].s = -qe2; E2[i].k = 0;
    }
    int32_t best_s = 0, best_j = 0, best_k = 0;
    for (int32_t j = 1; j <= tlen; ++j) {
        const int tc = code_of(target[j - 1]);
        cell_t fresh; fresh.s = 0; fresh.k = j;         /* an alignment whose first aligned target base is j */
        cell_t hdiag = H[0];                            /* H(0, j-1) = (0, j-1) */
        H[0] = fresh;                                   /* H(0, j)   = (0, j)   */
        cell_t f1 = cadd(fresh, -qe1), f2 = cadd(fresh, -qe2);   /* F(1, j) from H(0, j) */
        for (int32_t i = 1; i <= qlen; ++i) {
            int32_t s;
            if (tc > 3 || qc[i - 1] > 3) s = -sc->ambiguous;
            else s = (tc == qc[i - 1]) ? sc->match : -sc->mismatch;
            cell_t h = cadd(hdiag, s);
            h = cmax(h, E1[i]); h = cmax(h, E2[i]);
            h = cmax(h, f1);    h = cmax(h, f2);
            h = cmax(h, fresh);
            hdiag = H[i];
            H[i] = h;
            /* E(i, j+1) and F(i+1, j) */
            E1[i] = cmax(cadd(h, -qe1), cadd(E1[i], -sc->gap_ext1));
            E2[i] = cmax(cadd(h, -qe2), cadd(E2[i], -sc->gap_ext2));
            f1 = cmax(cadd(h, -qe1), cadd(f1, -sc->gap_ext1));
            f2 = cmax(cadd(h, -qe2), cadd(f2, -sc->gap_ext2));
            /* result order: score desc, then tend asc (strict > keeps the earliest column),
               then start desc within the same column */
            if (h.s > best_s) { best_s = h.s; best_j = j; best_k = h.k; }
            else if (h.s == best_s && best_s > 0 && j == best_j && h.k > best_k) best_k = h.k;
        }
    }
    out->score = best_s;
    if (best_s > 0) { out->tstart = best_k; out->tend = best_j; }
    free(H); free(E1); free(E2); free(qc);
    return 0;
}

/*
 * Same contract as nro_align_tuple, written for speed (this is what the CPU baseline times): each
 * (score, start) pair lives in one int64 as score * 2^32 + start, so the lexicographic max is a plain
 * integer max and "add d to the score" is "+ d * 2^32".  tests/test_oracle.py checks it against
 * nro_align_tuple and against the pure-Python restatement.
 */
int nro_align(const nro_scoring_t* sc, const char* query, int32_t qlen,
              const char* target, int32_t tlen, nro_aln_t* out)
{
    out->score = 0; out->tstart = 0; out->tend = 0;
    if (qlen <= 0 || tlen <= 0) return 0;
    typedef int64_t v_t;
    #define V(s, k) ((v_t)(((int64_t)(s)) * 4294967296LL + (int64_t)(k)))
    #define VMAX(a, b) ((a) > (b) ? (a) : (b))
    v_t* H  = (v_t*)malloc(sizeof(v_t) * (size_t)(qlen + 1) * 3);
    int64_t* prof = (int64_t*)malloc(sizeof(int64_t) * (size_t)qlen * 5);
    if (!H || !prof) { free(H); free(prof); return -1; }
    v_t* E1 = H + (qlen + 1);
    v_t* E2 = E1 + (qlen + 1);
    /* query profile: prof[c * qlen + i] = substitution score (already shifted) of query[i] vs code c */
    for (int c = 0; c < 5; ++c)
        for (int32_t i = 0; i < qlen; ++i) {
            int qc = code_of(query[i]);
            int32_t s = (c > 3 || qc > 3) ? -sc->ambiguous : (c == qc ? sc->match : -sc->mismatch);
            prof[(size_t)c * qlen + i] = V(s, 0);
        }
    const v_t qe1 = V(sc->gap_open1 + sc->gap_ext1, 0), qe2 = V(sc->gap_open2 + sc->gap_ext2, 0);
    const v_t e1 = V(sc->gap_ext1, 0), e2 = V(sc->gap_ext2, 0);
    for (int32_t i = 0; i <= qlen; ++i) { H[i] = 0; E1[i] = -qe1; E2[i] = -qe2; }
    int32_t best_s = 0, best_j = 0, best_k = 0;
    for (int32_t j = 1; j <= tlen; ++j) {
        const int64_t* pr = prof + (size_t)code_of(target[j - 1]) * qlen - 1;   /* pr[i], i = 1..qlen */
        const v_t fresh = V(0, j);
        v_t hdiag = H[0];
        H[0] = fresh;
        v_t f1 = fresh - qe1, f2 = fresh - qe2;
        v_t colmax = 0;
        for (int32_t i = 1; i <= qlen; ++i) {
            v_t h = hdiag + pr[i];
            v_t ee1 = E1[i], ee2 = E2[i];
            h = VMAX(h, ee1); h = VMAX(h, ee2); h = VMAX(h, f1); h = VMAX(h, f2); h = VMAX(h, fresh);
            hdiag = H[i];
            H[i] = h;
            v_t ho1 = h - qe1, ho2 = h - qe2;
            ee1 -= e1; ee2 -= e2; f1 -= e1; f2 -= e2;
            E1[i] = VMAX(ho1, ee1); E2[i] = VMAX(ho2, ee2);
            f1 = VMAX(ho1, f1);     f2 = VMAX(ho2, f2);
            colmax = VMAX(colmax, h);
        }
        int32_t cs = (int32_t)(colmax >> 32);
        if (cs > best_s) { best_s = cs; best_j = j; best_k = (int32_t)(colmax & 0xffffffffLL); }
    }
    #undef V
    #undef VMAX
    out->score = best_s;
    if (best_s > 0) { out->tstart = best_k; out->tend = best_j; }
    free(H); free(prof);
    return 0;
}

/* ---- threading: plain pthreads work queue (this image's default gcc wrapper lacks libgomp.spec) ---- */
#include <pthread.h>
#include <unistd.h>

typedef struct {
    const nro_scoring_t* sc;
    int32_t n_items;
    int32_t n_reads;            /* ladder mode: reads (n_items = rungs) */
    volatile int32_t next;      /* claimed with __sync_fetch_and_add */
    volatile int rc;
    /* batch mode */
    const char* const* queries; const int32_t* qlen;
    const char* const* targets; const int32_t* tlen;
    /* ladder mode */
    const char* const* cores; const int32_t* core_len;
    const char* left; int32_t n_left; const char* right; int32_t n_right;
    const char* motif; int32_t m;
    const int32_t* kmin; const int32_t* kmax; const int64_t* rung_offset;
    nro_aln_t* out;
    int ladder;
} work_t;

int nro_align_ladder(const nro_scoring_t* sc, const char* core, int32_t core_len,
                     const char* left, int32_t n_left, const char* right, int32_t n_right,
                     const char* motif, int32_t m, int32_t kmin, int32_t kmax, nro_aln_t* out);

static void* worker(void* arg) {
    work_t* w = (work_t*)arg;
    for (;;) {
        int32_t t = __sync_fetch_and_add(&w->next, 1);
        if (t >= w->n_items) break;
        int r;
        if (w->ladder) {
            /* one work item per RUNG (a long expanded allele has up to 301 of them, 10^8 cells each): find the read
               whose rung range holds item t, then score that one rung as its own rectangle */
            int32_t lo = 0, hi = w->n_reads;             /* rung_offset[lo] <= t < rung_offset[hi] */
            while (hi - lo > 1) {
                int32_t mid = lo + (hi - lo) / 2;
                if (w->rung_offset[mid] <= (int64_t)t) lo = mid; else hi = mid;
            }
            int32_t k = w->kmin[lo] + (int32_t)((int64_t)t - w->rung_offset[lo]);
            r = nro_align_ladder(w->sc, w->cores[lo], w->core_len[lo], w->left, w->n_left, w->right, w->n_right,
                                 w->motif, w->m, k, k, w->out + t);
        } else
            r = nro_align(w->sc, w->queries[t], w->qlen[t], w->targets[t], w->tlen[t], &w->out[t]);
        if (r != 0) w->rc = -1;
    }
    return NULL;
}

static int run_pool(work_t* w, int32_t n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if (n_threads > w->n_items) n_threads = w->n_items > 0 ? w->n_items : 1;
    pthread_t th[256];
    int started = 0;
    for (int i = 1; i < n_threads; ++i)
        if (pthread_create(&th[started], NULL, worker, w) == 0) ++started;
    worker(w);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    return w->rc;
}

/* Batch of independent tasks, n_threads host threads. */
int nro_align_batch(const nro_scoring_t* sc, int32_t n_tasks,
                    const char* const* queries, const int32_t* qlen,
                    const char* const* targets, const int32_t* tlen,
                    nro_aln_t* out, int32_t n_threads)
{
    work_t w; memset(&w, 0, sizeof w);
    w.sc = sc; w.n_items = n_tasks; w.queries = queries; w.qlen = qlen; w.targets = targets; w.tlen = tlen;
    w.out = out; w.ladder = 0;
    return run_pool(&w, n_threads);
}

/*
 * Ladder helper for the round-3 shape (reference nanoRepeat_bam.py:474-481): scores `core` against
 * left + motif*k + right for k = kmin..kmax, each rung as an independent full rectangle.
 * out[k - kmin] receives the rung's result.  Template strings are built here so Python callers do not
 * have to materialise n_rungs large strings per read.
 */
int nro_align_ladder(const nro_scoring_t* sc, const char* core, int32_t core_len,
                     const char* left, int32_t n_left, const char* right, int32_t n_right,
                     const char* motif, int32_t m, int32_t kmin, int32_t kmax, nro_aln_t* out)
{
    if (kmax < kmin) return 0;
    size_t cap = (size_t)n_left + (size_t)m * (size_t)kmax + (size_t)n_right + 1;
    char* tpl = (char*)malloc(cap);
    if (!tpl) return -1;
    int rc = 0;
    memcpy(tpl, left, (size_t)n_left);
    for (int32_t k = 0; k < kmax; ++k) memcpy(tpl + n_left + (size_t)k * m, motif, (size_t)m);
    for (int32_t k = kmax; k >= kmin; --k) {
        /* descending k: the right flank written for rung k only clobbers units >= k, which smaller rungs
           never read */
        size_t body = (size_t)n_left + (size_t)m * (size_t)k;
        memcpy(tpl + body, right, (size_t)n_right);
        if (nro_align(sc, core, core_len, tpl, (int32_t)(body + (size_t)n_right), &out[k - kmin]) != 0) rc = -1;
    }
    free(tpl);
    return rc;
}

/* Batched ladders (one per read) over n_threads host threads. rung_offset[r] indexes out[] for read r. */
int nro_align_ladders(const nro_scoring_t* sc, int32_t n_reads,
                      const char* const* cores, const int32_t* core_len,
                      const char* left, int32_t n_left, const char* right, int32_t n_right,
                      const char* motif, int32_t m,
                      const int32_t* kmin, const int32_t* kmax, const int64_t* rung_offset,
                      nro_aln_t* out, int32_t n_threads)
{
    work_t w; memset(&w, 0, sizeof w);
    if (n_reads <= 0 || rung_offset[n_reads] <= 0) return 0;
    if (rung_offset[n_reads] > 0x7fffffffLL) return -1;
    w.sc = sc; w.n_items = (int32_t)rung_offset[n_reads]; w.n_reads = n_reads; w.cores = cores; w.core_len = core_len;
    w.left = left; w.n_left = n_left; w.right = right; w.n_right = n_right; w.motif = motif; w.m = m;
    w.kmin = kmin; w.kmax = kmax; w.rung_offset = rung_offset; w.out = out; w.ladder = 1;
    return run_pool(&w, n_threads);
}

int nro_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Joint path: alignment score AND the score of the alignment inside a template window, in one DP.
 *
 * The reference re-scores the CIGAR of every alignment inside the window [a, b) of the template with
 * tk.target_region_alignment_stats_from_cigar (src/NanoRepeat/tk.py:435-500): '=' +2, 'X' -4 per base inside the
 * window; a deletion by the part of it inside the window, -4 - 2 (part - 1); an insertion in full, -4 - 2 (len - 1),
 * when its template position p satisfies a < p < b - 1.  Which of several optimal alignments minimap2 reports is not
 * pinned by anything in the reference tree (PARITY UNPINNED, see the top of this file), so the contract names one:
 * DP values are pairs (score, payload) compared lexicographically, payload = window score collected along the path, i.e.
 * among all alignments of the best score the one with the HIGHEST window score is the alignment.  A move's payload:
 *   diagonal into template position p:             +2 / -4 (bases equal / not) if a <= p < b
 *   deletion step consuming template position p:   -4 if it opens the run or p == a, else -2, if a <= p < b
 *   insertion step at template position p:         -4 if it opens the run, else -2, if a < p < b - 1
 * nro_align_window returns (score, payload); nro_align_window_cigar also walks the chosen path back and writes its
 * CIGAR (=, X, I, D), so that tests can hand it to the reference's own function: payload == its .score.
 * reverse != 0: the reverse complement of the query is aligned (minimap2 reports '-' strand hits in template
 * coordinates; the window arithmetic is the same).
 */
typedef struct { int32_t score, window_score, tstart, tend; } nro_win_t;

static inline int comp_code(int c) { return c > 3 ? c : 3 - c; }      /* A0 C1 G2 T3 */

typedef struct { int64_t h, e1, e2; } wcol_t;

#define WV(s, p) ((int64_t)(s) * 4294967296LL + (int64_t)(p))
static inline int64_t wmax(int64_t a, int64_t b) { return a > b ? a : b; }

/* tb != NULL: one byte per cell (row-major, (qlen + 1) x (tlen + 1)):
 * bits 0-2 source of H (0 fresh start, 1 diagonal, 2 E1, 3 E2, 4 F1, 5 F2), bit 3: E1 extended (else opened),
 * bit 4: E2 extended, bit 5: F1 extended, bit 6: F2 extended -- for the states LEAVING this cell. */
static int window_dp(const nro_scoring_t* sc, const uint8_t* qc, int32_t qlen, const uint8_t* tc, int32_t tlen,
                     int32_t a, int32_t b, uint8_t* tb, nro_win_t* out, int32_t* best_i, int32_t* best_j)
{
    const int64_t qe1 = sc->gap_open1 + sc->gap_ext1, qe2 = sc->gap_open2 + sc->gap_ext2;
    const int64_t x1 = sc->gap_ext1, x2 = sc->gap_ext2;
    wcol_t* col = (wcol_t*)malloc(sizeof(wcol_t) * (size_t)(qlen + 1));
    if (!col) return -1;
    for (int32_t i = 0; i <= qlen; ++i) { col[i].h = 0; col[i].e1 = WV(-qe1, 0); col[i].e2 = WV(-qe2, 0); }
    int64_t best = 0;
    *best_i = 0; *best_j = 0;
    for (int32_t j = 1; j <= tlen; ++j) {
        const int32_t p = j - 1;                                  /* template position of this column */
        const int in_diag = p >= a && p < b;
        const int in_next = p + 1 >= a && p + 1 < b;              /* E leaving this column consumes position p + 1 */
        const int in_ins = j > a && j < b - 1;                    /* insertions behind this column sit at position j */
        const int64_t h_open_pay = in_next ? -4 : 0, h_ext_pay = in_next ? (p + 1 == a ? -4 : -2) : 0;
        const int64_t v_open_pay = in_ins ? -4 : 0, v_ext_pay = in_ins ? -2 : 0;
        int64_t hdiag = col[0].h;                                 /* H(0, j - 1) = 0 */
        col[0].h = 0;
        int64_t f1 = WV(-(1 << 28), 0), f2 = WV(-(1 << 28), 0);   /* nothing above row 1 */
        for (int32_t i = 1; i <= qlen; ++i) {
            const int q = qc[i - 1], t = tc[p];
            int32_t s; int64_t pay = 0;
            if (q > 3 || t > 3) s = -sc->ambiguous; else s = q == t ? sc->match : -sc->mismatch;
            if (in_diag) pay = (q == t && q <= 3) ? 2 : -4;
            int64_t cand[6] = {0, hdiag + WV(s, pay), col[i].e1, col[i].e2, f1, f2};
            int src = 0; int64_t h = 0;
            for (int k = 1; k < 6; ++k) if (cand[k] > h) { h = cand[k]; src = k; }
            hdiag = col[i].h;
            col[i].h = h;
            const int64_t oe1 = h + WV(-qe1, h_open_pay), xe1 = col[i].e1 + WV(-x1, h_ext_pay);
            const int64_t oe2 = h + WV(-qe2, h_open_pay), xe2 = col[i].e2 + WV(-x2, h_ext_pay);
            const int64_t of1 = h + WV(-qe1, v_open_pay), xf1 = f1 + WV(-x1, v_ext_pay);
            const int64_t of2 = h + WV(-qe2, v_open_pay), xf2 = f2 + WV(-x2, v_ext_pay);
            col[i].e1 = wmax(oe1, xe1); col[i].e2 = wmax(oe2, xe2);
            f1 = wmax(of1, xf1); f2 = wmax(of2, xf2);
            if (tb) tb[(size_t)i * (size_t)(tlen + 1) + (size_t)j] =
                (uint8_t)(src | ((xe1 > oe1) << 3) | ((xe2 > oe2) << 4) | ((xf1 > of1) << 5) | ((xf2 > of2) << 6));
            if (h > best) { best = h; *best_i = i; *best_j = j; }
        }
    }
    free(col);
    int64_t s = (best + 2147483648LL) >> 32;                      /* payload is a signed 32-bit value around the score field */
    out->score = (int32_t)s;
    out->window_score = (int32_t)(best - s * 4294967296LL);
    out->tstart = 0; out->tend = *best_j;
    return 0;
}

static uint8_t* codes_of(const char* s, int32_t n, int reverse) {
    uint8_t* c = (uint8_t*)malloc((size_t)(n > 0 ? n : 1));
    if (!c) return NULL;
    for (int32_t i = 0; i < n; ++i) c[i] = (uint8_t)(reverse ? comp_code(code_of(s[n - 1 - i])) : code_of(s[i]));
    return c;
}

int nro_align_window(const nro_scoring_t* sc, const char* query, int32_t qlen, const char* target, int32_t tlen,
                     int32_t win_a, int32_t win_b, int32_t reverse, nro_win_t* out)
{
    out->score = out->window_score = out->tstart = out->tend = 0;
    if (qlen <= 0 || tlen <= 0) return 0;
    uint8_t* qc = codes_of(query, qlen, reverse); uint8_t* tc = codes_of(target, tlen, 0);
    int32_t bi, bj;
    int rc = (qc && tc) ? window_dp(sc, qc, qlen, tc, tlen, win_a, win_b, NULL, out, &bi, &bj) : -1;
    free(qc); free(tc);
    return rc;
}

/* cigar: caller's buffer of cigar_cap bytes; returns 0, -1 out of memory, -2 buffer too small */
int nro_align_window_cigar(const nro_scoring_t* sc, const char* query, int32_t qlen, const char* target, int32_t tlen,
                           int32_t win_a, int32_t win_b, int32_t reverse, nro_win_t* out, char* cigar, int32_t cigar_cap)
{
    out->score = out->window_score = out->tstart = out->tend = 0;
    if (cigar_cap > 0) cigar[0] = 0;
    if (qlen <= 0 || tlen <= 0) return 0;
    uint8_t* qc = codes_of(query, qlen, reverse); uint8_t* tc = codes_of(target, tlen, 0);
    uint8_t* tb = (uint8_t*)calloc((size_t)(qlen + 1) * (size_t)(tlen + 1), 1);
    int32_t bi = 0, bj = 0;
    int rc = (qc && tc && tb) ? window_dp(sc, qc, qlen, tc, tlen, win_a, win_b, tb, out, &bi, &bj) : -1;
    if (rc == 0 && out->score > 0) {
        /* walk back from the best cell; ops are collected last-to-first */
        size_t cap = (size_t)qlen + (size_t)tlen + 2, n = 0;
        char* ops = (char*)malloc(cap);
        if (!ops) rc = -1;
        int32_t i = bi, j = bj;
        int state = 0;                   /* 0: in H, 2/3: in E1/E2 (arrived at (i, j) by a deletion), 4/5: in F1/F2 */
        while (rc == 0) {
            const uint8_t c = tb[(size_t)i * (size_t)(tlen + 1) + (size_t)j];
            if (state == 0) {
                const int src = c & 7;
                if (src == 0) break;                                          /* fresh start: the alignment begins here */
                if (src == 1) { ops[n++] = (qc[i - 1] == tc[j - 1] && qc[i - 1] <= 3) ? '=' : 'X'; --i; --j; }
                else state = src;                                             /* H(i, j) was taken from a gap state */
            } else if (state == 2 || state == 3) {
                /* E(i, j) came from cell (i, j - 1): opened from its H or extended from its E */
                ops[n++] = 'D'; --j;
                const uint8_t pc = tb[(size_t)i * (size_t)(tlen + 1) + (size_t)j];
                if (j <= 0 || !((pc >> (state == 2 ? 3 : 4)) & 1)) state = 0;
            } else {
                /* F(i, j) came from cell (i - 1, j) */
                ops[n++] = 'I'; --i;
                const uint8_t pc = tb[(size_t)i * (size_t)(tlen + 1) + (size_t)j];
                if (i <= 0 || !((pc >> (state == 4 ? 5 : 6)) & 1)) state = 0;
            }
            if (n + 1 >= cap) { rc = -1; break; }
            if (i <= 0 || j <= 0) break;
        }
        if (rc == 0) {
            out->tstart = j;
            /* run-length encode, first-to-last */
            int32_t w = 0;
            for (size_t k = n; k > 0;) {
                const char op = ops[k - 1];
                size_t run = 0;
                while (k > 0 && ops[k - 1] == op) { ++run; --k; }
                int len = snprintf(cigar + w, (size_t)(cigar_cap - w), "%zu%c", run, op);
                if (len < 0 || len >= cigar_cap - w) { rc = -2; break; }
                w += len;
            }
        }
        free(ops);
    }
    free(qc); free(tc); free(tb);
    return rc;
}
