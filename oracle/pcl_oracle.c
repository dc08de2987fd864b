/*
 * pcl_oracle.c -- CPU restatement (fp64) of the reference decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under polarcode_and_ldpc_b200/ may import,
 * link or execute this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py do, and there only as the
 * checker or the CPU timing baseline.
 *
 * What it restates (paths relative to /root/reference):
 *   src/polar/decoder.py   SCDecoder  (:12-173)  -> oracle_polar_sc
 *                          SCLDecoder (:176-444) -> oracle_polar_scl
 *   src/polar/utils.py     bit_reverse (:11-26), crc_check (:128-163)
 *   src/ldpc/decoder.py    BPDecoder (:11-205), MSDecoder (:208-355) -> oracle_ldpc
 *
 * Pinning: tests/test_oracle_golden.py checks every function here against
 * the .npz files under tests/golden/, which tests/golden/gen_golden.py produced by importing and
 * running the reference's own decoders in the build container (decoded bits,
 * path metrics, leaf LLRs, iteration counts, total LLRs).  The one exception is
 * use_crc=1 selection: the reference ignores use_crc (decoder.py:202-203,259),
 * so that rule has no reference behaviour -> "parity unpinned" for CRC only.
 *
 * Formulation.  The reference decodes leaves in bit-reversed order over (N, n+1)
 * matrices.  Index i (decode step) <-> reference bit index l = bit_reverse(i).
 * Stage s of the reference (decoder.py:80-94) pairs rows at distance 2^s; in
 * decode-step ("natural") order that is level d = s+1 whose node halves are
 * contiguous.  Each f/g below therefore consumes exactly the operands the
 * reference's _upper_llr/_lower_llr consume, in the same order, so LLRs are
 * bit-identical; only storage differs (N>>d values per level instead of N).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_OK 0
#define ORACLE_EINVAL 1
#define ORACLE_EDEG1 2 /* MSDecoder on a degree-1 check: reference raises ValueError */

/* src/polar/utils.py:11-26 */
static inline int bit_reverse(int v, int nbits)
{
    int r = 0;
    for (int i = 0; i < nbits; i++) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

static inline int ilog2(int N) { int n = 0; while ((1 << n) < N) n++; return n; }

/* decoder.py:121-127 / :408-410 : np.sign(l1)*np.sign(l2)*min(|l1|,|l2|) */
static inline double f_upper(double a, double b)
{
    double sa = (double)((a > 0) - (a < 0));
    double sb = (double)((b > 0) - (b < 0));
    double fa = fabs(a), fb = fabs(b);
    return sa * sb * (fa <= fb ? fa : fb);
}

/* decoder.py:129-144 / :412-417 : btm + top (bit 0) or btm - top (bit 1) */
static inline double g_lower(double btm, double top, int bit)
{
    return bit == 0 ? btm + top : btm - top;
}

/* decoder.py:374-406, four branches kept as written */
static inline double log_likelihood(double llr, int bit)
{
    if (bit == 0) {
        if (llr >= 0) return -log1p(exp(-llr));
        return llr - log1p(exp(llr));
    } else {
        if (llr >= 0) return -llr - log1p(exp(-llr));
        return -log1p(exp(llr));
    }
}

/* Per-path compact state: llr level d (1..n) has N>>d values at off[d];
 * left[d] holds the partial sums of the completed left child at level d. */
typedef struct {
    double *llr;   /* N-1 values, level d at (N - (N >> (d-1)))  */
    uint8_t *left; /* N-1 bits (one per byte), same offsets        */
    uint8_t *u;    /* N decisions in decode-step order             */
    double *leaf;  /* N leaf LLRs in decode-step order             */
} path_t;

static inline int lvl_off(int N, int d) { return N - (N >> (d - 1)); }

/* decoder.py:73-94 (_update_llrs) for decode step i: recompute levels
 * start..n where start = n - ALL(l) + 1 = n - ctz(i) (i>0), 1 for i==0. */
static double update_llrs(const path_t *p, const double *y, int N, int n, int i)
{
    int start = 1;
    if (i != 0) { int c = 0; while (((i >> c) & 1) == 0) c++; start = n - c; }
    for (int d = start; d <= n; d++) {
        int sz = N >> d;
        const double *src = (d == 1) ? y : p->llr + lvl_off(N, d - 1);
        double *dst = p->llr + lvl_off(N, d);
        const uint8_t *ub = p->left + lvl_off(N, d);
        int bit = (i >> (n - d)) & 1;
        if (bit == 0)
            for (int k = 0; k < sz; k++) dst[k] = f_upper(src[k], src[k + sz]);
        else
            for (int k = 0; k < sz; k++) dst[k] = g_lower(src[k + sz], src[k], ub[k]);
    }
    return p->llr[lvl_off(N, n)];
}

/* decoder.py:96-115 (_update_bits): after deciding step i with bit u, fold
 * completed right children upwards: parent = [left ^ right, right]. */
static void update_bits(path_t *p, int N, int n, int i, int u, uint8_t *tmp)
{
    int d = n, sz = 1, idx = i;
    tmp[0] = (uint8_t)u;
    while (d > 0 && (idx & 1)) {
        const uint8_t *lf = p->left + lvl_off(N, d);
        for (int k = 0; k < sz; k++) { tmp[sz + k] = tmp[k]; }
        for (int k = 0; k < sz; k++) { tmp[k] = lf[k] ^ tmp[sz + k]; }
        sz <<= 1; d--; idx >>= 1;
    }
    if (d > 0) memcpy(p->left + lvl_off(N, d), tmp, (size_t)sz);
}

/* src/polar/utils.py:128-163, bit-serial MSB-first, init 0 */
static int crc_check_bits(const uint8_t *bits, int len, uint32_t poly, int crc_len)
{
    uint64_t crc = 0, top = 1ull << (crc_len - 1), mask = (1ull << crc_len) - 1;
    for (int k = 0; k < len; k++) {
        crc ^= ((uint64_t)bits[k]) << (crc_len - 1);
        if (crc & top) crc = (crc << 1) ^ poly; else crc = crc << 1;
        crc &= mask;
    }
    return crc == 0;
}

/* ------------------------------------------------------------------ SC -- */
/* SCDecoder.decode, decoder.py:38-71.  frozen_ref[l]=1 if reference index l is
 * frozen.  u_ref_out[f*N + l] = B[l, n]; leaf_out[f*N + l] = L[l, n]. */
int oracle_polar_sc(int N, const uint8_t *frozen_ref, const double *llr, int64_t F,
                    uint8_t *u_ref_out, double *leaf_out, int nthreads)
{
    if (N < 2 || (N & (N - 1))) return ORACLE_EINVAL;
    int n = ilog2(N);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        path_t p;
        p.llr = (double *)malloc(sizeof(double) * N);
        p.left = (uint8_t *)calloc(N, 1);
        p.u = NULL; p.leaf = NULL;
        double *y = (double *)malloc(sizeof(double) * N);
        uint8_t *tmp = (uint8_t *)malloc(N);
#pragma omp for schedule(static)
        for (int64_t f = 0; f < F; f++) {
            const double *in = llr + f * N;
            for (int i = 0; i < N; i++) y[i] = in[bit_reverse(i, n)];
            for (int i = 0; i < N; i++) {
                int l = bit_reverse(i, n);
                double x = update_llrs(&p, y, N, n, i);
                int u = frozen_ref[l] ? 0 : (x >= 0 ? 0 : 1); /* :61-64, :117-119 */
                u_ref_out[f * N + l] = (uint8_t)u;
                if (leaf_out) leaf_out[f * N + l] = x;
                update_bits(&p, N, n, i, u, tmp);
            }
        }
        free(p.llr); free(p.left); free(y); free(tmp);
    }
    return ORACLE_OK;
}

/* ----------------------------------------------------------------- SCL -- */
typedef struct { double m; int parent; int bit; } cand_t;

/* SCLDecoder.decode, decoder.py:225-262 (+ :264-339).
 * pm_out[f*L + slot] = path_metrics (-inf for inactive slots).
 * use_crc: see header; info bits are taken in ascending reference index. */
int oracle_polar_scl(int N, int L, const uint8_t *frozen_ref, const double *llr, int64_t F,
                     uint8_t *u_ref_out, double *pm_out, double *leaf_out,
                     int use_crc, uint32_t crc_poly, int crc_len, int nthreads)
{
    if (N < 2 || (N & (N - 1)) || L < 1) return ORACLE_EINVAL;
    int n = ilog2(N);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        path_t *cur = (path_t *)malloc(sizeof(path_t) * L);
        path_t *nxt = (path_t *)malloc(sizeof(path_t) * L);
        for (int q = 0; q < L; q++) {
            cur[q].llr = (double *)malloc(sizeof(double) * N);
            cur[q].left = (uint8_t *)calloc(N, 1);
            cur[q].u = (uint8_t *)calloc(N, 1);
            cur[q].leaf = (double *)calloc(N, sizeof(double));
            nxt[q].llr = (double *)malloc(sizeof(double) * N);
            nxt[q].left = (uint8_t *)calloc(N, 1);
            nxt[q].u = (uint8_t *)calloc(N, 1);
            nxt[q].leaf = (double *)calloc(N, sizeof(double));
        }
        double *pm = (double *)malloc(sizeof(double) * L);
        double *pm2 = (double *)malloc(sizeof(double) * L);
        cand_t *cand = (cand_t *)malloc(sizeof(cand_t) * 2 * L);
        cand_t *srt = (cand_t *)malloc(sizeof(cand_t) * 2 * L);
        double *y = (double *)malloc(sizeof(double) * N);
        uint8_t *tmp = (uint8_t *)malloc(N);
        uint8_t *ibits = (uint8_t *)malloc(N);
#pragma omp for schedule(static)
        for (int64_t f = 0; f < F; f++) {
            const double *in = llr + f * N;
            for (int i = 0; i < N; i++) y[i] = in[bit_reverse(i, n)];
            int nact = 1;                       /* :238-241 */
            for (int q = 0; q < L; q++) pm[q] = -INFINITY;
            pm[0] = 0.0;
            for (int i = 0; i < N; i++) {
                int l = bit_reverse(i, n);
                if (frozen_ref[l]) {            /* _decode_frozen_bit :264-281 */
                    for (int q = 0; q < nact; q++) {
                        double x = update_llrs(&cur[q], y, N, n, i);
                        cur[q].u[i] = 0; cur[q].leaf[i] = x;
                        pm[q] += log_likelihood(x, 0);
                        update_bits(&cur[q], N, n, i, 0, tmp);
                    }
                } else {                        /* _decode_info_bit :283-339 */
                    for (int q = 0; q < nact; q++) {
                        double x = update_llrs(&cur[q], y, N, n, i);
                        cur[q].leaf[i] = x;
                        cand[q].m = pm[q] + log_likelihood(x, 0);
                        cand[q].parent = q; cand[q].bit = 0;
                        cand[nact + q].m = pm[q] + log_likelihood(x, 1);
                        cand[nact + q].parent = q; cand[nact + q].bit = 1;
                    }
                    int nc = 2 * nact;
                    /* stable sort, metric descending (:306-307): insertion sort
                     * keeps list order among equal keys, as list.sort does. */
                    for (int a = 0; a < nc; a++) {
                        cand_t c = cand[a]; int b = a;
                        while (b > 0 && srt[b - 1].m < c.m) { srt[b] = srt[b - 1]; b--; }
                        srt[b] = c;
                    }
                    int ns = nc < L ? nc : L;   /* :310-311 */
                    for (int r = 0; r < ns; r++) {   /* :323-339: full state copy */
                        const path_t *o = &cur[srt[r].parent];
                        memcpy(nxt[r].llr, o->llr, sizeof(double) * N);
                        memcpy(nxt[r].left, o->left, N);
                        memcpy(nxt[r].u, o->u, (size_t)i + 1);
                        memcpy(nxt[r].leaf, o->leaf, sizeof(double) * ((size_t)i + 1));
                        nxt[r].u[i] = (uint8_t)srt[r].bit;
                        pm2[r] = srt[r].m;
                        update_bits(&nxt[r], N, n, i, srt[r].bit, tmp);
                    }
                    for (int r = ns; r < L; r++) pm2[r] = -INFINITY;
                    { path_t *t = cur; cur = nxt; nxt = t; }
                    { double *t = pm; pm = pm2; pm2 = t; }
                    nact = ns;
                }
            }
            /* np.argmax(path_metrics) :259 -> lowest slot holding the maximum */
            int best = 0;
            for (int q = 1; q < L; q++) if (pm[q] > pm[best]) best = q;
            if (use_crc) {
                /* visit active slots by (metric desc, slot asc); first CRC pass wins */
                int *ord = (int *)malloc(sizeof(int) * L);
                for (int q = 0; q < nact; q++) {
                    int b = q;
                    while (b > 0 && pm[ord[b - 1]] < pm[q]) { ord[b] = ord[b - 1]; b--; }
                    ord[b] = q;
                }
                for (int r = 0; r < nact; r++) {
                    int K = 0;
                    for (int l = 0; l < N; l++)
                        if (!frozen_ref[l]) ibits[K++] = cur[ord[r]].u[bit_reverse(l, n)];
                    if (crc_check_bits(ibits, K, crc_poly, crc_len)) { best = ord[r]; break; }
                }
                free(ord);
            }
            for (int i = 0; i < N; i++) {
                int l = bit_reverse(i, n);
                u_ref_out[f * N + l] = cur[best].u[i];
                if (leaf_out) leaf_out[f * N + l] = cur[best].leaf[i];
            }
            if (pm_out) for (int q = 0; q < L; q++) pm_out[f * L + q] = pm[q];
        }
        for (int q = 0; q < L; q++) {
            free(cur[q].llr); free(cur[q].left); free(cur[q].u); free(cur[q].leaf);
            free(nxt[q].llr); free(nxt[q].left); free(nxt[q].u); free(nxt[q].leaf);
        }
        free(cur); free(nxt); free(pm); free(pm2); free(cand); free(srt);
        free(y); free(tmp); free(ibits);
    }
    return ORACLE_OK;
}

/* ---------------------------------------------------------------- LDPC -- */
/* numpy add.reduce over a contiguous float64 vector (pairwise_sum in
 * numpy/core/src/umath/loops_utils.h): n<8 sequential from 0.0; n<=128 eight
 * accumulators; else split.  ldpc/decoder.py:116 and :337 call np.sum on the
 * incoming messages, so this IS the reference's association order. */
static double np_sum(const double *a, int n)
{
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; i++) r += a[i];
        return r;
    } else if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; j++) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return np_sum(a, n2) + np_sum(a + n2, n - n2);
    }
}

static inline double clip1(double v) /* np.clip(v, -0.999999, 0.999999) */
{
    if (v < -0.999999) return -0.999999;
    if (v > 0.999999) return 0.999999;
    return v;
}

/*
 * mode 0: BPDecoder.decode (ldpc/decoder.py:124-202, check rule :62-96,
 *         variable rule :98-122).  mode 1: MSDecoder.decode (:289-352, check
 *         rule :257-287).
 * Edges e = 0..E-1 are in check-major order (np.nonzero(H), row-major), i.e.
 * check_neighbors[c] ascending v (:43-47).  vperm lists edge ids in
 * variable-major order (var_neighbors[v] ascending c).
 * bits_out[f*n + v], iters_out[f], total_out[f*n + v] (last iteration's totals).
 */
int oracle_ldpc(int mode, int m, int n, int E, const int32_t *cptr, const int32_t *col,
                const int32_t *vptr, const int32_t *vperm, double norm, int max_iter,
                int early_stop, const double *llr, int64_t F, uint8_t *bits_out,
                int32_t *iters_out, double *total_out, int nthreads)
{
    if (max_iter < 1) return ORACLE_EINVAL; /* reference: UnboundLocalError */
    if (mode == 1)
        for (int c = 0; c < m; c++)
            if (cptr[c + 1] - cptr[c] == 1) return ORACLE_EDEG1;
    int maxd = 1;
    for (int c = 0; c < m; c++) if (cptr[c + 1] - cptr[c] > maxd) maxd = cptr[c + 1] - cptr[c];
    for (int v = 0; v < n; v++) if (vptr[v + 1] - vptr[v] > maxd) maxd = vptr[v + 1] - vptr[v];
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        double *v2c = (double *)malloc(sizeof(double) * (E + 1));
        double *c2v = (double *)malloc(sizeof(double) * (E + 1));
        double *t = (double *)malloc(sizeof(double) * (maxd + 1));
        double *tot = (double *)malloc(sizeof(double) * n);
        uint8_t *dec = (uint8_t *)malloc(n);
#pragma omp for schedule(dynamic, 4)
        for (int64_t f = 0; f < F; f++) {
            const double *ch = llr + f * n;
            for (int e = 0; e < E; e++) v2c[e] = ch[col[e]];   /* :144-146 */
            int iters = max_iter;                               /* :149 */
            for (int it = 0; it < max_iter; it++) {
                for (int c = 0; c < m; c++) {
                    int e0 = cptr[c], d = cptr[c + 1] - e0;
                    if (mode == 0) {
                        for (int j = 0; j < d; j++) t[j] = clip1(tanh(v2c[e0 + j] / 2.0));
                        for (int i = 0; i < d; i++) {
                            double p = 1.0; int first = 1;    /* np.prod: left to right */
                            for (int j = 0; j < d; j++) {
                                if (j == i) continue;
                                if (first) { p = t[j]; first = 0; } else p *= t[j];
                            }
                            p = clip1(p);
                            double o = 2.0 * atanh(p);
                            if (isnan(o)) o = 0.0;            /* np.nan_to_num :94 */
                            else if (isinf(o)) o = o > 0 ? 20.0 : -20.0;
                            c2v[e0 + i] = o;
                        }
                    } else {
                        for (int i = 0; i < d; i++) {
                            double sp = 1.0, mn = INFINITY; int first = 1;
                            for (int j = 0; j < d; j++) {
                                if (j == i) continue;
                                double x = v2c[e0 + j];
                                double s = (double)((x > 0) - (x < 0));
                                if (first) { sp = s; first = 0; } else sp *= s;
                                double a = fabs(x);
                                if (a < mn) mn = a;
                            }
                            c2v[e0 + i] = sp * mn * norm;      /* :285 */
                        }
                    }
                }
                for (int v = 0; v < n; v++) {
                    int j0 = vptr[v], d = vptr[v + 1] - j0;
                    for (int j = 0; j < d; j++) t[j] = c2v[vperm[j0 + j]];
                    double total = ch[v] + np_sum(t, d);       /* :116 / :337 */
                    tot[v] = total;
                    for (int j = 0; j < d; j++) v2c[vperm[j0 + j]] = total - t[j];
                }
                for (int v = 0; v < n; v++) dec[v] = tot[v] <= 0;  /* :191 */
                if (early_stop) {                               /* :194-198 */
                    int ok = 1;
                    for (int c = 0; c < m && ok; c++) {
                        int par = 0;
                        for (int e = cptr[c]; e < cptr[c + 1]; e++) par ^= dec[col[e]];
                        if (par) ok = 0;
                    }
                    if (ok) { iters = it + 1; break; }
                }
            }
            memcpy(bits_out + f * n, dec, (size_t)n);
            if (iters_out) iters_out[f] = iters;
            if (total_out) memcpy(total_out + f * n, tot, sizeof(double) * n);
        }
        free(v2c); free(c2v); free(t); free(tot); free(dec);
    }
    return ORACLE_OK;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
