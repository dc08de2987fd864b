// polar_scl.cuh -- batched SC / SCL polar decoder, one warp per frame.
//
// Replaces SCDecoder.decode (/root/reference/src/polar/decoder.py:38-71) and
// SCLDecoder.decode (:225-262, with _decode_frozen_bit :264-281,
// _decode_info_bit :283-339, _log_likelihood :374-406) for a whole batch.
//
// Formulation (validated against the reference's own outputs, see tests/):
// decode step i <-> reference bit index l = bit_reverse(i); level d = stage s+1;
// in decode-step order every node's halves are contiguous, so the reference's
// stage-s butterfly at distance 2^s becomes dst[k] = f/g(src[k], src[k+sz]) with
// sz = N >> d.  Level 0 (the channel LLRs) is read straight from the input with
// the bit-reversed address, so no permuted copy is ever materialised:
// y[k] = llr[br(k)], y[k + N/2] = llr[br(k) + 1].
//
// Mapping.  LP = list size padded to a power of two (<= 32).  Lane = kk*LP + p:
// p = list slot, kk = sub-lane (S = 32/LP lanes cooperate on one path).  Level d
// of all slots lives in one array laid out [k][slot], so a warp access touches
// 32 consecutive words -> bank-conflict free in shared memory, fully coalesced
// for the top G levels that live in an L2-resident global scratch (they are
// touched 2^d times per frame only; keeping them out of shared memory is what
// buys occupancy).
//
// Lazy path copy.  The reference copies the whole (N, n+1) LLR and bit matrices
// of every survivor (decoder.py:314-329).  Here a survivor inherits two packed
// words of per-level slot pointers from its parent with one shuffle each.  All
// live paths recompute level d in the same step, so a path can always write its
// own slot r at level d and point field d at r: no reference counts, no copies.
// Partial sums ("left" arrays, decoder.py:96-115) for node sizes >= 32 are
// bit-packed words handled the same way; the sizes 16..1 share one 32-bit
// register per path laid out at their natural bit position (size s at bit
// 32-2s), so a path's codeword estimate is simply the concatenation of its left
// arrays and u = x * F^{(x)n} is recovered at the end with a bit-parallel
// butterfly -- no per-path decision history is stored or copied.
//
// Path metric: PM += ll(x,u), ll = -([u != hard(x)]*|x| + log1p(exp(-|x|)))
// (decoder.py:391-406), accumulated in fp64 for both compute types.  Prune:
// candidates (PM_p + ll(x_p,b), p, b), stable descending sort == total order
// (metric desc, bit asc, parent asc) (decoder.py:306-311) evaluated as an
// all-pairs rank; survivor of rank r takes slot r (decoder.py:323-339).
#pragma once
#include "pcl_common.cuh"

struct PolarLayout {
    int N, n, K, L;      // code length, log2 N, info bits, list size (L <= LP)
    int G;               // levels 1..G of the LLR tree live in global scratch
    int NW;              // max(1, N/32) words per codeword estimate
    int nb;              // number of bit-packed "big" left levels = max(0, n-5)
    // per-warp shared memory byte offsets
    int off_cm, off_newpm, off_sel, off_llr, off_bw, off_uw, warp_bytes;
    int uw_slots;        // 1, or LP when CRC selection needs every path's u
    int hdr_bytes;       // TM variant: per-block header (TMEM base, group tickets) ahead of the warps' regions
    int64_t scratch_per_warp;  // reals of global scratch per resident warp
};

template <typename real>
struct PolarParams {
    PolarLayout lay;
    const real* llr;               // [F][N] channel LLRs, reference index order
    uint8_t* bits;                 // [F][K] decoded info bits, ascending reference index
    double* pm_out;                // [F][L] final path metrics (optional)
    real* dbg_leaf;                // [F][N][LP] leaf LLR of every slot at step i (optional)
    uint8_t* dbg_parent;           // [F][N][LP] parent slot chosen at step i (optional)
    const uint32_t* frozen_words;  // decode-step-order frozen mask, bit (i + shift)
    const uint16_t* info_pos;      // [K] decode step of the k-th info bit
    real* scratch;                 // global scratch for levels 1..G
    int64_t F;
    unsigned long long* next;      // TM variant: ticket counter the groups of warps pull frame chunks from
    unsigned long long ticket_base;  // value of *next when this launch started (the counter is never reset)
    int want_pm;                   // L == 1: maintain the metric only when asked
    int use_crc, crc_len;
    uint32_t crc_poly;
};

template <int LP> struct pcl_log2 { static const int v = 1 + pcl_log2<LP / 2>::v; };
template <> struct pcl_log2<1> { static const int v = 0; };

template <int LP, typename real>
__global__ void __launch_bounds__(128) polar_scl_kernel(PolarParams<real> P)
{
    constexpr int PB = pcl_log2<LP>::v;
    constexpr int S = 32 / LP;
    const PolarLayout& Y = P.lay;
    const int N = Y.N, n = Y.n, K = Y.K, L = Y.L, G = Y.G, NW = Y.NW, nb = Y.nb;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int p = lane & (LP - 1);
    const int kk = lane >> PB;
    const int shift = (N < 32) ? 32 - N : 0;

    unsigned char* wsm = pcl_dyn_smem() + (size_t)warp * Y.warp_bytes;
    double* cm = (double*)(wsm + Y.off_cm);
    double* newpm = (double*)(wsm + Y.off_newpm);
    int* sel = (int*)(wsm + Y.off_sel);
    real* sl = (real*)(wsm + Y.off_llr);          // levels G+1 .. n-1
    uint32_t* bw = (uint32_t*)(wsm + Y.off_bw);   // big left levels 1 .. nb
    uint32_t* uw = (uint32_t*)(wsm + Y.off_uw);   // final u words
    real* gl = P.scratch + (int64_t)(blockIdx.x * wpb + warp) * Y.scratch_per_warp;

    const real RINF = pcl_math<real>::inf();
    const double NEG_INF = -(double)RINF;

    for (int64_t f = (int64_t)blockIdx.x * wpb + warp; f < P.F; f += (int64_t)gridDim.x * wpb) {
        const real* y = P.llr + f * N;
        int nact = 1;
        bool act = (p == 0);
        double pm = act ? 0.0 : NEG_INF;
        uint64_t ptrL = 0, ptrB = 0;
        uint32_t small = 0, ulast = 0;
        uint32_t fw = 0;

        for (int i = 0; i < N; i++) {
            if (((i + shift) & 31) == 0 || i == 0) fw = P.frozen_words[(i + shift) >> 5];
            const bool frozen = (fw >> ((i + shift) & 31)) & 1u;

            // ---- LLR levels start..n (decoder.py:73-94 / :341-356) ------------
            const int start = (i == 0) ? 1 : n - (__ffs(i) - 1);
            real x = 0;
            for (int d = start; d <= n; d++) {
                const int sz = N >> d;
                const int bit = (i >> (n - d)) & 1;
                const real* src = nullptr;
                int q = 0;
                if (d > 1) {
                    q = pcl_get_field<PB>(ptrL, d - 2);
                    src = (d - 1 <= G) ? gl + (int64_t)LP * (N - (N >> (d - 2)))
                                       : sl + LP * ((N >> G) - (N >> (d - 2)));
                }
                real* dst = nullptr;
                if (d < n)
                    dst = (d <= G) ? gl + (int64_t)LP * (N - (N >> (d - 1)))
                                   : sl + LP * ((N >> G) - (N >> (d - 1)));
                int qb = 0;
                const uint32_t* bsrc = nullptr;
                if (bit && d <= nb) {
                    qb = pcl_get_field<PB>(ptrB, d - 1);
                    bsrc = bw + LP * ((N >> 5) - (N >> (d + 4)));
                }
                if (act) {
                    for (int k = kk; k < sz; k += S) {
                        real a, b;
                        if (d == 1) {
                            int r = (int)(__brev((unsigned)k) >> (32 - n));
                            a = y[r];
                            b = y[r + 1];
                        } else {
                            a = src[k * LP + q];
                            b = src[(k + sz) * LP + q];
                        }
                        real v;
                        if (bit) {
                            uint32_t ub;
                            if (d <= nb) ub = (bsrc[(k >> 5) * LP + qb] >> (k & 31)) & 1u;
                            else ub = (small >> (32 - 2 * sz + k)) & 1u;
                            v = ub ? b - a : b + a;            // decoder.py:141-144
                        } else {
                            v = pcl_math<real>::f(a, b);       // decoder.py:127
                        }
                        if (d < n) dst[k * LP + p] = v; else x = v;
                    }
                }
                if (d < n) {
                    if (act) ptrL = pcl_set_field<PB>(ptrL, d - 1, p);
                    __syncwarp();
                }
            }
            // Every path has now read its (possibly borrowed) source arrays; order those
            // reads before the owners overwrite them at a later step.
            __syncwarp();
            if (S > 1) x = __shfl_sync(PCL_FULL_MASK, x, p);   // lane p holds kk == 0

            // ---- leaf decision ------------------------------------------------
            const real ax = fabs(x);
            const bool hard = !(x >= (real)0);                 // decoder.py:117-119
            uint32_t u = 0;
            int parent = p;
            if (LP == 1) {
                u = frozen ? 0u : (hard ? 1u : 0u);
                if (P.want_pm) {
                    double sp = (double)pcl_math<real>::softplus_neg_abs(ax);
                    pm -= ((u != (uint32_t)hard) ? (double)ax : 0.0) + sp;
                }
            } else if (frozen) {                               // decoder.py:264-281
                if (act) {
                    double sp = (double)pcl_math<real>::softplus_neg_abs(ax);
                    pm -= (hard ? (double)ax : 0.0) + sp;
                }
            } else {                                           // decoder.py:283-339
                double m0 = NEG_INF, m1 = NEG_INF;
                if (act) {
                    double sp = (double)pcl_math<real>::softplus_neg_abs(ax);
                    m0 = pm - ((hard ? (double)ax : 0.0) + sp);
                    m1 = pm - ((hard ? 0.0 : (double)ax) + sp);
                }
                if (kk == 0) { cm[p] = m0; cm[LP + p] = m1; }
                __syncwarp();
                const int ns = (2 * nact < L) ? 2 * nact : L;
                for (int c = lane; c < 2 * LP; c += 32) {
                    const double mc = cm[c];
                    int rank = 0;
                    for (int j = 0; j < 2 * LP; j++) {
                        const double mj = cm[j];
                        rank += (mj > mc) || (mj == mc && j < c);
                    }
                    if (rank < ns) { sel[rank] = c; newpm[rank] = mc; }
                }
                __syncwarp();
                act = p < ns;
                if (act) {
                    int c = sel[p];
                    parent = c & (LP - 1);
                    u = (uint32_t)(c >> PB);
                    pm = newpm[p];
                } else {
                    pm = NEG_INF;
                }
                nact = ns;
                const int srcl = (lane & ~(LP - 1)) | parent;
                ptrL = pcl_shfl_u64(ptrL, srcl);
                ptrB = pcl_shfl_u64(ptrB, srcl);
                small = __shfl_sync(PCL_FULL_MASK, small, srcl);
                __syncwarp();
            }
            if (P.dbg_leaf != nullptr && kk == 0) {
                P.dbg_leaf[(f * N + i) * LP + p] = x;
                P.dbg_parent[(f * N + i) * LP + p] = (uint8_t)parent;
            }

            // ---- partial sums (decoder.py:96-115 / :358-372) --------------------
            if (i == N - 1) {
                ulast = u;
            } else if ((i & 1) == 0) {
                small = pcl_bfi(small, u, 30, 1);
            } else {
                uint32_t c = u;
                int s = 1, t = i;
                while ((t & 1) && s < 32) {
                    uint32_t left = pcl_bfe(small, 32 - 2 * s, s);
                    c = (left ^ c) | (c << s);
                    s <<= 1;
                    t >>= 1;
                }
                if (!(t & 1)) {
                    if (s < 32) {
                        small = pcl_bfi(small, c, 32 - 2 * s, s);
                    } else {                    // one full word: level n-5
                        const int d = n - 5;
                        if (act) {
                            if (kk == 0) bw[LP * ((N >> 5) - (N >> (d + 4))) + p] = c;
                            ptrB = pcl_set_field<PB>(ptrB, d - 1, p);
                        }
                        __syncwarp();
                    }
                } else {                        // keep folding word-wise, in place
                    const int cto = __ffs(~i) - 1;          // trailing ones of i
                    const int d = n - cto;                  // level of the left child reached
                    const int Wd = N >> (d + 5);
                    uint32_t* dest = bw + LP * ((N >> 5) - (N >> (d + 4)));
                    if (act && kk == 0) dest[(Wd - 1) * LP + p] = c;
                    __syncwarp();
                    for (int l = n - 5; l > d; l--) {
                        const int w = N >> (l + 5);
                        const int ql = pcl_get_field<PB>(ptrB, l - 1);
                        const uint32_t* lsrc = bw + LP * ((N >> 5) - (N >> (l + 4)));
                        if (act)
                            for (int j = kk; j < w; j += S)
                                dest[(Wd - 2 * w + j) * LP + p] =
                                    lsrc[j * LP + ql] ^ dest[(Wd - w + j) * LP + p];
                        __syncwarp();
                    }
                    if (act) ptrB = pcl_set_field<PB>(ptrB, d - 1, p);
                }
            }
        }

        // ---- final selection (decoder.py:259-262) -------------------------------
        int best = 0;
        if (LP > 1) {
            if (kk == 0) newpm[p] = pm;
            __syncwarp();
            double bm = newpm[0];
            for (int q = 1; q < LP; q++) {
                double v = newpm[q];
                if (v > bm) { bm = v; best = q; }           // first maximum, like np.argmax
            }
        }
        if (P.pm_out != nullptr && kk == 0 && p < L) P.pm_out[f * L + p] = pm;

        // u words of the slots we need: concatenated left arrays + butterfly.
        const int nslots = P.use_crc ? nact : 1;
        for (int sidx = 0; sidx < nslots; sidx++) {
            const int slot = P.use_crc ? sidx : best;
            uint32_t* U = uw + (P.use_crc ? sidx * NW : 0);
            const uint64_t pB = pcl_shfl_u64(ptrB, slot);
            const uint32_t sm = __shfl_sync(PCL_FULL_MASK, small, slot);
            const uint32_t ul = __shfl_sync(PCL_FULL_MASK, ulast, slot);
            for (int w = lane; w < NW; w += 32) {
                uint32_t v;
                if (w == NW - 1) {
                    v = pcl_bfi(sm, ul, 31, 1);
                    // sizes 16..1 share this word: stride t only inside blocks >= 2t
                    v ^= (v >> 1) & 0x15555555u;
                    v ^= (v >> 2) & 0x03333333u;
                    v ^= (v >> 4) & 0x000F0F0Fu;
                    v ^= (v >> 8) & 0x000000FFu;
                } else {
                    const int r = NW - w;                     // 2 .. NW
                    const int Wl = 1 << (31 - __clz(r - 1));  // block words: Wl < r <= 2 Wl
                    const int l = (31 - __clz(NW)) - (31 - __clz(Wl));
                    const int j = w - (NW - 2 * Wl);
                    v = bw[LP * ((N >> 5) - (N >> (l + 4))) + j * LP + pcl_get_field<PB>(pB, l - 1)];
                    v ^= (v >> 1) & 0x55555555u;
                    v ^= (v >> 2) & 0x33333333u;
                    v ^= (v >> 4) & 0x0F0F0F0Fu;
                    v ^= (v >> 8) & 0x00FF00FFu;
                    v ^= (v >> 16) & 0x0000FFFFu;
                }
                U[w] = v;
            }
            __syncwarp();
            for (int t = 1; t < NW; t <<= 1) {                // strides of 32 t bits
                for (int w = lane; w < NW - 1; w += 32) {
                    const int r = NW - w;
                    const int Wl = 1 << (31 - __clz(r - 1));
                    const int j = w - (NW - 2 * Wl);
                    if (t < Wl && (j & t) == 0) U[w] ^= U[w + t];
                }
                __syncwarp();
            }
        }

        if (P.use_crc) {
            // first path in (metric desc, slot asc) order whose info bits pass the
            // CRC register test (src/polar/utils.py:128-163); else the best metric.
            bool pass = false;
            if (kk == 0 && p < nact) {
                const uint32_t* U = uw + p * NW;
                const uint32_t top = 1u << (P.crc_len - 1);
                const uint32_t msk = (P.crc_len >= 32) ? 0xffffffffu : ((1u << P.crc_len) - 1u);
                uint32_t reg = 0;
                for (int k = 0; k < K; k++) {
                    const int pos = (int)P.info_pos[k] + shift;
                    const uint32_t b = (U[pos >> 5] >> (pos & 31)) & 1u;
                    reg ^= b << (P.crc_len - 1);
                    reg = (reg & top) ? ((reg << 1) ^ P.crc_poly) : (reg << 1);
                    reg &= msk;
                }
                pass = (reg == 0);
            }
            const unsigned pmask = __ballot_sync(PCL_FULL_MASK, pass);
            if (pmask != 0) {
                int bsel = -1;
                double bm = 0;
                for (int q = 0; q < nact; q++) {
                    if (!((pmask >> q) & 1u)) continue;
                    double v = newpm[q];
                    if (bsel < 0 || v > bm) { bm = v; bsel = q; }
                }
                best = bsel;
            }
        }

        {   // decoded = u[info_bits] (decoder.py:70-71 / :260-262)
            const uint32_t* U = uw + (P.use_crc ? best * NW : 0);
            uint8_t* out = P.bits + f * K;
            for (int k = lane; k < K; k += 32) {
                const int pos = (int)P.info_pos[k] + shift;
                out[k] = (uint8_t)((U[pos >> 5] >> (pos & 31)) & 1u);
            }
        }
        __syncwarp();
    }
}
