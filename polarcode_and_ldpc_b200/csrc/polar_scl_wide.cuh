// polar_scl_wide.cuh -- SCL polar decoder for list sizes above 32: one BLOCK per frame, one
// thread per list slot (LP = list size padded to a power of two, 64 .. 1024 threads).
//
// The reference accepts any list_size >= 1 (/root/reference/src/polar/decoder.py:194-196;
// benchmarks/sc_vs_scl.py sweeps it); the warp kernels (polar_scl.cuh, polar_scl_fast.cuh)
// stop at 32 slots, the width of a warp.  Same formulation as polar_scl.cuh (read that header
// first): decode step i <-> reference bit bit_reverse(i), level d = stage d-1, lazy path copy by
// per-level slot pointers, bit-packed partial sums, all-pairs rank for the prune with the
// reference's stable order (decoder.py:306-311), fp64 path metric.  What changes with the width:
//   * every LLR level 1 .. n-1 and every bit-packed left level lives in a global scratch
//     ([element][slot], so the LP threads of a block touch consecutive words); a list this wide
//     does not fit shared memory (L = 64, N = 1024: 256 KB of LLRs);
//   * a thread owns a whole path (no sub-lanes), so inside a level walk a path reads what it
//     wrote itself and the block meets only once per leaf (after the walk) and around the prune;
//   * the slot pointers are arrays in shared memory, double buffered: a survivor copies its
//     parent's column into the other buffer (a register word cannot hold n fields of 6-10 bits).
#pragma once
#include "polar_scl.cuh"

// shared memory of one block (bytes); the kernel carves the same regions in the same order
PCL_HOST_DEVICE int pcl_wide_smem_bytes(int LP, int n, int nb)
{
    int off = 0;
    off += 2 * LP * 8;                 // cm: candidate metrics
    off += LP * 8;                     // newpm
    off += LP * 4;                     // sel
    off += LP * 4;                     // smallx
    off += LP * 4;                     // ulx / pass flags
    off += 2 * (n > 1 ? n - 1 : 1) * LP * 2;    // ptrL[2][n-1][LP]
    off += 2 * (nb > 0 ? nb : 1) * LP * 2;      // ptrB[2][nb][LP]
    off = (off + 15) / 16 * 16;
    off += 16;                                  // ctl: vote word of the in-place test
    return off;
}

// global scratch of one resident block, in bytes: LLR levels + big left levels + u words of every slot
PCL_HOST_DEVICE int64_t pcl_wide_scratch_bytes(int LP, int N, int rsz)
{
    const int64_t NW = N >= 32 ? N / 32 : 1;
    return (int64_t)LP * N * rsz + (int64_t)LP * NW * 4 * 2;
}

template <typename real>
__global__ void __launch_bounds__(1024) polar_scl_wide_kernel(PolarParams<real> P)
{
    const PolarLayout& Y = P.lay;
    const int N = Y.N, n = Y.n, K = Y.K, L = Y.L, NW = Y.NW, nb = Y.nb;
    const int LP = (int)blockDim.x;
    const int PB = __ffs(LP) - 1;
    const int p = (int)threadIdx.x;
    const int shift = (N < 32) ? 32 - N : 0;
    const int nl = n > 1 ? n - 1 : 1, nbb = nb > 0 ? nb : 1;

    unsigned char* sm = pcl_dyn_smem();
    double* cm = (double*)sm;                      sm += 2 * LP * 8;
    double* newpm = (double*)sm;                   sm += LP * 8;
    int* sel = (int*)sm;                           sm += LP * 4;
    uint32_t* smallx = (uint32_t*)sm;              sm += LP * 4;
    uint32_t* ulx = (uint32_t*)sm;                 sm += LP * 4;
    uint16_t* ptrL = (uint16_t*)sm;                sm += 2 * nl * LP * 2;
    uint16_t* ptrB = (uint16_t*)sm;                sm += 2 * nbb * LP * 2;
    volatile int* ctl = (volatile int*)(pcl_dyn_smem() + pcl_wide_smem_bytes(LP, n, nb) - 16);

    unsigned char* gs = (unsigned char*)P.scratch + (int64_t)blockIdx.x * pcl_wide_scratch_bytes(LP, N, (int)sizeof(real));
    real* gl = (real*)gs;                                              // level d at LP * (N - (N >> (d - 1)))
    uint32_t* bw = (uint32_t*)(gs + (int64_t)LP * N * sizeof(real));   // big left level l at LP * ((N >> 5) - (N >> (l + 4)))
    uint32_t* Ug = bw + (int64_t)LP * NW;                              // u words [slot][NW]

    const real RINF = pcl_math<real>::inf();
    const double NEG_INF = -(double)RINF;

    for (int64_t f = blockIdx.x; f < P.F; f += gridDim.x) {
        const real* y = P.llr + f * N;
        int nact = 1;
        bool act = (p == 0);
        double pm = act ? 0.0 : NEG_INF;
        uint32_t small = 0, ulast = 0, fw = 0;
        int cur = 0;                               // which pointer buffer is live
        uint16_t* PL = ptrL;
        uint16_t* PBp = ptrB;

        for (int i = 0; i < N; i++) {
            if (((i + shift) & 31) == 0 || i == 0) fw = P.frozen_words[(i + shift) >> 5];
            const bool frozen = (fw >> ((i + shift) & 31)) & 1u;

            // ---- LLR levels start..n (decoder.py:341-356) -------------------------------------
            const int start = (i == 0) ? 1 : n - (__ffs(i) - 1);
            real x = 0;
            if (act) {
                for (int d = start; d <= n; d++) {
                    const int sz = N >> d;
                    const int bit = (i >> (n - d)) & 1;
                    // source and destination are different levels of the scratch: the loads of the next elements
                    // need not wait for the stores of the previous ones
                    const real* __restrict__ src = nullptr;
                    int q = 0;
                    if (d > 1) {
                        q = PL[(d - 2) * LP + p];
                        src = gl + (int64_t)LP * (N - (N >> (d - 2)));
                    }
                    real* __restrict__ dst = (d < n) ? gl + (int64_t)LP * (N - (N >> (d - 1))) : nullptr;
                    int qb = 0;
                    const uint32_t* bsrc = nullptr;
                    if (bit && d <= nb) {
                        qb = PBp[(d - 1) * LP + p];
                        bsrc = bw + (int64_t)LP * ((N >> 5) - (N >> (d + 4)));
                    }
#pragma unroll 4
                    for (int k = 0; k < sz; k++) {
                        real a, b;
                        if (d == 1) {
                            const int r = (int)(__brev((unsigned)k) >> (32 - n));
                            a = y[r];
                            b = y[r + 1];
                        } else {
                            a = src[(int64_t)k * LP + q];
                            b = src[(int64_t)(k + sz) * LP + q];
                        }
                        real v;
                        if (bit) {
                            uint32_t ub;
                            if (d <= nb) ub = (bsrc[(int64_t)(k >> 5) * LP + qb] >> (k & 31)) & 1u;
                            else ub = (small >> (32 - 2 * sz + k)) & 1u;
                            v = ub ? b - a : b + a;            // decoder.py:141-144
                        } else {
                            v = pcl_math<real>::f(a, b);       // decoder.py:127
                        }
                        if (d < n) dst[(int64_t)k * LP + p] = v; else x = v;
                    }
                    if (d < n) PL[(d - 1) * LP + p] = (uint16_t)p;   // own slot from here on (read back by this thread only)
                }
            }
            // every path has read its (possibly borrowed) source arrays: the owners may overwrite
            // them in a later step
            __syncthreads();

            // ---- leaf decision -----------------------------------------------------------------
            const real ax = fabs(x);
            const bool hard = !(x >= (real)0);                 // decoder.py:117-119
            uint32_t u = 0;
            int parent = p;
            if (frozen) {                                      // decoder.py:264-281
                if (act) {
                    const double sp = (double)pcl_math<real>::softplus_neg_abs(ax);
                    pm -= (hard ? (double)ax : 0.0) + sp;
                }
            } else {                                           // decoder.py:283-339
                double m0 = NEG_INF, m1 = NEG_INF;
                if (act) {
                    const double sp = (double)pcl_math<real>::softplus_neg_abs(ax);
                    m0 = pm - ((hard ? (double)ax : 0.0) + sp);
                    m1 = pm - ((hard ? 0.0 : (double)ax) + sp);
                }
                cm[p] = m0;
                cm[LP + p] = m1;
                smallx[p] = small;
                // Reliable bit on a full list (the common case): the likely candidates are still in slot order
                // (strictly, so no tie rule is involved) and every unlikely one lies strictly below the last of
                // them -> the ranks are the slots, every path continues in place with its likely bit: one vote
                // instead of the all-pairs ranking, no pointer columns to copy.  Same survivors and slots.
                const bool full = nact >= L;
                const double mlik = hard ? m1 : m0, munl = hard ? m0 : m1;
                if (full) {
                    newpm[p] = mlik;
                    if (p == 0) ctl[0] = 1;
                }
                __syncthreads();
                const int ns = (2 * nact < L) ? 2 * nact : L;
                bool in_place = false;
                if (full) {
                    if (p < L && !((p == 0 || newpm[p - 1] > mlik) && munl < newpm[L - 1])) ctl[0] = 0;
                    __syncthreads();
                    in_place = ctl[0] != 0;
                }
                if (in_place) {
                    if (act) { pm = mlik; u = hard ? 1u : 0u; }
                } else {
                if (act) {
                    // rank of this path's two candidates among the 2 nact live ones: (metric desc, bit asc,
                    // parent asc); an inactive slot holds -inf and can never be ahead of a live candidate
                    int r0 = 0, r1 = 0;
                    for (int j = 0; j < nact; j++) {
                        const double a0 = cm[j], a1 = cm[LP + j];
                        r0 += (a0 > m0) || (a0 == m0 && j < p);
                        r0 += (a1 > m0);
                        r1 += (a0 > m1) || (a0 == m1);
                        r1 += (a1 > m1) || (a1 == m1 && j < p);
                    }
                    if (r0 < ns) { sel[r0] = p; newpm[r0] = m0; }
                    if (r1 < ns) { sel[r1] = LP + p; newpm[r1] = m1; }
                }
                __syncthreads();
                act = p < ns;
                uint16_t* PLn = ptrL + (cur ^ 1) * nl * LP;
                uint16_t* PBn = ptrB + (cur ^ 1) * nbb * LP;
                if (act) {
                    const int c = sel[p];
                    parent = c & (LP - 1);
                    u = (uint32_t)(c >> PB);
                    pm = newpm[p];
                    small = smallx[parent];
                    for (int d = 0; d < n - 1; d++) PLn[d * LP + p] = PL[d * LP + parent];
                    for (int l = 0; l < nb; l++) PBn[l * LP + p] = PBp[l * LP + parent];
                } else {
                    pm = NEG_INF;
                }
                nact = ns;
                cur ^= 1;
                PL = PLn;
                PBp = PBn;
                __syncthreads();
                }
            }
            if (P.dbg_leaf != nullptr) {
                P.dbg_leaf[(f * N + i) * LP + p] = x;
                P.dbg_parent[(f * N + i) * LP + p] = (uint8_t)(parent & 0xff);
            }

            // ---- partial sums (decoder.py:358-372) ----------------------------------------------
            if (i == N - 1) {
                ulast = u;
            } else if ((i & 1) == 0) {
                small = pcl_bfi(small, u, 30, 1);
            } else {
                uint32_t c = u;
                int s = 1, t = i;
                while ((t & 1) && s < 32) {
                    const uint32_t left = pcl_bfe(small, 32 - 2 * s, s);
                    c = (left ^ c) | (c << s);
                    s <<= 1;
                    t >>= 1;
                }
                if (!(t & 1)) {
                    if (s < 32) {
                        small = pcl_bfi(small, c, 32 - 2 * s, s);
                    } else if (act) {           // one full word: level n-5
                        const int d = n - 5;
                        bw[(int64_t)LP * ((N >> 5) - (N >> (d + 4))) + p] = c;
                        PBp[(d - 1) * LP + p] = (uint16_t)p;
                    }
                } else if (act) {               // keep folding word-wise into the own slot of level d
                    const int cto = __ffs(~i) - 1;
                    const int d = n - cto;
                    const int Wd = N >> (d + 5);
                    uint32_t* dest = bw + (int64_t)LP * ((N >> 5) - (N >> (d + 4)));
                    dest[(int64_t)(Wd - 1) * LP + p] = c;
                    for (int l = n - 5; l > d; l--) {
                        const int w = N >> (l + 5);
                        const int ql = PBp[(l - 1) * LP + p];
                        const uint32_t* lsrc = bw + (int64_t)LP * ((N >> 5) - (N >> (l + 4)));
                        for (int j = 0; j < w; j++)
                            dest[(int64_t)(Wd - 2 * w + j) * LP + p] =
                                lsrc[(int64_t)j * LP + ql] ^ dest[(int64_t)(Wd - w + j) * LP + p];
                    }
                    PBp[(d - 1) * LP + p] = (uint16_t)p;
                }
            }
        }

        // ---- final selection (decoder.py:259-262) ------------------------------------------------
        newpm[p] = pm;
        smallx[p] = small;
        ulx[p] = ulast;
        __syncthreads();
        int best = 0;
        {
            double bm = newpm[0];
            for (int q = 1; q < LP; q++) {
                const double v = newpm[q];
                if (v > bm) { bm = v; best = q; }           // first maximum, like np.argmax
            }
        }
        if (P.pm_out != nullptr && p < L) P.pm_out[f * L + p] = pm;

        // u words: every live slot computes its own (CRC selection needs them all; otherwise only `best` is read)
        if (act && (P.use_crc || p == best)) {
            uint32_t* U = Ug + (int64_t)p * NW;
            for (int w = 0; w < NW; w++) {
                uint32_t v;
                if (w == NW - 1) {
                    v = pcl_bfi(small, ulast, 31, 1);
                    v ^= (v >> 1) & 0x15555555u;
                    v ^= (v >> 2) & 0x03333333u;
                    v ^= (v >> 4) & 0x000F0F0Fu;
                    v ^= (v >> 8) & 0x000000FFu;
                } else {
                    const int r = NW - w;                     // 2 .. NW
                    const int Wl = 1 << (31 - __clz(r - 1));  // block words: Wl < r <= 2 Wl
                    const int l = (31 - __clz(NW)) - (31 - __clz(Wl));
                    const int j = w - (NW - 2 * Wl);
                    v = bw[(int64_t)LP * ((N >> 5) - (N >> (l + 4))) + (int64_t)j * LP + PBp[(l - 1) * LP + p]];
                    v ^= (v >> 1) & 0x55555555u;
                    v ^= (v >> 2) & 0x33333333u;
                    v ^= (v >> 4) & 0x0F0F0F0Fu;
                    v ^= (v >> 8) & 0x00FF00FFu;
                    v ^= (v >> 16) & 0x0000FFFFu;
                }
                U[w] = v;
            }
            for (int t = 1; t < NW; t <<= 1)                  // strides of 32 t bits
                for (int w = 0; w < NW - 1; w++) {
                    const int r = NW - w;
                    const int Wl = 1 << (31 - __clz(r - 1));
                    const int j = w - (NW - 2 * Wl);
                    if (t < Wl && (j & t) == 0) U[w] ^= U[w + t];
                }
        }

        if (P.use_crc) {
            // first path in (metric desc, slot asc) order whose info bits pass the CRC register test
            // (src/polar/utils.py:128-163); else the best metric
            uint32_t pass = 0;
            if (act) {
                const uint32_t* U = Ug + (int64_t)p * NW;
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
                pass = (reg == 0) ? 1u : 0u;
            }
            ulx[p] = pass;
            __syncthreads();
            int bsel = -1;
            double bm = 0;
            for (int q = 0; q < nact; q++) {
                if (!ulx[q]) continue;
                const double v = newpm[q];
                if (bsel < 0 || v > bm) { bm = v; bsel = q; }
            }
            if (bsel >= 0) best = bsel;
        }
        __syncthreads();                                      // the chosen slot's u words are visible to the block

        {   // decoded = u[info_bits] (decoder.py:260-262)
            const uint32_t* U = Ug + (int64_t)best * NW;
            uint8_t* out = P.bits + f * K;
            for (int k = p; k < K; k += LP) {
                const int pos = (int)P.info_pos[k] + shift;
                out[k] = (uint8_t)((U[pos >> 5] >> (pos & 31)) & 1u);
            }
        }
        __syncthreads();
    }
}
