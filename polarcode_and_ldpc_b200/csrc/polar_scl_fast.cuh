// polar_scl_fast.cuh -- SC / SCL decoder with a register-resident tree bottom and several
// frames per warp.  Production kernel for N >= 16 (polar_scl.cuh is the fallback).
//
// Same algorithm, list semantics and pointer scheme as polar_scl.cuh (read that header
// first).  What changed, each step driven by an ncu capture (profiles/r01a .. r01h):
//   * The generic kernel was issue-bound (515 k warp instructions per N=1024 L=8 frame):
//     7 of every 8 level visits touch a node of <= 4 LLRs per path and paid loop, address and
//     packed-pointer arithmetic plus a shared-memory round trip each.  Here the tree is cut at
//     height 3: the decoder walks N/8 blocks of 8 leaves; the block root (8 LLRs per path) is
//     produced straight into registers and the levels of size 4, 2, 1 below it stay in
//     registers.  The leaf loop is rolled (an unrolled body overflowed the instruction cache);
//     f / g per stage is still a compile-time choice behind three uniform branches on j.
//   * S lanes own a path and a warp decodes FPW = 32 / (LP * S) frames side by side
//     (default S = 1: a lane owns a whole path, 4 frames per warp at L = 8).  All frames of a
//     warp follow the same schedule (frozen pattern, number of live paths), so every branch is
//     warp-uniform and the per-leaf bookkeeping instructions are shared by FPW frames.
//   * A surviving path inherits its parent's live registers with shuffles (only those a later
//     leaf of the block still reads) together with two packed 32-bit pointer words (LLR levels,
//     partial-sum levels) and the 32-bit small partial-sum word.
//   * Level 1 is never stored: level 2 recomputes its two level-1 operands from ONE aligned
//     4-element channel load, walked in bit-reversed order so the loads are sequential.
//     Levels 2..G live in an L2/HBM scratch and are streamed in software-pipelined batches of
//     UNR independent row loads; levels G+1..n-4 stay in shared memory ([k][column] layout:
//     conflict free / fully coalesced).
//   * Prune: all-pairs rank.  fp32 build: the candidate id replaces the lowest mantissa bits of
//     the fp64 metric, so one DSETP + one predicated add per pair gives both the metric order
//     and the reference's tie order; fp64 build keeps the exact two-key comparison.
// Path metric in fp64; fp32 build evaluates log1p(exp(-|x|)) as 2 atanh(u / (2 + u)) with
// u = 2^(-|x| log2 e) (one ex2, one rcp, 6 FMA; |error| < 1e-7).
#pragma once
#include "pcl_common.cuh"
#include "polar_scl.cuh"

// ---- fast softplus(-|x|) ------------------------------------------------------------
PCL_DEVICE float pcl_ex2f(float x)
{
#ifdef PCL_EMU
    return exp2f(x);
#else
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}
PCL_DEVICE float pcl_rcpf(float x)
{
#ifdef PCL_EMU
    return 1.0f / x;
#else
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}

template <typename real> struct pcl_fast;
template <> struct pcl_fast<float> {
    static PCL_DEVICE float softplus_neg_abs(float ax)
    {
        const float u = pcl_ex2f(ax * -1.4426950408889634f);      // exp(-|x|) in (0, 1]
        const float s = u * pcl_rcpf(2.0f + u);                   // log1p(u) = 2 atanh(s), s <= 1/3
        const float s2 = s * s;
        float pl = fmaf(s2, 0.07692308f, 0.09090909f);
        pl = fmaf(s2, pl, 0.11111111f);
        pl = fmaf(s2, pl, 0.14285715f);
        pl = fmaf(s2, pl, 0.2f);
        pl = fmaf(s2, pl, 0.33333334f);
        pl = fmaf(s2, pl, 1.0f);
        return 2.0f * s * pl;
    }
    // b + (bit ? -a : a) with the sign flip done on the bit pattern
    static PCL_DEVICE float g(float a, float b, uint32_t bit)
    {
        return b + __uint_as_float(__float_as_uint(a) ^ (bit << 31));
    }
};
template <> struct pcl_fast<double> {
    static PCL_DEVICE double softplus_neg_abs(double ax) { return log1p(exp(-ax)); }
    static PCL_DEVICE double g(double a, double b, uint32_t bit) { return bit ? b - a : b + a; }
};

template <typename real>
PCL_DEVICE real pcl_shfl_real(real v, int src);
template <>
PCL_DEVICE float pcl_shfl_real<float>(float v, int src) { return __shfl_sync(PCL_FULL_MASK, v, src); }
template <>
PCL_DEVICE double pcl_shfl_real<double>(double v, int src)
{
    return __longlong_as_double((long long)pcl_shfl_u64((uint64_t)__double_as_longlong(v), src));
}

// One butterfly stage inside the register-resident block: level of size Z from the level
// of size 2Z.  S lanes own a path; element k of a level of size z lives in lane (k % S),
// register (k / S) when z >= S, and in lane k, register 0 when z < S.  LPF = 32 / S is the
// lane distance between consecutive sub-lanes of a path.
template <int LPF, int S, int Z, bool IS_G, typename real>
PCL_DEVICE void pcl_block_stage(real* dst, const real* src, uint32_t small, int kk, int lane)
{
    if (Z >= S) {
        constexpr int CNT = (Z >= S) ? Z / S : 1;
#pragma unroll
        for (int t = 0; t < CNT; t++) {
            const real a = src[t], b = src[t + CNT];
            if (IS_G) dst[t] = pcl_fast<real>::g(a, b, (small >> (32 - 2 * Z + kk + S * t)) & 1u);
            else dst[t] = pcl_math<real>::f(a, b);
        }
    } else {
        const real a = src[0];
        const real b = pcl_shfl_real<real>(a, lane + Z * LPF);    // element k + Z lives Z sub-lanes up
        if (IS_G) dst[0] = pcl_fast<real>::g(a, b, (small >> ((32 - 2 * Z + kk) & 31)) & 1u);
        else dst[0] = pcl_math<real>::f(a, b);
    }
}

// Prune keys.  fp64 (validation) build: exact total order (metric desc, candidate id asc),
// the stable sort of decoder.py:306-307.  fp32 build: the candidate id replaces the lowest
// mantissa bits of the (always negative) fp64 metric, so one compare decides both the
// metric order and the reference's tie order; the perturbation is < 2^-46 relative.
template <int NC, bool EXACT>
PCL_DEVICE double pcl_prune_key(double m, int c)
{
    if (EXACT) return m;
    long long b = __double_as_longlong(m);
    b = (b & ~(long long)(NC - 1)) | (long long)c;
    return __longlong_as_double(b);
}

template <bool EXACT>
PCL_DEVICE int pcl_beats(double kj, double key, int jj, int c)
{
    if (EXACT) return (kj > key) || (kj == key && jj < c);
    return kj > key;
}

// rank += (kj > key) as one compare and one predicated add
PCL_DEVICE void pcl_rank_acc(int& rank, double kj, double key)
{
#ifdef PCL_EMU
    rank += kj > key;
#else
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(rank) : "d"(kj), "d"(key));
#endif
}

template <typename real>
PCL_DEVICE void pcl_load_quad(const real* ptr, real* v);
template <>
PCL_DEVICE void pcl_load_quad<float>(const float* ptr, float* v)
{
    const float4 q = *reinterpret_cast<const float4*>(ptr);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
template <>
PCL_DEVICE void pcl_load_quad<double>(const double* ptr, double* v)
{
    const double2 q0 = *reinterpret_cast<const double2*>(ptr);
    const double2 q1 = *reinterpret_cast<const double2*>(ptr + 2);
    v[0] = q0.x; v[1] = q0.y; v[2] = q1.x; v[3] = q1.y;
}

template <typename real>
PCL_DEVICE void pcl_store_quad(real* ptr, const real* v);
template <>
PCL_DEVICE void pcl_store_quad<float>(float* ptr, const float* v)
{
    float4 q;
    q.x = v[0]; q.y = v[1]; q.z = v[2]; q.w = v[3];
    *reinterpret_cast<float4*>(ptr) = q;
}
template <>
PCL_DEVICE void pcl_store_quad<double>(double* ptr, const double* v)
{
    double2 q0, q1;
    q0.x = v[0]; q0.y = v[1]; q1.x = v[2]; q1.y = v[3];
    *reinterpret_cast<double2*>(ptr) = q0;
    *reinterpret_cast<double2*>(ptr + 2) = q1;
}

template <typename real>
PCL_DEVICE void pcl_load_pair(const real* ptr, real& a, real& b);
template <>
PCL_DEVICE void pcl_load_pair<float>(const float* ptr, float& a, float& b)
{
    const float2 v = *reinterpret_cast<const float2*>(ptr);
    a = v.x;
    b = v.y;
}
template <>
PCL_DEVICE void pcl_load_pair<double>(const double* ptr, double& a, double& b)
{
    const double2 v = *reinterpret_cast<const double2*>(ptr);
    a = v.x;
    b = v.y;
}

// LP = list slots per frame (power of two), S = lanes per path; a warp decodes
// FPW = 32 / (LP * S) frames side by side.  All frames of a warp follow the same
// schedule (frozen pattern, number of live paths), so every branch stays warp-uniform and
// the per-leaf bookkeeping instructions are shared by FPW frames.
#ifndef PCL_POLAR_MINB
#define PCL_POLAR_MINB 5      // resident 128-thread blocks per SM the register allocation aims for (96 regs)
#endif
template <int LP, int S, typename real>
__global__ void __launch_bounds__(128, (sizeof(real) == 4) ? PCL_POLAR_MINB : 3) polar_scl_fast_kernel(PolarParams<real> P)
{
    constexpr int PB = pcl_log2<LP>::v;
    constexpr int LPF = 32 / S;                  // columns = (frame, slot) pairs per warp
    constexpr int CB = pcl_log2<LPF>::v;
    constexpr int FPW = LPF / LP;                // frames per warp
    constexpr int NC = 2 * LP;                   // prune candidates per frame
    constexpr bool EXACT = sizeof(real) == 8;
    // S == 1: stored levels use the layout [k / 4][column][k % 4], so a lane moves four
    // consecutive elements of its path with one 16-byte access (a warp: 512 contiguous bytes)
    constexpr bool VEC = (S == 1);
    constexpr int E3 = (8 >= S) ? 8 / S : 1;     // registers per lane for the size-8 level
    constexpr int E2 = (4 >= S) ? 4 / S : 1;
    constexpr int E1 = (2 >= S) ? 2 / S : 1;
    // prune work split: lane (kk, col) owns candidate (kk & 1) * LP + p and compares it with
    // the CH candidates of segment kk >> 1 (S >= 2); with S == 1 a lane owns two candidates.
    constexpr int NSEG_RAW = (S >= 2) ? S / 2 : 1;
    constexpr int NSEG = (NSEG_RAW > NC) ? NC : NSEG_RAW;
    constexpr int CH = NC / NSEG;
    const PolarLayout& Y = P.lay;
    const int N = Y.N, n = Y.n, K = Y.K, L = Y.L, G = Y.G, NW = Y.NW, nb = Y.nb;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int col = lane & (LPF - 1);
    const int kk = lane >> CB;
    const int p = col & (LP - 1);
    const int fr = col >> PB;
    const int cbase = col - p;                    // first column of this lane's frame
    const int shift = (N < 32) ? 32 - N : 0;
    const int NB = N >> 3;

    unsigned char* wsm = pcl_dyn_smem() + (size_t)warp * Y.warp_bytes;
    double* cm = (double*)(wsm + Y.off_cm) + fr * (NC + 2);    // this frame's candidate keys (padded: banks)
    double* newpm = (double*)(wsm + Y.off_newpm) + cbase;
    int* sel = (int*)(wsm + Y.off_sel) + cbase;
    real* sl = (real*)(wsm + Y.off_llr);          // levels max(G,1)+1 .. n-4, [k][col]
    uint32_t* bw = (uint32_t*)(wsm + Y.off_bw);   // big left levels 1 .. nb, [w][col]
    uint32_t* uw = (uint32_t*)(wsm + Y.off_uw);   // final u words
    real* gl = P.scratch + (int64_t)(blockIdx.x * wpb + warp) * Y.scratch_per_warp;

    const double NEG_INF = -(double)pcl_math<real>::inf();
    const double DEAD = -1.0e300;                 // key of an inactive slot (sorts last, stays finite)

    for (int64_t f0 = ((int64_t)blockIdx.x * wpb + warp) * FPW; f0 < P.F; f0 += (int64_t)gridDim.x * wpb * FPW) {
        const int64_t f = f0 + fr;
        const bool valid = f < P.F;
        const real* y = P.llr + (valid ? f : f0) * N;
        int nact = 1;
        bool act = (p == 0) && valid;
        double pm = act ? 0.0 : NEG_INF;
        uint32_t ptrL = 0, ptrB = 0;              // packed slot pointers: LLR levels / left levels
        uint32_t small = 0, ulast = 0;
        uint32_t fw = 0;

        for (int blk = 0; blk < NB; blk++) {
            const int i0 = blk << 3;
            if (((i0 + shift) & 31) == 0 || blk == 0) fw = P.frozen_words[(i0 + shift) >> 5];
            const uint32_t fz8 = (fw >> ((i0 + shift) & 31)) & 0xffu;

            // ---- levels above the cut: start .. n-3; the last one lands in registers ----
            // Level 1 is never stored: level 2 recomputes its two level-1 operands from the
            // channel LLRs on the fly (y[k] = llr[br(k)], y[k + N/2] = llr[br(k) + 1]), which
            // removes the largest scratch array and its HBM round trips.
            real R3[E3];
#pragma unroll
            for (int t = 0; t < E3; t++) R3[t] = (real)0;
            const int start = (blk == 0) ? 1 : n - (__ffs(i0) - 1);
            const int dfirst = (n - 3 >= 2) ? 2 : 1;
            const int bit1 = (i0 >> (n - 1)) & 1;
            const uint32_t* b1src = bw + cbase + (ptrB & (LP - 1));      // left array of level 1 (nb >= 1)
            auto lvl1 = [&](int m) -> real {
                const int r = (int)(__brev((unsigned)m) >> (32 - n));      // even: the pair is one aligned load
                real y0, y1;
                pcl_load_pair<real>(y + r, y0, y1);
                if (bit1) {
                    const uint32_t ub = (nb >= 1) ? (b1src[(m >> 5) * LPF] >> (m & 31)) & 1u
                                                  : (small >> ((32 - N + m) & 31)) & 1u;
                    return pcl_fast<real>::g(y0, y1, ub);
                }
                return pcl_math<real>::f(y0, y1);
            };
            for (int d = (start > dfirst ? start : dfirst); d <= n - 3; d++) {
                const int sz = N >> d;
                const int bit = (i0 >> (n - d)) & 1;
                const int hi = sz * LPF;                       // word distance of the partner element
                const real* src = nullptr;                     // + sub-lane and borrowed column
                if (d > 2) {
                    const int q = (ptrL >> ((d - 2) * PB)) & (LP - 1);
                    src = ((d - 1 <= G) ? gl + (int64_t)LPF * ((N >> 1) - (N >> (d - 2)))
                                        : sl + LPF * ((N >> G) - (N >> (d - 2)))) +
                          (VEC ? 4 * (cbase + q) : kk * LPF + cbase + q);
                }
                const uint32_t* bsrc = nullptr;
                if (bit && d <= nb)
                    bsrc = bw + LPF * ((N >> 5) - (N >> (d + 4))) + cbase + ((ptrB >> ((d - 1) * PB)) & (LP - 1));
                const uint32_t smf = small >> ((32 - 2 * sz) & 31);        // small-field partial sums of level d
                if (d < n - 3) {
                    real* dst = ((d <= G) ? gl + (int64_t)LPF * ((N >> 1) - (N >> (d - 1)))
                                          : sl + LPF * ((N >> G) - (N >> (d - 1)))) + (VEC ? 4 * col : kk * LPF + col);
                    // element k = kk + S t sits at word offset 32 t of the [k][col] array
                    if (act) {
                        if (d == 2 && VEC) {
                            // quad i4 of level 2 = elements 4 i4 + r; element k comes from the 4 channel
                            // values at 4 br(k), and br(4 i4 + r) = br(i4) + br2(r) * sz/4: four sequential
                            // streams of 16-byte loads, one 16-byte store
                            const int nq = sz >> 2;
#pragma unroll 2
                            for (int tl = 0; tl < nq; tl++) {
                                const int i4 = (int)(__brev((unsigned)tl) >> (36 - n));    // (n-4)-bit reversal
                                real yv[4][4];
#pragma unroll
                                for (int mm = 0; mm < 4; mm++) pcl_load_quad<real>(y + 4 * (tl + mm * nq), yv[mm]);
                                const int k4 = 4 * i4;
                                uint32_t n1 = 0, n2 = 0, n0 = 0;
                                if (bit1) {
                                    if (nb >= 1) {
                                        n1 = b1src[(k4 >> 5) * LPF] >> (k4 & 31);
                                        n2 = b1src[((k4 + sz) >> 5) * LPF] >> ((k4 + sz) & 31);
                                    } else {
                                        n1 = small >> ((32 - N + k4) & 31);
                                        n2 = small >> ((32 - N + k4 + sz) & 31);
                                    }
                                }
                                if (bit) n0 = (d <= nb) ? bsrc[(k4 >> 5) * LPF] >> (k4 & 31) : smf >> k4;
                                real out[4];
#pragma unroll
                                for (int mm = 0; mm < 4; mm++) {
                                    const int r = ((mm & 1) << 1) | (mm >> 1);             // 2-bit reversal
                                    real a, b;
                                    if (bit1) {
                                        a = pcl_fast<real>::g(yv[mm][0], yv[mm][1], (n1 >> r) & 1u);
                                        b = pcl_fast<real>::g(yv[mm][2], yv[mm][3], (n2 >> r) & 1u);
                                    } else {
                                        a = pcl_math<real>::f(yv[mm][0], yv[mm][1]);
                                        b = pcl_math<real>::f(yv[mm][2], yv[mm][3]);
                                    }
                                    out[r] = bit ? pcl_fast<real>::g(a, b, (n0 >> r) & 1u) : pcl_math<real>::f(a, b);
                                }
                                pcl_store_quad<real>(dst + i4 * 128, out);
                            }
                        } else if (VEC) {
                            // quads of 4 consecutive elements, two quads per batch, ping-pong register sets
                            const int nq = sz >> 2;                              // quads per half
                            real a[2][4], b[2][4], a2[2][4], b2[2][4];
                            auto emitq = [&](real (*av)[4], real (*bv)[4], int i) {
#pragma unroll
                                for (int u = 0; u < 2; u++) {
                                    real out[4];
                                    if (bit) {
                                        const int k4 = 4 * (i + u);
                                        const uint32_t nbits = (d <= nb) ? bsrc[(k4 >> 5) * LPF] >> (k4 & 31) : smf >> k4;
#pragma unroll
                                        for (int e = 0; e < 4; e++) out[e] = pcl_fast<real>::g(av[u][e], bv[u][e], (nbits >> e) & 1u);
                                    } else {
#pragma unroll
                                        for (int e = 0; e < 4; e++) out[e] = pcl_math<real>::f(av[u][e], bv[u][e]);
                                    }
                                    pcl_store_quad<real>(dst + (i + u) * 128, out);
                                }
                            };
#pragma unroll
                            for (int u = 0; u < 2; u++) {
                                pcl_load_quad<real>(src + u * 128, a[u]);
                                pcl_load_quad<real>(src + (u + nq) * 128, b[u]);
                            }
                            for (int i = 0; i < nq; i += 4) {
                                if (i + 2 < nq) {
#pragma unroll
                                    for (int u = 0; u < 2; u++) {
                                        pcl_load_quad<real>(src + (i + 2 + u) * 128, a2[u]);
                                        pcl_load_quad<real>(src + (i + 2 + u + nq) * 128, b2[u]);
                                    }
                                }
                                emitq(a, b, i);
                                if (i + 2 < nq) {
                                    if (i + 4 < nq) {
#pragma unroll
                                        for (int u = 0; u < 2; u++) {
                                            pcl_load_quad<real>(src + (i + 4 + u) * 128, a[u]);
                                            pcl_load_quad<real>(src + (i + 4 + u + nq) * 128, b[u]);
                                        }
                                    }
                                    emitq(a2, b2, i + 2);
                                }
                            }
                        } else if (d == 2) {
                            // Level-2 element k needs level-1 elements k and k + N/4, i.e. the channel
                            // pairs at br(k) and br(k) + 2: ONE aligned 4-element load.  Walking t = br(k)
                            // upwards makes those loads sequential in memory.
                            real* dst2 = dst - kk * LPF;
#pragma unroll 4
                            for (int t = kk; t < sz; t += S) {
                                const int k = (int)(__brev((unsigned)t) >> (34 - n));      // (n-2)-bit reversal
                                real yv[4];
                                pcl_load_quad<real>(y + 4 * t, yv);
                                real a, b;
                                if (bit1) {
                                    uint32_t ua, ub;
                                    if (nb >= 1) {
                                        ua = (b1src[(k >> 5) * LPF] >> (k & 31)) & 1u;
                                        ub = (b1src[((k + sz) >> 5) * LPF] >> ((k + sz) & 31)) & 1u;
                                    } else {
                                        ua = (small >> ((32 - N + k) & 31)) & 1u;
                                        ub = (small >> ((32 - N + k + sz) & 31)) & 1u;
                                    }
                                    a = pcl_fast<real>::g(yv[0], yv[1], ua);
                                    b = pcl_fast<real>::g(yv[2], yv[3], ub);
                                } else {
                                    a = pcl_math<real>::f(yv[0], yv[1]);
                                    b = pcl_math<real>::f(yv[2], yv[3]);
                                }
                                real v;
                                if (bit) v = pcl_fast<real>::g(a, b, (d <= nb) ? (bsrc[(k >> 5) * LPF] >> (k & 31)) & 1u
                                                                              : (smf >> k) & 1u);
                                else v = pcl_math<real>::f(a, b);
                                dst2[k * LPF] = v;
                            }
                        } else {
                            // Software-pipelined batches of independent loads: these arrays may live in
                            // L2 / HBM.  sz >= 16, so sz / S is a multiple of UNR (no tail); the UNR
                            // elements of a batch take their partial-sum bits from one word.
#ifndef PCL_POLAR_UNR
#define PCL_POLAR_UNR 8
#endif
                            constexpr int UNR = (S >= 4 && PCL_POLAR_UNR > 4) ? 4 : PCL_POLAR_UNR;
                            constexpr int STEP = 32 * UNR;
                            real a[UNR], b[UNR], a2[UNR], b2[UNR];
                            // compute + store one batch whose operands are already in registers
                            auto emit = [&](const real* av, const real* bv, int o0) {
                                if (bit) {
                                    const int k0 = kk + S * (o0 >> 5);          // k of element e: k0 + S e
                                    const uint32_t wbits = (d <= nb) ? bsrc[(k0 >> 5) * LPF] >> (k0 & 31) : smf >> k0;
#pragma unroll
                                    for (int e = 0; e < UNR; e++)
                                        dst[o0 + 32 * e] = pcl_fast<real>::g(av[e], bv[e], (wbits >> (S * e)) & 1u);
                                } else {
#pragma unroll
                                    for (int e = 0; e < UNR; e++) dst[o0 + 32 * e] = pcl_math<real>::f(av[e], bv[e]);
                                }
                            };
#pragma unroll
                            for (int e = 0; e < UNR; e++) {
                                a[e] = src[32 * e];
                                b[e] = src[32 * e + hi];
                            }
                            // ping-pong between two register sets: the loads of the next batch are in
                            // flight while the current one is computed, without register moves
                            for (int o0 = 0; o0 < hi; o0 += 2 * STEP) {
                                const int o1 = o0 + STEP, o2 = o0 + 2 * STEP;
                                if (o1 < hi) {
#pragma unroll
                                    for (int e = 0; e < UNR; e++) {
                                        a2[e] = src[o1 + 32 * e];
                                        b2[e] = src[o1 + 32 * e + hi];
                                    }
                                }
                                emit(a, b, o0);
                                if (o1 < hi) {
                                    if (o2 < hi) {
#pragma unroll
                                        for (int e = 0; e < UNR; e++) {
                                            a[e] = src[o2 + 32 * e];
                                            b[e] = src[o2 + 32 * e + hi];
                                        }
                                    }
                                    emit(a2, b2, o1);
                                }
                            }
                        }
                        ptrL = (ptrL & ~((uint32_t)(LP - 1) << ((d - 1) * PB))) | ((uint32_t)p << ((d - 1) * PB));
                    }
                    __syncwarp();
                } else if (act && VEC && d > 2) {      // d == n-3 from four 16-byte loads of level n-4
                    real qa[2][4], qb[2][4];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        pcl_load_quad<real>(src + u * 128, qa[u]);
                        pcl_load_quad<real>(src + (u + 2) * 128, qb[u]);
                    }
#pragma unroll
                    for (int t = 0; t < E3; t++) {
                        const real a = qa[(t >> 2) & 1][t & 3], b = qb[(t >> 2) & 1][t & 3];
                        const real vg = pcl_fast<real>::g(a, b, (small >> (16 + t)) & 1u);
                        const real vf = pcl_math<real>::f(a, b);
                        R3[t] = bit ? vg : vf;
                    }
                } else if (act) {                 // d == n-3: sz == 8, straight into registers
#pragma unroll
                    for (int t = 0; t < E3; t++) {
                        const int k = kk + S * t;
                        if (k < 8) {
                            real a, b;
                            if (d == 1) {
                                const int r = (int)(__brev((unsigned)k) >> (32 - n));
                                a = y[r];
                                b = y[r + 1];
                            } else if (d == 2) {
                                a = lvl1(k);
                                b = lvl1(k + 8);
                            } else if (VEC) {
                                a = src[(t >> 2) * 128 + (t & 3)];           // level n-4: quads 0,1 | 2,3
                                b = src[((t >> 2) + 2) * 128 + (t & 3)];
                            } else {
                                a = src[32 * t];
                                b = src[32 * t + 8 * LPF];
                            }
                            const real vg = pcl_fast<real>::g(a, b, (small >> (16 + k)) & 1u);
                            const real vf = pcl_math<real>::f(a, b);
                            R3[t] = bit ? vg : vf;
                        }
                    }
                }
            }
            // all borrowed source arrays have been read: order before later overwrites
            __syncwarp();

            // ---- the 8 leaves of the block (rolled: the body must stay I-cache resident) ----
            real R2[E2], R1[E1];
#pragma unroll
            for (int t = 0; t < E2; t++) R2[t] = (real)0;
#pragma unroll
            for (int t = 0; t < E1; t++) R1[t] = (real)0;
#pragma unroll 1
            for (int j = 0; j < 8; j++) {
                const int i = i0 + j;
                const bool frozen = (fz8 >> j) & 1u;
                // heights to recompute: j == 0 -> 2,1,0; else ctz(j) .. 0; f or g per bit of j
                real x;
                if (j & 1) {
                    pcl_block_stage<LPF, S, 1, true, real>(&x, R1, small, kk, lane);
                } else {
                    if ((j & 2) == 0) {
                        if (j & 4) pcl_block_stage<LPF, S, 4, true, real>(R2, R3, small, kk, lane);
                        else pcl_block_stage<LPF, S, 4, false, real>(R2, R3, small, kk, lane);
                        pcl_block_stage<LPF, S, 2, false, real>(R1, R2, small, kk, lane);
                    } else {
                        pcl_block_stage<LPF, S, 2, true, real>(R1, R2, small, kk, lane);
                    }
                    pcl_block_stage<LPF, S, 1, false, real>(&x, R1, small, kk, lane);
                }
                if (S > 1) x = pcl_shfl_real<real>(x, col);        // lane `col` is sub-lane 0 of the path
                if (!act) x = (real)0;

                // ---- leaf decision (polar_scl.cuh for the rules) -------------------------
                const real ax = fabs(x);
                const bool hard = !(x >= (real)0);
                uint32_t u = 0;
                int parent = p;
                bool forked = false;
                if (LP == 1) {
                    u = frozen ? 0u : (hard ? 1u : 0u);
                    if (P.want_pm) {
                        const real pen = pcl_fast<real>::softplus_neg_abs(ax) + ((u != (uint32_t)hard) ? ax : (real)0);
                        if (act) pm -= (double)pen;
                    }
                } else if (frozen) {
                    const real pen = pcl_fast<real>::softplus_neg_abs(ax) + (hard ? ax : (real)0);
                    if (act) pm -= (double)pen;
                } else {
                    forked = true;
                    const double base = pm - (double)pcl_fast<real>::softplus_neg_abs(ax);
                    const double other = base - (double)ax;
                    const int ns = (2 * nact < L) ? 2 * nact : L;
                    if (S >= 2) {
                        // own candidate: bit (kk & 1) of path p
                        const int c = (kk & 1) * LP + p;
                        const int seg = kk >> 1;
                        double mc = ((kk & 1) != 0) == hard ? base : other;
                        if (!act) mc = DEAD;
                        const double key = pcl_prune_key<NC, EXACT>(mc, c);
                        if (kk < 2) cm[c] = key;
                        __syncwarp();
                        int rank = 0;
                        if (seg < NSEG) {
#pragma unroll
                            for (int e = 0; e < CH; e++) {
                                if (EXACT) rank += pcl_beats<EXACT>(cm[seg * CH + e], key, seg * CH + e, c);
                                else pcl_rank_acc(rank, cm[seg * CH + e], key);
                            }
                        }
#pragma unroll
                        for (int o = 1; o < NSEG; o <<= 1) rank += __shfl_xor_sync(PCL_FULL_MASK, rank, 2 * LPF * o);
                        if (kk < 2 && rank < ns) { sel[rank] = c; newpm[rank] = mc; }
                    } else {
                        double mca = hard ? other : base;             // bit 0
                        double mcb = hard ? base : other;             // bit 1
                        if (!act) { mca = DEAD; mcb = DEAD; }
                        const double ka = pcl_prune_key<NC, EXACT>(mca, p);
                        const double kb = pcl_prune_key<NC, EXACT>(mcb, LP + p);
                        cm[p] = ka;
                        cm[LP + p] = kb;
                        __syncwarp();
                        int ra = 0, rb = 0;
#pragma unroll 8
                        for (int jj = 0; jj < NC; jj += 2) {
                            const double2 kp = *reinterpret_cast<const double2*>(cm + jj);   // 16-byte aligned
                            if (EXACT) {
                                ra += pcl_beats<EXACT>(kp.x, ka, jj, p) + pcl_beats<EXACT>(kp.y, ka, jj + 1, p);
                                rb += pcl_beats<EXACT>(kp.x, kb, jj, LP + p) + pcl_beats<EXACT>(kp.y, kb, jj + 1, LP + p);
                            } else {
                                pcl_rank_acc(ra, kp.x, ka);
                                pcl_rank_acc(rb, kp.x, kb);
                                pcl_rank_acc(ra, kp.y, ka);
                                pcl_rank_acc(rb, kp.y, kb);
                            }
                        }
                        if (ra < ns) { sel[ra] = p; newpm[ra] = mca; }
                        if (rb < ns) { sel[rb] = LP + p; newpm[rb] = mcb; }
                    }
                    __syncwarp();
                    act = (p < ns) && valid;
                    if (p < ns) {
                        const int c = sel[p];
                        parent = c & (LP - 1);
                        u = (uint32_t)(c >> PB);
                        pm = newpm[p];
                    }
                    if (!act) pm = NEG_INF;
                    nact = ns;
                }
                if (forked) {
                    // a survivor takes over its parent's pointer words and live registers
                    const int srcl = (lane & ~(LP - 1)) | parent;
                    ptrL = __shfl_sync(PCL_FULL_MASK, ptrL, srcl);
                    ptrB = __shfl_sync(PCL_FULL_MASK, ptrB, srcl);
                    small = __shfl_sync(PCL_FULL_MASK, small, srcl);
                    if (j < 4) {
#pragma unroll
                        for (int t = 0; t < E3; t++) R3[t] = pcl_shfl_real<real>(R3[t], srcl);
                    }
                    if ((j & 3) < 2) {
#pragma unroll
                        for (int t = 0; t < E2; t++) R2[t] = pcl_shfl_real<real>(R2[t], srcl);
                    }
                    if ((j & 1) == 0) {
#pragma unroll
                        for (int t = 0; t < E1; t++) R1[t] = pcl_shfl_real<real>(R1[t], srcl);
                    }
                    __syncwarp();
                }
                if (P.dbg_leaf != nullptr && kk == 0 && valid) {
                    P.dbg_leaf[(f * N + i) * LP + p] = x;
                    P.dbg_parent[(f * N + i) * LP + p] = (uint8_t)parent;
                }

                // ---- partial sums: fields of sizes 1, 2, 4 at bits 30, 28, 24 ------------
                if ((j & 1) == 0) {
                    small = (small & ~(1u << 30)) | (u << 30);
                } else {
                    const uint32_t c2 = (((small >> 30) & 1u) ^ u) | (u << 1);
                    if ((j & 2) == 0) {
                        small = (small & ~(3u << 28)) | (c2 << 28);
                    } else {
                        const uint32_t c4 = (((small >> 28) & 3u) ^ c2) | (c2 << 2);
                        if (j == 3) {
                            small = (small & ~(15u << 24)) | (c4 << 24);
                        } else if (blk == NB - 1) {
                            ulast = u;               // last leaf: the fields stay as they are
                        } else {
                            // block complete: fold upwards while the node is a right child
                            uint32_t c = (((small >> 24) & 15u) ^ c4) | (c4 << 4);
                            int s = 8, tt = blk;
                            while ((tt & 1) && s < 32) {
                                const uint32_t left = pcl_bfe(small, 32 - 2 * s, s);
                                c = (left ^ c) | (c << s);
                                s <<= 1;
                                tt >>= 1;
                            }
                            if (!(tt & 1)) {
                                if (s < 32) {
                                    small = pcl_bfi(small, c, 32 - 2 * s, s);
                                } else {
                                    const int d = n - 5;
                                    if (act) {
                                        if (kk == 0) bw[LPF * ((N >> 5) - (N >> (d + 4))) + col] = c;
                                        ptrB = (ptrB & ~((uint32_t)(LP - 1) << ((d - 1) * PB))) | ((uint32_t)p << ((d - 1) * PB));
                                    }
                                    __syncwarp();
                                }
                            } else {
                                const int cto = __ffs(~i) - 1;
                                const int d = n - cto;
                                const int Wd = N >> (d + 5);
                                uint32_t* dest = bw + LPF * ((N >> 5) - (N >> (d + 4)));
                                if (act && kk == 0) dest[(Wd - 1) * LPF + col] = c;
                                __syncwarp();
                                for (int l = n - 5; l > d; l--) {
                                    const int w = N >> (l + 5);
                                    const int ql = cbase + ((ptrB >> ((l - 1) * PB)) & (LP - 1));
                                    const uint32_t* lsrc = bw + LPF * ((N >> 5) - (N >> (l + 4)));
                                    if (act)
                                        for (int jw = kk; jw < w; jw += S)
                                            dest[(Wd - 2 * w + jw) * LPF + col] =
                                                lsrc[jw * LPF + ql] ^ dest[(Wd - w + jw) * LPF + col];
                                    __syncwarp();
                                }
                                if (act) ptrB = (ptrB & ~((uint32_t)(LP - 1) << ((d - 1) * PB))) | ((uint32_t)p << ((d - 1) * PB));
                            }
                        }
                    }
                }
            }
        }

        // ---- final selection (decoder.py:259-262), per frame of the warp ------------------------
        int best = 0;
        if (LP > 1) {
            if (kk == 0) newpm[p] = pm;
            __syncwarp();
            double bm = newpm[0];
            for (int q = 1; q < LP; q++) {
                const double v = newpm[q];
                if (v > bm) { bm = v; best = q; }           // first maximum, like np.argmax
            }
        }
        if (P.pm_out != nullptr && kk == 0 && p < L && valid) P.pm_out[f * L + p] = pm;

        const int nslots = P.use_crc ? nact : 1;
        for (int fq = 0; fq < FPW; fq++) {
            const int fbest = __shfl_sync(PCL_FULL_MASK, best, fq * LP);
            for (int sidx = 0; sidx < nslots; sidx++) {
                const int slot = P.use_crc ? sidx : fbest;
                uint32_t* U = uw + (P.use_crc ? (fq * LP + sidx) * NW : fq * NW);
                const uint32_t pB = __shfl_sync(PCL_FULL_MASK, ptrB, fq * LP + slot);
                const uint32_t sm = __shfl_sync(PCL_FULL_MASK, small, fq * LP + slot);
                const uint32_t ul = __shfl_sync(PCL_FULL_MASK, ulast, fq * LP + slot);
                for (int w = lane; w < NW; w += 32) {
                    uint32_t v;
                    if (w == NW - 1) {
                        v = pcl_bfi(sm, ul, 31, 1);
                        v ^= (v >> 1) & 0x15555555u;
                        v ^= (v >> 2) & 0x03333333u;
                        v ^= (v >> 4) & 0x000F0F0Fu;
                        v ^= (v >> 8) & 0x000000FFu;
                    } else {
                        const int r = NW - w;
                        const int Wl = 1 << (31 - __clz(r - 1));
                        const int l = (31 - __clz(NW)) - (31 - __clz(Wl));
                        const int jw = w - (NW - 2 * Wl);
                        v = bw[LPF * ((N >> 5) - (N >> (l + 4))) + jw * LPF + fq * LP + ((pB >> ((l - 1) * PB)) & (LP - 1))];
                        v ^= (v >> 1) & 0x55555555u;
                        v ^= (v >> 2) & 0x33333333u;
                        v ^= (v >> 4) & 0x0F0F0F0Fu;
                        v ^= (v >> 8) & 0x00FF00FFu;
                        v ^= (v >> 16) & 0x0000FFFFu;
                    }
                    U[w] = v;
                }
                __syncwarp();
                for (int t = 1; t < NW; t <<= 1) {
                    for (int w = lane; w < NW - 1; w += 32) {
                        const int r = NW - w;
                        const int Wl = 1 << (31 - __clz(r - 1));
                        const int jw = w - (NW - 2 * Wl);
                        if (t < Wl && (jw & t) == 0) U[w] ^= U[w + t];
                    }
                    __syncwarp();
                }
            }
        }

        if (P.use_crc) {
            // first path in (metric desc, slot asc) order whose info bits pass the CRC register
            // test (src/polar/utils.py:128-163); else the best metric.
            bool pass = false;
            if (kk == 0 && p < nact && valid) {
                const uint32_t* U = uw + col * NW;
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
            const unsigned pmask = (__ballot_sync(PCL_FULL_MASK, pass) >> cbase) & ((LP >= 32) ? 0xffffffffu : ((1u << LP) - 1u));
            if (pmask != 0) {
                int bsel = -1;
                double bm = 0;
                for (int q = 0; q < nact; q++) {
                    if (!((pmask >> q) & 1u)) continue;
                    const double v = newpm[q];
                    if (bsel < 0 || v > bm) { bm = v; bsel = q; }
                }
                best = bsel;
            }
        }
        for (int fq = 0; fq < FPW; fq++) {   // decoded = u[info_bits] (decoder.py:70-71 / :260-262)
            const int fbest = __shfl_sync(PCL_FULL_MASK, best, fq * LP);
            if (f0 + fq < P.F) {
                const uint32_t* U = uw + (P.use_crc ? (fq * LP + fbest) * NW : fq * NW);
                uint8_t* out = P.bits + (f0 + fq) * K;
                for (int k = lane; k < K; k += 32) {
                    const int pos = (int)P.info_pos[k] + shift;
                    out[k] = (uint8_t)((U[pos >> 5] >> (pos & 31)) & 1u);
                }
            }
        }
        __syncwarp();
    }
}
