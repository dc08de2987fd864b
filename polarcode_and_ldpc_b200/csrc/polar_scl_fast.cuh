// polar_scl_fast.cuh -- SC / SCL decoder with a register-resident tree bottom and several
// frames per warp.  Production kernel for N >= 16 (polar_scl.cuh is the fallback).
//
// Same algorithm, list semantics and pointer scheme as polar_scl.cuh (read that header
// first).  What differs, each step driven by an ncu capture (profiles/r01a .. r01o):
//   * A lane owns a whole path and a warp decodes FPW = 32 / LP frames side by side (4 frames
//     at L = 8).  All frames of a warp follow the same schedule (frozen pattern, number of live
//     paths), so every branch is warp-uniform and the per-leaf bookkeeping instructions are
//     shared by FPW frames.
//   * The tree is cut at height 3: the decoder walks N/8 blocks of 8 leaves; the block root
//     (8 LLRs per path) is produced straight into registers and the levels of size 4, 2, 1
//     below it stay in registers.  The leaf loop is rolled (the body must stay resident in the
//     32 KB instruction cache); f / g per stage sits behind uniform branches on j.
//   * A surviving path inherits its parent's live registers with shuffles (only those a later
//     leaf of the block still reads) together with two packed 32-bit pointer words (LLR levels,
//     partial-sum levels) and the 32-bit small partial-sum word.
//   * Level 1 is never stored: level 2 recomputes its two level-1 operands from ONE aligned
//     4-element channel load, walked in bit-reversed order so the loads are sequential.
//     Levels 2..G live in an L2/HBM scratch, levels G+1..n-4 in shared memory, both in the
//     layout [k / 4][column][k % 4]: a lane moves four consecutive elements of its path with one
//     16-byte access (a warp: 512 contiguous bytes) and two batches of loads are in flight
//     while the previous one is computed.
//   * NL / GL: the headline code sizes are compiled with log2 N and G as constants, which turns
//     the level walk into straight-line code per level (trip counts, array offsets and the
//     memory space of every access are immediates; f and g loops are separate).  NL = 0 keeps
//     the run-time n for every other code length.
//   * Prune: all-pairs rank.  fp32 build: the candidate id replaces the lowest mantissa bits of
//     the fp64 metric, so one DSETP + one predicated add per pair gives both the metric order
//     and the reference's tie order, and the survivor scatter moves the key alone (the id comes
//     back out of it); fp64 build keeps the exact two-key comparison and the exact metric.
//     Reliable info bits never get that far (fp32 build): on a full list whose likely keys are
//     still in slot order, with every unlikely key below the last of them, each path continues
//     in place (80 % of the info leaves at 2 dB); otherwise the ranks against the likely keys
//     come first and the unlikely half is skipped when none of them can survive.  Both give
//     exactly the survivors and slots of the full ranking.
//   * The warps of a block make the same number of frame passes and meet at a barrier before
//     each: identical instruction streams keep them at the same tree position, so a block
//     shares one instruction working set (the hot code is ~40 KB, the instruction cache 32 KB).
// Path metric in fp64; fp32 build evaluates log1p(exp(-|x|)) as 2 atanh(u / (2 + u)) with
// u = 2^(-|x| log2 e) (one ex2, one rcp, 6 FMA; |error| < 1e-7).
#pragma once
#include "pcl_common.cuh"
#include "pcl_tmem.cuh"
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

PCL_DEVICE float pcl_lg2f(float x)
{
#ifdef PCL_EMU
    return log2f(x);
#else
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}

#ifndef PCL_SOFTPLUS_LG2
#define PCL_SOFTPLUS_LG2 1
#endif
template <typename real> struct pcl_fast;
template <> struct pcl_fast<float> {
    static PCL_DEVICE float softplus_neg_abs(float ax)
    {
        const float u = pcl_ex2f(ax * -1.4426950408889634f);      // exp(-|x|) in (0, 1]
#if PCL_SOFTPLUS_LG2
        return pcl_lg2f(1.0f + u) * 0.6931471805599453f;          // two MUFU ops, 5 instructions
#endif
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
    // b + (bit ? -a : a) with the sign flip done on the bit pattern; `sbit` has the bit at position 31
    static PCL_DEVICE float gs(float a, float b, uint32_t sbit)
    {
        return b + __uint_as_float(__float_as_uint(a) ^ (sbit & 0x80000000u));
    }
    static PCL_DEVICE float g(float a, float b, uint32_t bit) { return gs(a, b, bit << 31); }
};
template <> struct pcl_fast<double> {
    static PCL_DEVICE double softplus_neg_abs(double ax) { return log1p(exp(-ax)); }
    static PCL_DEVICE double gs(double a, double b, uint32_t sbit) { return (sbit & 0x80000000u) ? b - a : b + a; }
    static PCL_DEVICE double g(double a, double b, uint32_t bit) { return bit ? b - a : b + a; }
};

PCL_DEVICE double pcl_shfl_f64(double v, int src)
{
    return __longlong_as_double((long long)pcl_shfl_u64((uint64_t)__double_as_longlong(v), src));
}

template <typename real>
PCL_DEVICE real pcl_shfl_real(real v, int src);
template <>
PCL_DEVICE float pcl_shfl_real<float>(float v, int src) { return __shfl_sync(PCL_FULL_MASK, v, src); }
template <>
PCL_DEVICE double pcl_shfl_real<double>(double v, int src)
{
    return __longlong_as_double((long long)pcl_shfl_u64((uint64_t)__double_as_longlong(v), src));
}

// Prune keys.  fp64 (validation) build: exact total order (metric desc, candidate id asc),
// the stable sort of decoder.py:306-307.  fp32 build: the candidate id replaces the lowest
// mantissa bits of the (always negative) fp64 metric, so one compare decides both the
// metric order and the reference's tie order; the perturbation is < 2^-46 relative.
template <int NC>
PCL_DEVICE double pcl_prune_key(double m, int c)
{
    long long b = __double_as_longlong(m);
    b = (b & ~(long long)(NC - 1)) | (long long)c;
    return __longlong_as_double(b);
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

// channel LLRs are read once per visit of level 2 and never written: streaming loads keep them
// from displacing the scratch levels in L2
template <typename real>
PCL_DEVICE void pcl_load_quad_stream(const real* ptr, real* v);
template <>
PCL_DEVICE void pcl_load_quad_stream<float>(const float* ptr, float* v)
{
#if defined(PCL_EMU) || !defined(PCL_STREAM_LLR)
    pcl_load_quad<float>(ptr, v);
#else
    const float4 q = __ldcs(reinterpret_cast<const float4*>(ptr));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
#endif
}
template <>
PCL_DEVICE void pcl_load_quad_stream<double>(const double* ptr, double* v)
{
    pcl_load_quad<double>(ptr, v);
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

// Per-frame prune scratch in shared memory: 2 LP candidate keys | LP survivor keys / metrics |
// LP survivor ids (fp64 build).  One base address per lane, every access at an immediate offset.
// LP = 1 (SC) never prunes and needs none.
PCL_HOST_DEVICE constexpr int pcl_fast_frame_bytes(int LP) { return LP == 1 ? 0 : ((LP * 28 + 15) / 16) * 16; }

// Scratch levels in global memory: optional L2-only accesses (-DPCL_SCRATCH_CG)
template <bool GLOBAL, typename real>
PCL_DEVICE void pcl_ldq(const real* ptr, real* v)
{
#if !defined(PCL_EMU) && defined(PCL_SCRATCH_CG)
    if (GLOBAL && sizeof(real) == 4) {
        const float4 q = __ldcg(reinterpret_cast<const float4*>(ptr));
        v[0] = (real)q.x; v[1] = (real)q.y; v[2] = (real)q.z; v[3] = (real)q.w;
        return;
    }
#endif
    pcl_load_quad<real>(ptr, v);
}
template <bool GLOBAL, typename real>
PCL_DEVICE void pcl_stq(real* ptr, const real* v)
{
#if !defined(PCL_EMU) && defined(PCL_SCRATCH_CG)
    if (GLOBAL && sizeof(real) == 4) {
        float4 q;
        q.x = (float)v[0]; q.y = (float)v[1]; q.z = (float)v[2]; q.w = (float)v[3];
        __stcg(reinterpret_cast<float4*>(ptr), q);
        return;
    }
#endif
    pcl_store_quad<real>(ptr, v);
}

// One quad (4 consecutive elements of a path) of a butterfly stage.
template <bool BIT, typename real>
PCL_DEVICE void pcl_quad_op(real* out, const real* a, const real* b, uint32_t nbits)
{
#pragma unroll
    for (int e = 0; e < 4; e++) {
        if (BIT) out[e] = pcl_fast<real>::gs(a[e], b[e], nbits << (31 - e));
        else out[e] = pcl_math<real>::f(a[e], b[e]);
    }
}

// Stored level d (3 <= d <= n-4) of one path from level d-1: quads i and i + nq of the source
// give quad i of the destination; quad stride is 128 elements ([k / 4][column][k % 4]).
// Two batches of two quads: the loads of the next batch are in flight while the current one is
// computed.  The partial-sum bits of 8 consecutive quads sit in one word: `bsrc` (stride 32
// words, levels of >= 32 elements) or the register field `smf` (16 elements).
template <bool BIT, bool GS, bool GD, typename real>
PCL_DEVICE void pcl_level_vec(real* dst, const real* src, int nq, const uint32_t* bsrc, uint32_t smf)
{
    real a[2][4], b[2][4], a2[2][4], b2[2][4];
    uint32_t wb = smf;
#pragma unroll
    for (int u = 0; u < 2; u++) {
        pcl_ldq<GS, real>(src + u * 128, a[u]);
        pcl_ldq<GS, real>(src + (u + nq) * 128, b[u]);
    }
#pragma unroll 2
    for (int i = 0; i < nq; i += 4) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            pcl_ldq<GS, real>(src + (i + 2 + u) * 128, a2[u]);
            pcl_ldq<GS, real>(src + (i + 2 + u + nq) * 128, b2[u]);
        }
        if (BIT && bsrc != nullptr && (i & 7) == 0) wb = bsrc[(i >> 3) * 32];
        const uint32_t w4 = wb >> (4 * (i & 7));
#pragma unroll
        for (int u = 0; u < 2; u++) {
            real out[4];
            pcl_quad_op<BIT, real>(out, a[u], b[u], w4 >> (4 * u));
            pcl_stq<GD, real>(dst + (i + u) * 128, out);
        }
        if (i + 4 < nq) {
#pragma unroll
            for (int u = 0; u < 2; u++) {
                pcl_ldq<GS, real>(src + (i + 4 + u) * 128, a[u]);
                pcl_ldq<GS, real>(src + (i + 4 + u + nq) * 128, b[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            real out[4];
            pcl_quad_op<BIT, real>(out, a2[u], b2[u], w4 >> (8 + 4 * u));
            pcl_stq<GD, real>(dst + (i + 2 + u) * 128, out);
        }
    }
}

// Level 2 of one path straight from the channel LLRs (level 1 is never stored).  Quad i4 of
// level 2 = elements 4 i4 + r; element k comes from the 4 channel values at 4 br(k), and
// br(4 i4 + r) = br(i4) + br2(r) * sz / 4: four sequential streams of 16-byte loads, one
// 16-byte store.  b1 = left array of level 1, b2 = left array of level 2 (nullptr: `smf`).
template <bool BIT1, bool BIT, bool GD, typename real>
PCL_DEVICE void pcl_level2_vec(real* dst, const real* y, int n, const uint32_t* b1, const uint32_t* b2, uint32_t smf)
{
    const int sz = 1 << (n - 2);
    const int nq = sz >> 2;
#pragma unroll 2
    for (int tl = 0; tl < nq; tl++) {
        const int i4 = (int)(__brev((unsigned)tl) >> (36 - n));    // (n-4)-bit reversal
        real yv[4][4];
#pragma unroll
        for (int mm = 0; mm < 4; mm++) pcl_load_quad_stream<real>(y + 4 * (tl + mm * nq), yv[mm]);
        const int k4 = 4 * i4;
        uint32_t n1 = 0, n2 = 0, n0 = 0;
        if (BIT1) {
            n1 = b1[(k4 >> 5) * 32] >> (k4 & 31);
            n2 = b1[((k4 + sz) >> 5) * 32] >> ((k4 + sz) & 31);
        }
        if (BIT) n0 = (b2 != nullptr) ? b2[(k4 >> 5) * 32] >> (k4 & 31) : smf >> k4;
        real out[4];
#pragma unroll
        for (int mm = 0; mm < 4; mm++) {
            const int r = ((mm & 1) << 1) | (mm >> 1);             // 2-bit reversal
            real a, b;
            if (BIT1) {
                a = pcl_fast<real>::gs(yv[mm][0], yv[mm][1], n1 << (31 - r));
                b = pcl_fast<real>::gs(yv[mm][2], yv[mm][3], n2 << (31 - r));
            } else {
                a = pcl_math<real>::f(yv[mm][0], yv[mm][1]);
                b = pcl_math<real>::f(yv[mm][2], yv[mm][3]);
            }
            out[r] = BIT ? pcl_fast<real>::gs(a, b, n0 << (31 - r)) : pcl_math<real>::f(a, b);
        }
        pcl_stq<GD, real>(dst + i4 * 128, out);
    }
}


// Level 3 of one path straight from the channel LLRs: neither level 1 nor level 2 is ever stored
// (TM variant).  Element k of level 3 comes from the EIGHT consecutive channel values at
// 8 br(k): pairs (0,1) (2,3) (4,5) (6,7) are level-1 elements k, k + 2s, k + s, k + 3s (s = N/8), the
// first two and the last two pairs make level-2 elements k and k + s.  Quad i4 = br(t) of level 3
// needs the 32-byte groups at 8 (t + mm N/32), mm = 0 .. 3: iteration t walks four sequential streams.
// The 8 (or 32) lanes of a frame need the same 128 bytes per iteration, so every lane fetches ONE
// 16-byte piece with cp.async into a ring in shared memory (the region of the levels that this very
// walk recomputes afterwards) and all lanes read the pieces back as broadcasts: an eighth of the
// global loads, DEPTH rounds in flight, no registers held across the wait.
template <int LP, bool BIT1, bool BIT2, bool BIT3>
PCL_DEVICE void pcl_level3_fused(float* dst, const float* yf, int n, float* stage, int lane,
                                 const uint32_t* b1, const uint32_t* b2, const uint32_t* b3)
{
    static_assert(LP >= 8, "a frame needs at least 8 lanes to fetch its 8 pieces");
    constexpr int R = LP / 8;                     // iterations one round of 32 pieces covers
    constexpr int DEPTH = 8;                      // rounds in the ring (8 x 512 bytes)
    const int nq3 = 1 << (n - 5);                 // quads of level 3
    const int sz3 = 4 * nq3;
    const int nrounds = nq3 / R;
    const int piece = lane & 7, grp8 = lane >> 3;
    const int sub = (lane & (LP - 1)) >> 3;       // which iteration of the round this lane fetches for
    const int fr = lane / LP;
    const float* src0 = yf + 8 * (sub + (piece >> 1) * nq3) + 4 * (piece & 1);
    float* mine = stage + grp8 * 32 + piece * 4;
#pragma unroll 1
    for (int rr = 0; rr < DEPTH - 1; rr++) {
        if (rr < nrounds) pcl_cp_async16(mine + rr * 128, src0 + 8 * R * rr);
        pcl_cp_async_commit();
    }
#pragma unroll 1
    for (int rr = 0; rr < nrounds; rr++) {
        pcl_cp_async_wait<DEPTH - 2>();           // this lane's piece of round rr has landed ...
        __syncwarp();                             // ... and so have the others'; round rr - 1 is consumed
        const int nx = rr + DEPTH - 1;
        if (nx < nrounds) pcl_cp_async16(mine + (nx % DEPTH) * 128, src0 + 8 * R * nx);
        pcl_cp_async_commit();
#pragma unroll
        for (int su = 0; su < R; su++) {
            const int t = rr * R + su;
            const float* sp = stage + (rr % DEPTH) * 128 + (fr * R + su) * 32;
            const int i4 = (int)(__brev((unsigned)t) >> (37 - n));          // (n-5)-bit reversal
            const int k4 = 4 * i4;
            const int sh = k4 & 31;
            uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0, l0 = 0, l1 = 0, t0 = 0;
            if (BIT1) {
                m0 = b1[(k4 >> 5) * 32] >> sh;
                m1 = b1[((k4 + sz3) >> 5) * 32] >> sh;
                m2 = b1[((k4 + 2 * sz3) >> 5) * 32] >> sh;
                m3 = b1[((k4 + 3 * sz3) >> 5) * 32] >> sh;
            }
            if (BIT2) {
                l0 = b2[(k4 >> 5) * 32] >> sh;
                l1 = b2[((k4 + sz3) >> 5) * 32] >> sh;
            }
            if (BIT3) t0 = b3[(k4 >> 5) * 32] >> sh;
            float out[4];
#pragma unroll
            for (int mm = 0; mm < 4; mm++) {
                const int r = ((mm & 1) << 1) | (mm >> 1);                  // 2-bit reversal
                float v[8];
                pcl_load_quad<float>(sp + 8 * mm, v);
                pcl_load_quad<float>(sp + 8 * mm + 4, v + 4);
                float p01, p23, p45, p67;
                if (BIT1) {
                    p01 = pcl_fast<float>::gs(v[0], v[1], m0 << (31 - r));
                    p23 = pcl_fast<float>::gs(v[2], v[3], m2 << (31 - r));
                    p45 = pcl_fast<float>::gs(v[4], v[5], m1 << (31 - r));
                    p67 = pcl_fast<float>::gs(v[6], v[7], m3 << (31 - r));
                } else {
                    p01 = pcl_math<float>::f(v[0], v[1]);
                    p23 = pcl_math<float>::f(v[2], v[3]);
                    p45 = pcl_math<float>::f(v[4], v[5]);
                    p67 = pcl_math<float>::f(v[6], v[7]);
                }
                const float q0 = BIT2 ? pcl_fast<float>::gs(p01, p23, l0 << (31 - r)) : pcl_math<float>::f(p01, p23);
                const float q1 = BIT2 ? pcl_fast<float>::gs(p45, p67, l1 << (31 - r)) : pcl_math<float>::f(p45, p67);
                out[r] = BIT3 ? pcl_fast<float>::gs(q0, q1, t0 << (31 - r)) : pcl_math<float>::f(q0, q1);
            }
            pcl_stq<true, float>(dst + i4 * 128, out);
        }
    }
    pcl_cp_async_wait<0>();
    __syncwarp();                                 // the ring region goes back to the level that owns it
}

// ---- TM variant (fp32, compiled code length, n >= 9): the three levels below the global scratch
// stay on chip.  Level n-6 (64 elements per path) and level n-4 (16) live in TENSOR MEMORY, one
// TMEM lane per path (pcl_tmem.cuh); level n-5 (32) in shared memory.  A path always writes its own
// lane; a path that has to read another slot's array (the first level of a walk after a prune
// moved it) lets every lane load its own array and fetches the values with shuffles.
template <bool BIT>
PCL_DEVICE void pcl_level_g2t(uint32_t tdst, const float* src, const uint32_t* bsrc)
{
    constexpr int nq = 16;                        // destination quads; source quads i and i + 16
    float a[2][4], b[2][4], a2[2][4], b2[2][4];
    uint32_t wb = 0;
#pragma unroll
    for (int u = 0; u < 2; u++) {
        pcl_ldq<true, float>(src + u * 128, a[u]);
        pcl_ldq<true, float>(src + (u + nq) * 128, b[u]);
    }
#pragma unroll 2
    for (int i = 0; i < nq; i += 4) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            pcl_ldq<true, float>(src + (i + 2 + u) * 128, a2[u]);
            pcl_ldq<true, float>(src + (i + 2 + u + nq) * 128, b2[u]);
        }
        if (BIT && (i & 7) == 0) wb = bsrc[(i >> 3) * 32];
        const uint32_t w4 = wb >> (4 * (i & 7));
        {
            float out[8];
            uint32_t o[8];
#pragma unroll
            for (int u = 0; u < 2; u++) pcl_quad_op<BIT, float>(out + 4 * u, a[u], b[u], w4 >> (4 * u));
#pragma unroll
            for (int k = 0; k < 8; k++) o[k] = __float_as_uint(out[k]);
            pcl_tmem_st<8>(tdst + 4 * i, o);
        }
        if (i + 4 < nq) {
#pragma unroll
            for (int u = 0; u < 2; u++) {
                pcl_ldq<true, float>(src + (i + 4 + u) * 128, a[u]);
                pcl_ldq<true, float>(src + (i + 4 + u + nq) * 128, b[u]);
            }
        }
        {
            float out[8];
            uint32_t o[8];
#pragma unroll
            for (int u = 0; u < 2; u++) pcl_quad_op<BIT, float>(out + 4 * u, a2[u], b2[u], w4 >> (8 + 4 * u));
#pragma unroll
            for (int k = 0; k < 8; k++) o[k] = __float_as_uint(out[k]);
            pcl_tmem_st<8>(tdst + 4 * i + 8, o);
        }
    }
    pcl_tmem_wait_st();
}

// level n-5 (32 elements, shared memory, own column) from level n-6 in tensor memory
template <bool BIT>
PCL_DEVICE void pcl_level_t2s(float* dst, uint32_t tsrc, int srcl, bool ident, uint32_t wb)
{
#pragma unroll
    for (int h = 0; h < 4; h++) {
        uint32_t a[8], b[8];
        pcl_tmem_ld<8>(tsrc + 8 * h, a);
        pcl_tmem_ld<8>(tsrc + 32 + 8 * h, b);
        pcl_tmem_wait_ld();
        if (!ident) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                a[k] = __shfl_sync(PCL_FULL_MASK, a[k], srcl);
                b[k] = __shfl_sync(PCL_FULL_MASK, b[k], srcl);
            }
        }
#pragma unroll
        for (int qd = 0; qd < 2; qd++) {
            float fa[4], fb[4], out[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                fa[e] = __uint_as_float(a[4 * qd + e]);
                fb[e] = __uint_as_float(b[4 * qd + e]);
            }
            pcl_quad_op<BIT, float>(out, fa, fb, wb >> (8 * h + 4 * qd));
            pcl_store_quad<float>(dst + (2 * h + qd) * 128, out);
        }
    }
}

// level n-4 (16 elements, tensor memory) from level n-5 in shared memory (column of the source slot)
template <bool BIT>
PCL_DEVICE void pcl_level_s2t(uint32_t tdst, const float* src, uint32_t smf)
{
    float a[4][4], b[4][4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        pcl_load_quad<float>(src + u * 128, a[u]);
        pcl_load_quad<float>(src + (u + 4) * 128, b[u]);
    }
    uint32_t o[16];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        float out[4];
        pcl_quad_op<BIT, float>(out, a[u], b[u], smf >> (4 * u));
#pragma unroll
        for (int e = 0; e < 4; e++) o[4 * u + e] = __float_as_uint(out[e]);
    }
    pcl_tmem_st<16>(tdst, o);
    pcl_tmem_wait_st();
}

// SC (list size 1) without a metric: the 8 leaves of a block as one unrolled recursion -- no leaf loop, no
// softplus, a frozen subtree costs nothing (same rules as polar_sc.cuh: f, g, bit = [x < 0]).  Returns the
// node's partial sums, u = its decisions.
template <typename real, int SZ>
struct pcl_sc_node {
    static PCL_DEVICE uint32_t run(const real* a, uint32_t fz, uint32_t& u)
    {
        constexpr uint32_t FULL = (1u << SZ) - 1u;
        constexpr int H = SZ / 2;
        constexpr uint32_t HALF = (1u << H) - 1u;
        if ((fz & FULL) == FULL) { u = 0; return 0; }
        real t[H];
        uint32_t ul = 0, ur = 0, cl = 0;
        if ((fz & HALF) != HALF) {
#pragma unroll
            for (int k = 0; k < H; k++) t[k] = pcl_math<real>::f(a[k], a[k + H]);
            cl = pcl_sc_node<real, H>::run(t, fz & HALF, ul);
        }
#pragma unroll
        for (int k = 0; k < H; k++) t[k] = pcl_fast<real>::gs(a[k], a[k + H], cl << (31 - k));
        const uint32_t cr = pcl_sc_node<real, H>::run(t, (fz >> H) & HALF, ur);
        u = ul | (ur << H);
        return (cl ^ cr) | (cr << H);
    }
};
template <typename real>
struct pcl_sc_node<real, 1> {
    static PCL_DEVICE uint32_t run(const real* a, uint32_t fz, uint32_t& u)
    {
        u = (fz & 1u) ? 0u : (!(a[0] >= (real)0) ? 1u : 0u);     // decoder.py:58-66
        return u;
    }
};

// LP = list slots per frame (power of two); a warp decodes FPW = 32 / LP frames side by side.
// NL = log2 N as a compile-time constant (0: read it from the layout), GL = G for that NL.
#ifndef PCL_PRUNE_SHORTCUT
#define PCL_PRUNE_SHORTCUT 1
#endif
#ifndef PCL_PRUNE_QUICK
#define PCL_PRUNE_QUICK 1
#endif
#ifndef PCL_LEAF_V2
#define PCL_LEAF_V2 1         // 32-bit quick test for in-place survivors, lean in-place update
#endif
#ifndef PCL_FROZEN_BLOCK
#define PCL_FROZEN_BLOCK 1    // an all-frozen block of 8 leaves as straight-line code (no leaf loop)
#endif
#ifndef PCL_RAW_BLOCK_BITS
#define PCL_RAW_BLOCK_BITS 1  // decisions of the current block as raw bits 24 .. 31 of `small` (one OR per leaf); 0: folded fields per leaf
#endif
#ifndef PCL_SC_BLOCK
#define PCL_SC_BLOCK 1        // list size 1 without a metric: unrolled 8-leaf recursion instead of the leaf loop
#endif
#ifndef PCL_POLAR_MINB
#define PCL_POLAR_MINB 6      // resident 128-thread blocks per SM the register allocation aims for (80 regs)
#endif
#ifndef PCL_POLAR_TM_GROUP
#define PCL_POLAR_TM_GROUP 2      // warps that pull a chunk of frames together and stay in step (2: 10.59, 4: 10.47 Gbps)
#endif
#ifndef PCL_POLAR_TM_THREADS
#define PCL_POLAR_TM_THREADS 640   // TM variant: ONE block per SM, 20 warps = 5 groups of 4 (96 registers; 768 / 80 is 7 % slower)
#endif
template <int LP, typename real, int NL, int GL, int TM = 0>
__global__ void __launch_bounds__(TM ? PCL_POLAR_TM_THREADS : 128, TM ? 1 : ((sizeof(real) == 4) ? PCL_POLAR_MINB : 3))
polar_scl_fast_kernel(PolarParams<real> P)
{
    static_assert(!TM || (NL >= 9 && GL == NL - 7 && sizeof(real) == 4), "TM variant: fp32, compiled code length >= 512");
    constexpr int PB = pcl_log2<LP>::v;
    constexpr int FPW = 32 / LP;                 // frames per warp
    constexpr int NC = 2 * LP;                   // prune candidates per frame
    constexpr bool EXACT = sizeof(real) == 8;
    // per-leaf LLR / parent dumps (SCDecoder.L, the debug scripts) only in the run-time-N variants: a handle
    // with a compiled code length keeps a run-time-N twin for those calls (pcl_api.cu), and the production
    // leaf loop carries no test for them
    constexpr bool DBG = (NL == 0);
    constexpr int FB = pcl_fast_frame_bytes(LP);
    const PolarLayout& Y = P.lay;
    const int n = NL ? NL : Y.n;
    const int N = 1 << n;
    const int G = NL ? GL : Y.G;
    const int nb = n > 5 ? n - 5 : 0;
    const int NW = N >= 32 ? N >> 5 : 1;
    const int K = Y.K, L = Y.L;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int p = lane & (LP - 1);
    const int fr = lane >> PB;
    const int cbase = lane - p;                   // first column of this lane's frame
    const int shift = (N < 32) ? 32 - N : 0;
    const int NB = N >> 3;

    unsigned char* wsm = pcl_dyn_smem() + (TM ? Y.hdr_bytes : 0) + (size_t)warp * Y.warp_bytes;
    unsigned char* fb = wsm + Y.off_cm + fr * FB; // this frame's prune scratch
    double* cm = (double*)fb;                     // candidate keys: 2p -> bit 0, 2p + 1 -> bit 1 of path p
    double* newpm = (double*)(fb + NC * 8);       // survivor keys (fp32 build) / metrics, by rank
    int* sel = (int*)(fb + NC * 8 + LP * 8);      // survivor candidate ids (fp64 build)
    real* sl = (real*)(wsm + Y.off_llr);          // levels max(G,1)+1 .. n-4
    uint32_t* bw = (uint32_t*)(wsm + Y.off_bw);   // big left levels 1 .. nb, [w][col]
    uint32_t* uw = (uint32_t*)(wsm + Y.off_uw);   // final u words
    real* gl = P.scratch + (int64_t)(blockIdx.x * wpb + warp) * Y.scratch_per_warp;

    const double DEAD = -1.0e300;                 // metric of an inactive slot (sorts last, stays finite)

    // TM variant: one block per SM owns the SM's tensor memory; warp w works in lane quarter w % 4,
    // columns (w / 4) * TMW .. : level n-6 at +0 (64 columns), level n-4 at +64 (16 columns).
    // Groups of 4 warps pull chunks of 4 FPW frames from a ticket counter and meet at their own
    // named barrier (the whole-block barrier of the other variant would couple all 24 warps).
    uint32_t tL4 = 0, tL6 = 0;
    const int grp = warp / PCL_POLAR_TM_GROUP, gw = warp % PCL_POLAR_TM_GROUP;
    uint32_t* hdr = (uint32_t*)pcl_dyn_smem();    // [0] TMEM base, [4 + 4 g + 2 k ..] ticket of group g, parity k (64 bit)
    if (TM) {
        if (warp == 0) pcl_tmem_alloc_all(hdr);
        pcl_tmem_fence_before();
        __syncthreads();
        pcl_tmem_fence_after();
        const int tmw = (PCL_TMEM_COLS / ((wpb + 3) >> 2)) & ~15;
        tL4 = pcl_tmem_addr(hdr[0], warp, (warp >> 2) * tmw);
        tL6 = tL4 + 64;
    }

    // Every warp of a block makes the same number of passes (a pass past the end of the batch
    // works on masked-off frames) and the block meets at a barrier before each pass: all frames
    // follow the same instruction stream, so warps that start together stay close to the same
    // tree position and share their instruction working set (measured +3 %; barriers inside the
    // pass, every 1 .. 64 blocks of 8 leaves, gain less the more often they come).
    int64_t fbase = (int64_t)blockIdx.x * wpb * FPW;
    for (int pass = 0;; pass++) {
        if (TM) {
            unsigned long long* tk = (unsigned long long*)(hdr + 4) + 2 * grp + (pass & 1);
            if (gw == 0 && lane == 0) *tk = atomicAdd(P.next, 1ull) - P.ticket_base;
            pcl_named_barrier(1 + grp, 32 * PCL_POLAR_TM_GROUP);
            fbase = (int64_t)(*(volatile unsigned long long*)tk) * (PCL_POLAR_TM_GROUP * FPW);
            if (fbase >= P.F) break;
        } else {
            if (pass) fbase += (int64_t)gridDim.x * wpb * FPW;
            if (fbase >= P.F) break;
            __syncthreads();
        }
        const int64_t f0 = fbase + (int64_t)(TM ? gw : warp) * FPW;
        const int64_t f = f0 + fr;
        const bool valid = f < P.F;
        const real* y = P.llr + (valid ? f : fbase) * N;
        int nact = 1;
        bool act = (p == 0) && valid;
        double pm = act ? 0.0 : DEAD;
        uint32_t ptrL = 0, ptrB = 0;              // packed slot pointers: LLR levels / left levels
        uint32_t small = 0, ulast = 0;
        uint32_t fw = 0;

        for (int blk = 0; blk < NB; blk++) {
            const int i0 = blk << 3;
            if (((i0 + shift) & 31) == 0 || blk == 0) fw = P.frozen_words[(i0 + shift) >> 5];
            const uint32_t fz8 = (fw >> ((i0 + shift) & 31)) & 0xffu;

            // ---- levels above the cut: start .. n-3; the last one lands in registers ----
            real R3[8];
#pragma unroll
            for (int t = 0; t < 8; t++) R3[t] = (real)0;
            const int bit3 = (i0 >> 3) & 1;       // f or g at level n-3
            if constexpr (TM != 0) {
                const int start = (blk == 0) ? 2 : n - (__ffs(i0) - 1);     // <= n-3
                // Levels 1 and 2 are never stored: level 3 comes straight from the channel LLRs, every
                // lane of the warp fetching a share of them (pcl_level3_fused); global level d >= 3 of
                // the scratch sits at 32 (N/4 - N >> (d-1)).
                if (start <= 3) {
                    const int bit1 = (i0 >> (n - 1)) & 1, bit2 = (i0 >> (n - 2)) & 1, bit3l = (i0 >> (n - 3)) & 1;
                    const uint32_t* b1src = bw + cbase + (ptrB & (LP - 1));
                    const uint32_t* b2src = bw + 32 * ((N >> 5) - (N >> 6)) + cbase + ((ptrB >> PB) & (LP - 1));
                    const uint32_t* b3src = bw + 32 * ((N >> 5) - (N >> 7)) + cbase + ((ptrB >> (2 * PB)) & (LP - 1));
                    float* dst = (float*)gl + 4 * lane;
                    float* ring = (float*)sl;
                    const float* yf = (const float*)y;
                    switch (bit1 * 4 + bit2 * 2 + bit3l) {
                        case 0: pcl_level3_fused<LP, false, false, false>(dst, yf, n, ring, lane, b1src, b2src, b3src); break;
                        case 1: pcl_level3_fused<LP, false, false, true>(dst, yf, n, ring, lane, b1src, b2src, b3src); break;
                        case 2: pcl_level3_fused<LP, false, true, false>(dst, yf, n, ring, lane, b1src, b2src, b3src); break;
                        case 3: pcl_level3_fused<LP, false, true, true>(dst, yf, n, ring, lane, b1src, b2src, b3src); break;
                        case 4: pcl_level3_fused<LP, true, false, false>(dst, yf, n, ring, lane, b1src, b2src, b3src); break;
                        case 5: pcl_level3_fused<LP, true, false, true>(dst, yf, n, ring, lane, b1src, b2src, b3src); break;
                        case 6: pcl_level3_fused<LP, true, true, false>(dst, yf, n, ring, lane, b1src, b2src, b3src); break;
                        default: pcl_level3_fused<LP, true, true, true>(dst, yf, n, ring, lane, b1src, b2src, b3src); break;
                    }
                    ptrL = (ptrL & ~((uint32_t)(LP - 1) << (2 * PB))) | ((uint32_t)p << (2 * PB));
                }
#pragma unroll
                for (int d = 4; d <= n - 7; d++) {                          // global scratch -> global scratch
                    if (d >= start && act) {
                        const int nq = (N >> d) >> 2;
                        const int bit = (i0 >> (n - d)) & 1;
                        const int q = (ptrL >> ((d - 2) * PB)) & (LP - 1);
                        const uint32_t* bsrc = bw + 32 * ((N >> 5) - (N >> (d + 4))) + cbase + ((ptrB >> ((d - 1) * PB)) & (LP - 1));
                        const real* src = gl + 32 * ((N >> 2) - (N >> (d - 2))) + 4 * (cbase + q);
                        real* dst = gl + 32 * ((N >> 2) - (N >> (d - 1))) + 4 * lane;
                        if (bit) pcl_level_vec<true, true, true, real>(dst, src, nq, bsrc, 0u);
                        else pcl_level_vec<false, true, true, real>(dst, src, nq, bsrc, 0u);
                        ptrL = (ptrL & ~((uint32_t)(LP - 1) << ((d - 1) * PB))) | ((uint32_t)p << ((d - 1) * PB));
                    }
                }
                // The on-chip levels run on every lane of the warp: tcgen05.ld / st are warp-wide, an
                // inactive lane computes on whatever its slot holds and nobody ever reads its result.
                if (n - 6 >= start) {                                       // level n-6: global -> tensor memory
                    constexpr int d = NL - 6;
                    const int bit = (i0 >> (n - d)) & 1;
                    const int q = (ptrL >> ((d - 2) * PB)) & (LP - 1);
                    const uint32_t* bsrc = bw + 32 * ((N >> 5) - (N >> (d + 4))) + cbase + ((ptrB >> ((d - 1) * PB)) & (LP - 1));
                    const float* src = (const float*)gl + 32 * ((N >> 2) - (N >> (d - 2))) + 4 * (cbase + q);
                    if (bit) pcl_level_g2t<true>(tL4, src, bsrc);
                    else pcl_level_g2t<false>(tL4, src, bsrc);
                    ptrL = (ptrL & ~((uint32_t)(LP - 1) << ((d - 1) * PB))) | ((uint32_t)p << ((d - 1) * PB));
                }
                if (n - 5 >= start) {                                       // level n-5: tensor memory -> shared memory
                    constexpr int d = NL - 5;
                    const int bit = (i0 >> (n - d)) & 1;
                    const int q = (ptrL >> ((d - 2) * PB)) & (LP - 1);
                    const bool ident = __all_sync(PCL_FULL_MASK, !act || q == p);
                    const uint32_t wb = bw[32 * ((N >> 5) - (N >> (d + 4))) + cbase + ((ptrB >> ((d - 1) * PB)) & (LP - 1))];
                    if (bit) pcl_level_t2s<true>((float*)sl + 4 * lane, tL4, cbase + q, ident, wb);
                    else pcl_level_t2s<false>((float*)sl + 4 * lane, tL4, cbase + q, ident, 0u);
                    ptrL = (ptrL & ~((uint32_t)(LP - 1) << ((d - 1) * PB))) | ((uint32_t)p << ((d - 1) * PB));
                }
                if (n - 4 >= start) {                                       // level n-4: shared memory -> tensor memory
                    constexpr int d = NL - 4;
                    const int bit = (i0 >> (n - d)) & 1;
                    const int q = (ptrL >> ((d - 2) * PB)) & (LP - 1);
                    const float* src = (const float*)sl + 4 * (cbase + q);
                    if (bit) pcl_level_s2t<true>(tL6, src, small);
                    else pcl_level_s2t<false>(tL6, src, small);
                    ptrL = (ptrL & ~((uint32_t)(LP - 1) << ((d - 1) * PB))) | ((uint32_t)p << ((d - 1) * PB));
                }
                {                                                           // level n-3 (registers) from level n-4
                    const int q = (ptrL >> ((n - 5) * PB)) & (LP - 1);
                    const bool ident = __all_sync(PCL_FULL_MASK, !act || q == p);
                    uint32_t v[16];
                    pcl_tmem_ld<16>(tL6, v);
                    pcl_tmem_wait_ld();
                    if (!ident) {
#pragma unroll
                        for (int k = 0; k < 16; k++) v[k] = __shfl_sync(PCL_FULL_MASK, v[k], cbase + q);
                    }
                    if (bit3) {
#pragma unroll
                        for (int t = 0; t < 8; t++)
                            R3[t] = (real)pcl_fast<float>::gs(__uint_as_float(v[t]), __uint_as_float(v[t + 8]), small << (15 - t));
                    } else {
#pragma unroll
                        for (int t = 0; t < 8; t++) R3[t] = (real)pcl_math<float>::f(__uint_as_float(v[t]), __uint_as_float(v[t + 8]));
                    }
                }
            } else if (n >= 6) {
                const int start = (blk == 0) ? 2 : n - (__ffs(i0) - 1);     // <= n-3
                const uint32_t* b1src = bw + cbase + (ptrB & (LP - 1));     // left array of level 1
                if (start <= 2) {
                    if (act) {
                        const int bit1 = (i0 >> (n - 1)) & 1;
                        const int bit = (i0 >> (n - 2)) & 1;
                        const uint32_t* b2src = (2 <= nb) ? bw + 32 * ((N >> 5) - (N >> 6)) + cbase + ((ptrB >> PB) & (LP - 1))
                                                          : nullptr;
                        const uint32_t smf = small;                         // n == 6: level 2 has 16 elements
                        auto run2 = [&](auto gd, real* dst) {
                            constexpr bool GD = decltype(gd)::value;
                            if (bit1) {
                                if (bit) pcl_level2_vec<true, true, GD, real>(dst, y, n, b1src, b2src, smf);
                                else pcl_level2_vec<true, false, GD, real>(dst, y, n, b1src, b2src, smf);
                            } else {
                                if (bit) pcl_level2_vec<false, true, GD, real>(dst, y, n, b1src, b2src, smf);
                                else pcl_level2_vec<false, false, GD, real>(dst, y, n, b1src, b2src, smf);
                            }
                        };
                        if (2 <= G) run2(pcl_true(), gl + 4 * lane);
                        else run2(pcl_false(), sl + 4 * lane);
                        ptrL = (ptrL & ~((uint32_t)(LP - 1) << PB)) | ((uint32_t)p << PB);
                    }
                }
#pragma unroll(NL ? 16 : 1)
                for (int d = 3; d <= n - 4; d++) {
                    if (d >= start && act) {
                        const int sz = N >> d;
                        const int nq = sz >> 2;
                        const int bit = (i0 >> (n - d)) & 1;
                        const int q = (ptrL >> ((d - 2) * PB)) & (LP - 1);
                        const uint32_t* bsrc = (d <= nb) ? bw + 32 * ((N >> 5) - (N >> (d + 4))) + cbase +
                                                               ((ptrB >> ((d - 1) * PB)) & (LP - 1))
                                                         : nullptr;
                        const uint32_t smf = small;                         // d == n-4: 16 elements at bits 0..15
                        const int soff = 4 * (cbase + q), doff = 4 * lane;
                        if (d <= G) {
                            const real* src = gl + 32 * ((N >> 1) - (N >> (d - 2))) + soff;
                            real* dst = gl + 32 * ((N >> 1) - (N >> (d - 1))) + doff;
                            if (bit) pcl_level_vec<true, true, true, real>(dst, src, nq, bsrc, smf);
                            else pcl_level_vec<false, true, true, real>(dst, src, nq, bsrc, smf);
                        } else if (d - 1 <= G) {
                            const real* src = gl + 32 * ((N >> 1) - (N >> (d - 2))) + soff;
                            real* dst = sl + 32 * ((N >> G) - (N >> (d - 1))) + doff;
                            if (bit) pcl_level_vec<true, true, false, real>(dst, src, nq, bsrc, smf);
                            else pcl_level_vec<false, true, false, real>(dst, src, nq, bsrc, smf);
                        } else {
                            const real* src = sl + 32 * ((N >> G) - (N >> (d - 2))) + soff;
                            real* dst = sl + 32 * ((N >> G) - (N >> (d - 1))) + doff;
                            if (bit) pcl_level_vec<true, false, false, real>(dst, src, nq, bsrc, smf);
                            else pcl_level_vec<false, false, false, real>(dst, src, nq, bsrc, smf);
                        }
                        ptrL = (ptrL & ~((uint32_t)(LP - 1) << ((d - 1) * PB))) | ((uint32_t)p << ((d - 1) * PB));
                    }
                }
                if (act) {                        // level n-3 from the four quads of level n-4
                    const int q = (ptrL >> ((n - 5) * PB)) & (LP - 1);
                    real qa[2][4], qb[2][4];
                    if (n - 4 <= G) {
                        const real* src = gl + 32 * ((N >> 1) - 32) + 4 * (cbase + q);
#pragma unroll
                        for (int u = 0; u < 2; u++) {
                            pcl_ldq<true, real>(src + u * 128, qa[u]);
                            pcl_ldq<true, real>(src + (u + 2) * 128, qb[u]);
                        }
                    } else {
                        const real* src = sl + 32 * ((N >> G) - 32) + 4 * (cbase + q);
#pragma unroll
                        for (int u = 0; u < 2; u++) {
                            pcl_ldq<false, real>(src + u * 128, qa[u]);
                            pcl_ldq<false, real>(src + (u + 2) * 128, qb[u]);
                        }
                    }
                    if (bit3) {
#pragma unroll
                        for (int t = 0; t < 8; t++) R3[t] = pcl_fast<real>::gs(qa[t >> 2][t & 3], qb[t >> 2][t & 3], small << (15 - t));
                    } else {
#pragma unroll
                        for (int t = 0; t < 8; t++) R3[t] = pcl_math<real>::f(qa[t >> 2][t & 3], qb[t >> 2][t & 3]);
                    }
                }
            } else if (act) {
                // N = 16 / 32: the block root comes straight from the channel (level 1) or from
                // two level-1 values recomputed on the fly (level 2)
                const int bit1 = (i0 >> (n - 1)) & 1;
                auto lvl1 = [&](int m) -> real {
                    const int r = (int)(__brev((unsigned)m) >> (32 - n));
                    real y0, y1;
                    pcl_load_pair<real>(y + r, y0, y1);
                    if (bit1) return pcl_fast<real>::g(y0, y1, (small >> ((32 - N + m) & 31)) & 1u);
                    return pcl_math<real>::f(y0, y1);
                };
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    real a, b;
                    if (n == 4) {
                        const int r = (int)(__brev((unsigned)t) >> (32 - n));
                        a = y[r];
                        b = y[r + 1];
                    } else {
                        a = lvl1(t);
                        b = lvl1(t + 8);
                    }
                    R3[t] = bit3 ? pcl_fast<real>::g(a, b, (small >> (16 + t)) & 1u) : pcl_math<real>::f(a, b);
                }
            }
            // all borrowed source arrays have been read: order before later overwrites
            __syncwarp();

            // Partial sums of the block being decoded.  PCL_RAW_BLOCK_BITS: leaf j ORs its decision into bit 24 + j of
            // `small` (bit 31 = leaf 7 = what the final u-word wants there) and the three places that need folded
            // sums -- g at leaves 2 / 6, g at leaf 4, the block's own sums c8 at its end -- fold them there, once;
            // a survivor inherits the raw bits with `small`.  (Before: the fields of sizes 1, 2, 4 at bits 30, 28, 24
            // were re-folded after every leaf behind three tests of j, 14 instructions per leaf.)
            uint32_t c4hi = 0;                    // folded fields only: partial sums of leaves 4 .. 7
            uint32_t c8 = 0;                      // the block's 8 partial sums, folded upwards below
            if (PCL_FROZEN_BLOCK && !EXACT && LP > 1 && fz8 == 0xffu && (!DBG || P.dbg_leaf == nullptr)) {
                // ---- all-frozen block (46 of the 128 blocks of the headline code): no decisions, every
                // g is a plain sum, so the 8 leaf LLRs are one straight-line butterfly and the 8 penalties
                // are independent chains; added to the metric in leaf order (same sums as leaf by leaf).
                real xs[8];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    real H[4];
#pragma unroll
                    for (int t = 0; t < 4; t++) H[t] = h ? R3[t] + R3[t + 4] : pcl_math<real>::f(R3[t], R3[t + 4]);
                    const real c0 = pcl_math<real>::f(H[0], H[2]), c1 = pcl_math<real>::f(H[1], H[3]);
                    const real d0 = H[0] + H[2], d1 = H[1] + H[3];
                    xs[4 * h + 0] = pcl_math<real>::f(c0, c1);
                    xs[4 * h + 1] = c0 + c1;
                    xs[4 * h + 2] = pcl_math<real>::f(d0, d1);
                    xs[4 * h + 3] = d0 + d1;
                }
                double pmn = pm;
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const real ax = fabs(xs[t]);
                    const real pen = pcl_fast<real>::softplus_neg_abs(ax) + (!(xs[t] >= (real)0) ? ax : (real)0);
                    pmn -= (double)pen;
                }
                pm = act ? pmn : pm;              // an inactive slot keeps DEAD whatever its registers hold
#if PCL_RAW_BLOCK_BITS
                small &= 0x00ffffffu;             // eight zero decisions
#else
                small &= ~(127u << 24);           // fields of sizes 4, 2, 1: all zero
                ulast = 0;
#endif
            } else if (PCL_SC_BLOCK && LP == 1 && !P.want_pm && (!DBG || P.dbg_leaf == nullptr)) {
                // ---- SC, bits only: the block as an unrolled recursion (pcl_sc_node) -----------------
                uint32_t u8 = 0;
                c8 = pcl_sc_node<real, 8>::run(R3, fz8, u8);
#if PCL_RAW_BLOCK_BITS
                small = (small & 0x00ffffffu) | (u8 << 24);
#else
                c4hi = c8 >> 4;
                // fields as the leaf loop leaves them: size 4 = sums of leaves 0 .. 3, size 2 = (u4 ^ u5, u5), size 1 = u6
                small = (small & ~(127u << 24)) | (((c8 ^ (c8 >> 4)) & 15u) << 24) |
                        ((((u8 >> 4) ^ (u8 >> 5)) & 1u) << 28) | (((u8 >> 5) & 1u) << 29) | (((u8 >> 6) & 1u) << 30);
                ulast = (u8 >> 7) & 1u;
#endif
            } else {
            // ---- the 8 leaves of the block (rolled: the body must stay I-cache resident) ----
            real R2[4], R1[2];
#pragma unroll
            for (int t = 0; t < 4; t++) R2[t] = (real)0;
            R1[0] = R1[1] = (real)0;
#if PCL_RAW_BLOCK_BITS
            small &= 0x00ffffffu;
#endif
#pragma unroll 1
            for (int j = 0; j < 8; j++) {
                // heights to recompute: j == 0 -> 2,1,0; else ctz(j) .. 0; f or g per bit of j.
                // small: size-4 field at bits 24..27, size-2 at 28..29, size-1 at 30
                real x;
                if (j & 1) {
#if PCL_RAW_BLOCK_BITS
                    x = pcl_fast<real>::gs(R1[0], R1[1], small << (8 - j));          // u[j-1] sits at bit 23 + j
#else
                    x = pcl_fast<real>::gs(R1[0], R1[1], small << 1);
#endif
                } else {
                    if ((j & 2) == 0) {
                        if (j & 4) {
#if PCL_RAW_BLOCK_BITS
                            uint32_t c4 = small >> 24;                               // u0 .. u3 -> their partial sums
                            c4 ^= (c4 >> 1) & 5u;
                            c4 ^= (c4 >> 2) & 3u;
#pragma unroll
                            for (int t = 0; t < 4; t++) R2[t] = pcl_fast<real>::gs(R3[t], R3[t + 4], c4 << (31 - t));
#else
#pragma unroll
                            for (int t = 0; t < 4; t++) R2[t] = pcl_fast<real>::gs(R3[t], R3[t + 4], small << (7 - t));
#endif
                        } else {
#pragma unroll
                            for (int t = 0; t < 4; t++) R2[t] = pcl_math<real>::f(R3[t], R3[t + 4]);
                        }
                        R1[0] = pcl_math<real>::f(R2[0], R2[2]);
                        R1[1] = pcl_math<real>::f(R2[1], R2[3]);
                    } else {
#if PCL_RAW_BLOCK_BITS
                        const uint32_t s1 = small << (8 - j);                        // u[j-1] at bit 31
                        R1[0] = pcl_fast<real>::gs(R2[0], R2[2], s1 ^ (small << (9 - j)));   // u[j-2] ^ u[j-1]
                        R1[1] = pcl_fast<real>::gs(R2[1], R2[3], s1);
#else
                        R1[0] = pcl_fast<real>::gs(R2[0], R2[2], small << 3);
                        R1[1] = pcl_fast<real>::gs(R2[1], R2[3], small << 2);
#endif
                    }
                    x = pcl_math<real>::f(R1[0], R1[1]);
                }
                if (!act) x = (real)0;

                // ---- leaf decision (polar_scl.cuh for the rules) -------------------------
                const real ax = fabs(x);
                const bool hard = !(x >= (real)0);
                const real sp = pcl_fast<real>::softplus_neg_abs(ax);
                uint32_t u = 0;
                int parent = p;
                if (LP == 1) {
                    u = ((fz8 >> j) & 1u) ? 0u : (hard ? 1u : 0u);
                    if (P.want_pm) pm -= (double)(sp + ((u != (uint32_t)hard) ? ax : (real)0));
                } else if ((fz8 >> j) & 1u) {
                    if ((j & 1) == 0 && ((fz8 >> j) & 3u) == 3u && (!DBG || P.dbg_leaf == nullptr)) {
                    // Frozen pair (2 j, 2 j + 1): the right leaf needs no decision of the left one
                    // (u = 0: x1 = R1[0] + R1[1]), so both penalties are evaluated side by side and the
                    // loop moves on by two leaves; same operations in the same order as leaf by leaf.
                    real x1 = R1[0] + R1[1];
                    if (!act) x1 = (real)0;
                    const real ax1 = fabs(x1);
                    const real sp1 = pcl_fast<real>::softplus_neg_abs(ax1);
                    pm -= (double)(sp + (hard ? ax : (real)0));
                    pm -= (double)(sp1 + (!(x1 >= (real)0) ? ax1 : (real)0));
#if !PCL_RAW_BLOCK_BITS
                    small &= ~(1u << 30);
#endif
                    j++;
                    } else {
                    pm -= (double)(sp + (hard ? ax : (real)0));    // DEAD absorbs the penalty
                    }
                } else {
                    const int ns = (2 * nact < L) ? 2 * nact : L;
                    int ra = 0, rb = 0;
                    double ka, kb;
                    if (EXACT) {
                        // decoder.py:391-406: PM + (-(|x| + log1p(exp(-|x|)))) for the unlikely bit
                        const double base = pm - (double)sp, other = pm - ((double)ax + (double)sp);
                        ka = hard ? other : base;                  // bit 0
                        kb = hard ? base : other;                  // bit 1
                        if (!act) { ka = DEAD; kb = DEAD; }
                    } else {
                        // ka: the likely bit (u = hard), kb: the other one.  The ids (bit, parent) go into the lowest
                        // mantissa bits only when the list really has to be ranked (below): the in-place test reads the
                        // high words, and a path that continues in place keeps its metric untouched.
                        ka = pm - (double)sp;
                        kb = ka - (double)ax;
                    }
                    bool in_place = false;        // every survivor is its own parent's likely bit, same slot
                    if (!EXACT && PCL_PRUNE_QUICK && nact >= L) {
                        // Reliable bit on a full list (the common case): the likely keys are still in
                        // slot order and every unlikely key lies below the last of them, so the ranks
                        // are the slots themselves -- two shuffled compares instead of the ranking.
#if PCL_LEAF_V2
                        // Decided on the HIGH words alone (sign, exponent, 20 mantissa bits): every key is
                        // negative (at most one likely key of a frame can be +0, and then its word is the
                        // smallest), so a strictly larger unsigned high word means a strictly smaller key.
                        // Equal high words say nothing: the list takes the full ranking below, which gives
                        // the same survivors and slots.
                        const uint32_t ha = (uint32_t)((uint64_t)__double_as_longlong(ka) >> 32);
                        const uint32_t hb = (uint32_t)((uint64_t)__double_as_longlong(kb) >> 32);
                        const uint32_t ha_prev = __shfl_up_sync(PCL_FULL_MASK, ha, 1, LP);
                        const uint32_t ha_last = __shfl_sync(PCL_FULL_MASK, ha, cbase + ns - 1);
                        in_place = __all_sync(PCL_FULL_MASK, p >= ns || ((p == 0 || ha > ha_prev) && hb > ha_last));
#else
                        const double ka_prev = pcl_shfl_f64(ka, p == 0 ? lane : lane - 1);
                        const double ka_last = pcl_shfl_f64(ka, cbase + ns - 1);
                        in_place = __all_sync(PCL_FULL_MASK, p >= ns || ((p == 0 || ka < ka_prev) && kb < ka_last));
#endif
                    }
                    if (!in_place) {
                    if (!EXACT) {
                        const int ida = (hard ? LP : 0) | p;
                        ka = pcl_prune_key<NC>(ka, ida);
                        kb = pcl_prune_key<NC>(kb, ida ^ LP);
                    }
                    double2 kv;
                    kv.x = ka;
                    kv.y = kb;
                    *reinterpret_cast<double2*>(cm + 2 * p) = kv;
                    __syncwarp();
                    if (EXACT || !PCL_PRUNE_SHORTCUT) {
#pragma unroll
                        for (int jj = 0; jj < NC; jj += 2) {
                            const double2 kp = *reinterpret_cast<const double2*>(cm + jj);   // 16-byte aligned
                            if (EXACT) {
                                // position jj holds candidate (bit 0, path jj/2), jj + 1 (bit 1, path jj/2)
                                const int q = jj >> 1;
                                ra += (kp.x > ka) || (kp.x == ka && q < p);
                                ra += (kp.y > ka);
                                rb += (kp.x > kb) || (kp.x == kb);
                                rb += (kp.y > kb) || (kp.y == kb && q < p);
                            } else {
                                pcl_rank_acc(ra, kp.x, ka);
                                pcl_rank_acc(rb, kp.x, kb);
                                pcl_rank_acc(ra, kp.y, ka);
                                pcl_rank_acc(rb, kp.y, kb);
                            }
                        }
                    } else {
                        // Ranks against the LP likely keys first.  An unlikely candidate with ns likely
                        // keys above it is pruned, and if that holds for all of them (reliable bit: |x|
                        // exceeds the spread of the list) no unlikely key can sit above a surviving
                        // likely one either: the ranks among the likely keys are final and the second
                        // half of the comparisons is skipped.  If the likely keys are moreover still in
                        // slot order, every path simply continues in place: no scatter, no shuffles.
#pragma unroll
                        for (int q = 0; q < LP; q++) {
                            const double kq = cm[2 * q];
                            pcl_rank_acc(ra, kq, ka);
                            pcl_rank_acc(rb, kq, kb);
                        }
                        if (__all_sync(PCL_FULL_MASK, rb >= ns)) {
                            rb = NC;
                            in_place = __all_sync(PCL_FULL_MASK, p >= ns || ra == p);
                        } else {
#pragma unroll
                            for (int q = 0; q < LP; q++) {
                                const double kq = cm[2 * q + 1];
                                pcl_rank_acc(ra, kq, ka);
                                pcl_rank_acc(rb, kq, kb);
                            }
                        }
                    }
                    }
                    nact = ns;
                    if (in_place) {
#if PCL_LEAF_V2
                        // only reached on a full list: ns == nact, nobody's `act` changes, slots p >= ns stay DEAD
                        if (p < ns) pm = ka;
#else
                        act = (p < ns) && valid;
                        pm = (p < ns) ? ka : DEAD;
#endif
                        u = hard ? 1u : 0u;
                    } else {
                        if (EXACT) {
                            if (ra < ns) { sel[ra] = p; newpm[ra] = ka; }
                            if (rb < ns) { sel[rb] = LP + p; newpm[rb] = kb; }
                        } else {
                            if (ra < ns) newpm[ra] = ka;
                            if (rb < ns) newpm[rb] = kb;
                        }
                        __syncwarp();
                        act = (p < ns) && valid;
                        pm = DEAD;
                        if (p < ns) {
                            pm = newpm[p];
                            const int c = EXACT ? sel[p] : (int)(__double_as_longlong(pm) & (NC - 1));
                            parent = c & (LP - 1);
                            u = (uint32_t)(c >> PB);
                        }
                        // a survivor takes over its parent's pointer words and live registers
                        const int srcl = cbase | parent;
                        ptrL = __shfl_sync(PCL_FULL_MASK, ptrL, srcl);
                        ptrB = __shfl_sync(PCL_FULL_MASK, ptrB, srcl);
                        small = __shfl_sync(PCL_FULL_MASK, small, srcl);
                        if (j < 4) {
#pragma unroll
                            for (int t = 0; t < 8; t++) R3[t] = pcl_shfl_real<real>(R3[t], srcl);
                        }
                        if ((j & 3) < 2) {
#pragma unroll
                            for (int t = 0; t < 4; t++) R2[t] = pcl_shfl_real<real>(R2[t], srcl);
                        }
                        if ((j & 1) == 0) {
                            R1[0] = pcl_shfl_real<real>(R1[0], srcl);
                            R1[1] = pcl_shfl_real<real>(R1[1], srcl);
                        }
                    }
                }
                if (DBG && P.dbg_leaf != nullptr && valid) {
                    P.dbg_leaf[(f * N + i0 + j) * LP + p] = x;
                    P.dbg_parent[(f * N + i0 + j) * LP + p] = (uint8_t)parent;
                }

#if PCL_RAW_BLOCK_BITS
                small |= u << (24 + j);
#else
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
                        } else {
                            c4hi = c4;               // leaf 7: the block is complete (folded below)
                            ulast = u;
                        }
                    }
                }
#endif
            }
#if PCL_RAW_BLOCK_BITS
            c8 = small >> 24;                     // u0 .. u7 -> the block's partial sums
            c8 ^= (c8 >> 1) & 0x55u;
            c8 ^= (c8 >> 2) & 0x33u;
            c8 ^= (c8 >> 4) & 0x0fu;
#else
            c8 = (((small >> 24) & 15u) ^ c4hi) | (c4hi << 4);
#endif
            }

            // ---- block complete: fold its 8 partial sums upwards while the node is a right child
            // (the last block keeps its fields: the final u-word is assembled from them) -----------
            if (blk != NB - 1) {
                uint32_t c = c8;
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
                            bw[32 * ((N >> 5) - (N >> (d + 4))) + lane] = c;
                            ptrB = (ptrB & ~((uint32_t)(LP - 1) << ((d - 1) * PB))) | ((uint32_t)p << ((d - 1) * PB));
                        }
                        __syncwarp();
                    }
                } else {
                    const int cto = __ffs(~(i0 + 7)) - 1;
                    const int d = n - cto;
                    const int Wd = N >> (d + 5);
                    uint32_t* dest = bw + 32 * ((N >> 5) - (N >> (d + 4)));
                    if (act) dest[(Wd - 1) * 32 + lane] = c;
                    __syncwarp();
                    for (int l = n - 5; l > d; l--) {
                        const int w = N >> (l + 5);
                        const int ql = cbase + ((ptrB >> ((l - 1) * PB)) & (LP - 1));
                        const uint32_t* lsrc = bw + 32 * ((N >> 5) - (N >> (l + 4)));
                        if (act)
                            for (int jw = 0; jw < w; jw++)
                                dest[(Wd - 2 * w + jw) * 32 + lane] =
                                    lsrc[jw * 32 + ql] ^ dest[(Wd - w + jw) * 32 + lane];
                        __syncwarp();
                    }
                    if (act) ptrB = (ptrB & ~((uint32_t)(LP - 1) << ((d - 1) * PB))) | ((uint32_t)p << ((d - 1) * PB));
                }
            }
        }

        // ---- final selection (decoder.py:259-262), per frame of the warp ------------------------
        int best = 0;
        __syncwarp();
        if (LP > 1) {
            newpm[p] = pm;
            __syncwarp();
            double bm = newpm[0];
            for (int q = 1; q < LP; q++) {
                const double v = newpm[q];
                if (v > bm) { bm = v; best = q; }           // first maximum, like np.argmax
            }
        }
        if (P.pm_out != nullptr && p < L && valid)
            P.pm_out[f * L + p] = (pm < -1.0e299) ? -(double)pcl_math<real>::inf() : pm;

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
#if PCL_RAW_BLOCK_BITS
                        v = sm;                               // the top byte holds the last block's decisions as they are
                        v ^= (v >> 1) & 0x00555555u;
                        v ^= (v >> 2) & 0x00333333u;
#else
                        v = pcl_bfi(sm, ul, 31, 1);
                        v ^= (v >> 1) & 0x15555555u;
                        v ^= (v >> 2) & 0x03333333u;
#endif
                        v ^= (v >> 4) & 0x000F0F0Fu;
                        v ^= (v >> 8) & 0x000000FFu;
                    } else {
                        const int r = NW - w;
                        const int Wl = 1 << (31 - __clz(r - 1));
                        const int l = (31 - __clz(NW)) - (31 - __clz(Wl));
                        const int jw = w - (NW - 2 * Wl);
                        v = bw[32 * ((N >> 5) - (N >> (l + 4))) + jw * 32 + fq * LP + ((pB >> ((l - 1) * PB)) & (LP - 1))];
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
            if (p < nact && valid) {
                const uint32_t* U = uw + lane * NW;
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
    if (TM) {
        pcl_tmem_fence_before();
        __syncthreads();
        if (warp == 0) pcl_tmem_free_all(hdr[0]);
    }
}
