// ldpc_bp.cuh -- batched flooding BP / Min-Sum LDPC decoder, one warp per frame.
//
// Replaces BPDecoder.decode (/root/reference/src/ldpc/decoder.py:124-202, check
// rule :62-96, variable rule :98-122) and MSDecoder.decode (:289-352, check rule
// :257-287) for a whole batch, with the reference's schedule: all checks, then all
// variables, hard decision total <= 0, syndrome early stop.
//
// Layout.  Edges are enumerated check-major (row-major scan of H == 1), which is
// exactly check_neighbors[c] ascending (decoder.py:43-47); `vperm` lists the same
// edge ids variable-major with checks ascending (var_neighbors[v]).  A frame's E
// edge messages live in ONE shared-memory array for the whole decode: the check
// pass overwrites v2c[e] with c2v[e] in place (a check owns its contiguous edge
// run), the variable pass gathers through vperm and overwrites c2v[e] with v2c[e].
// HBM sees only the n channel LLRs in and n hard bits out per frame; the Tanner
// tables are shared by every frame and stay in L1/L2.
//
// fp64 build: the reference's arithmetic and association order exactly
// (tanh -> clip -> left-to-right leave-one-out product -> clip -> 2 atanh;
// numpy's pairwise np.sum order for the variable sum).
// fp32 build: same schedule; the check rule works on u = exp(-|x|) and the even / odd
// elementary symmetric sums of the u_j (cn_bp_core below: no cancellation for saturated or
// weak messages, three MUFU ops per edge); signs travel as an XOR of sign bits.  Regular
// (3, 6) codes get compile-time degrees (REG) and, in the fp32 build, the conflict-free layout of
// ldpc_banked.cuh; this kernel serves the fp64 validation build, irregular codes and
// PCL_LDPC_BANKED=0.  COOP: codes too large for ~16 resident frames per SM are decoded by a
// block of 4 warps per frame with block barriers between the passes.
#pragma once
#include "pcl_common.cuh"

struct LdpcLayout {
    int m, n, E;
    int max_iter, early_stop;
    int nhw;                    // hard-decision words = ceil(n / 32)
    int off_msg, off_llr, off_hard, off_ctl, warp_bytes;
    int coop;                   // 1: the whole block decodes one frame (large codes), 0: one warp per frame
    int banked, nR, NP, NS;     // banked layout (ldpc_banked_kernel): check rounds, variable positions, slots
    int paired;                 // banked layout with two messages of a check per 8-byte word pair
};

template <typename real>
struct LdpcParams {
    LdpcLayout lay;
    const real* llr;            // [F][n]
    uint8_t* bits;              // [F][n]
    int32_t* iters;             // [F] or null
    real* total;                // [F][n] or null
    const int32_t* cptr;        // [m+1] edge offsets per check
    const uint16_t* col;        // [E] variable of edge e
    const int32_t* vptr;        // [n+1]
    const uint16_t* vperm;      // [E] edge ids, variable-major
    const unsigned long long* vpack;  // REG: per variable, its 3 edge ids packed 16 bits each
    const unsigned long long* bpack;  // banked: per variable position, the byte offsets of its 3 message words (16 bits each)
    const uint16_t* varof;      // banked: variable at position pi (0xffff: hole)
    const uint16_t* posof;      // banked: position of variable v
    const uint16_t* cpos;       // banked: variable position behind slot s (0xffff: hole)
    unsigned long long* next;   // dynamic frame counter (never reset: a launch counts from ticket_base)
    unsigned long long ticket_base;
    int64_t F;
    real norm;                  // Min-Sum normalisation
};

template <typename real> struct ldpc_const;
template <> struct ldpc_const<float> {
    static PCL_DEVICE float clipv() { return 0.999999f; }
    static PCL_DEVICE float wmin() { return 1e-6f; }
};
template <> struct ldpc_const<double> {
    static PCL_DEVICE double clipv() { return 0.999999; }
    static PCL_DEVICE double wmin() { return 1e-6; }
};

// ---- check node, reference arithmetic (decoder.py:62-96) -----------------------
template <int DMAX>
PCL_DEVICE void cn_bp_exact(double* msg, int d)
{
    double t[DMAX];
#pragma unroll
    for (int j = 0; j < DMAX; j++) {
        if (j < d) {
            double v = tanh(msg[j] / 2.0);
            v = fmin(fmax(v, -0.999999), 0.999999);
            t[j] = v;
        } else {
            t[j] = 1.0;
        }
    }
#pragma unroll
    for (int i = 0; i < DMAX; i++) {
        if (i < d) {
            double pr = 1.0;
            bool first = true;
#pragma unroll
            for (int j = 0; j < DMAX; j++) {
                if (j < d && j != i) {
                    if (first) { pr = t[j]; first = false; } else pr = pr * t[j];
                }
            }
            pr = fmin(fmax(pr, -0.999999), 0.999999);
            double o = 2.0 * atanh(pr);
            if (o != o) o = 0.0;                               // np.nan_to_num, :94
            msg[i] = o;
        }
    }
}

// ---- check node, fp32 production rule (even / odd symmetric sums of u = exp(-|x|)) ----
// tanh(|x|/2) = (1 - u) / (1 + u) with u = exp(-|x|), so for the other edges of a check
//   prod t_j = (E - O) / (E + O),   prod (1 + u_j) = E + O,   prod (1 - u_j) = E - O,
// E / O = the even / odd elementary symmetric sums of the u_j, and the outgoing message is
//   2 atanh(prod t_j) = ln((1 + p) / (1 - p)) = ln(E / O).
// E and O are sums of positive terms only: no cancellation anywhere, full relative precision
// both for saturated messages (all u ~ 1e-6: O ~ sum u_j) and for weak ones (u ~ 1: E ~ O).
// A factor (1 + u) maps (E, O) -> (E + u O, O + u E); the leave-one-out pair comes from a
// prefix and a suffix pair: E = Ep Es + Op Os, O = Ep Os + Op Es.  Three MUFU ops per edge
// (ex2 in; rcp, lg2 out).  The reference's clips (+-0.999999 on t and on the product,
// decoder.py:82,88) become u >= (1 - 0.999999) / (1 + 0.999999) and E / O <= 1999999; an
// absent edge (irregular codes) is u = 0, the identity factor; a degree-1 check gives
// E / O = 1 / 0 -> clipped -> ln(1999999) = 14.5087, as the reference does.
PCL_DEVICE float pcl_ex2(float x)
{
#ifdef PCL_EMU
    return exp2f(x);
#else
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}
PCL_DEVICE float pcl_lg2(float x)
{
#ifdef PCL_EMU
    return log2f(x);
#else
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}
PCL_DEVICE float pcl_rcp(float x)
{
#ifdef PCL_EMU
    return 1.0f / x;
#else
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}

// Register-level rule: xb = bit patterns of the DMAX incoming messages (ALL: every one exists,
// else the first d), out = outgoing messages.
// LOG2: the messages are carried in units of ln 2 (y = x / ln 2), so u = 2^-|y| and the output is
// lg2(E / O) with no scaling multiply on either side (the variable pass is linear, the channel
// LLRs are scaled once per frame).  With ALL and DMAX >= 3 the output clip can never act: every
// u >= 5e-7 makes O >= 1e-6 E, i.e. E / O < 1999999 (the reference's second clip, decoder.py:88,
// is likewise unreachable after the first one for checks of degree >= 3).
template <int DMAX, bool ALL, bool LOG2 = false>
PCL_DEVICE void cn_bp_core(const uint32_t* xb, float* out, int d)
{
    const float UMIN = 5.00000250000125e-07f, RMAX = 1999999.0f;
    float u[DMAX];
    uint32_t sall = 0;
#pragma unroll
    for (int j = 0; j < DMAX; j++) {
        const float ax = fabsf(__uint_as_float(xb[j]));
        const float e = LOG2 ? pcl_ex2(-ax) : pcl_ex2(ax * -1.4426950408889634f);   // exp(-|x|), rel. err ~ |x| * 6e-8
        u[j] = (ALL || j < d) ? fmaxf(e, UMIN) : 0.0f;
        sall ^= xb[j];
    }
    // suffix pairs, then a running prefix pair
    float se[DMAX + 1], so[DMAX + 1];
    se[DMAX] = 1.0f;
    so[DMAX] = 0.0f;
#pragma unroll
    for (int j = DMAX - 1; j >= 1; j--) {
        se[j] = fmaf(u[j], so[j + 1], se[j + 1]);
        so[j] = fmaf(u[j], se[j + 1], so[j + 1]);
    }
    float pe = 1.0f, po = 0.0f;
#pragma unroll
    for (int i = 0; i < DMAX; i++) {
        const float E = fmaf(po, so[i + 1], pe * se[i + 1]);
        const float O = fmaf(po, se[i + 1], pe * so[i + 1]);
        const float r0 = E * pcl_rcp(O);
        const float ratio = (ALL && DMAX >= 3 && LOG2) ? r0 : fminf(r0, RMAX);
        const float mag = LOG2 ? pcl_lg2(ratio) : 0.6931471805599453f * pcl_lg2(ratio);
        out[i] = __uint_as_float(__float_as_uint(mag) | ((sall ^ xb[i]) & 0x80000000u));
        const float npe = fmaf(u[i], po, pe);
        po = fmaf(u[i], pe, po);
        pe = npe;
    }
}

// ---- two checks per lane on the packed fp32x2 pipe (FFMA2 / FMUL2, sm_100) ---------------------
// The even / odd recurrences are pure multiply-adds, so two independent checks ride in the two
// halves of 64-bit register pairs: the FMA-pipe instruction count of the rule halves, and one
// reciprocal serves both checks of a pair (r_a = E_a O_b / (O_a O_b), r_b = E_b O_a / (O_a O_b)):
// 2.5 MUFU per edge instead of 3.  All messages exist (regular code), LOG2 units.
#ifdef PCL_EMU
PCL_DEVICE float2 pcl_fma2(float2 a, float2 b, float2 c) { float2 r; r.x = fmaf(a.x, b.x, c.x); r.y = fmaf(a.y, b.y, c.y); return r; }
PCL_DEVICE float2 pcl_mul2(float2 a, float2 b) { float2 r; r.x = a.x * b.x; r.y = a.y * b.y; return r; }
#else
PCL_DEVICE float2 pcl_fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
PCL_DEVICE float2 pcl_mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
#endif

template <int DMAX>
PCL_DEVICE void cn_bp_core2(const uint32_t* xa, const uint32_t* xb, float* outa, float* outb)
{
    // The first and last steps of both recurrences meet the identity pair (E, O) = (1, 0): those
    // multiply-adds are written out by hand (x * 1 + 0 is exact, so the results are the same bits;
    // the compiler may not drop them itself because 0 * inf is not 0).
    static_assert(DMAX >= 4, "pair rule is written for check degree >= 4");
    const float UMIN = 5.00000250000125e-07f;
    float2 u[DMAX];
    uint32_t sa = 0, sb = 0;
#pragma unroll
    for (int j = 0; j < DMAX; j++) {
        u[j].x = fmaxf(pcl_ex2(-fabsf(__uint_as_float(xa[j]))), UMIN);
        u[j].y = fmaxf(pcl_ex2(-fabsf(__uint_as_float(xb[j]))), UMIN);
        sa ^= xa[j];
        sb ^= xb[j];
    }
    float2 one;
    one.x = one.y = 1.0f;
    // suffix pairs over edges j .. DMAX-1; (se, so)[DMAX-1] = (1, u[DMAX-1])
    float2 se[DMAX], so[DMAX];
    se[DMAX - 1] = one;
    so[DMAX - 1] = u[DMAX - 1];
#pragma unroll
    for (int j = DMAX - 2; j >= 1; j--) {
        se[j] = pcl_fma2(u[j], so[j + 1], se[j + 1]);
        so[j] = pcl_fma2(u[j], se[j + 1], so[j + 1]);
    }
    float2 pe = one, po = u[0];                  // prefix pair over edges 0 .. i-1 (i >= 1)
#pragma unroll
    for (int i = 0; i < DMAX; i++) {
        float2 E, O;
        if (i == 0) {                            // no prefix
            E = se[1];
            O = so[1];
        } else if (i == DMAX - 1) {              // no suffix
            E = pe;
            O = po;
        } else if (i == 1) {                     // prefix = (1, u0)
            E = pcl_fma2(po, so[2], se[2]);
            O = pcl_fma2(po, se[2], so[2]);
        } else if (i == DMAX - 2) {              // suffix = (1, u[DMAX-1])
            E = pcl_fma2(po, so[i + 1], pe);
            O = pcl_fma2(pe, so[i + 1], po);
        } else {
            E = pcl_fma2(po, so[i + 1], pcl_mul2(pe, se[i + 1]));
            O = pcl_fma2(po, se[i + 1], pcl_mul2(pe, so[i + 1]));
        }
        const float inv = pcl_rcp(O.x * O.y);
        const float ra = (E.x * O.y) * inv, rb = (E.y * O.x) * inv;
        outa[i] = __uint_as_float(__float_as_uint(pcl_lg2(ra)) | ((sa ^ xa[i]) & 0x80000000u));
        outb[i] = __uint_as_float(__float_as_uint(pcl_lg2(rb)) | ((sb ^ xb[i]) & 0x80000000u));
        if (i >= 1 && i < DMAX - 1) {
            const float2 npe = pcl_fma2(u[i], po, pe);
            po = pcl_fma2(u[i], pe, po);
            pe = npe;
        }
    }
}

// VEC2: the check's DMAX messages start at an 8-byte boundary and all exist (regular codes):
// 8-byte shared-memory accesses, conflict-free at a stride of 6 words.
template <int DMAX, bool VEC2>
PCL_DEVICE void cn_bp_f32(float* msg, int d)
{
    uint32_t xb[DMAX];
    float out[DMAX];
    if (VEC2) {
#pragma unroll
        for (int j = 0; j < DMAX; j += 2) {
            const float2 v = reinterpret_cast<const float2*>(msg)[j >> 1];
            xb[j] = __float_as_uint(v.x);
            xb[j + 1] = __float_as_uint(v.y);
        }
    } else {
#pragma unroll
        for (int j = 0; j < DMAX; j++) xb[j] = (j < d) ? __float_as_uint(msg[j]) : 0x7f800000u;   // +inf: u = 0
    }
    cn_bp_core<DMAX, VEC2>(xb, out, d);
    if (VEC2) {
#pragma unroll
        for (int j = 0; j < DMAX; j += 2) {
            float2 v;
            v.x = out[j];
            v.y = out[j + 1];
            reinterpret_cast<float2*>(msg)[j >> 1] = v;
        }
    } else {
#pragma unroll
        for (int j = 0; j < DMAX; j++)
            if (j < d) msg[j] = out[j];
    }
}

// ---- check node of ANY degree, in place, no per-thread arrays (checks wider than 32 edges) ----
// Phase 1 replaces every message by its transformed value (fp64: the clipped tanh; fp32: u with the
// sign of the message); phase 2 walks the edges in order with the running PREFIX in registers and
// folds the not-yet-overwritten entries behind edge i onto it.  fp64: exactly the reference's
// left-to-right product that skips i (decoder.py:84-90), O(d^2) multiplies; fp32: the even / odd pair.
PCL_DEVICE void cn_bp_loop(double* msg, int d)
{
    for (int j = 0; j < d; j++) {
        double v = tanh(msg[j] / 2.0);
        msg[j] = fmin(fmax(v, -0.999999), 0.999999);
    }
    double pre = 1.0;
    bool has_pre = false;
    for (int i = 0; i < d; i++) {
        double pr = pre;
        bool first = !has_pre;
        for (int j = i + 1; j < d; j++) {
            if (first) { pr = msg[j]; first = false; } else pr = pr * msg[j];
        }
        if (first) pr = 1.0;                                   // degree-1 check: empty product
        const double ti = msg[i];
        pr = fmin(fmax(pr, -0.999999), 0.999999);
        double o = 2.0 * atanh(pr);
        if (o != o) o = 0.0;
        msg[i] = o;
        if (has_pre) pre = pre * ti; else { pre = ti; has_pre = true; }
    }
}
PCL_DEVICE void cn_bp_loop(float* msg, int d)
{
    const float UMIN = 5.00000250000125e-07f, RMAX = 1999999.0f;
    uint32_t sall = 0;
    for (int j = 0; j < d; j++) {
        const uint32_t xb = __float_as_uint(msg[j]);
        const float u = fmaxf(pcl_ex2(fabsf(msg[j]) * -1.4426950408889634f), UMIN);
        sall ^= xb;
        msg[j] = __uint_as_float(__float_as_uint(u) | (xb & 0x80000000u));
    }
    float pe = 1.0f, po = 0.0f;
    for (int i = 0; i < d; i++) {
        float e = pe, o = po;
        for (int j = i + 1; j < d; j++) {
            const float u = fabsf(msg[j]);
            const float ne = fmaf(u, o, e);
            o = fmaf(u, e, o);
            e = ne;
        }
        const uint32_t xi = __float_as_uint(msg[i]);
        const float ui = fabsf(msg[i]);
        const float mag = 0.6931471805599453f * pcl_lg2(fminf(e * pcl_rcp(o), RMAX));
        msg[i] = __uint_as_float(__float_as_uint(mag) | ((sall ^ xi) & 0x80000000u));
        const float npe = fmaf(ui, po, pe);
        po = fmaf(ui, pe, po);
        pe = npe;
    }
}

// ---- check node, Min-Sum (decoder.py:257-287), any degree, in place -------------
template <typename real>
PCL_DEVICE void cn_ms(real* msg, int d, real norm)
{
    real m1 = pcl_math<real>::inf(), m2 = pcl_math<real>::inf();
    int i1 = -1, nneg = 0, nzero = 0;
    for (int j = 0; j < d; j++) {
        const real x = msg[j];
        const real a = fabs(x);
        if (a < m1) { m2 = m1; m1 = a; i1 = j; } else if (a < m2) { m2 = a; }
        nneg += (x < (real)0);
        nzero += (x == (real)0);
    }
    for (int i = 0; i < d; i++) {
        const real x = msg[i];
        const real mn = (i == i1) ? m2 : m1;
        const int zeros_other = nzero - (x == (real)0);
        const int neg_other = nneg - (x < (real)0);
        real sp = (zeros_other > 0) ? (real)0 : ((neg_other & 1) ? (real)-1 : (real)1);
        msg[i] = sp * mn * norm;                                // :285
    }
}

// Regular codes: all DMAX messages of the check exist and start at an 8-byte boundary.  The two
// smallest magnitudes come from a branch-free min / max ladder; edge i takes the second one iff
// its own magnitude is the smallest (equal minima make the two the same number, so comparing
// values needs no index); sign = product of the other signs as an XOR of sign bits.  A zero
// among the other inputs makes the minimum 0, which is what np.sign(0) = 0 gives (:270-285).
template <typename real> struct ldpc_bits;
template <> struct ldpc_bits<float> {
    typedef uint32_t u;
    static PCL_DEVICE u of(float x) { return __float_as_uint(x); }
    static PCL_DEVICE float to(u b) { return __uint_as_float(b); }
    static PCL_DEVICE u sign() { return 0x80000000u; }
};
template <> struct ldpc_bits<double> {
    typedef unsigned long long u;
    static PCL_DEVICE u of(double x) { return (u)__double_as_longlong(x); }
    static PCL_DEVICE double to(u b) { return __longlong_as_double((long long)b); }
    static PCL_DEVICE u sign() { return 0x8000000000000000ull; }
};

template <typename real, int DMAX>
PCL_DEVICE void cn_ms_core(const real* x, real* out, real norm)
{
    typedef typename ldpc_bits<real>::u bits_t;
    if constexpr (sizeof(real) == 4 && DMAX == 6) {
        // fp32, degree 6: the six leave-one-out minima straight from 3-input minima (FMNMX3 with |.| operands on
        // sm_100: nvcc fuses the nested fminf) -- two group minima and one 3-input minimum per edge, 8 instructions
        // instead of the min / second-min ladder and a compare-select per edge (30).  Signs travel as raw sign
        // bits: a -0.0 among the others makes their minimum 0, so the sign of that zero is all it can change.
        uint32_t xb[6], sall = 0;
        float a[6], mn[6];
#pragma unroll
        for (int j = 0; j < 6; j++) {
            xb[j] = __float_as_uint((float)x[j]);
            a[j] = fabsf((float)x[j]);
            sall ^= xb[j];
        }
        const float t012 = fminf(fminf(a[0], a[1]), a[2]), t345 = fminf(fminf(a[3], a[4]), a[5]);
        mn[0] = fminf(fminf(a[1], a[2]), t345);
        mn[1] = fminf(fminf(a[0], a[2]), t345);
        mn[2] = fminf(fminf(a[0], a[1]), t345);
        mn[3] = fminf(fminf(a[4], a[5]), t012);
        mn[4] = fminf(fminf(a[3], a[5]), t012);
        mn[5] = fminf(fminf(a[3], a[4]), t012);
        // the sign rides on the normalisation factor: min * (+-norm), one LOP3 and one FMUL per edge ((+-1 * min) * norm, :285)
        const uint32_t sn = (sall & 0x80000000u) ^ __float_as_uint((float)norm);
#pragma unroll
        for (int i = 0; i < 6; i++) out[i] = (real)(mn[i] * __uint_as_float(sn ^ (xb[i] & 0x80000000u)));
        return;
    }
    real a[DMAX];
    bits_t sall = 0;
    real m1 = pcl_math<real>::inf(), m2 = pcl_math<real>::inf();
#pragma unroll
    for (int j = 0; j < DMAX; j++) {
        a[j] = fabs(x[j]);
        // np.sign: a NaN-free input is negative iff x < 0 (-0.0 is not)
        sall ^= (x[j] < (real)0) ? ldpc_bits<real>::sign() : (bits_t)0;
        m2 = fmin(m2, fmax(m1, a[j]));
        m1 = fmin(m1, a[j]);
    }
#pragma unroll
    for (int i = 0; i < DMAX; i++) {
        const real mn = (a[i] > m1) ? m1 : m2;
        const bits_t sg = sall ^ ((x[i] < (real)0) ? ldpc_bits<real>::sign() : (bits_t)0);
        out[i] = ldpc_bits<real>::to(ldpc_bits<real>::of(mn * norm) ^ sg);       // (+-1 * min) * norm, :285
    }
}

template <typename real, int DMAX>
PCL_DEVICE void cn_ms_reg(real* msg, real norm)
{
    real x[DMAX], out[DMAX];
    if (sizeof(real) == 4 && DMAX % 2 == 0) {
#pragma unroll
        for (int j = 0; j < DMAX; j += 2) {
            const float2 v = reinterpret_cast<const float2*>(msg)[j >> 1];
            x[j] = (real)v.x;
            x[j + 1] = (real)v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < DMAX; j++) x[j] = msg[j];
    }
    cn_ms_core<real, DMAX>(x, out, norm);
    if (sizeof(real) == 4 && DMAX % 2 == 0) {
#pragma unroll
        for (int j = 0; j < DMAX; j += 2) {
            float2 v;
            v.x = (float)out[j];
            v.y = (float)out[j + 1];
            reinterpret_cast<float2*>(msg)[j >> 1] = v;
        }
    } else {
#pragma unroll
        for (int j = 0; j < DMAX; j++) msg[j] = out[j];
    }
}

// numpy add.reduce association order (pairwise_sum) over gathered messages
template <typename real>
PCL_DEVICE real vn_sum_np(const real* msg, const uint16_t* ed, int d)
{
    if (d < 8) {
        real r = (real)0;
        for (int j = 0; j < d; j++) r += msg[ed[j]];
        return r;
    }
    real r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = msg[ed[j]];
    int i = 8;
    for (; i < d - (d % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] += msg[ed[i + j]];
    }
    real res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < d; i++) res += msg[ed[i]];
    return res;
}

// REG != 0: every check has degree DMAX and every variable degree 3 (the regular (3, DMAX)
// Gallager codes of the benchmarks): degrees and edge offsets become compile-time constants,
// no pointer-table loads and no per-edge degree predicates.
template <typename real, int MODE, int DMAX, int REG, int COOP>
__global__ void __launch_bounds__(256) ldpc_decode_kernel(LdpcParams<real> P)
{
    const LdpcLayout& Y = P.lay;
    const int m = Y.m, n = Y.n, E = Y.E;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // Large codes (a frame's messages take a big share of the SM's shared memory) are decoded by
    // the whole block: `tid` of `T` threads strides over checks / variables and the passes are
    // separated by block barriers; small codes keep one warp per frame and warp barriers.
    constexpr bool coop = COOP != 0;
    const int T = coop ? (int)blockDim.x : 32;
    const int tid = coop ? (int)threadIdx.x : lane;
    unsigned char* wsm = pcl_dyn_smem() + (coop ? (size_t)0 : (size_t)warp * Y.warp_bytes);
    real* msg = (real*)(wsm + Y.off_msg);
    real* sllr = (real*)(wsm + Y.off_llr);
    uint32_t* hard = (uint32_t*)(wsm + Y.off_hard);
    unsigned long long* ctl = (unsigned long long*)(wsm + Y.off_ctl);   // coop: frame index, stop flag
    auto sync = [&]() {
        if (coop) __syncthreads();
        else __syncwarp();
    };

    for (;;) {
        unsigned long long fq = 0;
        if (coop) {
            if (tid == 0) ctl[0] = atomicAdd(P.next, 1ull) - P.ticket_base;
            __syncthreads();
            fq = ctl[0];
        } else {
            if (lane == 0) fq = atomicAdd(P.next, 1ull) - P.ticket_base;
            fq = pcl_shfl_u64(fq, 0);
        }
        if ((int64_t)fq >= P.F) break;
        const int64_t f = (int64_t)fq;

        const real* ch = P.llr + f * n;
        for (int v = tid; v < n; v += T) sllr[v] = ch[v];
        sync();
        for (int e = tid; e < E; e += T) msg[e] = sllr[P.col[e]];   // decoder.py:144-146
        sync();

        int iters = Y.max_iter;                                       // :149
        for (int it = 0; it < Y.max_iter; it++) {
            // 1. check nodes (:152-168)
            for (int c = tid; c < m; c += T) {
                const int e0 = REG ? c * DMAX : P.cptr[c];
                const int d = REG ? DMAX : P.cptr[c + 1] - e0;
                if constexpr (MODE == 1) {
                    if constexpr (REG != 0) cn_ms_reg<real, DMAX>(msg + e0, P.norm);
                    else cn_ms<real>(msg + e0, d, P.norm);
                } else {
                    if constexpr (DMAX == 0) cn_bp_loop(msg + e0, d);                     // checks wider than 32 edges
                    else if constexpr (sizeof(real) == 8) cn_bp_exact<DMAX>((double*)(msg + e0), d);
                    else cn_bp_f32<DMAX, (REG != 0 && DMAX % 2 == 0)>((float*)(msg + e0), d);
                }
            }
            sync();
            // 2. variable nodes (:173-188) + 3. hard decision (:191)
            // (hard decisions are only materialised when something reads them: the syndrome test
            // or the output after the last iteration)
            const bool want_hard = Y.early_stop || it == Y.max_iter - 1;
            for (int vb = coop ? warp * 32 : 0; vb < n; vb += T) {
                const int v = vb + lane;
                bool bit = false;
                if (v < n) {
                    real total;
                    if (REG) {
                        const unsigned long long pk = P.vpack[v];
                        const int ea = (int)(pk & 0xffffu), eb = (int)((pk >> 16) & 0xffffu), ec = (int)((pk >> 32) & 0xffffu);
                        const real ma = msg[ea], mb = msg[eb], mc = msg[ec];
                        total = sllr[v] + ((((real)0 + ma) + mb) + mc);
                        msg[ea] = total - ma;
                        msg[eb] = total - mb;
                        msg[ec] = total - mc;
                    } else {
                        const int j0 = P.vptr[v];
                        const int d = P.vptr[v + 1] - j0;
                        const uint16_t* ed = P.vperm + j0;
                        if (d == 3) {
                            const int ea = ed[0], eb = ed[1], ec = ed[2];
                            const real ma = msg[ea], mb = msg[eb], mc = msg[ec];
                            total = sllr[v] + ((((real)0 + ma) + mb) + mc);
                            msg[ea] = total - ma;
                            msg[eb] = total - mb;
                            msg[ec] = total - mc;
                        } else {
                            total = sllr[v] + vn_sum_np<real>(msg, ed, d);
                            for (int j = 0; j < d; j++) {
                                const int e = ed[j];
                                msg[e] = total - msg[e];
                            }
                        }
                    }
                    bit = (total <= (real)0);
                    if (P.total != nullptr) P.total[f * n + v] = total;
                }
                if (want_hard) {
                    const unsigned bal = __ballot_sync(PCL_FULL_MASK, bit);
                    if (lane == 0) hard[vb >> 5] = bal;
                }
            }
            sync();
            // 4. syndrome early stop (:194-198)
            if (Y.early_stop) {
                bool bad = false;
                for (int c = tid; c < m; c += T) {
                    const int e0 = REG ? c * DMAX : P.cptr[c], e1 = REG ? e0 + DMAX : P.cptr[c + 1];
                    unsigned par = 0;
                    for (int e = e0; e < e1; e++) {
                        const int v = P.col[e];
                        par ^= hard[v >> 5] >> (v & 31);
                    }
                    bad |= (par & 1u) != 0;
                }
                bool any_bad = __any_sync(PCL_FULL_MASK, bad);
                if (coop) {
                    if (tid == 0) ctl[1] = 0ull;
                    __syncthreads();
                    if (any_bad && lane == 0) ctl[1] = 1ull;
                    __syncthreads();
                    any_bad = ctl[1] != 0ull;
                }
                if (!any_bad) { iters = it + 1; break; }
            }
        }
        uint8_t* out = P.bits + f * n;
        for (int v = tid; v < n; v += T) out[v] = (uint8_t)((hard[v >> 5] >> (v & 31)) & 1u);
        if (P.iters != nullptr && tid == 0) P.iters[f] = iters;
        sync();
    }
}
