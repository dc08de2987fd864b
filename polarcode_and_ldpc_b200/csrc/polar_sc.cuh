// polar_sc.cuh -- dedicated successive-cancellation (list size 1) decoders for N = 256, the
// reference's quick-start configuration (BASELINE configs[0]; SCDecoder.decode,
// /root/reference/src/polar/decoder.py:38-115), and for N = 1024 as four length-256 codes in a row
// (N = 512 / 1024 / 2048: polar_sc_big_kernel<M>, further down).
//
// SC needs no path metric, no prune and no path copies, so the list kernel's machinery is dead
// weight.  Here a LANE decodes a whole frame and the decoder's entire state lives in registers:
//   * the tree is cut at the 32-leaf level: level 3 (32 LLRs) comes straight from the channel --
//     element k from the eight consecutive values at 8 br(k), as in pcl_level3_fused -- and the five
//     levels below it are a fully unrolled recursion (sc_node<32>), every array index a constant;
//   * partial sums are bit words: one word per 32-leaf node, folded upwards into the left arrays of
//     levels 3, 2 and 1 (1 + 2 + 4 registers);
//   * a frozen subtree costs nothing at all: its decisions are zero and, without a path metric,
//     nobody needs its LLRs (the reference computes them and throws them away) -- the whole visit,
//     channel loads included, is skipped behind a warp-uniform branch;
//   * the decisions leave as one 32-bit word per node, u[f][N / 32]; a second, HBM-bound kernel
//     gathers the K information bits into the byte rows of the C ABI.
// Arithmetic: f = sign sign min (one FMNMX.XORSIGN), g = b + (1 - 2u) a, decision bit = [x < 0]
// exactly as decoder.py:58-66, :121-144 -- in fp32; the fp64 validation build keeps the list kernel.
#pragma once
#include "pcl_common.cuh"
#include "pcl_tmem.cuh"

struct PolarScParams {
    const float* llr;               // [F][N] channel LLRs, reference index order
    uint32_t* uwords;               // [F][N / 32] decisions, decode-step order
    const uint32_t* frozen_words;   // decode-step-order frozen mask (N / 32 words)
    unsigned long long* next;       // ticket counter: a warp pulls 32 frames per ticket
    unsigned long long ticket_base;
    int64_t F;
};

PCL_DEVICE float4 pcl_ldg_f4(const float4* p)
{
#ifdef PCL_EMU
    return *p;
#else
    return __ldg(p);
#endif
}

PCL_DEVICE float2 pcl_ldg_f2(const float2* p)
{
#ifdef PCL_EMU
    return *p;
#else
    return __ldg(p);
#endif
}

PCL_DEVICE float sc_g(float a, float b, uint32_t bit_at_31)
{
    return b + __uint_as_float(__float_as_uint(a) ^ (bit_at_31 & 0x80000000u));
}

// One node of SZ leaves whose LLRs a[0 .. SZ) sit in registers.  fz: SZ frozen flags (uniform over the
// warp), returns the node's partial sums (SZ bits), u: its decisions.
template <int SZ>
struct sc_node {
    static PCL_DEVICE uint32_t run(const float* a, uint32_t fz, uint32_t& u)
    {
        constexpr uint32_t FULL = (SZ >= 32) ? 0xffffffffu : ((1u << SZ) - 1u);
        constexpr int H = SZ / 2;
        constexpr uint32_t HALF = (1u << H) - 1u;
        if ((fz & FULL) == FULL) { u = 0; return 0; }          // rate-0 subtree: nothing to compute
        float t[H];
        uint32_t ul = 0, ur = 0, cl = 0;
        if ((fz & HALF) != HALF) {
#pragma unroll
            for (int k = 0; k < H; k++) t[k] = pcl_math<float>::f(a[k], a[k + H]);
            cl = sc_node<H>::run(t, fz & HALF, ul);
        }
#pragma unroll
        for (int k = 0; k < H; k++) t[k] = sc_g(a[k], a[k + H], cl << (31 - k));
        const uint32_t cr = sc_node<H>::run(t, (fz >> H) & HALF, ur);
        u = ul | (ur << H);
        return (cl ^ cr) | (cr << H);
    }
};
template <>
struct sc_node<1> {
    static PCL_DEVICE uint32_t run(const float* a, uint32_t fz, uint32_t& u)
    {
        u = (fz & 1u) ? 0u : ((a[0] < 0.0f) ? 1u : 0u);        // decoder.py:58-66
        return u;
    }
};

// N = 256: eight nodes of 32 leaves.  b1 / b2 / b3: left arrays of levels 1 / 2 / 3 (128 / 64 / 32 bits).
// A lane reading its own frame straight from global memory touches 32 different lines per load
// instruction (65 Gbps, L1 wavefront bound); instead the warp copies its 32 frames ONCE, coalesced
// (cp.async, 512 contiguous bytes per instruction), into shared-memory rows of N + 4 floats -- the
// 16-byte pad makes the lanes' 16-byte reads of their own rows conflict free -- and every later
// visit of level 3 reads the channel from there.
#define PCL_SC256_ROW 260
#define PCL_SC256_WPB 3

// The eight 32-leaf nodes of a length-256 code whose "channel" row y (256 floats, reference index order) sits in
// shared memory: decisions go to uw[0 .. 8) (global, decode-step order), X receives the code's 256 partial sums in
// natural order (the left array a parent level would keep; dead code when the caller ignores it).
PCL_DEVICE void sc256_decode_row(const float* y, const uint32_t* fzw, uint32_t* uw, bool valid, uint32_t* X)
{
    uint32_t b1[4] = {0, 0, 0, 0}, b2[2] = {0, 0}, b3 = 0;
#pragma unroll 1
    for (int sb = 0; sb < 8; sb++) {
        const uint32_t fz = fzw[sb];
        uint32_t u = 0, c = 0;
        if (fz != 0xffffffffu) {
            const int bit1 = sb >> 2, bit2 = (sb >> 1) & 1, bit3 = sb & 1;
            float R[32];
#pragma unroll
            for (int k = 0; k < 32; k++) {
                // element k of level 3 <- channel values 8 br5(k) .. + 7 (pairs (0,1) (2,3) (4,5) (6,7) are the
                // level-1 elements k, k + 64, k + 32, k + 96)
                const int br = ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
                const float4 v0 = *reinterpret_cast<const float4*>(y + 8 * br);
                const float4 v1 = *reinterpret_cast<const float4*>(y + 8 * br + 4);
                float p01, p23, p45, p67, q0, q1;
                if (bit1) {
                    p01 = sc_g(v0.x, v0.y, b1[0] << (31 - k));
                    p23 = sc_g(v0.z, v0.w, b1[2] << (31 - k));
                    p45 = sc_g(v1.x, v1.y, b1[1] << (31 - k));
                    p67 = sc_g(v1.z, v1.w, b1[3] << (31 - k));
                } else {
                    p01 = pcl_math<float>::f(v0.x, v0.y);
                    p23 = pcl_math<float>::f(v0.z, v0.w);
                    p45 = pcl_math<float>::f(v1.x, v1.y);
                    p67 = pcl_math<float>::f(v1.z, v1.w);
                }
                if (bit2) {
                    q0 = sc_g(p01, p23, b2[0] << (31 - k));
                    q1 = sc_g(p45, p67, b2[1] << (31 - k));
                } else {
                    q0 = pcl_math<float>::f(p01, p23);
                    q1 = pcl_math<float>::f(p45, p67);
                }
                R[k] = bit3 ? sc_g(q0, q1, b3 << (31 - k)) : pcl_math<float>::f(q0, q1);
            }
            c = sc_node<32>::run(R, fz, u);
        }
        if (valid) uw[sb] = u;
        // fold the node's partial sums upwards (decoder.py:96-115): a left child parks its word, a right
        // child combines with its parked sibling and hands the pair on
        if ((sb & 1) == 0) {
            b3 = c;
        } else {
            const uint32_t x0 = b3 ^ c, x1 = c;                       // level-2 node: 64 bits
            if ((sb & 2) == 0) {
                b2[0] = x0; b2[1] = x1;
            } else if (sb == 3) {                                     // level-1 left node: 128 bits
                b1[0] = b2[0] ^ x0; b1[1] = b2[1] ^ x1; b1[2] = x0; b1[3] = x1;
            } else {                                                  // sb == 7: the right node, and with it the whole code
                const uint32_t r0 = b2[0] ^ x0, r1 = b2[1] ^ x1;
                X[0] = b1[0] ^ r0; X[1] = b1[1] ^ r1; X[2] = b1[2] ^ x0; X[3] = b1[3] ^ x1;
                X[4] = r0; X[5] = r1; X[6] = x0; X[7] = x1;
            }
        }
    }
}

__global__ void __launch_bounds__(32 * PCL_SC256_WPB) polar_sc256_kernel(PolarScParams P)
{
    constexpr int N = 256;
    const int lane = threadIdx.x & 31;
    float* rows = (float*)pcl_dyn_smem() + (size_t)(threadIdx.x >> 5) * 32 * PCL_SC256_ROW;
    for (;;) {
        unsigned long long tk = 0;
        if (lane == 0) tk = atomicAdd(P.next, 1ull) - P.ticket_base;
        tk = pcl_shfl_u64(tk, 0);
        const int64_t f0 = (int64_t)tk * 32;
        if (f0 >= P.F) break;
        const int64_t f = f0 + lane;
        const bool valid = f < P.F;
        __syncwarp();                                            // the previous pass has read its rows
#pragma unroll 4
        for (int r = 0; r < 64; r++) {                           // frame r / 2, half r % 2: 512 contiguous bytes
            const int64_t fr = (f0 + (r >> 1) < P.F) ? f0 + (r >> 1) : f0;
            pcl_cp_async16(rows + (r >> 1) * PCL_SC256_ROW + (r & 1) * 128 + 4 * lane, P.llr + fr * N + (r & 1) * 128 + 4 * lane);
        }
        pcl_cp_async_commit();
        pcl_cp_async_wait<0>();
        __syncwarp();
        uint32_t X[8];
        sc256_decode_row(rows + lane * PCL_SC256_ROW, P.frozen_words, P.uwords + (valid ? f : f0) * (N / 32), valid, X);
    }
}

// ---- N = 1024: four length-256 codes in a row ---------------------------------------------------------------
// The two top levels of the tree never exist as arrays.  Quarter q of the code (leaves 256 q ..) is itself a
// length-256 code whose channel values are the level-2 LLRs of that quarter, and in reference index order those
// come from FOUR CONSECUTIVE channel LLRs each:  row[j] = op2(op1(llr[4j], llr[4j+1]), op1(llr[4j+2], llr[4j+3])),
// op1 = f or g by bit 1 of q, op2 by bit 0 -- a sequential, fully coalesced stream, 512 bytes per instruction.
// So the warp (32 frames, a lane each, as above) makes four passes: all lanes together compute the 32 rows of a
// quarter into the same padded shared-memory rows the N = 256 kernel uses, then every lane decodes its own row with
// sc256_decode_row.  What g needs are the partial sums of the quarters already decoded, at natural index br8(j): each
// lane bit-reverses its frame's 256-bit sums once per quarter (three delta swaps between word pairs and a byte swap)
// and parks them in shared memory, [frame][24 words]: C = the previous quarter of the same half (level 2), A / B =
// the two halves of the level-1 left array, (X0 ^ X1, X1).
PCL_DEVICE void sc_bitrev256(uint32_t* X)
{
#pragma unroll
    for (int w = 0; w < 8; w++) X[w] = __byte_perm(X[w], 0, 0x3120);          // address bits 3 <-> 4
#pragma unroll
    for (int i = 0; i < 3; i++) {                                              // address bit i <-> word bit 2 - i
        const uint32_t m = (i == 0) ? 0x55555555u : (i == 1) ? 0x33333333u : 0x0f0f0f0fu;
        const int sh = 1 << i, d = 4 >> i;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            if ((w & d) == 0) {
                const uint32_t t = ((X[w] >> sh) ^ X[w + d]) & m;
                X[w + d] ^= t;
                X[w] ^= t << sh;
            }
        }
    }
}


// ---- N = 256 M, M = 2 / 4 / 8 / 16: M length-256 codes in a row ----------------------------------------------------
// The m = log2 M top levels never exist as arrays: sub-block q is a length-256 code whose channel values, in
// reference index order, come from M CONSECUTIVE channel LLRs each through a tree of adjacent pairs -- level l
// (1 = top) combines e[2t], e[2t+1] with f or g by bit (m - l) of q.  The partial sums g needs are the level-l
// LEFT array of the node that contains q: 2^(m-l) chunks of 256 bits, element t reading chunk rev(t) at natural
// index br8(j).  Every lane keeps its frame's left arrays bit-reversed inside each chunk (so position j reads bit j)
// in shared memory, [frame][8 (M - 1) words]: level l at word 8 (2^(m-l) - 1); after a sub-block the lane folds its
// 256 new sums upwards through the parked arrays -- (left ^ right, right) chunk by chunk -- until it meets a level
// where the node is a left child, and parks them there.  M = 4 is the N = 1024 layout described above.
template <int M>
__global__ void __launch_bounds__(M == 4 ? 96 : 192) polar_sc_big_kernel(PolarScParams P)
{
    constexpr int m = (M == 2) ? 1 : (M == 4) ? 2 : (M == 8) ? 3 : 4;
    constexpr int N = 256 * M;
    constexpr int PW = 8 * (M - 1);                  // parked words per frame
    constexpr int U = 32 / M;                        // lane-positions per pipeline stage (32 floats per lane)
    static_assert(M == 2 || M == 4 || M == 8 || M == 16, "256 M with M = 2, 4, 8, 16");
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float* rows = (float*)pcl_dyn_smem() + (size_t)warp * 32 * PCL_SC256_ROW;
    uint32_t* park = (uint32_t*)((float*)pcl_dyn_smem() + (size_t)(blockDim.x >> 5) * 32 * PCL_SC256_ROW) + (size_t)warp * 32 * PW;
    for (;;) {
        unsigned long long tk = 0;
        if (lane == 0) tk = atomicAdd(P.next, 1ull) - P.ticket_base;
        tk = pcl_shfl_u64(tk, 0);
        const int64_t f0 = (int64_t)tk * 32;
        if (f0 >= P.F) break;
        const int64_t f = f0 + lane;
        const bool valid = f < P.F;
#pragma unroll 1
        for (int q = 0; q < M; q++) {
            __syncwarp();                            // rows and parked words of the previous sub-block are consumed / written
            // ---- the 32 rows of this sub-block.  Lane-position lp = 8 r + it: frame r, position j = 32 it + lane,
            // M consecutive floats at M j; stage s + 1 is in flight while stage s is computed ----
            auto fetch = [&](int st, float* v) {
#pragma unroll
                for (int g8 = 0; g8 < (U + 7) / 8; g8++) {               // one frame per group of up to 8 lane-positions
                    const int lp0 = st * U + 8 * g8, r = lp0 >> 3, it0 = lp0 & 7;
                    const int64_t fr = (f0 + r < P.F) ? f0 + r : f0;
                    const float* src = P.llr + fr * N + M * (32 * it0 + lane);
#pragma unroll
                    for (int i = 0; i < (U < 8 ? U : 8); i++) {
                        float* o = v + M * (8 * g8 + i);
                        if (M == 2) {
                            const float2 t = pcl_ldg_f2(reinterpret_cast<const float2*>(src + M * 32 * i));
                            o[0] = t.x; o[1] = t.y;
                        } else {
#pragma unroll
                            for (int h = 0; h < M / 4; h++) {
                                const float4 t = pcl_ldg_f4(reinterpret_cast<const float4*>(src + M * 32 * i) + h);
                                o[4 * h] = t.x; o[4 * h + 1] = t.y; o[4 * h + 2] = t.z; o[4 * h + 3] = t.w;
                            }
                        }
                    }
                }
            };
            auto emit = [&](int st, const float* v) {
                // the words g reads, shifted so that this lane's bit sits at bit 31: level l, result t -> chunk rev(t)
                uint32_t wsh[U][M - 1];
#pragma unroll
                for (int i = 0; i < U; i++) {
                    const int lp = st * U + i, r = lp >> 3, it = lp & 7;
                    const uint32_t* pk = park + r * PW + it;
                    int idx = 0;
#pragma unroll
                    for (int l = 1; l <= m; l++) {
                        const bool isg = (q >> (m - l)) & 1;
#pragma unroll
                        for (int t = 0; t < (M >> l); t++, idx++) {
                            int c = 0;
#pragma unroll
                            for (int bb = 0; bb < m - l; bb++) c |= ((t >> bb) & 1) << (m - l - 1 - bb);
                            wsh[i][idx] = isg ? (pk[8 * ((1 << (m - l)) - 1) + 8 * c] << (31 - lane)) : 0u;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < U; i++) {
                    const int lp = st * U + i, r = lp >> 3, it = lp & 7;
                    float e[M];
#pragma unroll
                    for (int t = 0; t < M; t++) e[t] = v[M * i + t];
                    int idx = 0;
#pragma unroll
                    for (int l = 1; l <= m; l++) {                       // level l: M >> l results from adjacent pairs
                        const bool isg = (q >> (m - l)) & 1;
#pragma unroll
                        for (int t = 0; t < (M >> l); t++, idx++) {
                            if (isg) e[t] = sc_g(e[2 * t], e[2 * t + 1], wsh[i][idx]);
                            else e[t] = pcl_math<float>::f(e[2 * t], e[2 * t + 1]);
                        }
                    }
                    rows[r * PCL_SC256_ROW + 32 * it + lane] = e[0];
                }
            };
            {
                constexpr int NST = 256 / U;                             // stages per sub-block
                // two stages (2 x 32 floats per lane, 8 KB per warp) in flight while a third is combined: the ncu capture
                // of the two-buffer version showed 6 warps per SM waiting on these loads (long_scoreboard 1.7 cycles / issue)
                float va[32], vb[32], vc[32];
                fetch(0, va);
                fetch(1, vb);
#pragma unroll 1
                for (int st = 0; st < NST; st += 3) {
                    if (st + 2 < NST) fetch(st + 2, vc);
                    emit(st, va);
                    if (st + 3 < NST) fetch(st + 3, va);
                    if (st + 1 < NST) emit(st + 1, vb);
                    if (st + 4 < NST) fetch(st + 4, vb);
                    if (st + 2 < NST) emit(st + 2, vc);
                }
            }
            __syncwarp();
            // ---- every lane decodes its own row ----
            uint32_t X[8];
#pragma unroll
            for (int w = 0; w < 8; w++) X[w] = 0;
            sc256_decode_row(rows + lane * PCL_SC256_ROW, P.frozen_words + 8 * q, P.uwords + (valid ? f : f0) * (N / 32) + 8 * q,
                             valid, X);
            if (q == M - 1) break;
            // ---- fold the new sums upwards and park them (own frame only: no barrier needed before the next staging
            // except the one at the top of the loop) ----
            sc_bitrev256(X);
            uint32_t* mine = park + lane * PW;
            uint32_t cur[8 * (M / 2)];
#pragma unroll
            for (int w = 0; w < 8; w++) cur[w] = X[w];
            bool done = false;
#pragma unroll
            for (int l = m; l >= 1; l--) {
                const int nch = 1 << (m - l);                            // chunks of the node at this level (compile time after unrolling)
                const int off = 8 * (nch - 1);
                if (!done) {
                    if (((q >> (m - l)) & 1) == 0) {                     // left child: park
#pragma unroll
                        for (int w = 0; w < 8 * (M / 2); w++)
                            if (w < 8 * nch) mine[off + w] = cur[w];
                        done = true;
                    } else if (l > 1) {                                  // right child: (left ^ right, right), twice the chunks
#pragma unroll
                        for (int w = 0; w < 8 * (M / 2); w++)
                            if (w < 8 * nch && 8 * nch + w < 8 * (M / 2)) {
                                cur[8 * nch + w] = cur[w];
                                cur[w] ^= mine[off + w];
                            }
                    }
                }
            }
        }
    }
}

// u words (decode-step order) -> K information bits per frame, ascending reference index (decoder.py:70-71)
__global__ void __launch_bounds__(256) polar_sc_extract_kernel(const uint32_t* uwords, const uint16_t* info_pos, int64_t F,
                                                               int NW, int K, uint8_t* bits)
{
    const int64_t total = F * (int64_t)K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = i / K;
        const int k = (int)(i - f * K);
        const int pos = info_pos[k];
        bits[i] = (uint8_t)((uwords[f * NW + (pos >> 5)] >> (pos & 31)) & 1u);
    }
}
