// framegen.cuh -- on-device frame generation for BER / FER sweeps.
//
// One warp builds one frame: random message -> encode -> BPSK + AWGN -> LLR, the pipeline
// the reference's callers run per frame on the host
// (/root/reference/benchmarks/benchmark_scl.py:95-103, test_snr_curves.py:121-130):
//   * message bits and noise come from Philox4x32-10, a counter-based generator keyed by the
//     seed and addressed by (GLOBAL frame index, element block, stream): a frame's content
//     does not depend on how a sweep is sharded over ranks, chunks or launches;
//   * polar encode x = u F^{(x)n} (src/polar/utils.py:193-229: x[i] ^= x[i + stride], stride
//     1, 2, .., N/2, natural order) on bit-packed words: the five in-word stages are
//     shift-and-mask steps, the word stages XOR whole words;
//   * LDPC encode c = m G mod 2 (src/ldpc/encoder.py:88-90) as the XOR of the bit-packed rows
//     of G selected by the message bits;
//   * BPSK 0 -> +1, 1 -> -1, y = s + sigma z, LLR = 2 y / sigma^2 (src/channel/awgn.py:47,75,88),
//     z from Box-Muller on two Philox words;
//   * Rayleigh fading with perfect channel knowledge (src/channel/fading.py:26-52): y = |h| s +
//     sigma z, LLR = 2 y |h| / sigma^2, |h|^2 = hr^2 + hi^2 with hr, hi ~ N(0, 1/2), i.e. |h|^2 is
//     exponential with mean 1: |h| = sqrt(-ln u);
//   * binary symmetric channel (src/channel/bsc.py:24-39): the bit flips with probability p; the
//     decoders are fed LLR = +-ln((1 - p) / p) of the received bit.
// The host numpy path (channel/awgn.py: transmit_batch) stays the one that reproduces the
// reference's np.random stream exactly; this one removes the host RNG and the PCIe copy from
// sweeps.
#pragma once
#include "pcl_common.cuh"

struct pcl_philox4 { uint32_t x, y, z, w; };

PCL_HOST_DEVICE uint32_t pcl_mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1)
PCL_HOST_DEVICE pcl_philox4 pcl_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = pcl_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = pcl_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    pcl_philox4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

enum { PCL_STREAM_MSG = 0, PCL_STREAM_NOISE = 1, PCL_STREAM_FADE = 2 };
#ifndef PCL_CH_AWGN          // include/pcl.h
#define PCL_CH_AWGN 0
#define PCL_CH_RAYLEIGH 1
#define PCL_CH_BSC 2
#endif

struct GenParams {
    int kind;                     // 0 polar, 1 LDPC
    int N, K, NW;                 // code length, message length, ceil(N / 32)
    const uint32_t* info_words;   // polar: info-position mask, bit i of the codeword index space
    const uint16_t* info_rank;    // polar: number of info positions before word w
    const uint32_t* G;            // LDPC: [K][NW] bit-packed rows of the generator matrix
    int64_t F, frame0;
    uint32_t seed_lo, seed_hi;
    float sigma, scale;           // noise std and 2 / sigma^2
    double scale64, sigma64;
    int channel;                  // PCL_CH_*
    float bsc_p, bsc_llr;         // BSC: crossover probability and ln((1 - p) / p)
    int f64;                      // LLR output type
    uint8_t* msg;                 // [F][K] message bytes (optional)
    uint8_t* cw;                  // [F][N] codeword bytes (optional)
    void* llr;                    // [F][N] float or double
};

// word j of the frame's random message bit string
PCL_DEVICE uint32_t pcl_gen_msg_word(const GenParams& P, uint64_t frame, int j)
{
    const pcl_philox4 r = pcl_philox4x32_10((uint32_t)frame, (uint32_t)(frame >> 32), (uint32_t)(j >> 2), PCL_STREAM_MSG,
                                            P.seed_lo, P.seed_hi);
    const int s = j & 3;
    return s == 0 ? r.x : s == 1 ? r.y : s == 2 ? r.z : r.w;
}

__global__ void __launch_bounds__(128) framegen_kernel(GenParams P)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    uint32_t* U = (uint32_t*)pcl_dyn_smem() + (size_t)warp * P.NW;       // this frame's codeword words
    const int NW = P.NW, N = P.N, K = P.K;
    for (int64_t f = (int64_t)blockIdx.x * wpb + warp; f < P.F; f += (int64_t)gridDim.x * wpb) {
        const uint64_t frame = (uint64_t)(P.frame0 + f);
        if (P.kind == 0) {
            // u[info positions] = message bits in ascending index order (src/polar/encoder.py:74-80)
            for (int w = lane; w < NW; w += 32) {
                uint32_t m = P.info_words[w];
                int k = P.info_rank[w];
                uint32_t word = 0, rnd = 0;
                int have = -1;
                while (m) {
                    const int b = __ffs((int)m) - 1;
                    m &= m - 1;
                    if ((k >> 5) != have) { have = k >> 5; rnd = pcl_gen_msg_word(P, frame, have); }
                    const uint32_t bit = (rnd >> (k & 31)) & 1u;
                    word |= bit << b;
                    if (P.msg != nullptr) P.msg[f * K + k] = (uint8_t)bit;
                    k++;
                }
                // in-word stages: stride 1, 2, 4, 8, 16
                word ^= (word >> 1) & 0x55555555u;
                word ^= (word >> 2) & 0x33333333u;
                word ^= (word >> 4) & 0x0F0F0F0Fu;
                word ^= (word >> 8) & 0x00FF00FFu;
                word ^= (word >> 16) & 0x0000FFFFu;
                U[w] = word;
            }
            __syncwarp();
            for (int s = 1; s < NW; s <<= 1) {                        // word stages: stride 32 s
                for (int w = lane; w < NW; w += 32)
                    if ((w & s) == 0) U[w] ^= U[w + s];
                __syncwarp();
            }
        } else {
            // c = m G mod 2: XOR of the rows of G selected by the message bits
            for (int w0 = 0; w0 < NW; w0 += 32) {
                const int w = w0 + lane;
                uint32_t acc = 0;
                for (int j = 0; j < (K + 31) >> 5; j++) {
                    uint32_t rnd = pcl_gen_msg_word(P, frame, j);
                    if (j == (K >> 5)) rnd &= (1u << (K & 31)) - 1u;   // K % 32 != 0 here
                    if (P.msg != nullptr && w0 == 0) {
                        const int k = 32 * j + lane;
                        if (k < K) P.msg[f * K + k] = (uint8_t)((rnd >> lane) & 1u);
                    }
                    while (rnd) {
                        const int b = __ffs((int)rnd) - 1;
                        rnd &= rnd - 1;
                        if (w < NW) acc ^= P.G[(size_t)(32 * j + b) * NW + w];
                    }
                }
                if (w < NW) U[w] = acc;
            }
            __syncwarp();
        }
        // channel -> LLR; element i = 32 t + lane, four elements per Philox call
        for (int t0 = 0; 32 * t0 < N; t0 += 4) {
            const uint32_t blk = (uint32_t)((t0 >> 2) * 32 + lane);
            const pcl_philox4 r = pcl_philox4x32_10((uint32_t)frame, (uint32_t)(frame >> 32), blk, PCL_STREAM_NOISE,
                                                    P.seed_lo, P.seed_hi);
            float z[4], hm[4];
            if (P.channel == PCL_CH_BSC) {
                z[0] = (float)r.x; z[1] = (float)r.y; z[2] = (float)r.z; z[3] = (float)r.w;
            } else {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t a = h ? r.z : r.x, b = h ? r.w : r.y;
                    const float u1 = fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // (a + 0.5) 2^-32
                    const float rad = sqrtf(-2.0f * logf(u1));
                    float sn, cs;
                    sincospif((float)b * 4.656612873077393e-10f, &sn, &cs);                             // 2 pi b 2^-32
                    z[2 * h] = rad * cs;
                    z[2 * h + 1] = rad * sn;
                }
            }
            hm[0] = hm[1] = hm[2] = hm[3] = 1.0f;
            if (P.channel == PCL_CH_RAYLEIGH) {
                const pcl_philox4 g = pcl_philox4x32_10((uint32_t)frame, (uint32_t)(frame >> 32), blk, PCL_STREAM_FADE,
                                                        P.seed_lo, P.seed_hi);
                const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                for (int q = 0; q < 4; q++)
                    hm[q] = sqrtf(-logf(fmaf((float)gw[q], 2.3283064365386963e-10f, 1.1641532182693481e-10f)));
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int t = t0 + q;
                const int i = 32 * t + lane;
                if (i < N) {
                    const uint32_t bit = (U[t] >> lane) & 1u;
                    if (P.cw != nullptr) P.cw[f * N + i] = (uint8_t)bit;
                    const float s = bit ? -1.0f : 1.0f;
                    if (P.channel == PCL_CH_BSC) {
                        const bool flip = z[q] * 2.3283064365386963e-10f < P.bsc_p;      // uniform [0, 1)
                        const float v = (flip ? -s : s) * P.bsc_llr;
                        if (P.f64) ((double*)P.llr)[f * N + i] = (double)v;
                        else ((float*)P.llr)[f * N + i] = v;
                    } else if (P.f64) {
                        ((double*)P.llr)[f * N + i] = ((double)hm[q] * (double)s + P.sigma64 * (double)z[q]) * (double)hm[q] * P.scale64;
                    } else {
                        ((float*)P.llr)[f * N + i] = fmaf(P.sigma, z[q], hm[q] * s) * hm[q] * P.scale;
                    }
                }
            }
        }
        __syncwarp();
    }
}
