/*
 * pcl.h -- C ABI of libpcl.so, the B200 (sm_100a) batched Polar / LDPC decoder.
 *
 * The reference (B1ear/PolarCode_and_LDPC) is pure Python and has no FFI layer;
 * its boundary for the decode hot path is the Python class API.  Each entry
 * point below names the reference interface it stands behind (paths relative to
 * the reference root).  The Python drop-in classes in polarcode_and_ldpc_b200/
 * bind these with ctypes (see INTEGRATION.md for the stub a reference
 * maintainer would add).
 *
 * Conventions: plain pointers and sizes, no torch types.  *_dev pointers are
 * device pointers borrowed for the duration of the call (never freed here);
 * *_host pointers are host memory (pinned memory makes the copies asynchronous).
 * `stream` is a cudaStream_t passed as void* (NULL = default stream).  Every
 * function returns 0 on success, a PCL_E* code otherwise; pcl_last_error() gives
 * the text.  A handle may be used from one stream at a time.  There is no CPU
 * fallback: without a CUDA device every compute call fails with PCL_ECUDA.
 */
#ifndef PCL_H
#define PCL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCL_OK 0
#define PCL_EINVAL 1      /* bad argument (AssertionError on the Python side) */
#define PCL_ECUDA 2       /* CUDA runtime error / no device */
#define PCL_EUNSUPPORTED 3
#define PCL_EDEGREE1 4    /* Min-Sum on a degree-1 check: reference raises ValueError */

#define PCL_F32 0         /* production compute type */
#define PCL_F64 1         /* validation build: bit-exact decoded bits vs the reference */
#define PCL_F16 2         /* TRANSPORT format of host LLRs only (half the PCIe bytes); widened to fp32 on the device */

#define PCL_OUT_BYTES 0   /* uint8 [F][W], one byte per bit */
#define PCL_OUT_PACKED 1  /* uint32 [F][ceil(W/32)], bit k of a row at word k / 32, bit k % 32 (packed on the device) */
#define PCL_OUT_INT64 2   /* int64 [F][W]: what F calls of the reference's decode() return (np.int64) */

#define PCL_LDPC_BP 0     /* BPDecoder  (src/ldpc/decoder.py:11)  */
#define PCL_LDPC_MS 1     /* MSDecoder  (src/ldpc/decoder.py:208) */

typedef struct pcl_polar pcl_polar_t;
typedef struct pcl_ldpc pcl_ldpc_t;
typedef struct pcl_gen pcl_gen_t;

int pcl_version(void);
const char* pcl_last_error(void);
int pcl_device_count(void);

/*
 * SCDecoder.__init__ (src/polar/decoder.py:16-36) / SCLDecoder.__init__ (:191-223).
 * frozen_mask[l] != 0 marks reference bit index l as frozen (N entries); K must
 * equal the number of zeros.  list_size 1 with want_metric 0 is the SC decoder.
 * crc_len 0 disables CRC-aided selection (the reference never applies it,
 * decoder.py:259); otherwise crc_poly/crc_len follow src/polar/utils.py:128-163.
 * Limits (PCL_EUNSUPPORTED beyond): N <= 65536 (16-bit leaf positions; a configuration whose per-frame bit arrays
 * do not fit one block's shared memory is refused at creation), list_size <= 1024.
 */
int pcl_polar_create(pcl_polar_t** out, int N, int K, int list_size, const uint8_t* frozen_mask,
                     int crc_len, uint32_t crc_poly, int dtype);
void pcl_polar_destroy(pcl_polar_t* h);

/*
 * SCDecoder.decode (src/polar/decoder.py:38-71) / SCLDecoder.decode (:225-262) over
 * F frames.  llr_dev: [F][N] in the handle's dtype.  bits_dev: [F][K] bytes, the
 * info bits in ascending reference index (what decode() returns).  Optional
 * outputs (NULL to skip): pm_dev [F][list_size] doubles = SCLDecoder.path_metrics;
 * leaf_dev [F][N][LP] (dtype) + parent_dev [F][N][LP] bytes = per-step leaf LLR
 * and survivor parent of every list slot (LP = list_size rounded up to a power of
 * two, pcl_polar_lp()), from which the host rebuilds L_paths[best, :, n].
 */
int pcl_polar_decode_batch(pcl_polar_t* h, const void* llr_dev, int64_t F, uint8_t* bits_dev,
                           double* pm_dev, void* leaf_dev, uint8_t* parent_dev, void* stream);

/* Same call with HOST buffers: chunks the batch, overlaps H2D / decode / D2H. */
int pcl_polar_decode_host(pcl_polar_t* h, const void* llr_host, int64_t F, uint8_t* bits_host,
                          void* stream);
/*
 * The reference call shape end to end: decode(llr) takes any float array (np.float64 in every
 * caller, src/polar/decoder.py:47) and returns np.int64[K] (:70-71).  llr_host [F][N] in
 * `llr_dtype` -- PCL_F64 from pageable memory is narrowed to fp32 by host threads into pinned
 * staging (fp32 handles; a float64 handle copies it as is), PCL_F32 is copied directly, PCL_F16 is
 * an opt-in transport format widened on the device -- and out_host in `out_format`
 * (PCL_OUT_BYTES / PCL_OUT_PACKED / PCL_OUT_INT64).  Chunked; conversion, H2D, decode, D2H and
 * unpacking of different chunks overlap.
 */
int pcl_polar_decode_host_ex(pcl_polar_t* h, const void* llr_host, int llr_dtype, int64_t F, void* out_host,
                             int out_format, void* stream);
int pcl_polar_lp(const pcl_polar_t* h);
/* Launch geometry of the last decode (for gpu_launches / occupancy reporting): grid, block,
 * dynamic shared memory, number of tree levels kept in the L2 scratch, and the kernel in use:
 * 0 = generic kernel, 1 = register-resident-bottom kernel (a lane owns a path, 32 / LP frames
 * per warp), 2 = the same with log2 N and the level split compiled in as constants, 3 = the
 * one-block-per-SM variant with the mid tree levels in tensor / shared memory, 4 = the
 * register-resident SC kernels (N = 256; N = 512 … 4096 as 2 … 16 length-256 codes in a row), 5 = the block-per-frame kernel for list sizes 33 .. 1024
 * (any list_size >= 1 is what the reference accepts, src/polar/decoder.py:194-196). */
int pcl_polar_launch_info(const pcl_polar_t* h, int* grid, int* block, int* smem_bytes, int* glevels,
                          int* fast);

/*
 * BPDecoder.__init__ / MSDecoder.__init__ + _build_tanner_graph
 * (src/ldpc/decoder.py:18-60, :215-255).  H_dense: m*n row-major bytes; entries
 * equal to 1 are edges.  normalization is MSDecoder's factor (ignored for BP).
 */
int pcl_ldpc_create(pcl_ldpc_t** out, int m, int n, const uint8_t* H_dense, int mode,
                    double normalization, int max_iter, int early_stop, int dtype);
void pcl_ldpc_destroy(pcl_ldpc_t* h);

/*
 * BPDecoder.decode (src/ldpc/decoder.py:124-202) / MSDecoder.decode (:289-352) over
 * F frames.  llr_dev [F][n] (dtype); bits_dev [F][n] bytes (whole codeword, as the
 * reference returns); optional iters_dev [F] int32 (return_iterations=True) and
 * total_dev [F][n] (dtype) = total LLRs of the last executed iteration.
 */
int pcl_ldpc_decode_batch(pcl_ldpc_t* h, const void* llr_dev, int64_t F, uint8_t* bits_dev,
                          int32_t* iters_dev, void* total_dev, void* stream);
int pcl_ldpc_decode_host(pcl_ldpc_t* h, const void* llr_host, int64_t F, uint8_t* bits_host,
                         int32_t* iters_host, void* stream);
/* Host-buffer call with the input / output formats of pcl_polar_decode_host_ex (BPDecoder.decode,
 * src/ldpc/decoder.py:124: float64 in, np.int64[n] out). */
int pcl_ldpc_decode_host_ex(pcl_ldpc_t* h, const void* llr_host, int llr_dtype, int64_t F, void* out_host,
                            int out_format, int32_t* iters_host, void* stream);
int pcl_ldpc_num_edges(const pcl_ldpc_t* h);
int pcl_ldpc_launch_info(const pcl_ldpc_t* h, int* grid, int* block, int* smem_bytes);
/* Shared-memory layout in use: *banked = 1 for the conflict-free layout of regular (3, 6) codes in
 * the fp32 build (*residual = message fetches that still share a bank, per iteration), *coop = 1
 * when a whole block decodes one frame (large codes). */
int pcl_ldpc_layout_info(const pcl_ldpc_t* h, int* banked, int* residual, int* coop);

/*
 * On-device frame generation for BER / FER sweeps: the per-frame pipeline of the reference's
 * callers (benchmarks/benchmark_scl.py:95-103, test_snr_curves.py:121-130):
 *   message = randint(0, 2, K); codeword = encoder.encode(message); llr = AWGNChannel(snr).transmit(codeword)
 * with PolarEncoder.encode (src/polar/encoder.py:52-95, transform src/polar/utils.py:193-229),
 * LDPCEncoder.encode (src/ldpc/encoder.py:76-93, c = m G mod 2) and AWGNChannel.transmit
 * (src/channel/awgn.py:47,75,88).  Random bits and noise come from Philox4x32-10 keyed by
 * `seed` and addressed by the GLOBAL frame index frame0 + f, so a frame's content does not
 * depend on how a sweep is sharded.  frozen_mask[N] as in pcl_polar_create; G_dense [k][n]
 * bytes, row-major.  Outputs (device): msg_dev [F][K] bytes (optional), cw_dev [F][N] bytes
 * (optional), llr_dev [F][N] in `dtype`.
 */
int pcl_gen_polar_create(pcl_gen_t** out, int N, int K, const uint8_t* frozen_mask);
int pcl_gen_ldpc_create(pcl_gen_t** out, int n, int k, const uint8_t* G_dense);
void pcl_gen_destroy(pcl_gen_t* h);
int pcl_gen_frames(pcl_gen_t* h, int64_t F, int64_t frame0, unsigned long long seed, double snr_db,
                   int dtype, uint8_t* msg_dev, uint8_t* cw_dev, void* llr_dev, void* stream);
/* Same with a channel choice: PCL_CH_AWGN (param = SNR in dB), PCL_CH_RAYLEIGH (param = average
 * SNR in dB; RayleighFadingChannel.transmit, src/channel/fading.py:26-52: y = |h| s + n,
 * LLR = 2 y |h| / sigma^2) or PCL_CH_BSC (param = crossover probability in (0, 1);
 * BSCChannel.transmit, src/channel/bsc.py:24-39, delivered as LLR = +-ln((1 - p) / p)). */
#define PCL_CH_AWGN 0
#define PCL_CH_RAYLEIGH 1
#define PCL_CH_BSC 2
int pcl_gen_frames_channel(pcl_gen_t* h, int64_t F, int64_t frame0, unsigned long long seed, int channel,
                           double param, int dtype, uint8_t* msg_dev, uint8_t* cw_dev, void* llr_dev,
                           void* stream);
/* One Philox4x32-10 block evaluated on the host from the same source (known-answer tests). */
void pcl_philox4x32_10_host(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/*
 * np.sum(message != decoded) loops of the callers (benchmarks/test_snr_curves.py:
 * 133-138).  Adds to counters_dev[4] (uint64): bit errors, frame errors, frames,
 * bits compared.  Rows are `width` bytes; only the first `ncmp` bytes of each row
 * are compared (LDPC callers compare decoded[:k]).  The counter tensor is what the
 * one NCCL allreduce of a multi-GPU sweep reduces.
 */
int pcl_count_errors(const uint8_t* bits_dev, const uint8_t* ref_dev, int64_t F, int width, int ncmp,
                     unsigned long long* counters_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCL_H */
