// pcl_api.cu -- C ABI (include/pcl.h) over the sm_100a decoder kernels.
//
// Host side of the drop-in boundary: validates arguments the way the reference
// constructors do, precomputes the code tables once per handle (decode-step-order
// frozen mask and info gather map for polar; check-major / variable-major edge
// tables for LDPC), sizes shared memory / scratch / grid for the B200 and launches
// the kernels.  No torch types, no CPU decode path.
#include "pcl_common.cuh"
#include "polar_scl.cuh"
#include "polar_scl_fast.cuh"
#include "polar_scl_wide.cuh"
#include "polar_sc.cuh"
#include "framegen.cuh"
#include "ldpc_banked.cuh"
#include "ldpc_layout.h"
#include "ldpc_bp.cuh"
#include "../../include/pcl.h"
#include "pcl_host_pipe.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <map>
#include <mutex>
#include <vector>

#define PCL_VERSION_NUM 100

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                  \
    do {                                                                                \
        cudaError_t e_ = (expr);                                                        \
        if (e_ != cudaSuccess)                                                          \
            return fail(PCL_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));     \
    } while (0)

#ifdef PCL_EMU
#define PCL_LAUNCH(kern, grid, block, smem, stream, arg) \
    simt::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kern(arg); })
static int env_int(const char* name, int dflt) { const char* s = getenv(name); return s ? atoi(s) : dflt; }
#else
#define PCL_LAUNCH(kern, grid, block, smem, stream, arg) \
    kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(arg)
static int env_int(const char* name, int dflt) { const char* s = getenv(name); return s ? atoi(s) : dflt; }
#endif

extern "C" int pcl_version(void) { return PCL_VERSION_NUM; }
extern "C" const char* pcl_last_error(void) { return g_err.c_str(); }

extern "C" int pcl_device_count(void)
{
#ifdef PCL_EMU
    return 0;
#else
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
#endif
}

struct DeviceInfo { int sms; int smem_per_sm; int smem_per_block; };

static int device_info(DeviceInfo* di)
{
#ifdef PCL_EMU
    di->sms = 2; di->smem_per_sm = 228 * 1024; di->smem_per_block = 227 * 1024;
    return PCL_OK;
#else
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&di->sms, cudaDevAttrMultiProcessorCount, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&di->smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&di->smem_per_block, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    return PCL_OK;
#endif
}

#ifndef PCL_EMU
// The small helper kernels (error counters, frame generator) ask for the same shared-memory
// carve-out as the decoders they run between: an SM that has to change its L1 / shared split
// between two launches was seen to take the next decode kernel's blocks at the OLD split (fewer
// resident blocks; the rest of the grid runs as a tail), 8.0 instead of 10.0 Gbps (SCL-8) and 3.6
// instead of 9.0 Gbps (BP) in bench runs of one binary.  PCL_AUX_CARVEOUT=-1 restores the default.
static void aux_carveout(const void* kern)
{
    const int pct = env_int("PCL_AUX_CARVEOUT", 100);
    if (pct >= 0) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
}
#endif

static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }
static inline int ilog2i(int v) { int n = 0; while ((1 << n) < v) n++; return n; }
static inline int bit_reverse_i(int v, int nbits)
{
    int r = 0;
    for (int i = 0; i < nbits; i++) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

// =============================================================== polar =========
#define PCL_NSTAGE 3

struct pcl_polar {
    int N, n, K, L, LP, dtype;
    int crc_len; uint32_t crc_poly;
    PolarLayout lay;
    uint32_t* d_frozen_words = nullptr;
    uint16_t* d_info_pos = nullptr;
    void* d_scratch[PCL_NSTAGE] = {};
    size_t scratch_bytes = 0;
    int wpb, grid_max, smem_bytes;
    int last_grid = 0;
    int fast = 0;               // register-resident tree bottom (polar_scl_fast.cuh)
    int wide = 0;               // list size above 32: one block per frame, a thread per slot (polar_scl_wide.cuh)
    int fpw = 1;                // fast kernel: frames per warp
    int NL = 0, GL = 0;         // fast kernel: compiled-in log2 N and G (0: run-time values)
    int TM = 0;                 // fast kernel, TM variant: one block per SM, mid levels in tensor / shared memory
    int sc256 = 0;              // list size 1, N = 256 / 512 / 1024 / 2048, fp32: the register-resident SC kernels (polar_sc.cuh)
    uint32_t* d_uwords[PCL_NSTAGE] = {};   // its decision words [F][8], per host-pipeline stage
    int64_t uwords_cap[PCL_NSTAGE] = {};
    unsigned long long* d_next[PCL_NSTAGE] = {};   // TM variant: ticket counters (one per host-pipeline stage)
    unsigned long long tickets[PCL_NSTAGE] = {};   // their current values (never reset: no memset between launches)
#ifndef PCL_EMU
    HostPipe pipe;              // host-buffer pipeline (pcl_host_pipe.cuh)
#endif
    std::vector<uint8_t> frozen_copy;              // the caller's mask (twin creation)
    pcl_polar* twin = nullptr;  // run-time-N handle for per-leaf dumps when this one has a compiled code length
};

static size_t real_size(int dtype) { return dtype == PCL_F64 ? 8 : 4; }

static void polar_layout(PolarLayout& Y, int N, int K, int L, int LP, int G, int rsz, bool crc, bool fast = false,
                         int fpw = 1, bool tm = false)
{
    Y.hdr_bytes = 0;
    if (tm) {
        // TM variant (polar_scl_fast.cuh): levels 2 .. n-7 in the global scratch, n-6 and n-4 in tensor
        // memory, n-5 (32 elements per column) in shared memory; the final u words reuse that region.
        const int cols = 32;
        Y.N = N; Y.n = ilog2i(N); Y.K = K; Y.L = L; Y.G = Y.n - 7;
        Y.NW = N / 32; Y.nb = Y.n - 5;
        Y.uw_slots = crc ? cols : fpw;
        Y.hdr_bytes = 256;                          // TMEM base + two 8-byte ticket slots per group of warps
        int off = 0;
        Y.off_cm = off;     off += fpw * pcl_fast_frame_bytes(LP);
        Y.off_newpm = off;
        Y.off_llr = off;    off += std::max(cols * 32 * rsz, Y.uw_slots * Y.NW * 4);
        Y.off_uw = Y.off_llr;                      // written after the last level walk of a frame
        Y.off_sel = off;
        Y.off_bw = off;     off += cols * (N / 32 - 1) * 4;
        Y.warp_bytes = align_up(off, 16);
        Y.scratch_per_warp = (int64_t)cols * ((N >> 2) - (N >> Y.G));      // levels 3 .. n-7 (1 and 2 are never stored)
        return;
    }
    // `cols` = (frame, slot) columns a warp carries: LP for the generic kernel, LP * fpw for the fast one
    const int cols = LP * fpw;
    Y.N = N; Y.n = ilog2i(N); Y.K = K; Y.L = L; Y.G = G;
    Y.NW = N >= 32 ? N / 32 : 1;
    Y.nb = Y.n > 5 ? Y.n - 5 : 0;
    Y.uw_slots = crc ? cols : fpw;
    int off = 0;
    if (fast) {
        // one prune scratch block per frame of the warp (pcl_fast_frame_bytes)
        Y.off_cm = off;     off += fpw * pcl_fast_frame_bytes(LP);
        Y.off_newpm = off;
    } else {
        Y.off_cm = off;     off += (2 * cols + 2 * fpw) * 8;   // per frame: 2 LP keys + 2 pad (bank spread)
        Y.off_newpm = off;  off += cols * 8;
    }
    // generic kernel keeps levels G+1 .. n-1 in shared memory, the fast one G+1 .. n-4
    // (the fast kernel never stores level 1, so its G is at least 1)
    int llr_vals = fast ? ((G >= Y.n - 4) ? 0 : cols * ((N >> G) - 16))
                        : ((G >= Y.n - 1) ? 0 : cols * ((N >> G) - 2));
    Y.off_llr = off;    off += align_up(llr_vals * rsz, 8);
    Y.off_sel = off;    off += cols * 4;
    Y.off_bw = off;     off += cols * (N >= 64 ? (N / 32 - 1) : 0) * 4;
    Y.off_uw = off;     off += Y.uw_slots * Y.NW * 4;
    Y.warp_bytes = align_up(off, 16);
    Y.scratch_per_warp = fast ? (int64_t)cols * ((N >> 1) - (N >> G))     // levels 2 .. G
                              : (int64_t)cols * (N - (N >> G));           // levels 1 .. G
}

// Kernel variants.  Generic: one frame per warp, every level in shared memory.  Fast: a lane
// owns a path, 32 / LP frames per warp (polar_scl_fast.cuh); X(LP, log2 N, G) with log2 N = 0
// for the run-time code length, else the code length and G compiled in as constants.
#ifdef PCL_QUICK   // experiment builds (scripts/build_variants.py): headline kernels only
#define PCL_POLAR_FAST_VARIANTS(X) X(8, 0, 0, 0) X(8, 10, 5, 0) X(8, 10, 3, 1)
#else
#define PCL_POLAR_FAST_VARIANTS(X) \
    X(1, 0, 0, 0) X(2, 0, 0, 0) X(4, 0, 0, 0) X(8, 0, 0, 0) X(16, 0, 0, 0) X(32, 0, 0, 0) \
    X(8, 7, 1, 0) X(8, 8, 2, 0) X(8, 9, 4, 0) X(8, 10, 5, 0) X(8, 11, 7, 0) X(8, 12, 8, 0) X(32, 10, 5, 0) X(1, 8, 2, 0) \
    X(8, 10, 3, 1) X(16, 10, 3, 1) X(32, 10, 3, 1) X(8, 11, 4, 1) X(8, 12, 5, 1)
#endif
// (SCL-8 gains 1.2-1.9 x from a compiled code length at every N = 128 .. 4096.  SC, LP = 1, does
// not in general -- N = 1024: 29 Gbps compiled vs 40 at run time; with no prune its time is all
// level walk and the unrolled walk is bigger code -- so only the N = 256 quick-start size of
// BASELINE configs[0] keeps one: 47 vs 41 Gbps at the bench's batch of 524 288 frames.)

static bool polar_fast_variant_exists(int LP, int nl, int gl, int tm = 0)
{
#define X(lp, n_, g_, tm_) if (LP == lp && nl == n_ && gl == g_ && tm == tm_) return true;
    PCL_POLAR_FAST_VARIANTS(X)
#undef X
    return false;
}

template <typename real, typename Fn>
static int polar_with_kernel(pcl_polar* h, Fn&& fn)
{
#ifndef PCL_QUICK
    if (h->wide) return fn(polar_scl_wide_kernel<real>);
#endif
    if (h->fast) {
#define X(lp, n_, g_, tm_)                                                                     \
    if constexpr (n_ == 0 || sizeof(real) == 4) {                                                   \
        if (h->LP == lp && h->NL == n_ && h->GL == g_ && h->TM == tm_)                              \
            return fn(polar_scl_fast_kernel<lp, real, n_, g_, tm_>);                                \
    }
        PCL_POLAR_FAST_VARIANTS(X)
#undef X
    } else {
        switch (h->LP) {
#ifndef PCL_QUICK
            case 1: return fn(polar_scl_kernel<1, real>);
            case 2: return fn(polar_scl_kernel<2, real>);
            case 4: return fn(polar_scl_kernel<4, real>);
            case 16: return fn(polar_scl_kernel<16, real>);
            case 32: return fn(polar_scl_kernel<32, real>);
#endif
            case 8: return fn(polar_scl_kernel<8, real>);
        }
    }
    return fail(PCL_EUNSUPPORTED, "no kernel for list size %d", h->L);
}

template <typename Fn>
static int polar_with_kernel_dtype(pcl_polar* h, Fn&& fn)
{
    return h->dtype == PCL_F64 ? polar_with_kernel<double>(h, fn) : polar_with_kernel<float>(h, fn);
}

template <typename real>
static int polar_launch(pcl_polar* h, const PolarParams<real>& P, int grid, void* stream)
{
    (void)stream;
    return polar_with_kernel<real>(h, [&](auto kern) -> int {
        PCL_LAUNCH(kern, grid, h->wpb * 32, h->smem_bytes, stream, P);
        return PCL_OK;
    });
}

#ifndef PCL_EMU
template <typename real>
static int polar_occ(pcl_polar* h, int threads, int smem, int* bps)
{
    return polar_with_kernel<real>(h, [&](auto kern) -> int {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        // the occupancy query honours the function's CURRENT carve-out preference: ask with the largest
        // shared-memory split, not with whatever an earlier handle of the same kernel left behind (seen: 4
        // instead of 6 resident blocks for a handle created after one with a bigger block)
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, kern, threads, smem));
        // Ask for the shared-memory carve-out the computed residency needs.  Left to its default the
        // driver may keep the split of the kernel that ran before (seen after the polar kernel: half
        // the LDPC blocks per SM and half the throughput in some runs of the same binary).
        DeviceInfo di2;
        if (device_info(&di2) == PCL_OK && *bps > 0) {
            long need = (long)(*bps) * (smem + 1024);
            int pct = (int)((need * 100 + di2.smem_per_sm - 1) / di2.smem_per_sm);
            if (pct > 100) pct = 100;
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        }
        return PCL_OK;
    });
}
#endif

static int polar_create_impl(pcl_polar_t** out, int N, int K, int list_size, const uint8_t* frozen_mask,
                             int crc_len, uint32_t crc_poly, int dtype, bool compiled_ok);

extern "C" int pcl_polar_create(pcl_polar_t** out, int N, int K, int list_size, const uint8_t* frozen_mask,
                                int crc_len, uint32_t crc_poly, int dtype)
{
    return polar_create_impl(out, N, K, list_size, frozen_mask, crc_len, crc_poly, dtype, true);
}

static int polar_create_impl(pcl_polar_t** out, int N, int K, int list_size, const uint8_t* frozen_mask,
                             int crc_len, uint32_t crc_poly, int dtype, bool compiled_ok)
{
    if (!out || !frozen_mask) return fail(PCL_EINVAL, "null argument");
    // src/polar/decoder.py:17-18, :194-196
    if (N <= 0 || (N & (N - 1)) != 0) return fail(PCL_EINVAL, "N must be a power of 2");
    if (!(K > 0 && K < N)) return fail(PCL_EINVAL, "K must be in (0, N)");
    if (list_size < 1) return fail(PCL_EINVAL, "list_size must be >= 1");
    if (list_size > 1024) return fail(PCL_EUNSUPPORTED, "list_size %d > 1024 is not supported (one thread per list slot, one block per frame)", list_size);
    if (N > 65536) return fail(PCL_EUNSUPPORTED, "N %d > 65536 is not supported (16-bit leaf positions)", N);
    if (dtype != PCL_F32 && dtype != PCL_F64) return fail(PCL_EINVAL, "bad dtype");
    if (crc_len < 0 || crc_len > 32) return fail(PCL_EINVAL, "bad crc_len");
    int n = ilog2i(N);
    int nfree = 0;
    for (int l = 0; l < N; l++) nfree += frozen_mask[l] == 0;
    if (nfree != K) return fail(PCL_EINVAL, "frozen mask leaves %d info positions, K=%d", nfree, K);

    pcl_polar* h = new pcl_polar();
    h->frozen_copy.assign(frozen_mask, frozen_mask + N);
    h->N = N; h->n = n; h->K = K; h->L = list_size; h->dtype = dtype;
    h->crc_len = crc_len; h->crc_poly = crc_poly;
    int LP = 1;
    while (LP < list_size) LP <<= 1;
    h->LP = LP;

    // decode-step-order tables: step i <-> reference index bit_reverse(i)
    const int shift = N < 32 ? 32 - N : 0;
    const int NW = N >= 32 ? N / 32 : 1;
    std::vector<uint32_t> fw(NW, 0);
    for (int i = 0; i < N; i++)
        if (frozen_mask[bit_reverse_i(i, n)]) fw[(i + shift) >> 5] |= 1u << ((i + shift) & 31);
    std::vector<uint16_t> ip;
    for (int l = 0; l < N; l++)
        if (!frozen_mask[l]) ip.push_back((uint16_t)bit_reverse_i(l, n));

    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) { delete h; return rc; }
    const int rsz = (int)real_size(dtype);
    int pb = 0;
    while ((1 << pb) < LP) pb++;
    // Kernel choice.  (1) fast kernel (tree bottom in registers): needs N >= 16 and the packed 32-bit slot
    // pointers to hold (n-4) LLR-level fields and (n-5) left-level fields of log2(LP) bits; (2) its TM
    // variant (one block per SM, mid levels in tensor / shared memory) where one is compiled; (3) the
    // generic all-shared-memory kernel.  A configuration whose shared memory does not fit a block retries
    // with fewer warps per block (4, 2, 1) and then falls back to the next kernel in the list.
    const bool can_fast = n >= 4 && (n - 4) * pb <= 32 && (n > 5 ? n - 5 : 0) * pb <= 32 && env_int("PCL_POLAR_GENERIC", 0) == 0;
    const bool nl_ok = compiled_ok && env_int("PCL_POLAR_NL", 1) != 0;
    const bool can_tm = can_fast && dtype == PCL_F32 && nl_ok && env_int("PCL_POLAR_TM", 1) != 0 &&
                        n >= 9 && polar_fast_variant_exists(LP, n, n - 7, 1) && (LP == 8 || env_int("PCL_POLAR_TM32", 1) != 0)   /* PCL_POLAR_TM32=0: lists of 16 / 32 back on the round-1 layout */;
    int bps = 1;
    bool placed = false;
    if (LP > 32) {
        // list wider than a warp: one block of LP threads per frame, every level in the global scratch
        h->wide = 1; h->fast = 0; h->TM = 0; h->fpw = 1; h->NL = 0; h->GL = 0;
        h->wpb = LP / 32;
        PolarLayout& Y = h->lay;
        Y = PolarLayout();
        Y.N = N; Y.n = n; Y.K = K; Y.L = list_size; Y.G = n - 1;
        Y.NW = N >= 32 ? N / 32 : 1;
        Y.nb = n > 5 ? n - 5 : 0;
        Y.uw_slots = LP;
        h->smem_bytes = pcl_wide_smem_bytes(LP, n, Y.nb);
        Y.warp_bytes = h->smem_bytes / h->wpb;
        Y.scratch_per_warp = pcl_wide_scratch_bytes(LP, N, rsz) / rsz / h->wpb;
        if (h->smem_bytes <= di.smem_per_block) {
#ifndef PCL_EMU
            rc = (dtype == PCL_F64) ? polar_occ<double>(h, h->wpb * 32, h->smem_bytes, &bps)
                                    : polar_occ<float>(h, h->wpb * 32, h->smem_bytes, &bps);
            if (rc) { delete h; return rc; }
#endif
            placed = bps >= 1;
        }
    }
    for (int attempt = can_tm ? 0 : (can_fast ? 1 : 2); attempt < 3 && !placed && !h->wide; attempt++) {
        h->TM = attempt == 0;
        h->fast = attempt <= 1;
        if (attempt == 1 && !can_fast) continue;
        // a lane owns a whole path, 32 / LP frames share a warp's instruction stream (measured best
        // on B200 for L = 8: 3.7 Gbps vs 2.9 / 2.0 with 2 / 4 lanes per path, profiles/r01d, r01e)
        h->fpw = h->fast ? 32 / LP : 1;
        h->NL = 0; h->GL = 0;
        if (h->TM) {
            polar_layout(h->lay, N, K, list_size, LP, n - 7, rsz, crc_len > 0, true, h->fpw, true);
            h->NL = n; h->GL = n - 7;
#ifdef PCL_EMU
            int w = env_int("PCL_POLAR_WPB", 8);       // two groups are enough to exercise the tickets
#else
            int w = env_int("PCL_POLAR_WPB", PCL_POLAR_TM_THREADS / 32);
#endif
            w = std::max(PCL_POLAR_TM_GROUP, std::min(w, PCL_POLAR_TM_THREADS / 32)) / PCL_POLAR_TM_GROUP * PCL_POLAR_TM_GROUP;
            while (w > PCL_POLAR_TM_GROUP && h->lay.hdr_bytes + h->lay.warp_bytes * w > di.smem_per_block) w -= PCL_POLAR_TM_GROUP;
            h->wpb = w;
            h->smem_bytes = h->lay.hdr_bytes + h->lay.warp_bytes * w;
            if (h->smem_bytes > di.smem_per_block) continue;
        } else {
            const int gmax = h->fast ? n - 4 : n - 1;
            const int budget = env_int("PCL_POLAR_SMEM_PER_WARP", h->fast ? 8192 : 9216);
            int G = env_int("PCL_POLAR_G", -1);
            const int gmin = h->fast ? 1 : 0;
            if (G < 0) {
                for (G = gmin; G < gmax; G++) {
                    polar_layout(h->lay, N, K, list_size, LP, G, rsz, crc_len > 0, h->fast, h->fpw);
                    if (h->lay.warp_bytes <= budget) break;
                }
            }
            G = std::max(gmin, std::min(G, gmax));
            polar_layout(h->lay, N, K, list_size, LP, G, rsz, crc_len > 0, h->fast, h->fpw);
            // code lengths with log2 N and G compiled in (fp32 build; the validation build reads them at run time)
            if (h->fast && dtype == PCL_F32 && nl_ok && polar_fast_variant_exists(LP, n, G, 0)) { h->NL = n; h->GL = G; }
            int w = env_int("PCL_POLAR_WPB", 4);
            if (w < 1 || w > 4) w = 4;
            while (w > 1 && h->lay.warp_bytes * w > di.smem_per_block) w >>= 1;
            h->wpb = w;
            h->smem_bytes = h->lay.warp_bytes * w;
            if (h->smem_bytes > di.smem_per_block) continue;
        }
        bps = 1;
#ifndef PCL_EMU
        rc = (dtype == PCL_F64) ? polar_occ<double>(h, h->wpb * 32, h->smem_bytes, &bps)
                                : polar_occ<float>(h, h->wpb * 32, h->smem_bytes, &bps);
        if (rc) { delete h; return rc; }
        if (bps < 1) continue;
#endif
        placed = true;
    }
    if (!placed) {
        const int need = h->smem_bytes;
        delete h;
        return fail(PCL_EUNSUPPORTED, "no kernel configuration fits: shared memory %d B per block against the device limit %d B",
                    need, di.smem_per_block);
    }
    {
        int cap = env_int("PCL_POLAR_BPS", 0);           // experiment knob: resident blocks per SM
        // Small lists outside the tensor-memory variant keep most tree levels in the global scratch and re-read
        // the lower ones constantly: leave >= 60 KB of the SM's 228 KB to L1 instead of filling it with blocks
        // (scripts/sweep_resident_blocks.py: SC N=1024 28.5 -> 41.9 Gbps at 5 instead of 6 blocks, N=2048
        // 23.7 -> 30.2, SCL-4 13.5 -> 16.1, SCL-8 N=2048 6.0 -> 7.05; lists of 16 and 32 prefer the blocks).
        if (cap <= 0 && h->fast && !h->TM && !h->wide && LP <= 8) {
            const int l1kb = env_int("PCL_POLAR_L1KB", 60);
            if (l1kb > 0) cap = std::max(1, (di.smem_per_sm - l1kb * 1024) / (h->smem_bytes + 1024));
        }
        if (cap > 0 && cap < bps) {
            bps = cap;
#ifndef PCL_EMU
            // fewer resident blocks: give the shared memory they do not use back to L1 (the level scratch is
            // read through it)
            const int smem_need = h->smem_bytes;
            rc = polar_with_kernel_dtype(h, [&](auto kern) -> int {
                long need = (long)bps * (smem_need + 1024);
                int pct = (int)((need * 100 + di.smem_per_sm - 1) / di.smem_per_sm);
                CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct));
                return PCL_OK;
            });
            if (rc) { delete h; return rc; }
#endif
        }
    }
    h->grid_max = di.sms * bps;
    h->scratch_bytes = (size_t)h->grid_max * h->wpb * h->lay.scratch_per_warp * rsz;
    if (h->wide) {
        const int64_t per_block = pcl_wide_scratch_bytes(LP, N, rsz);
        const int64_t cap = std::max<int64_t>(1, ((int64_t)4 << 30) / per_block);       // at most 4 GiB of scratch
        h->grid_max = (int)std::min<int64_t>(h->grid_max, cap);
        h->scratch_bytes = (size_t)h->grid_max * (size_t)per_block;
    }

    if (cudaMalloc((void**)&h->d_frozen_words, NW * 4) != cudaSuccess ||
        cudaMalloc((void**)&h->d_info_pos, (size_t)K * 2) != cudaSuccess) {
        pcl_polar_destroy(h);
        return fail(PCL_ECUDA, "cudaMalloc failed (tables)");
    }
    if (cudaMemcpy(h->d_frozen_words, fw.data(), NW * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->d_info_pos, ip.data(), (size_t)K * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
        pcl_polar_destroy(h);
        return fail(PCL_ECUDA, "cudaMemcpy failed (tables)");
    }
    // list size 1, fp32, bits only: the register-resident SC kernels (polar_sc.cuh) for N = 256 and N = 512 / 1024 / 2048
    // (PCL_POLAR_SC1024=0: the bigger ones back on the list kernel with L = 1)
    h->sc256 = (list_size == 1 && dtype == PCL_F32 && crc_len == 0 &&
                ((N == 256 && env_int("PCL_POLAR_SC256", 1) != 0) ||
                 ((N == 512 || N == 1024 || N == 2048 || N == 4096) && env_int("PCL_POLAR_SC1024", 1) != 0))) ? 1 : 0;
    if (h->TM || h->sc256) {
        if (cudaMalloc((void**)&h->d_next[0], 8) != cudaSuccess || cudaMemset(h->d_next[0], 0, 8) != cudaSuccess) {
            pcl_polar_destroy(h);
            return fail(PCL_ECUDA, "cudaMalloc failed (ticket counter)");
        }
    }
    if (h->scratch_bytes) {
        if (cudaMalloc(&h->d_scratch[0], h->scratch_bytes) != cudaSuccess) {
            pcl_polar_destroy(h);
            return fail(PCL_ECUDA, "cudaMalloc failed (scratch %zu B)", h->scratch_bytes);
        }
    }
    *out = h;
    return PCL_OK;
}

extern "C" void pcl_polar_destroy(pcl_polar_t* h)
{
    if (!h) return;
    if (h->twin) pcl_polar_destroy(h->twin);
    cudaFree(h->d_frozen_words);
    cudaFree(h->d_info_pos);
    for (int s = 0; s < PCL_NSTAGE; s++) {
        cudaFree(h->d_scratch[s]);
        cudaFree(h->d_next[s]);
        cudaFree(h->d_uwords[s]);
    }
#ifndef PCL_EMU
    h->pipe.destroy();
#endif
    delete h;
}

extern "C" int pcl_polar_lp(const pcl_polar_t* h) { return h ? h->LP : 0; }

extern "C" int pcl_polar_launch_info(const pcl_polar_t* h, int* grid, int* block, int* smem_bytes, int* glevels,
                                     int* fast)
{
    if (!h) return fail(PCL_EINVAL, "null handle");
    if (fast) *fast = h->wide ? 5 : (h->sc256 ? 4 : (h->fast ? (h->TM ? 3 : (h->NL ? 2 : 1)) : 0));
    if (grid) *grid = h->last_grid;
    if (block) *block = h->wpb * 32;
    if (smem_bytes) *smem_bytes = h->smem_bytes;
    if (glevels) *glevels = h->lay.G;
    return PCL_OK;
}

// SC, N = 256, fp32, nothing but the bits asked for: register-resident kernel + bit gather
static int polar_sc256_decode(pcl_polar* h, const void* llr_dev, int64_t F, uint8_t* bits_dev, void* stream, int stage)
{
    if (h->uwords_cap[stage] < F) {
        cudaFree(h->d_uwords[stage]);
        h->d_uwords[stage] = nullptr;
        h->uwords_cap[stage] = 0;
        const int64_t cap = std::max<int64_t>(F, 4096);
        CUDA_TRY(cudaMalloc((void**)&h->d_uwords[stage], (size_t)cap * (h->N / 32) * 4));
        h->uwords_cap[stage] = cap;
    }
    if (!h->d_next[stage]) {
        CUDA_TRY(cudaMalloc((void**)&h->d_next[stage], 8));
        CUDA_TRY(cudaMemset(h->d_next[stage], 0, 8));
        h->tickets[stage] = 0;
    }
    PolarScParams P;
    P.llr = (const float*)llr_dev;
    P.uwords = h->d_uwords[stage];
    P.frozen_words = h->d_frozen_words;
    P.next = h->d_next[stage];
    P.ticket_base = h->tickets[stage];
    P.F = F;
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const int M = h->N / 256;                                    // M length-256 codes in a row (polar_sc_big_kernel<M>)
    const bool big = M > 1;
    const int NWu = h->N / 32;
    // 32 padded frame rows per warp (+ the parked partial sums of the bigger codes: 8 (M - 1) words per frame);
    // warps per block so that the SM's shared memory is used up: 2 x 3 (N = 256, 1024), 6 (N = 512), 5 (N = 2048)
    const int wpb = (M == 2) ? 6 : (M == 8) ? 5 : (M == 16) ? 4 : PCL_SC256_WPB;
    const int smem = wpb * 32 * (PCL_SC256_ROW + (big ? 8 * (M - 1) : 0)) * 4;
    const int bps = std::max(1, di.smem_per_sm / (smem + 1024));
    const int64_t warps_needed = (F + 31) / 32;
    const int grid = (int)std::min<int64_t>((warps_needed + wpb - 1) / wpb, (int64_t)di.sms * bps);
    h->last_grid = grid;
#ifndef PCL_EMU
    static bool attr_set[17] = {};
    if (!attr_set[M]) {
        const void* kern = M == 2 ? (const void*)polar_sc_big_kernel<2> : M == 4 ? (const void*)polar_sc_big_kernel<4>
                         : M == 8 ? (const void*)polar_sc_big_kernel<8> : M == 16 ? (const void*)polar_sc_big_kernel<16>
                         : (const void*)polar_sc256_kernel;
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set[M] = true;
    }
#endif
    if (M == 2) { PCL_LAUNCH(polar_sc_big_kernel<2>, grid, wpb * 32, smem, stream, P); }
    else if (M == 4) { PCL_LAUNCH(polar_sc_big_kernel<4>, grid, wpb * 32, smem, stream, P); }
    else if (M == 8) { PCL_LAUNCH(polar_sc_big_kernel<8>, grid, wpb * 32, smem, stream, P); }
    else if (M == 16) { PCL_LAUNCH(polar_sc_big_kernel<16>, grid, wpb * 32, smem, stream, P); }
    else { PCL_LAUNCH(polar_sc256_kernel, grid, wpb * 32, smem, stream, P); }
    CUDA_TRY(cudaGetLastError());
    h->tickets[stage] += (unsigned long long)warps_needed + (unsigned long long)grid * wpb;
#ifdef PCL_EMU
    struct Ex { const uint32_t* u; const uint16_t* ip; int64_t F; int NW, K; uint8_t* b; } ex{h->d_uwords[stage], h->d_info_pos, F, NWu, h->K, bits_dev};
    auto run = [](Ex e) { polar_sc_extract_kernel(e.u, e.ip, e.F, e.NW, e.K, e.b); };
    PCL_LAUNCH(run, 2, 256, 0, stream, ex);
#else
    const int egrid = (int)std::min<int64_t>((F * h->K + 255) / 256, (int64_t)di.sms * 8);
    polar_sc_extract_kernel<<<egrid, 256, 0, (cudaStream_t)stream>>>(h->d_uwords[stage], h->d_info_pos, F, NWu, h->K, bits_dev);
    CUDA_TRY(cudaGetLastError());
#endif
    return PCL_OK;
}

template <typename real>
static int polar_decode_impl(pcl_polar* h, const void* llr_dev, int64_t F, uint8_t* bits_dev, double* pm_dev,
                             void* leaf_dev, uint8_t* parent_dev, void* scratch, void* stream, int stage = 0)
{
    if (h->sc256 && pm_dev == nullptr && leaf_dev == nullptr && sizeof(real) == 4)
        return polar_sc256_decode(h, llr_dev, F, bits_dev, stream, stage);
    if (leaf_dev != nullptr && h->fast && h->NL != 0) {
        // per-leaf dumps are compiled into the run-time-N kernels only (same results bit for bit)
        if (!h->twin) {
            int rc = polar_create_impl(&h->twin, h->N, h->K, h->L, h->frozen_copy.data(), h->crc_len, h->crc_poly, h->dtype, false);
            if (rc) return rc;
        }
        return polar_decode_impl<real>(h->twin, llr_dev, F, bits_dev, pm_dev, leaf_dev, parent_dev, h->twin->d_scratch[0], stream, 0);
    }
    PolarParams<real> P;
    P.lay = h->lay;
    P.llr = (const real*)llr_dev;
    P.bits = bits_dev;
    P.pm_out = pm_dev;
    P.dbg_leaf = (real*)leaf_dev;
    P.dbg_parent = parent_dev;
    P.frozen_words = h->d_frozen_words;
    P.info_pos = h->d_info_pos;
    P.scratch = (real*)scratch;
    P.F = F;
    P.want_pm = pm_dev != nullptr;
    P.use_crc = h->crc_len > 0;
    P.crc_len = h->crc_len;
    P.crc_poly = h->crc_poly;
    if (h->wide && leaf_dev != nullptr && h->LP > 256)
        return fail(PCL_EUNSUPPORTED, "per-leaf dumps carry the parent slot in one byte: list size %d > 256", h->L);
    const int64_t fpb = h->wide ? 1 : (int64_t)h->wpb * h->fpw;            // frames per block per pass
    int64_t need = (F + fpb - 1) / fpb;
    int grid = (int)std::min<int64_t>(need, h->grid_max);
    P.next = nullptr;
    P.ticket_base = 0;
    unsigned long long advance = 0;
    if (h->TM) {
        // groups of PCL_POLAR_TM_GROUP warps pull chunks of frames; a small batch spreads its chunks over the SMs
        const int64_t chunks = (F + PCL_POLAR_TM_GROUP * h->fpw - 1) / (PCL_POLAR_TM_GROUP * h->fpw);
        grid = (int)std::min<int64_t>(chunks, h->grid_max);
        P.next = h->d_next[stage];
        P.ticket_base = h->tickets[stage];
        // every group fetches until it draws a ticket past the end: chunks + one per group
        advance = (unsigned long long)chunks + (unsigned long long)grid * (h->wpb / PCL_POLAR_TM_GROUP);
    }
    h->last_grid = grid;
    int rc = polar_launch<real>(h, P, grid, stream);
    if (rc) return rc;
    CUDA_TRY(cudaGetLastError());
    h->tickets[stage] += advance;              // only a launch that went out draws tickets
    return PCL_OK;
}

extern "C" int pcl_polar_decode_batch(pcl_polar_t* h, const void* llr_dev, int64_t F, uint8_t* bits_dev,
                                      double* pm_dev, void* leaf_dev, uint8_t* parent_dev, void* stream)
{
    if (!h || F < 0) return fail(PCL_EINVAL, "bad handle or F");
    if (F == 0) return PCL_OK;
    if (!llr_dev || !bits_dev) return fail(PCL_EINVAL, "null buffer");
    if ((leaf_dev == nullptr) != (parent_dev == nullptr)) return fail(PCL_EINVAL, "leaf_dev and parent_dev go together");
    if (h->dtype == PCL_F64)
        return polar_decode_impl<double>(h, llr_dev, F, bits_dev, pm_dev, leaf_dev, parent_dev, h->d_scratch[0], stream);
    return polar_decode_impl<float>(h, llr_dev, F, bits_dev, pm_dev, leaf_dev, parent_dev, h->d_scratch[0], stream);
}

extern "C" int pcl_polar_decode_host_ex(pcl_polar_t* h, const void* llr_host, int llr_dtype, int64_t F, void* out_host,
                                        int out_format, void* stream)
{
    if (!h || F < 0) return fail(PCL_EINVAL, "bad handle or F");
    if (F == 0) return PCL_OK;
    if (!llr_host || !out_host) return fail(PCL_EINVAL, "null buffer");
#ifdef PCL_EMU
    (void)stream; (void)llr_dtype; (void)out_format;
    return fail(PCL_ECUDA, "no CUDA device");
#else
    for (int s = 0; s < PCL_NSTAGE; s++) {
        if (!h->d_scratch[s] && h->scratch_bytes) CUDA_TRY(cudaMalloc(&h->d_scratch[s], h->scratch_bytes));
        if (h->TM && !h->d_next[s]) {
            CUDA_TRY(cudaMalloc((void**)&h->d_next[s], 8));
            CUDA_TRY(cudaMemset(h->d_next[s], 0, 8));
            h->tickets[s] = 0;
        }
    }
    // a chunk fills the resident grid at least once and carries >= 32 MiB of LLRs
    const int64_t resident = h->wide ? (int64_t)h->grid_max : (int64_t)h->grid_max * h->wpb * h->fpw;
    int64_t chunk = std::max<int64_t>(resident, ((int64_t)32 << 20) / ((int64_t)h->N * 4));
    chunk = (chunk + 255) / 256 * 256;
    chunk = env_int("PCL_HOST_CHUNK", (int)std::min<int64_t>(chunk, 1 << 20));
    PipeLaunch launch = [h](int stage, cudaStream_t st, const void* d_llr, int64_t fc, uint8_t* d_bits, int32_t*) -> int {
        return (h->dtype == PCL_F64)
            ? polar_decode_impl<double>(h, d_llr, fc, d_bits, nullptr, nullptr, nullptr, h->d_scratch[stage], st, stage)
            : polar_decode_impl<float>(h, d_llr, fc, d_bits, nullptr, nullptr, nullptr, h->d_scratch[stage], st, stage);
    };
    return host_pipe_run(h->pipe, h->dtype, h->N, h->K, chunk, llr_host, llr_dtype, F, out_host, out_format, nullptr, stream,
                         launch, fail);
#endif
}

extern "C" int pcl_polar_decode_host(pcl_polar_t* h, const void* llr_host, int64_t F, uint8_t* bits_host, void* stream)
{
    if (!h) return fail(PCL_EINVAL, "bad handle or F");
    return pcl_polar_decode_host_ex(h, llr_host, h->dtype, F, bits_host, PCL_OUT_BYTES, stream);
}

// ================================================================ LDPC =========
struct pcl_ldpc {
    int m, n, E, mode, max_iter, early_stop, dtype, dmax;
    int regular6 = 0;           // every check degree 6, every variable degree 3
    double norm;
    LdpcLayout lay;
    int32_t* d_cptr = nullptr;
    uint16_t* d_col = nullptr;
    int32_t* d_vptr = nullptr;
    uint16_t* d_vperm = nullptr;
    unsigned long long* d_vpack = nullptr;
    unsigned long long* d_bpack = nullptr;      // banked layout tables (ldpc_layout.h)
    uint16_t* d_varof = nullptr;
    uint16_t* d_posof = nullptr;
    uint16_t* d_cpos = nullptr;
    int banked_residual = 0;
    unsigned long long* d_next[PCL_NSTAGE] = {};
    unsigned long long tickets[PCL_NSTAGE] = {};   // value of each counter (no memset between launches)
    int wpb, grid_max, smem_bytes, last_grid = 0;
#ifndef PCL_EMU
    HostPipe pipe;              // host-buffer pipeline (pcl_host_pipe.cuh)
#endif
};

template <typename real, int COOP, typename Fn>
static int ldpc_with_kernel_c(pcl_ldpc* h, Fn&& fn)
{
    if constexpr (sizeof(real) == 4) {
        if (h->lay.banked) {
            if (h->lay.paired) {
                if constexpr (COOP == 0) {
                    // 16 variable positions per lane (n = 481 .. 512) and at most 28 warps per block: channel values in
                    // registers (ldpc_banked.cuh, NPR; the block-per-frame mode measured no gain: 10.59 vs 10.58 Gbps at n = 2016)
                    if (h->lay.NP == 512 && h->wpb <= 28 && env_int("PCL_LDPC_REGLLR", 1) != 0) {
                        if (h->mode == PCL_LDPC_MS) return fn(ldpc_banked_kernel<1, 6, 0, 1, 16>);
                        return fn(ldpc_banked_kernel<0, 6, 0, 1, 16>);
                    }
                }
                if (h->mode == PCL_LDPC_MS) return fn(ldpc_banked_kernel<1, 6, COOP, 1>);
                return fn(ldpc_banked_kernel<0, 6, COOP, 1>);
            }
            if (h->mode == PCL_LDPC_MS) return fn(ldpc_banked_kernel<1, 6, COOP, 0>);
            return fn(ldpc_banked_kernel<0, 6, COOP, 0>);
        }
    }
    if (h->mode == PCL_LDPC_MS) {
        if (h->regular6) return fn(ldpc_decode_kernel<real, 1, 6, 1, COOP>);
        return fn(ldpc_decode_kernel<real, 1, 8, 0, COOP>);
    }
    if (h->regular6) return fn(ldpc_decode_kernel<real, 0, 6, 1, COOP>);
#ifdef PCL_QUICK
    return fail(PCL_EUNSUPPORTED, "experiment build");
#endif
    if (h->dmax <= 8) return fn(ldpc_decode_kernel<real, 0, 8, 0, COOP>);
    if (h->dmax <= 16) return fn(ldpc_decode_kernel<real, 0, 16, 0, COOP>);
    if (h->dmax <= 32) return fn(ldpc_decode_kernel<real, 0, 32, 0, COOP>);
    return fn(ldpc_decode_kernel<real, 0, 0, 0, COOP>);           // any check degree (cn_bp_loop)
}

template <typename real, typename Fn>
static int ldpc_with_kernel(pcl_ldpc* h, Fn&& fn)
{
    if (h->lay.coop) return ldpc_with_kernel_c<real, 1>(h, fn);
    return ldpc_with_kernel_c<real, 0>(h, fn);
}

template <typename real>
static int ldpc_launch(pcl_ldpc* h, const LdpcParams<real>& P, int grid, void* stream)
{
    (void)stream;
    return ldpc_with_kernel<real>(h, [&](auto kern) -> int {
        PCL_LAUNCH(kern, grid, h->wpb * 32, h->smem_bytes, stream, P);
        return PCL_OK;
    });
}

#ifndef PCL_EMU
template <typename real>
static int ldpc_occ(pcl_ldpc* h, int threads, int smem, int* bps)
{
    return ldpc_with_kernel<real>(h, [&](auto kern) -> int {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        // the occupancy query honours the function's CURRENT carve-out preference: ask with the largest
        // shared-memory split, not with whatever an earlier handle of the same kernel left behind (seen: 4
        // instead of 6 resident blocks for a handle created after one with a bigger block)
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, kern, threads, smem));
        // Ask for the shared-memory carve-out the computed residency needs.  Left to its default the
        // driver may keep the split of the kernel that ran before (seen after the polar kernel: half
        // the LDPC blocks per SM and half the throughput in some runs of the same binary).
        DeviceInfo di2;
        if (device_info(&di2) == PCL_OK && *bps > 0) {
            long need = (long)(*bps) * (smem + 1024);
            int pct = (int)((need * 100 + di2.smem_per_sm - 1) / di2.smem_per_sm);
            if (pct > 100) pct = 100;
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        }
        return PCL_OK;
    });
}
#endif

extern "C" int pcl_ldpc_create(pcl_ldpc_t** out, int m, int n, const uint8_t* H, int mode, double normalization,
                               int max_iter, int early_stop, int dtype)
{
    if (!out || !H) return fail(PCL_EINVAL, "null argument");
    if (m <= 0 || n <= 0) return fail(PCL_EINVAL, "bad H shape");
    if (mode != PCL_LDPC_BP && mode != PCL_LDPC_MS) return fail(PCL_EINVAL, "bad mode");
    if (dtype != PCL_F32 && dtype != PCL_F64) return fail(PCL_EINVAL, "bad dtype");
    if (max_iter < 1) return fail(PCL_EINVAL, "max_iter must be >= 1");
    if (n > 65535) return fail(PCL_EUNSUPPORTED, "n > 65535 not supported");

    // _build_tanner_graph (src/ldpc/decoder.py:35-60): row-major scan, H[i, j] == 1
    std::vector<int32_t> cptr(m + 1, 0), vptr(n + 1, 0), vdeg(n, 0);
    std::vector<uint16_t> col;
    for (int c = 0; c < m; c++) {
        for (int v = 0; v < n; v++)
            if (H[(size_t)c * n + v] == 1) { col.push_back((uint16_t)v); vdeg[v]++; }
        cptr[c + 1] = (int32_t)col.size();
    }
    const int E = (int)col.size();
    if (E > 65535) return fail(PCL_EUNSUPPORTED, "more than 65535 edges not supported");
    for (int v = 0; v < n; v++) vptr[v + 1] = vptr[v] + vdeg[v];
    std::vector<uint16_t> vperm(E ? E : 1);
    {
        std::vector<int32_t> fill(vptr.begin(), vptr.end() - 1);
        for (int c = 0; c < m; c++)
            for (int e = cptr[c]; e < cptr[c + 1]; e++) vperm[fill[col[e]]++] = (uint16_t)e;
    }
    int dmax = 0, vmax = 0;
    for (int c = 0; c < m; c++) {
        int d = cptr[c + 1] - cptr[c];
        dmax = std::max(dmax, d);
        if (mode == PCL_LDPC_MS && d == 1)
            return fail(PCL_EDEGREE1, "zero-size array to reduction operation minimum which has no identity");
    }
    for (int v = 0; v < n; v++) vmax = std::max(vmax, vdeg[v]);
    if (vmax > 128) return fail(PCL_EUNSUPPORTED, "variable degree %d > 128 not supported", vmax);

    pcl_ldpc* h = new pcl_ldpc();
    h->m = m; h->n = n; h->E = E; h->mode = mode; h->max_iter = max_iter; h->early_stop = early_stop ? 1 : 0;
    h->dtype = dtype; h->dmax = dmax; h->norm = normalization;
    {
        bool reg = env_int("PCL_LDPC_GENERIC", 0) == 0;
        for (int c = 0; c < m && reg; c++) reg = (cptr[c + 1] - cptr[c]) == 6;
        for (int v = 0; v < n && reg; v++) reg = vdeg[v] == 3;
        h->regular6 = reg ? 1 : 0;
    }
    const int rsz = (int)real_size(dtype);
    LdpcLayout& Y = h->lay;
    Y.m = m; Y.n = n; Y.E = E; Y.max_iter = max_iter; Y.early_stop = h->early_stop;
    Y.nhw = (n + 31) / 32;
    // fp32 production build, regular (3, 6) codes: conflict-free banked layout (ldpc_banked.cuh)
    BankedLayout bl;
    Y.banked = 0; Y.nR = 0; Y.NP = 0; Y.NS = 0; Y.paired = 0;
    if (h->regular6 && dtype == PCL_F32 && E > 0 && env_int("PCL_LDPC_BANKED", 1) != 0) {
        std::vector<int> var_checks(3 * (size_t)n), check_vars(6 * (size_t)m);
        for (int v = 0; v < n; v++)
            for (int j = 0; j < 3; j++) var_checks[3 * v + j] = vperm[vptr[v] + j] / 6;   // edge e belongs to check e / 6
        for (int e = 0; e < E; e++) check_vars[e] = col[e];
        const long budget = std::min<long>(16000000L, 2700L * E);
        const bool paired = env_int("PCL_LDPC_PAIRED", 1) != 0;
        const long iters = env_int("PCL_LDPC_ANNEAL", (int)budget);
        // the search is deterministic in (graph, options): keep its result for the next handle of the same code
        static std::mutex cache_mu;
        static std::map<unsigned long long, BankedLayout> cache;
        unsigned long long key = 1469598103934665603ull;
        auto mix = [&](unsigned long long x) { key = (key ^ x) * 1099511628211ull; };
        mix((unsigned long long)m); mix((unsigned long long)n); mix(paired ? 1 : 0); mix((unsigned long long)iters);
        for (int e = 0; e < E; e++) mix(col[e]);
        bool have = false;
        {
            std::lock_guard<std::mutex> lk(cache_mu);
            auto itc = cache.find(key);
            if (itc != cache.end()) { bl = itc->second; have = true; }
        }
        if (!have) {
            have = build_banked_layout(m, n, 6, var_checks, check_vars, iters, &bl, paired);
            if (have) {
                std::lock_guard<std::mutex> lk(cache_mu);
                if (cache.size() < 64) cache[key] = bl;
            }
        }
        if (have) {
            Y.banked = 1; Y.nR = bl.nR; Y.NP = bl.NP; Y.paired = paired ? 1 : 0;
            h->banked_residual = bl.residual;
            // Device form of the per-position table: BYTE offsets of the three message words (the variable
            // pass adds them to the frame's base, no shifts).  An empty position points at words of an EMPTY
            // check seat instead of carrying a flag the kernel would have to branch on: such words start at
            // zero, the check rule maps six zeros to six zeros (the paired BP rule: to lg2 of a ratio that is 1
            // up to rounding, ~1e-7), and a position whose channel value is zero writes their sums back, so they
            // stay at that level for the whole decode and no real variable ever reads them.  (A (3, 6) code with empty positions
            // always has an empty seat: n = 2 m.)  The word is taken from a bank the fetch leaves free when
            // there is one; otherwise that fetch costs one more wavefront.
            Y.NS = bl.NS;
            std::vector<int> zero_words;
            for (int sidx = 0; sidx < bl.NS; sidx++)
                if (bl.cpos[sidx] == 0xffff) zero_words.push_back(sidx);
            bool holes = false;
            for (int pi = 0; pi < bl.NP; pi++) holes = holes || (bl.bpack[pi] >> 48) != 0;
            if ((holes && zero_words.empty()) || 4 * Y.NS > 65535) have = false;
            for (int r = 0; r < bl.NP / 32 && have && holes; r++)
                for (int j = 0; j < 3; j++) {
                    bool used[32] = {};
                    for (int l = 0; l < 32; l++) {
                        const unsigned long long pk = bl.bpack[32 * r + l];
                        if ((pk >> 48) == 0) used[((pk >> (16 * j)) & 0xffffu) & 31] = true;
                    }
                    for (int l = 0; l < 32; l++) {
                        unsigned long long& pk = bl.bpack[32 * r + l];
                        if ((pk >> 48) == 0) continue;
                        int pick = zero_words[0];
                        for (int z : zero_words)
                            if (!used[z & 31]) { pick = z; break; }
                        used[pick & 31] = true;
                        pk = (pk & ~(0xffffull << (16 * j))) | ((unsigned long long)pick << (16 * j));
                    }
                }
            if (have) {
                for (auto& pk : bl.bpack) {
                    const unsigned long long a = pk & 0xffffu, b = (pk >> 16) & 0xffffu, c = (pk >> 32) & 0xffffu;
                    pk = (4 * a) | ((4 * b) << 16) | ((4 * c) << 32);
                }
            } else {
                Y.banked = 0; Y.nR = 0; Y.NP = 0; Y.NS = 0; Y.paired = 0;
            }
        }
    }
    int off = 0;
    Y.off_msg = off;  off += align_up((Y.banked ? Y.NS : std::max(E, 1)) * rsz, 8);
    Y.off_llr = off;  off += align_up((Y.banked ? Y.NP : n) * rsz, 8);
    Y.off_hard = off; off += align_up((Y.banked ? Y.NP / 32 : Y.nhw) * 4, 8);
    Y.off_ctl = off;  off += 16;
    Y.warp_bytes = align_up(off, 16);

    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) { delete h; return rc; }
    // warps per block: the split of the SM's shared memory that keeps most frames resident
    h->wpb = env_int("PCL_LDPC_WPB", 0);
    if (h->wpb < 1 || h->wpb > 8) {
        int best_w = 1, best_res = 0;
        for (int w = 1; w <= 4; w++) {
            if (Y.warp_bytes * w > di.smem_per_block) break;
            int res = (di.smem_per_sm / (Y.warp_bytes * w + 1024)) * w;
            if (res >= best_res) { best_res = res; best_w = w; }
        }
        h->wpb = best_w;
    }
    while (h->wpb > 1 && Y.warp_bytes * h->wpb > di.smem_per_block) h->wpb >>= 1;
    // Large codes: with one warp per frame fewer than ~16 warps fit on an SM; let a block of 4
    // warps decode one frame together instead (4 x the resident warps, a quarter of the latency).
    {
        const int resident = (di.smem_per_sm / (Y.warp_bytes * h->wpb + 1024)) * h->wpb;
        int coop = env_int("PCL_LDPC_COOP", -1);
        if (coop < 0) coop = resident < 16 ? 1 : 0;
        Y.coop = coop ? 1 : 0;
        if (Y.coop) {
            // warps per frame: 4, or 8 for the largest codes (Min-Sum n = 2016: 9.8 / 11.7 / 11.9 Gbps at 2 / 4 / 8)
            h->wpb = env_int("PCL_LDPC_COOP_WPB", n >= 1985 ? 8 : 4);
            if (h->wpb != 2 && h->wpb != 8) h->wpb = 4;
        }
    }
    if (!Y.coop && Y.banked && env_int("PCL_LDPC_FAT", 1) != 0 && env_int("PCL_LDPC_WPB", 0) < 1) {
        // warp-per-frame banked kernel: one fat block per SM when a single block can hold everything the
        // SM's shared memory has room for (deterministic residency, see ldpc_banked.cuh)
        const int fat = std::min(32, (std::min(di.smem_per_sm - 1024, di.smem_per_block)) / Y.warp_bytes);
        const int now = (di.smem_per_sm / (Y.warp_bytes * h->wpb + 1024)) * h->wpb;
        if (fat >= now && fat >= 1) h->wpb = fat;
    }
    h->smem_bytes = Y.coop ? Y.warp_bytes : Y.warp_bytes * h->wpb;
    if (h->smem_bytes > di.smem_per_block) {
        const int need = Y.warp_bytes;           // Y lives inside *h
        delete h;
        return fail(PCL_EUNSUPPORTED, "code too large: %d B of shared memory per frame", need);
    }
    int bps = 1;
#ifndef PCL_EMU
    rc = (dtype == PCL_F64) ? ldpc_occ<double>(h, h->wpb * 32, h->smem_bytes, &bps)
                            : ldpc_occ<float>(h, h->wpb * 32, h->smem_bytes, &bps);
    if (rc) { delete h; return rc; }
    if (bps < 1) { delete h; return fail(PCL_ECUDA, "kernel does not fit on an SM"); }
#endif
    h->grid_max = di.sms * bps;

    bool ok = cudaMalloc((void**)&h->d_cptr, (m + 1) * 4) == cudaSuccess &&
              cudaMalloc((void**)&h->d_col, std::max(E, 1) * 2) == cudaSuccess &&
              cudaMalloc((void**)&h->d_vptr, (n + 1) * 4) == cudaSuccess &&
              cudaMalloc((void**)&h->d_vperm, std::max(E, 1) * 2) == cudaSuccess &&
              cudaMalloc((void**)&h->d_next[0], 8) == cudaSuccess && cudaMemset(h->d_next[0], 0, 8) == cudaSuccess;
    ok = ok && cudaMemcpy(h->d_cptr, cptr.data(), (m + 1) * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(h->d_vptr, vptr.data(), (n + 1) * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok && E)
        ok = cudaMemcpy(h->d_col, col.data(), (size_t)E * 2, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(h->d_vperm, vperm.data(), (size_t)E * 2, cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok && h->regular6) {
        std::vector<unsigned long long> vpack(n);
        for (int v = 0; v < n; v++)
            vpack[v] = (unsigned long long)vperm[vptr[v]] | ((unsigned long long)vperm[vptr[v] + 1] << 16) |
                       ((unsigned long long)vperm[vptr[v] + 2] << 32);
        ok = cudaMalloc((void**)&h->d_vpack, (size_t)n * 8) == cudaSuccess &&
             cudaMemcpy(h->d_vpack, vpack.data(), (size_t)n * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (ok && Y.banked) {
        ok = cudaMalloc((void**)&h->d_bpack, (size_t)Y.NP * 8) == cudaSuccess &&
             cudaMalloc((void**)&h->d_varof, (size_t)Y.NP * 2) == cudaSuccess &&
             cudaMalloc((void**)&h->d_posof, (size_t)n * 2) == cudaSuccess &&
             cudaMalloc((void**)&h->d_cpos, (size_t)Y.NS * 2) == cudaSuccess &&
             cudaMemcpy(h->d_bpack, bl.bpack.data(), (size_t)Y.NP * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(h->d_varof, bl.varof.data(), (size_t)Y.NP * 2, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(h->d_posof, bl.posof.data(), (size_t)n * 2, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(h->d_cpos, bl.cpos.data(), (size_t)Y.NS * 2, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (!ok) { pcl_ldpc_destroy(h); return fail(PCL_ECUDA, "device table setup failed"); }
    *out = h;
    return PCL_OK;
}

extern "C" void pcl_ldpc_destroy(pcl_ldpc_t* h)
{
    if (!h) return;
    cudaFree(h->d_cptr); cudaFree(h->d_col); cudaFree(h->d_vptr); cudaFree(h->d_vperm); cudaFree(h->d_vpack);
    cudaFree(h->d_bpack); cudaFree(h->d_varof); cudaFree(h->d_posof); cudaFree(h->d_cpos);
    for (int s = 0; s < PCL_NSTAGE; s++) cudaFree(h->d_next[s]);
#ifndef PCL_EMU
    h->pipe.destroy();
#endif
    delete h;
}

extern "C" int pcl_ldpc_num_edges(const pcl_ldpc_t* h) { return h ? h->E : 0; }

extern "C" int pcl_ldpc_layout_info(const pcl_ldpc_t* h, int* banked, int* residual, int* coop)
{
    if (!h) return fail(PCL_EINVAL, "null handle");
    if (banked) *banked = h->lay.banked;
    if (residual) *residual = h->banked_residual;
    if (coop) *coop = h->lay.coop;
    return PCL_OK;
}

extern "C" int pcl_ldpc_launch_info(const pcl_ldpc_t* h, int* grid, int* block, int* smem_bytes)
{
    if (!h) return fail(PCL_EINVAL, "null handle");
    if (grid) *grid = h->last_grid;
    if (block) *block = h->wpb * 32;
    if (smem_bytes) *smem_bytes = h->smem_bytes;
    return PCL_OK;
}

template <typename real>
static int ldpc_decode_impl(pcl_ldpc* h, const void* llr_dev, int64_t F, uint8_t* bits_dev, int32_t* iters_dev,
                            void* total_dev, int stage, void* stream)
{
    unsigned long long* next = h->d_next[stage];
    LdpcParams<real> P;
    P.lay = h->lay;
    P.llr = (const real*)llr_dev;
    P.bits = bits_dev;
    P.iters = iters_dev;
    P.total = (real*)total_dev;
    P.cptr = h->d_cptr; P.col = h->d_col; P.vptr = h->d_vptr; P.vperm = h->d_vperm; P.vpack = h->d_vpack;
    P.bpack = h->d_bpack; P.varof = h->d_varof; P.posof = h->d_posof; P.cpos = h->d_cpos;
    P.next = next;
    P.F = F;
    P.norm = (real)h->norm;
    int64_t need = h->lay.coop ? F : (F + h->wpb - 1) / h->wpb;
    int grid = (int)std::min<int64_t>(need, h->grid_max);
    h->last_grid = grid;
    // every fetching unit (a warp, or a block in the block-per-frame mode) draws tickets until one is past
    // the end: F + units draws per launch, so the counter needs no reset (a memset between two decode
    // kernels is one more foreign launch that can leave the SMs in another shared-memory split)
    P.ticket_base = h->tickets[stage];
    const unsigned long long advance = (unsigned long long)F + (unsigned long long)grid * (h->lay.coop ? 1 : h->wpb);
    int rc = ldpc_launch<real>(h, P, grid, stream);
    if (rc) return rc;
    CUDA_TRY(cudaGetLastError());
    h->tickets[stage] += advance;              // only a launch that went out draws tickets
    return PCL_OK;
}

extern "C" int pcl_ldpc_decode_batch(pcl_ldpc_t* h, const void* llr_dev, int64_t F, uint8_t* bits_dev,
                                     int32_t* iters_dev, void* total_dev, void* stream)
{
    if (!h || F < 0) return fail(PCL_EINVAL, "bad handle or F");
    if (F == 0) return PCL_OK;
    if (!llr_dev || !bits_dev) return fail(PCL_EINVAL, "null buffer");
    if (h->dtype == PCL_F64)
        return ldpc_decode_impl<double>(h, llr_dev, F, bits_dev, iters_dev, total_dev, 0, stream);
    return ldpc_decode_impl<float>(h, llr_dev, F, bits_dev, iters_dev, total_dev, 0, stream);
}

extern "C" int pcl_ldpc_decode_host_ex(pcl_ldpc_t* h, const void* llr_host, int llr_dtype, int64_t F, void* out_host,
                                       int out_format, int32_t* iters_host, void* stream)
{
    if (!h || F < 0) return fail(PCL_EINVAL, "bad handle or F");
    if (F == 0) return PCL_OK;
    if (!llr_host || !out_host) return fail(PCL_EINVAL, "null buffer");
#ifdef PCL_EMU
    (void)stream; (void)iters_host; (void)llr_dtype; (void)out_format;
    return fail(PCL_ECUDA, "no CUDA device");
#else
    for (int s = 0; s < PCL_NSTAGE; s++)
        if (!h->d_next[s]) {
            CUDA_TRY(cudaMalloc((void**)&h->d_next[s], 8));
            CUDA_TRY(cudaMemset(h->d_next[s], 0, 8));
            h->tickets[s] = 0;
        }
    const int64_t resident = (int64_t)h->grid_max * (h->lay.coop ? 1 : h->wpb);
    int64_t chunk = std::max<int64_t>(4 * resident, ((int64_t)32 << 20) / ((int64_t)h->n * 4));
    chunk = (chunk + 255) / 256 * 256;
    chunk = env_int("PCL_HOST_CHUNK", (int)std::min<int64_t>(chunk, 1 << 20));
    PipeLaunch launch = [h](int stage, cudaStream_t st, const void* d_llr, int64_t fc, uint8_t* d_bits, int32_t* d_iters) -> int {
        return (h->dtype == PCL_F64)
            ? ldpc_decode_impl<double>(h, d_llr, fc, d_bits, d_iters, nullptr, stage, st)
            : ldpc_decode_impl<float>(h, d_llr, fc, d_bits, d_iters, nullptr, stage, st);
    };
    return host_pipe_run(h->pipe, h->dtype, h->n, h->n, chunk, llr_host, llr_dtype, F, out_host, out_format, iters_host, stream,
                         launch, fail);
#endif
}

extern "C" int pcl_ldpc_decode_host(pcl_ldpc_t* h, const void* llr_host, int64_t F, uint8_t* bits_host,
                                    int32_t* iters_host, void* stream)
{
    if (!h) return fail(PCL_EINVAL, "bad handle or F");
    return pcl_ldpc_decode_host_ex(h, llr_host, h->dtype, F, bits_host, PCL_OUT_BYTES, iters_host, stream);
}

// ============================================================ frame generator ===
struct pcl_gen {
    int kind = 0, N = 0, K = 0, NW = 0;
    uint32_t* d_info_words = nullptr;
    uint16_t* d_info_rank = nullptr;
    uint32_t* d_G = nullptr;
};

extern "C" void pcl_gen_destroy(pcl_gen_t* h)
{
    if (!h) return;
    cudaFree(h->d_info_words);
    cudaFree(h->d_info_rank);
    cudaFree(h->d_G);
    delete h;
}

extern "C" int pcl_gen_polar_create(pcl_gen_t** out, int N, int K, const uint8_t* frozen_mask)
{
    if (!out || !frozen_mask) return fail(PCL_EINVAL, "null argument");
    if (N <= 0 || (N & (N - 1)) != 0 || N > 65536) return fail(PCL_EINVAL, "N must be a power of 2 (<= 65536)");
    if (!(K > 0 && K < N)) return fail(PCL_EINVAL, "K must be in (0, N)");
    const int NW = (N + 31) / 32;
    std::vector<uint32_t> iw(NW, 0);
    std::vector<uint16_t> rank(NW, 0);
    int cnt = 0;
    for (int w = 0; w < NW; w++) {
        rank[w] = (uint16_t)cnt;
        for (int b = 0; b < 32 && 32 * w + b < N; b++)
            if (!frozen_mask[32 * w + b]) { iw[w] |= 1u << b; cnt++; }
    }
    if (cnt != K) return fail(PCL_EINVAL, "frozen mask leaves %d info positions, K=%d", cnt, K);
    pcl_gen* h = new pcl_gen();
    h->kind = 0; h->N = N; h->K = K; h->NW = NW;
    if (cudaMalloc((void**)&h->d_info_words, NW * 4) != cudaSuccess ||
        cudaMalloc((void**)&h->d_info_rank, NW * 2) != cudaSuccess ||
        cudaMemcpy(h->d_info_words, iw.data(), NW * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->d_info_rank, rank.data(), NW * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
        pcl_gen_destroy(h);
        return fail(PCL_ECUDA, "cudaMalloc / cudaMemcpy failed (generator tables)");
    }
    *out = h;
    return PCL_OK;
}

extern "C" int pcl_gen_ldpc_create(pcl_gen_t** out, int n, int k, const uint8_t* G_dense)
{
    if (!out || !G_dense) return fail(PCL_EINVAL, "null argument");
    if (n <= 0 || k <= 0 || k > n || n > 65536) return fail(PCL_EINVAL, "bad generator shape");
    const int NW = (n + 31) / 32;
    std::vector<uint32_t> g((size_t)k * NW, 0);
    for (int r = 0; r < k; r++)
        for (int c = 0; c < n; c++)
            if (G_dense[(size_t)r * n + c] & 1) g[(size_t)r * NW + (c >> 5)] |= 1u << (c & 31);
    pcl_gen* h = new pcl_gen();
    h->kind = 1; h->N = n; h->K = k; h->NW = NW;
    if (cudaMalloc((void**)&h->d_G, g.size() * 4) != cudaSuccess ||
        cudaMemcpy(h->d_G, g.data(), g.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        pcl_gen_destroy(h);
        return fail(PCL_ECUDA, "cudaMalloc / cudaMemcpy failed (generator matrix)");
    }
    *out = h;
    return PCL_OK;
}

extern "C" int pcl_gen_frames_channel(pcl_gen_t* h, int64_t F, int64_t frame0, unsigned long long seed, int channel,
                                      double param, int dtype, uint8_t* msg_dev, uint8_t* cw_dev, void* llr_dev,
                                      void* stream)
{
    if (!h || F < 0 || frame0 < 0) return fail(PCL_EINVAL, "bad handle, F or frame0");
    if (dtype != PCL_F32 && dtype != PCL_F64) return fail(PCL_EINVAL, "bad dtype");
    if (channel < PCL_CH_AWGN || channel > PCL_CH_BSC) return fail(PCL_EINVAL, "bad channel");
    if (channel == PCL_CH_BSC && !(param > 0.0 && param < 1.0))
        return fail(PCL_EINVAL, "BSC crossover probability must be in (0, 1) to give finite LLRs");
    if (F == 0) return PCL_OK;
    if (!llr_dev) return fail(PCL_EINVAL, "null buffer");
    GenParams P;
    P.kind = h->kind; P.N = h->N; P.K = h->K; P.NW = h->NW;
    P.info_words = h->d_info_words; P.info_rank = h->d_info_rank; P.G = h->d_G;
    P.F = F; P.frame0 = frame0;
    P.seed_lo = (uint32_t)seed; P.seed_hi = (uint32_t)(seed >> 32);
    P.channel = channel;
    // src/channel/awgn.py:27-32 (and fading.py:18-20): sigma = sqrt(1 / (2 snr_linear)); LLR = 2 y / sigma^2
    const double snr_lin = pow(10.0, (channel == PCL_CH_BSC ? 0.0 : param) / 10.0);
    P.sigma64 = sqrt(1.0 / (2.0 * snr_lin));
    P.scale64 = 2.0 / (P.sigma64 * P.sigma64);
    P.sigma = (float)P.sigma64; P.scale = (float)P.scale64;
    P.bsc_p = channel == PCL_CH_BSC ? (float)param : 0.0f;
    P.bsc_llr = channel == PCL_CH_BSC ? (float)log((1.0 - param) / param) : 0.0f;
    P.f64 = dtype == PCL_F64;
    P.msg = msg_dev; P.cw = cw_dev; P.llr = llr_dev;
    const int wpb = 4;
    const int grid = (int)std::min<int64_t>((F + wpb - 1) / wpb, 148 * 16);
    const int smem = wpb * h->NW * 4;
#ifndef PCL_EMU
    aux_carveout((const void*)framegen_kernel);
#endif
    PCL_LAUNCH(framegen_kernel, grid, wpb * 32, smem, stream, P);
    CUDA_TRY(cudaGetLastError());
    return PCL_OK;
}

extern "C" int pcl_gen_frames(pcl_gen_t* h, int64_t F, int64_t frame0, unsigned long long seed, double snr_db,
                              int dtype, uint8_t* msg_dev, uint8_t* cw_dev, void* llr_dev, void* stream)
{
    return pcl_gen_frames_channel(h, F, frame0, seed, PCL_CH_AWGN, snr_db, dtype, msg_dev, cw_dev, llr_dev, stream);
}

// Philox4x32-10 block for known-answer tests (host evaluation of the device function's source)
extern "C" void pcl_philox4x32_10_host(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    const pcl_philox4 r = pcl_philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// ============================================================ error counters ===
__global__ void __launch_bounds__(256) count_errors_kernel(const uint8_t* bits, const uint8_t* ref, int64_t F,
                                                           int width, int ncmp, unsigned long long* counters)
{
    // one warp per frame row; warp-aggregated, then one atomic per warp per counter
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long be = 0, fe = 0, fr = 0;
    for (int64_t f = wid; f < F; f += nw) {
        const uint8_t* a = bits + f * width;
        const uint8_t* b = ref + f * width;
        unsigned e = 0;
        for (int k = lane; k < ncmp; k += 32) e += (a[k] != b[k]);
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(PCL_FULL_MASK, e, o);
        be += e;
        fe += (e != 0);
        fr += 1;
    }
    if (lane == 0 && fr) {
        atomicAdd(&counters[0], be);
        atomicAdd(&counters[1], fe);
        atomicAdd(&counters[2], fr);
        atomicAdd(&counters[3], fr * (unsigned long long)ncmp);
    }
}

struct CountArgs { const uint8_t* bits; const uint8_t* ref; int64_t F; int width; int ncmp; unsigned long long* c; };
#ifdef PCL_EMU
static void count_errors_emu(CountArgs a) { count_errors_kernel(a.bits, a.ref, a.F, a.width, a.ncmp, a.c); }
#endif

extern "C" int pcl_count_errors(const uint8_t* bits_dev, const uint8_t* ref_dev, int64_t F, int width, int ncmp,
                                unsigned long long* counters_dev, void* stream)
{
    if (F < 0 || width <= 0 || ncmp < 0 || ncmp > width) return fail(PCL_EINVAL, "bad shape");
    if (F == 0) return PCL_OK;
    if (!bits_dev || !ref_dev || !counters_dev) return fail(PCL_EINVAL, "null buffer");
    int grid = (int)std::min<int64_t>((F + 7) / 8, 148 * 8);
#ifdef PCL_EMU
    (void)stream;
    CountArgs a{bits_dev, ref_dev, F, width, ncmp, counters_dev};
    PCL_LAUNCH(count_errors_emu, grid, 256, 0, stream, a);
#else
    aux_carveout((const void*)count_errors_kernel);
    count_errors_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(bits_dev, ref_dev, F, width, ncmp, counters_dev);
    CUDA_TRY(cudaGetLastError());
#endif
    return PCL_OK;
}

#ifdef PCL_EMU
extern "C" void simt_set_reverse(int r) { simt::reverse_order() = r != 0; }
#endif
