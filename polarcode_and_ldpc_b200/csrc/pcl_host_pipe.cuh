// pcl_host_pipe.cuh -- the host-buffer path of the C ABI (pcl_*_decode_host_ex): a batch that lives
// in HOST memory is cut into chunks and pushed through a ring of stages so that, at any moment,
//   host threads convert / unpack one chunk  |  the copy engines move two others  |  the SMs decode a fourth.
//
//   input   PCL_F32 (pinned or pageable)  -> cudaMemcpyAsync straight from the caller's buffer
//           PCL_F64 with an fp32 handle    -> host threads narrow it into a pinned fp32 staging chunk
//                                            (this is the reference call shape: np.float64[F, N], pageable)
//           PCL_F16 (transport format)     -> copied as is (half the PCIe bytes), widened on the device
//   output  PCL_OUT_BYTES                  -> one byte per bit, copied straight into the caller's buffer
//           PCL_OUT_PACKED                 -> rows bit-packed on the device (1 / 8 of the D2H bytes)
//           PCL_OUT_INT64                  -> packed on the device, unpacked by host threads into the int64
//                                            array decode() callers expect
// Everything the decoders need from it is the `launch` callback (decode chunk on a stage's stream).
#pragma once
#include "pcl_common.cuh"
#include "../../include/pcl.h"

#ifndef PCL_EMU
#include <cuda_fp16.h>
#include <algorithm>
#include <functional>
#include <thread>
#include <vector>

#define PCL_PIPE_STAGES 3

// ---- device helpers -------------------------------------------------------------------------------
// rows of `width` bytes (0 / 1) -> rows of ceil(width / 32) words, bit k of a row at word k / 32, bit k % 32
__global__ void __launch_bounds__(256) pack_bits_kernel(const uint8_t* bits, int64_t F, int width, int words, uint32_t* out)
{
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t job = wid; job < F * words; job += nw) {
        const int64_t f = job / words;
        const int w = (int)(job - f * words);
        const int k = 32 * w + lane;
        const unsigned b = __ballot_sync(0xffffffffu, k < width && bits[f * width + k] != 0);
        if (lane == 0) out[job] = b;
    }
}

// fp16 LLRs (transport format) -> fp32, 8 values per thread and step
__global__ void __launch_bounds__(256) widen_f16_kernel(const __half* in, float* out, int64_t n8)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 raw = reinterpret_cast<const uint4*>(in)[i];
        const __half2* h = reinterpret_cast<const __half2*>(&raw);
        float4 a, b;
        float2 t;
        t = __half22float2(h[0]); a.x = t.x; a.y = t.y;
        t = __half22float2(h[1]); a.z = t.x; a.w = t.y;
        t = __half22float2(h[2]); b.x = t.x; b.y = t.y;
        t = __half22float2(h[3]); b.z = t.x; b.w = t.y;
        reinterpret_cast<float4*>(out)[2 * i] = a;
        reinterpret_cast<float4*>(out)[2 * i + 1] = b;
    }
}

// ---- host helpers ---------------------------------------------------------------------------------
static int pipe_threads()
{
    const char* s = getenv("PCL_HOST_THREADS");
    int n = s ? atoi(s) : (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
    return std::max(1, std::min(n, 64));
}

template <typename Fn>
static void parallel_rows(int64_t rows, Fn&& fn)
{
    const int nt = (int)std::min<int64_t>(pipe_threads(), std::max<int64_t>(1, rows / 64));
    if (nt <= 1) { fn(0, rows); return; }
    std::vector<std::thread> th;
    const int64_t per = (rows + nt - 1) / nt;
    for (int t = 1; t < nt; t++) {
        const int64_t lo = std::min<int64_t>(rows, per * t), hi = std::min<int64_t>(rows, lo + per);
        if (lo < hi) th.emplace_back([=, &fn]() { fn(lo, hi); });
    }
    fn(0, std::min<int64_t>(rows, per));
    for (auto& t : th) t.join();
}

static void narrow_f64_rows(const double* src, float* dst, int64_t n)
{
    for (int64_t i = 0; i < n; i++) dst[i] = (float)src[i];
}

static void unpack_rows_i64(const uint32_t* packed, int words, int width, int64_t lo, int64_t hi, int64_t* out)
{
    for (int64_t f = lo; f < hi; f++) {
        const uint32_t* row = packed + f * words;
        int64_t* o = out + f * width;
        for (int k = 0; k < width; k++) o[k] = (row[k >> 5] >> (k & 31)) & 1u;
    }
}

struct HostPipe {
    int64_t chunk = 0;              // frames per chunk the buffers below are sized for
    int in_width = 0, out_width = 0, words = 0, rsz = 4;
    cudaStream_t st[PCL_PIPE_STAGES] = {};
    cudaEvent_t done[PCL_PIPE_STAGES] = {};
    void* d_llr[PCL_PIPE_STAGES] = {};
    void* d_raw[PCL_PIPE_STAGES] = {};      // fp16 transport chunk
    uint8_t* d_bits[PCL_PIPE_STAGES] = {};
    uint32_t* d_packed[PCL_PIPE_STAGES] = {};
    int32_t* d_iters[PCL_PIPE_STAGES] = {};
    float* h_in[PCL_PIPE_STAGES] = {};      // pinned: narrowed fp64 input
    uint32_t* h_packed[PCL_PIPE_STAGES] = {};   // pinned: packed output awaiting the host unpack
    bool want_iters = false;

    void release()
    {
        for (int s = 0; s < PCL_PIPE_STAGES; s++) {
            cudaFree(d_llr[s]); cudaFree(d_raw[s]); cudaFree(d_bits[s]); cudaFree(d_packed[s]); cudaFree(d_iters[s]);
            if (h_in[s]) cudaFreeHost(h_in[s]);
            if (h_packed[s]) cudaFreeHost(h_packed[s]);
            d_llr[s] = d_raw[s] = nullptr; d_bits[s] = nullptr; d_packed[s] = nullptr; d_iters[s] = nullptr;
            h_in[s] = nullptr; h_packed[s] = nullptr;
        }
        chunk = 0;
    }
    void destroy()
    {
        release();
        for (int s = 0; s < PCL_PIPE_STAGES; s++) {
            if (st[s]) cudaStreamDestroy(st[s]);
            if (done[s]) cudaEventDestroy(done[s]);
            st[s] = nullptr; done[s] = nullptr;
        }
    }
};

// launch(stage, stream, d_llr, frames, d_bits, d_iters) -> PCL_* status
typedef std::function<int(int, cudaStream_t, const void*, int64_t, uint8_t*, int32_t*)> PipeLaunch;

static int host_pipe_run(HostPipe& hp, int handle_dtype, int in_width, int out_width, int64_t want_chunk,
                         const void* llr_host, int llr_dtype, int64_t F, void* out_host, int out_format,
                         int32_t* iters_host, void* user_stream, const PipeLaunch& launch,
                         int (*fail_fn)(int, const char*, ...))
{
    const int rsz = handle_dtype == PCL_F64 ? 8 : 4;
    if (llr_dtype != PCL_F32 && llr_dtype != PCL_F64 && llr_dtype != PCL_F16) return fail_fn(PCL_EINVAL, "bad llr_dtype");
    if (out_format != PCL_OUT_BYTES && out_format != PCL_OUT_PACKED && out_format != PCL_OUT_INT64)
        return fail_fn(PCL_EINVAL, "bad out_format");
    if (handle_dtype == PCL_F64 && llr_dtype != PCL_F64)
        return fail_fn(PCL_EINVAL, "a float64 (validation) handle takes float64 LLRs");
    if (llr_dtype == PCL_F16 && (in_width % 8) != 0) return fail_fn(PCL_EINVAL, "float16 transport needs a row length divisible by 8");
    const int words = (out_width + 31) / 32;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(F, want_chunk));
    if (hp.chunk < chunk || hp.in_width != in_width || hp.out_width != out_width || hp.rsz != rsz) {
        hp.release();
        hp.in_width = in_width; hp.out_width = out_width; hp.words = words; hp.rsz = rsz;
        for (int s = 0; s < PCL_PIPE_STAGES; s++) {
            if (cudaMalloc(&hp.d_llr[s], (size_t)chunk * in_width * rsz) != cudaSuccess ||
                cudaMalloc((void**)&hp.d_bits[s], (size_t)chunk * out_width) != cudaSuccess ||
                cudaMalloc((void**)&hp.d_packed[s], (size_t)chunk * words * 4) != cudaSuccess ||
                cudaMalloc((void**)&hp.d_iters[s], (size_t)chunk * 4) != cudaSuccess)
                return fail_fn(PCL_ECUDA, "cudaMalloc failed (host pipeline, %lld frames per chunk)", (long long)chunk);
        }
        hp.chunk = chunk;
    }
    for (int s = 0; s < PCL_PIPE_STAGES; s++) {
        if (!hp.st[s] && cudaStreamCreateWithFlags(&hp.st[s], cudaStreamNonBlocking) != cudaSuccess)
            return fail_fn(PCL_ECUDA, "cudaStreamCreate failed");
        if (!hp.done[s] && cudaEventCreateWithFlags(&hp.done[s], cudaEventDisableTiming) != cudaSuccess)
            return fail_fn(PCL_ECUDA, "cudaEventCreate failed");
        if (llr_dtype == PCL_F16 && !hp.d_raw[s] && cudaMalloc(&hp.d_raw[s], (size_t)hp.chunk * in_width * 2) != cudaSuccess)
            return fail_fn(PCL_ECUDA, "cudaMalloc failed (fp16 chunk)");
        if (llr_dtype == PCL_F64 && rsz == 4 && !hp.h_in[s] &&
            cudaHostAlloc((void**)&hp.h_in[s], (size_t)hp.chunk * in_width * 4, cudaHostAllocDefault) != cudaSuccess)
            return fail_fn(PCL_ECUDA, "cudaHostAlloc failed (fp32 staging)");
        if (out_format == PCL_OUT_INT64 && !hp.h_packed[s] &&
            cudaHostAlloc((void**)&hp.h_packed[s], (size_t)hp.chunk * words * 4, cudaHostAllocDefault) != cudaSuccess)
            return fail_fn(PCL_ECUDA, "cudaHostAlloc failed (packed staging)");
    }
    if (cudaStreamSynchronize((cudaStream_t)user_stream) != cudaSuccess) return fail_fn(PCL_ECUDA, "stream sync failed");

    struct Pending { int64_t f0 = 0, fc = 0; bool live = false; };
    Pending pend[PCL_PIPE_STAGES];
    auto drain = [&](int s) -> int {                 // finish the host side of the chunk that used stage s
        if (!pend[s].live) return PCL_OK;
        if (cudaEventSynchronize(hp.done[s]) != cudaSuccess) return fail_fn(PCL_ECUDA, "decode chunk failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (out_format == PCL_OUT_INT64) {
            const uint32_t* pk = hp.h_packed[s];
            int64_t* out = (int64_t*)out_host + pend[s].f0 * out_width;
            parallel_rows(pend[s].fc, [&](int64_t lo, int64_t hi) { unpack_rows_i64(pk, words, out_width, lo, hi, out); });
        }
        pend[s].live = false;
        return PCL_OK;
    };
    int stage = 0;
    for (int64_t f0 = 0; f0 < F; f0 += chunk, stage = (stage + 1) % PCL_PIPE_STAGES) {
        const int64_t fc = std::min<int64_t>(chunk, F - f0);
        int rc = drain(stage);
        if (rc) return rc;
        cudaStream_t st = hp.st[stage];
        const size_t row = (size_t)in_width;
        if (llr_dtype == PCL_F16) {
            if (cudaMemcpyAsync(hp.d_raw[stage], (const char*)llr_host + (size_t)f0 * row * 2, (size_t)fc * row * 2,
                                cudaMemcpyHostToDevice, st) != cudaSuccess) return fail_fn(PCL_ECUDA, "H2D failed");
            const int64_t n8 = fc * (int64_t)row / 8;
            widen_f16_kernel<<<(int)std::min<int64_t>((n8 + 255) / 256, 148 * 8), 256, 0, st>>>((const __half*)hp.d_raw[stage],
                                                                                               (float*)hp.d_llr[stage], n8);
        } else if (llr_dtype == PCL_F64 && rsz == 4) {
            const double* src = (const double*)llr_host + (size_t)f0 * row;
            float* dst = hp.h_in[stage];
            parallel_rows(fc, [&](int64_t lo, int64_t hi) { narrow_f64_rows(src + lo * row, dst + lo * row, (hi - lo) * (int64_t)row); });
            if (cudaMemcpyAsync(hp.d_llr[stage], dst, (size_t)fc * row * 4, cudaMemcpyHostToDevice, st) != cudaSuccess)
                return fail_fn(PCL_ECUDA, "H2D failed");
        } else {
            if (cudaMemcpyAsync(hp.d_llr[stage], (const char*)llr_host + (size_t)f0 * row * rsz, (size_t)fc * row * rsz,
                                cudaMemcpyHostToDevice, st) != cudaSuccess) return fail_fn(PCL_ECUDA, "H2D failed");
        }
        rc = launch(stage, st, hp.d_llr[stage], fc, hp.d_bits[stage], hp.d_iters[stage]);
        if (rc) return rc;
        if (out_format == PCL_OUT_BYTES) {
            if (cudaMemcpyAsync((uint8_t*)out_host + (size_t)f0 * out_width, hp.d_bits[stage], (size_t)fc * out_width,
                                cudaMemcpyDeviceToHost, st) != cudaSuccess) return fail_fn(PCL_ECUDA, "D2H failed");
        } else {
            const int64_t jobs = fc * words;
            pack_bits_kernel<<<(int)std::min<int64_t>((jobs + 7) / 8, 148 * 8), 256, 0, st>>>(hp.d_bits[stage], fc, out_width, words,
                                                                                             hp.d_packed[stage]);
            void* dst = out_format == PCL_OUT_PACKED ? (void*)((uint32_t*)out_host + (size_t)f0 * words) : (void*)hp.h_packed[stage];
            if (cudaMemcpyAsync(dst, hp.d_packed[stage], (size_t)fc * words * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess)
                return fail_fn(PCL_ECUDA, "D2H failed");
        }
        if (iters_host && cudaMemcpyAsync(iters_host + f0, hp.d_iters[stage], (size_t)fc * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess)
            return fail_fn(PCL_ECUDA, "D2H failed (iterations)");
        if (cudaEventRecord(hp.done[stage], st) != cudaSuccess) return fail_fn(PCL_ECUDA, "event record failed");
        pend[stage] = {f0, fc, true};
    }
    for (int k = 0; k < PCL_PIPE_STAGES; k++, stage = (stage + 1) % PCL_PIPE_STAGES) {
        int rc = drain(stage);
        if (rc) return rc;
    }
    if (cudaGetLastError() != cudaSuccess) return fail_fn(PCL_ECUDA, "host pipeline: CUDA error");
    return PCL_OK;
}
#endif  // !PCL_EMU
