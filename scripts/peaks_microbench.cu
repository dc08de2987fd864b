// peaks_microbench.cu -- measured on-chip peaks of one B200 for the roofline denominators that
// MEASURED_PEAKS.json does not hold ("the builder must measure one with a
// micro-benchmark and record both"):
//   * shared-memory bandwidth: conflict-free LDS.128 / LDS.32 streams on every SM
//   * MUFU rate: independent ex2.approx streams (the BP check rule is MUFU-bound before SMEM)
//   * FP32 FMA issue rate (sanity line for the two above)
//   * tensor memory (TMEM) as a scratchpad: tcgen05.ld / tcgen05.st latency and bandwidth
// Build + run (scripts/run_peaks.sh):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o _variants/peaks_microbench
// scripts/peaks_microbench.cu && _variants/peaks_microbench > profiles/onchip_peaks.json
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

// ---------------------------------------------------------------- shared memory ---
template <int VEC>
__global__ void __launch_bounds__(256) lds_stream(uint32_t* out, int iters)
{
    extern __shared__ __align__(16) uint32_t sm[];
    const int t = threadIdx.x;
    for (int i = t; i < 8 * 512 * VEC; i += blockDim.x) sm[i] = i * 2654435761u;
    __syncthreads();
    uint32_t acc0 = 0, acc1 = 0;
    // each warp walks its own window of 16 rows; lane l reads VEC consecutive words at l * VEC of row k
    // (inline PTX with a memory clobber and an iteration-dependent row: nvcc hoisted or merged the
    // loads of two earlier versions of this loop, which then reported 4-8 x the real rate)
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm + (t >> 5) * 512 * VEC + (t & 31) * VEC);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t a = base + ((k + it) & 15) * 32 * VEC * 4;   // the row depends on `it`: nothing to hoist
            if (VEC == 4) {
                uint32_t x, y, z, w;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a) : "memory");
                acc0 += x ^ y;
                acc1 += z ^ w;
            } else {
                uint32_t x;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(a) : "memory");
                acc0 += x;
            }
        }
    }
    if ((acc0 ^ acc1) == 0x12345u) out[blockIdx.x * blockDim.x + t] = acc0;
}

// ------------------------------------------------------------------------ MUFU ---
__global__ void __launch_bounds__(256) mufu_stream(float* out, int iters, float seed)
{
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = seed * (float)(threadIdx.x + k) * 1e-3f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) ffma_stream(float* out, int iters, float seed)
{
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = seed * (float)(threadIdx.x + k);
    const float m = seed * 0.999f, c = seed * 1e-3f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(m), "f"(c));
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ------------------------------------------------------------------------ TMEM ---
// One CTA per SM allocates all 512 columns; warp w owns lanes 32 (w % 4) .. + 31 and the column
// window (w / 4) * cols_per_warp.  A lane reads / writes NX consecutive 32-bit columns of its own
// TMEM lane per instruction (shape 32x32b.xNX).
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t* v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t* v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t addr, uint32_t* v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mode 0: ld x16 throughput (4 loads in flight per wait); 1: st x16 throughput; 2: dependent ld x4 latency
__global__ void __launch_bounds__(768, 1) tmem_bench(uint32_t* out, long long* cycles, int iters, int mode)
{
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(&tbase)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int wq = (blockDim.x >> 5) / 4 > 0 ? (blockDim.x >> 5) / 4 : 1;     // warps per lane quarter
    const int cpw = (512 / wq) & ~15;                                        // columns per warp
    const uint32_t my = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * cpw);
    uint32_t v[16], w[16], x[16], y[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = lane * 1000 + k + warp * 100000;
    // fill the window, then read it back once as a correctness check
    for (int c = 0; c + 16 <= cpw; c += 16) {
#pragma unroll
        for (int k = 0; k < 16; k++) w[k] = v[k] + c * 7;
        tmem_st16(my + c, w);
    }
    tmem_wait_st();
    uint32_t bad = 0;
    for (int c = 0; c + 16 <= cpw; c += 16) {
        tmem_ld16(my + c, w);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 16; k++) bad += (w[k] != v[k] + c * 7);
    }
    __syncthreads();
    const long long t0 = clock64();
    uint32_t acc = 0;
    if (mode == 0) {
        for (int it = 0; it < iters; it++) {
            const int c = (it * 64) % (cpw - 63 > 0 ? cpw - 63 : 1) & ~15;
            tmem_ld16(my + c, v);
            tmem_ld16(my + c + 16, w);
            tmem_ld16(my + c + 32, x);
            tmem_ld16(my + c + 48, y);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 16; k += 4) acc ^= v[k] ^ w[k] ^ x[k] ^ y[k];
        }
    } else if (mode == 1) {
        for (int it = 0; it < iters; it++) {
            const int c = (it * 64) % (cpw - 63 > 0 ? cpw - 63 : 1) & ~15;
            v[0] += it;
            tmem_st16(my + c, v);
            tmem_st16(my + c + 16, v);
            tmem_st16(my + c + 32, v);
            tmem_st16(my + c + 48, v);
            tmem_wait_st();
        }
    } else {
        uint32_t c = 0;
        for (int it = 0; it < iters; it++) {
            tmem_ld4(my + (c & 31), v);
            tmem_wait_ld();
            c = (v[0] + v[1]) & 1;            // the next address depends on the loaded value
            acc += c;
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (lane == 0) cycles[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    if (acc == 0xdeadbeefu || bad) out[blockIdx.x] = bad ? 0xbad00000u + bad : acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(512) : "memory");
}

template <typename F>
static double time_ms(F&& launch, int reps = 5)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();                                  // warm-up
    CK(cudaDeviceSynchronize());
    std::vector<float> t;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        t.push_back(ms);
    }
    CK(cudaGetLastError());
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main(int argc, char** argv)
{
    const bool with_tmem = !(argc > 1 && atoi(argv[1]) == 0);
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    uint32_t* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, 64 << 20)); CK(cudaMemset(d_out, 0, 64 << 20));
    CK(cudaMalloc(&d_cyc, 1 << 20));
    printf("{\n \"gpu\": \"%s\", \"sms\": %d, \"sm_clock_max_mhz\": %.1f,\n", pr.name, sms, clk_khz / 1000.0);

    // ---- shared memory
    {
        const int iters = 20000, blocks = sms * 3, threads = 256, smem = 65536;
        CK(cudaFuncSetAttribute(lds_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(lds_stream<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        double ms4 = time_ms([&] { lds_stream<4><<<blocks, threads, smem>>>(d_out, iters); });
        double ms1 = time_ms([&] { lds_stream<1><<<blocks, threads, smem>>>(d_out, iters); });
        const double b4 = (double)blocks * threads * iters * 16 * 16, b1 = (double)blocks * threads * iters * 16 * 4;
        printf(" \"smem_lds128_gbs\": %.1f, \"smem_lds32_gbs\": %.1f,\n", b4 / ms4 / 1e6, b1 / ms1 / 1e6);
        printf(" \"smem_lds128_bytes_per_clk_per_sm_at_max_clock\": %.2f,\n", b4 / (ms4 * 1e-3) / sms / (clk_khz * 1e3));
    }
    // ---- MUFU / FMA
    {
        const int iters = 40000, blocks = sms * 8, threads = 256;
        double msm = time_ms([&] { mufu_stream<<<blocks, threads>>>((float*)d_out, iters, 1.0f); });
        double msf = time_ms([&] { ffma_stream<<<blocks, threads>>>((float*)d_out, iters, 1.0f); });
        const double ops = (double)blocks * threads * iters * 8;
        printf(" \"mufu_ex2_gops\": %.1f, \"mufu_per_clk_per_sm_at_max_clock\": %.2f,\n", ops / msm / 1e6,
               ops / (msm * 1e-3) / sms / (clk_khz * 1e3));
        printf(" \"ffma_gops\": %.1f, \"ffma_per_clk_per_sm_at_max_clock\": %.2f,\n", ops / msf / 1e6,
               ops / (msf * 1e-3) / sms / (clk_khz * 1e3));
    }
    // ---- TMEM
    if (with_tmem) {
        for (int warps : {4, 8, 16, 24}) {
            const int iters = 4000;
            double ms_ld = time_ms([&] { tmem_bench<<<sms, warps * 32>>>(d_out, d_cyc, iters, 0); }, 3);
            double ms_st = time_ms([&] { tmem_bench<<<sms, warps * 32>>>(d_out, d_cyc, iters, 1); }, 3);
            const double bytes = (double)sms * warps * iters * 4 * 16 * 32 * 4;
            printf(" \"tmem_ld_x16_gbs_%dwarps\": %.1f, \"tmem_st_x16_gbs_%dwarps\": %.1f,\n", warps, bytes / ms_ld / 1e6,
                   warps, bytes / ms_st / 1e6);
        }
        const int iters = 2000;
        tmem_bench<<<sms, 128>>>(d_out, d_cyc, iters, 2);
        CK(cudaDeviceSynchronize());
        long long c = 0;
        CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
        uint32_t flag = 0;
        CK(cudaMemcpy(&flag, d_out, 4, cudaMemcpyDeviceToHost));
        printf(" \"tmem_ld_x4_dependent_cycles\": %.1f, \"tmem_readback_ok\": %s,\n", (double)c / iters,
               (flag >> 20) == 0xbad ? "false" : "true");
    }
    printf(" \"note\": \"conflict-free LDS streams / independent ex2 and fma streams on all SMs; TMEM: one CTA per SM owning 512 columns, 32x32b shapes\"\n}\n");
    return 0;
}
