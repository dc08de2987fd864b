// pcl_common.cuh -- shared device/host helpers for the sm_100a decoder kernels.
//
// The same .cuh sources are compiled two ways:
//   * nvcc -gencode arch=compute_100a,code=sm_100a  -> libpcl.so (the product)
//   * g++ with tests/emu/simt_emu.h (PCL_EMU)        -> CPU logic tests only
#pragma once

#ifdef PCL_EMU
#include "simt_emu.h"
#else
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#endif

#define PCL_FULL_MASK 0xffffffffu

#ifdef PCL_EMU
#define PCL_DEVICE inline
#define PCL_HOST_DEVICE inline
__device__ inline unsigned char* pcl_dyn_smem() { return simt::dyn_smem(); }
#else
#define PCL_DEVICE __device__ __forceinline__
#define PCL_HOST_DEVICE __host__ __device__ inline
__device__ __forceinline__ unsigned char* pcl_dyn_smem()
{
    extern __shared__ __align__(16) unsigned char pcl_smem_raw[];
    return pcl_smem_raw;
}
#endif

struct pcl_true { static constexpr bool value = true; };
struct pcl_false { static constexpr bool value = false; };

// ---- bit-field helpers -------------------------------------------------------
PCL_DEVICE uint32_t pcl_bfe(uint32_t w, int pos, int len)
{
    return (w >> pos) & ((len >= 32) ? 0xffffffffu : ((1u << len) - 1u));
}
PCL_DEVICE uint32_t pcl_bfi(uint32_t w, uint32_t v, int pos, int len)
{
    uint32_t m = ((len >= 32) ? 0xffffffffu : ((1u << len) - 1u)) << pos;
    return (w & ~m) | ((v << pos) & m);
}

// Packed per-level slot pointers (lazy path copy): field idx holds PB bits.
template <int PB>
PCL_DEVICE int pcl_get_field(uint64_t w, int idx)
{
    if (PB == 0) return 0;
    return (int)((w >> (idx * PB)) & (uint64_t)((1u << PB) - 1u));
}
template <int PB>
PCL_DEVICE uint64_t pcl_set_field(uint64_t w, int idx, int v)
{
    if (PB == 0) return w;
    uint64_t m = (uint64_t)((1u << PB) - 1u) << (idx * PB);
    return (w & ~m) | ((uint64_t)v << (idx * PB));
}

PCL_DEVICE uint64_t pcl_shfl_u64(uint64_t v, int src)
{
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_sync(PCL_FULL_MASK, lo, src);
    hi = __shfl_sync(PCL_FULL_MASK, hi, src);
    return ((uint64_t)hi << 32) | lo;
}

// ---- scalar math in the compute type ------------------------------------------
template <typename real> struct pcl_math;

template <> struct pcl_math<float> {
    // f(a,b) = sign(a) sign(b) min(|a|,|b|)  (reference: polar/decoder.py:121-127)
    static PCL_DEVICE float f(float a, float b)
    {
#if defined(PCL_EMU) || defined(PCL_NO_XORSIGN)
        float mn = fminf(fabsf(a), fabsf(b));
        uint32_t s = (__float_as_uint(a) ^ __float_as_uint(b)) & 0x80000000u;
        return __uint_as_float(__float_as_uint(mn) | s);
#else
        // one FMNMX.XORSIGN instead of FMNMX + 2 LOP3: min(|a|,|b|) with sign(a) ^ sign(b), exact
        float r;
        asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
        return r;
#endif
    }
    // log(1 + exp(-|x|)), the sign-independent part of decoder.py:374-406
    static PCL_DEVICE float softplus_neg_abs(float ax) { return log1pf(expf(-ax)); }
    static PCL_DEVICE float inf() { return __uint_as_float(0x7f800000u); }
};

template <> struct pcl_math<double> {
    static PCL_DEVICE double f(double a, double b)
    {
        double mn = fmin(fabs(a), fabs(b));
        long long s = (__double_as_longlong(a) ^ __double_as_longlong(b)) & (long long)0x8000000000000000ull;
        return __longlong_as_double(__double_as_longlong(mn) | s);
    }
    static PCL_DEVICE double softplus_neg_abs(double ax) { return log1p(exp(-ax)); }
    static PCL_DEVICE double inf() { return __longlong_as_double(0x7ff0000000000000ll); }
};
