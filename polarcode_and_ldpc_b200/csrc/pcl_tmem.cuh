// pcl_tmem.cuh -- Blackwell tensor memory (TMEM) as a per-lane scratchpad for the decoder's
// mid-tree LLR levels, and named barriers for groups of warps inside one fat CTA.
//
// TMEM is 512 columns x 128 lanes x 32 bit per SM; warp w of a CTA reaches lanes
// 32 (w % 4) .. + 31 only, and with the 32x32b shape lane l of the warp reads / writes NX
// consecutive columns of TMEM lane 32 (w % 4) + l -- exactly "a path (lane) owns its own array,
// four / sixteen consecutive elements per access" of the decoder's level layout, at 12-43 cycles
// latency (profiles/onchip_peaks.json) instead of an L2 / HBM round trip.  Nothing here feeds a
// tensor core: tcgen05.alloc / st / ld / wait / dealloc only.  SASS: UTCALLOC, STTM, LDTM.
//
// The SIMT-emulator build (tests/emu) models the same semantics on a per-block array.
#pragma once
#include "pcl_common.cuh"

#define PCL_TMEM_COLS 512

#ifdef PCL_EMU
PCL_DEVICE void pcl_tmem_alloc_all(uint32_t* smem_slot) { *smem_slot = 0; }
PCL_DEVICE void pcl_tmem_free_all(uint32_t) {}
PCL_DEVICE void pcl_tmem_fence_before() {}
PCL_DEVICE void pcl_tmem_fence_after() {}
PCL_DEVICE void pcl_tmem_wait_ld() {}
PCL_DEVICE void pcl_tmem_wait_st() {}
template <int NX> PCL_DEVICE void pcl_tmem_ld(uint32_t taddr, uint32_t* v) { simt::tmem_access(taddr, v, NX, false); }
template <int NX> PCL_DEVICE void pcl_tmem_st(uint32_t taddr, const uint32_t* v) { simt::tmem_access(taddr, const_cast<uint32_t*>(v), NX, true); }
PCL_DEVICE void pcl_named_barrier(int id, int nthreads) { simt::named_barrier(id, nthreads); }
#else
// One warp allocates the SM's whole TMEM for the CTA (the decoder runs one CTA per SM) and gives
// up the allocation permit; every thread reads the base address after the fence / barrier pair.
PCL_DEVICE void pcl_tmem_alloc_all(uint32_t* smem_slot)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(smem_slot)), "r"(PCL_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
PCL_DEVICE void pcl_tmem_free_all(uint32_t tbase)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(PCL_TMEM_COLS) : "memory");
}
PCL_DEVICE void pcl_tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
PCL_DEVICE void pcl_tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
PCL_DEVICE void pcl_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
PCL_DEVICE void pcl_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int NX> PCL_DEVICE void pcl_tmem_ld(uint32_t taddr, uint32_t* v);
template <> PCL_DEVICE void pcl_tmem_ld<4>(uint32_t taddr, uint32_t* v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
template <> PCL_DEVICE void pcl_tmem_ld<8>(uint32_t taddr, uint32_t* v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
template <> PCL_DEVICE void pcl_tmem_ld<16>(uint32_t taddr, uint32_t* v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
template <int NX> PCL_DEVICE void pcl_tmem_st(uint32_t taddr, const uint32_t* v);
template <> PCL_DEVICE void pcl_tmem_st<4>(uint32_t taddr, const uint32_t* v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
template <> PCL_DEVICE void pcl_tmem_st<8>(uint32_t taddr, const uint32_t* v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
template <> PCL_DEVICE void pcl_tmem_st<16>(uint32_t taddr, const uint32_t* v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
// bar.sync id, nthreads: the warps of one group meet without stopping the rest of the CTA
PCL_DEVICE void pcl_named_barrier(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}
#endif

// ---- cp.async (LDGSTS): 16 bytes global -> shared without a register stop-over ----------------
#ifdef PCL_EMU
PCL_DEVICE void pcl_cp_async16(void* smem_dst, const void* gsrc) { memcpy(smem_dst, gsrc, 16); }
PCL_DEVICE void pcl_cp_async_commit() {}
template <int N> PCL_DEVICE void pcl_cp_async_wait() {}
#else
PCL_DEVICE void pcl_cp_async16(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
PCL_DEVICE void pcl_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> PCL_DEVICE void pcl_cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
#endif

// TMEM address of column `col` in this warp's lane quarter
PCL_DEVICE uint32_t pcl_tmem_addr(uint32_t tbase, int warp, int col)
{
    return tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)col;
}
