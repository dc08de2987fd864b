// simt_emu.h -- TEST-ONLY single-threaded SIMT emulator.
//
// Compiles the repo's .cuh kernel sources with g++ so that the kernel *logic*
// (indexing, lazy-copy bookkeeping, bit tricks, warp protocols) can be checked
// against the oracle on the CPU-only build container.  It is never part of the
// product: the shipped library is built by nvcc from the same .cuh files and the
// Python package loads only that library (and fails loudly without it).
//
// Model: every CUDA thread of a block is a ucontext fiber; fibers run one at a
// time and switch only inside barriers (__syncwarp, __syncthreads) and the
// warp-collective intrinsics built on them (__shfl_sync, __ballot_sync, ...).
// Lanes are resumed in ascending or descending order (simt::reverse_order) so a
// missing barrier shows up as a wrong answer in at least one of the two orders.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define PCL_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x) alignas(x)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_ { unsigned x, y, z; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };

namespace simt {

struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
    bool started = false;
    int wait_kind = 0;  // 0 runnable, 1 warp barrier, 2 block barrier, 3 named barrier
    unsigned wait_gen = 0;
    int wait_id = 0;
};

struct WarpState {
    unsigned gen = 0;
    int arrived = 0;
    int alive = 0;
    uint64_t xchg[32];
};

struct BlockState {
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    unsigned block_gen = 0;
    int block_arrived = 0;
    int block_alive = 0;
    ucontext_t sched;
    int current = -1;
    std::function<void()> body;
    std::vector<unsigned char> smem;
    std::vector<uint32_t> tmem;                 // tensor memory: 128 lanes x 512 columns (pcl_tmem.cuh)
    struct NamedBar { unsigned gen = 0; int arrived = 0; };
    NamedBar named[16];
};

inline BlockState*& cur_block() { static BlockState* b = nullptr; return b; }
inline bool& reverse_order() { static bool r = false; return r; }
inline uint3_& tidx() { static uint3_ v; return v; }
inline uint3_& bidx() { static uint3_ v; return v; }
inline dim3& bdim() { static dim3 v; return v; }
inline dim3& gdim() { static dim3 v; return v; }

inline void yield_to_sched() {
    BlockState* b = cur_block();
    swapcontext(&b->fibers[b->current].ctx, &b->sched);
}

inline void warp_barrier() {
    BlockState* b = cur_block();
    int t = b->current;
    WarpState& w = b->warps[t / 32];
    w.arrived++;
    if (w.arrived >= w.alive) { w.arrived = 0; w.gen++; return; }
    Fiber& f = b->fibers[t];
    f.wait_kind = 1; f.wait_gen = w.gen;
    while (b->warps[t / 32].gen == f.wait_gen) yield_to_sched();
    f.wait_kind = 0;
}

inline void block_barrier() {
    BlockState* b = cur_block();
    int t = b->current;
    b->block_arrived++;
    if (b->block_arrived >= b->block_alive) { b->block_arrived = 0; b->block_gen++; return; }
    Fiber& f = b->fibers[t];
    f.wait_kind = 2; f.wait_gen = b->block_gen;
    while (b->block_gen == f.wait_gen) yield_to_sched();
    f.wait_kind = 0;
}

// bar.sync id, nthreads: the first `nthreads` arrivals release the barrier
inline void named_barrier(int id, int nthreads) {
    BlockState* b = cur_block();
    int t = b->current;
    BlockState::NamedBar& nb = b->named[id & 15];
    nb.arrived++;
    if (nb.arrived >= nthreads) { nb.arrived = 0; nb.gen++; return; }
    Fiber& f = b->fibers[t];
    unsigned g = nb.gen;
    f.wait_kind = 3; f.wait_gen = g; f.wait_id = id & 15;
    while (b->named[id & 15].gen == g) yield_to_sched();
    f.wait_kind = 0;
}

// tcgen05.ld / st, shape 32x32b.xN: lane l of warp w moves N consecutive columns of TMEM lane
// 32 (w % 4) + l; the lane field of the address must name the warp's own quarter
inline void tmem_access(uint32_t taddr, uint32_t* v, int nx, bool store) {
    BlockState* b = cur_block();
    int t = b->current;
    if (b->tmem.empty()) b->tmem.assign(128 * 512, 0xCDCDCDCDu);
    int lane_field = (int)(taddr >> 16), col = (int)(taddr & 0xffffu);
    int quarter = ((t / 32) & 3) * 32;
    if (lane_field != quarter || col < 0 || col + nx > 512) {
        fprintf(stderr, "simt_emu: bad TMEM access (lane field %d, warp quarter %d, col %d x%d)\n", lane_field, quarter, col, nx);
        abort();
    }
    uint32_t* row = b->tmem.data() + (size_t)(quarter + t % 32) * 512 + col;
    for (int k = 0; k < nx; k++) {
        if (store) row[k] = v[k];
        else v[k] = row[k];
    }
}

inline void fiber_entry() {
    BlockState* b = cur_block();
    b->body();
    int t = b->current;
    b->fibers[t].done = true;
    // a finished lane counts as permanently arrived
    WarpState& w = b->warps[t / 32];
    w.alive--;
    b->block_alive--;
    if (w.alive > 0 && w.arrived >= w.alive) { w.arrived = 0; w.gen++; }
    if (b->block_alive > 0 && b->block_arrived >= b->block_alive) { b->block_arrived = 0; b->block_gen++; }
    swapcontext(&b->fibers[t].ctx, &b->sched);
}

inline void run_block(unsigned bx, dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    BlockState bs;
    int nt = (int)block.x;
    bs.fibers.resize(nt);
    bs.warps.resize((nt + 31) / 32);
    for (int t = 0; t < nt; t++) bs.warps[t / 32].alive++;
    bs.block_alive = nt;
    bs.body = body;
    bs.smem.assign(smem_bytes + 64, 0xCD);  // poison: uninitialised reads show up
    cur_block() = &bs;
    bidx() = {bx, 0, 0};
    bdim() = block;
    gdim() = grid;
    const size_t STK = 256 * 1024;
    for (int t = 0; t < nt; t++) {
        Fiber& f = bs.fibers[t];
        f.stack.resize(STK);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.data();
        f.ctx.uc_stack.ss_size = STK;
        f.ctx.uc_link = &bs.sched;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    int remaining = nt;
    long spins = 0;
    while (remaining > 0) {
        bool progressed = false;
        for (int k = 0; k < nt; k++) {
            int t = reverse_order() ? nt - 1 - k : k;
            Fiber& f = bs.fibers[t];
            if (f.done) continue;
            if (f.wait_kind == 1 && bs.warps[t / 32].gen == f.wait_gen) continue;
            if (f.wait_kind == 2 && bs.block_gen == f.wait_gen) continue;
            if (f.wait_kind == 3 && bs.named[f.wait_id].gen == f.wait_gen) continue;
            bs.current = t;
            tidx() = {(unsigned)t, 0, 0};
            swapcontext(&bs.sched, &f.ctx);
            progressed = true;
            if (f.done) remaining--;
        }
        if (!progressed) {
            if (++spins > 4) { fprintf(stderr, "simt_emu: deadlock (divergent barrier?)\n"); abort(); }
        } else spins = 0;
    }
    cur_block() = nullptr;
}

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    for (unsigned bx = 0; bx < grid.x; bx++) run_block(bx, grid, block, smem_bytes, body);
}

inline unsigned char* dyn_smem() {
    BlockState* b = cur_block();
    uintptr_t p = (uintptr_t)b->smem.data();
    return (unsigned char*)((p + 15) & ~(uintptr_t)15);
}

template <typename T>
inline T xchg_all(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    BlockState* b = cur_block();
    int t = b->current;
    WarpState& w = b->warps[t / 32];
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w.xchg[t % 32] = raw;
    warp_barrier();
    int base = (t / 32) * 32;
    int nlanes = (int)bdim().x - base; if (nlanes > 32) nlanes = 32;
    int s = ((src_lane % 32) + 32) % 32;
    T out = v;
    if (s < nlanes) memcpy(&out, &w.xchg[s], sizeof(T));
    warp_barrier();
    return out;
}

}  // namespace simt

#define threadIdx (simt::tidx())
#define blockIdx (simt::bidx())
#define blockDim (simt::bdim())
#define gridDim (simt::gdim())

inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_barrier(); }
inline void __syncthreads() { simt::block_barrier(); }

template <typename T> inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    int lane = (int)(threadIdx.x % 32);
    int s = (lane / width) * width + (src % width);
    return simt::xchg_all(v, s);
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int mask, int width = 32) {
    (void)width;
    return simt::xchg_all(v, (int)(threadIdx.x % 32) ^ mask);
}
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
    int lane = (int)(threadIdx.x % 32);
    int s = lane + (int)d;
    if ((s / width) != (lane / width)) s = lane;
    return simt::xchg_all(v, s);
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
    int lane = (int)(threadIdx.x % 32);
    int s = lane - (int)d;
    if (s < 0 || (s / width) != (lane / width)) s = lane;
    return simt::xchg_all(v, s);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned out = 0;
    unsigned mine = pred ? 1u : 0u;
    for (int l = 0; l < 32; l++) {
        // one exchange per lane keeps the emulator simple; cost is irrelevant here
        unsigned b = simt::xchg_all(mine, l);
        int base = ((int)threadIdx.x / 32) * 32;
        if (base + l < (int)blockDim.x) out |= (b & 1u) << l;
    }
    return out;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) {
    unsigned b = __ballot_sync(m, pred);
    int base = ((int)threadIdx.x / 32) * 32;
    int nl = (int)blockDim.x - base; if (nl > 32) nl = 32;
    unsigned full = nl == 32 ? 0xffffffffu : ((1u << nl) - 1);
    return (b & full) == full;
}

inline void sincospif(float x, float* s, float* c) { *s = sinf(3.14159265358979323846f * x); *c = cosf(3.14159265358979323846f * x); }
inline int __ffs(int v) { return v ? __builtin_ffs(v) : 0; }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    unsigned long long v = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xffu) << (8 * i);
    return r;
}
inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
inline long long __double_as_longlong(double d) { long long i; memcpy(&i, &d, 8); return i; }
inline double __longlong_as_double(long long i) { double d; memcpy(&d, &i, 8); return d; }
inline float __fdividef(float a, float b) { return a / b; }
inline float __expf(float a) { return expf(a); }
inline float __logf(float a) { return logf(a); }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }

template <typename T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
template <typename T> inline T __ldg(const T* p) { return *p; }

// ---- just enough of the runtime API for the shared host code ---------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
