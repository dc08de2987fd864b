// ldpc_banked.cuh -- flooding BP / Min-Sum for regular (3, DC) codes, fp32, with a
// shared-memory layout in which NEITHER pass has bank conflicts.
//
// Same schedule and arithmetic as ldpc_bp.cuh (read that header first); what changes is where
// an edge message lives.  ncu on the check-major layout (profiles/r01l): the shared-memory pipe
// is the top limiter (83 %), 43 % of its wavefronts being bank conflicts of the variable-node
// gather / scatter through `vperm` (random banks, ~2.6-way).  Here:
//   * a check sits at (round R, lane l); its k-th message is word 32 (DC R + k) + l: a check
//     round reads / writes DC rows of 32 consecutive words -- conflict free, no tables;
//   * a variable sits at position pi = 32 r + lane; its three messages are fetched by three
//     instructions, and the host chooses (a) the lane l of every check, (b) the round r of every
//     variable, (c) which of a variable's edges each of the three instructions takes, such that
//     the 32 words one instruction touches lie in 32 different banks (bank = l).  (a), (b) come
//     from a short annealing run at handle creation (a round's 96 edges must hit every bank
//     exactly 3 times), (c) from recolouring; the few edges that cannot be placed cost one extra
//     wavefront each (about 10 of the 1512 fetches of a pass for the n = 504 code).
// Hard decisions, the channel LLRs and the syndrome test live in position space; `varof` /
// `posof` translate at frame load and store.  The variable sum adds its three messages in
// instruction order, not in ascending check order, and BP messages are carried in units of ln 2
// (one multiply less on each side of the check rule); both fp32 build only -- the fp64 validation
// build keeps ldpc_bp.cuh, the reference's order and its units.
#pragma once
#include "ldpc_bp.cuh"

// Message words addressed by a 32-bit shared-space address + byte offset (the variable pass): the
// offsets come packed from a table and one integer add per word forms the address.
#ifdef PCL_EMU
typedef unsigned char* pcl_saddr;
PCL_DEVICE pcl_saddr pcl_saddr_of(void* p) { return (unsigned char*)p; }
PCL_DEVICE float pcl_lds_f32(pcl_saddr a) { return *reinterpret_cast<const float*>(a); }
PCL_DEVICE void pcl_sts_f32(pcl_saddr a, float v) { *reinterpret_cast<float*>(a) = v; }
#else
typedef uint32_t pcl_saddr;
PCL_DEVICE pcl_saddr pcl_saddr_of(void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
PCL_DEVICE float pcl_lds_f32(pcl_saddr a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
PCL_DEVICE void pcl_sts_f32(pcl_saddr a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
#endif

// PAIRED: the k-th message of the check at (R, lane) is word 64 (DC/2 R + k/2) + 2 lane + (k & 1):
// the check pass moves two messages per 8-byte access (conflict free), an edge's bank becomes
// 2 (lane % 16) + (k & 1), and the host search also decides which half of a check's edges sit
// at odd k.
// One warp per frame (COOP = 0): the library launches ONE block per SM holding every warp the SM's
// shared memory has room for (up to 32), so the residency cannot depend on which shared-memory /
// L1 split the SM happens to be in when the grid arrives.
// NPR > 0 (every lane owns exactly NPR variable positions: NP == 32 NPR in the warp-per-frame mode with at most 28 warps
// per block, NP == 128 NPR in the block-per-frame mode): a lane keeps the channel values of its positions in registers
// for the whole decode (72 registers instead of 55; NPR shared-memory reads and their address arithmetic less per
// iteration: BP n = 504 10.91 -> 11.12 Gbps).
template <int MODE, int DC, int COOP, int PAIRED, int NPR = 0>
__global__ void __launch_bounds__(COOP ? 256 : (NPR ? 896 : 1024)) ldpc_banked_kernel(LdpcParams<float> P)
{
    const LdpcLayout& Y = P.lay;
    const int n = Y.n, nR = Y.nR, NP = Y.NP, NS = Y.NS;
    const int lane = threadIdx.x & 31;
#ifdef PCL_LDPC_ONEWARP     // experiment: blocks of one warp (launch with PCL_LDPC_FAT=0 PCL_LDPC_WPB=1), constant shared-memory base
    const int warp = COOP ? (threadIdx.x >> 5) : 0;
#else
    const int warp = threadIdx.x >> 5;
#endif
    constexpr bool coop = COOP != 0;
    const int T = coop ? (int)blockDim.x : 32;
    const int tid = coop ? (int)threadIdx.x : lane;
    const int w0 = coop ? warp : 0, wstep = coop ? (int)(blockDim.x >> 5) : 1;
    unsigned char* wsm = pcl_dyn_smem() + (coop ? (size_t)0 : (size_t)warp * Y.warp_bytes);
    float* msg = (float*)(wsm + Y.off_msg);
    float* sllr = (float*)(wsm + Y.off_llr);
    const pcl_saddr msg_s = pcl_saddr_of(msg);
    uint32_t* hard = (uint32_t*)(wsm + Y.off_hard);
    unsigned long long* ctl = (unsigned long long*)(wsm + Y.off_ctl);
    auto sync = [&]() {
        if (coop) __syncthreads();
        else __syncwarp();
    };

    for (int s = tid; s < NS; s += T) msg[s] = 0.0f;                   // the words of empty check seats stay (numerically) zero from here on
    sync();
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

        const float* ch = P.llr + f * n;
        for (int pi = tid; pi < NP; pi += T) {
            const int v = P.varof[pi];
            // BP carries every message in units of ln 2 (cn_bp_core<.., LOG2>): scale the channel once
            sllr[pi] = (v != 0xffff) ? ch[v] * (MODE == 0 ? 1.4426950408889634f : 1.0f) : 0.0f;
        }
        sync();
        // decoder.py:144-146: every edge starts with its variable's channel value -- scattered from position
        // space through the same byte-offset table the variable pass uses (an empty position writes its zero to
        // the zero words of an empty seat; those were cleared once, before the first frame, and stay zero)
        for (int pi = tid; pi < NP; pi += T) {
            const unsigned long long pk = P.bpack[pi];
            const uint32_t plo = (uint32_t)pk;
            const float c = sllr[pi];
            pcl_sts_f32(msg_s + (plo & 0xffffu), c);
            pcl_sts_f32(msg_s + (plo >> 16), c);
            pcl_sts_f32(msg_s + (uint32_t)(pk >> 32), c);
        }
        float cl[NPR > 0 ? NPR : 1];
        if (NPR > 0) {
#pragma unroll
            for (int r = 0; r < NPR; r++) cl[r] = sllr[32 * (w0 + wstep * r) + lane];
        }
        sync();

        int iters = Y.max_iter;
        for (int it = 0; it < Y.max_iter; it++) {
            // 1. check nodes: round R, lane = the check's bank (an empty seat computes on zeros)
            int R = w0;
#if !defined(PCL_LDPC_NO_PAIR2)
            if (MODE == 0 && PAIRED) {
                // BP: two rounds per step, the two checks of a lane side by side on the fp32x2 pipe
                for (; R + wstep < nR; R += 2 * wstep) {
                    float* ba = msg + 32 * DC * R + 2 * lane;
                    float* bb = ba + 32 * DC * wstep;
                    uint32_t xa[DC], xb2[DC];
                    float oa[DC], ob[DC];
#pragma unroll
                    for (int k = 0; k < DC; k += 2) {
                        const float2 va = *reinterpret_cast<const float2*>(ba + 32 * k);
                        const float2 vb = *reinterpret_cast<const float2*>(bb + 32 * k);
                        xa[k] = __float_as_uint(va.x); xa[k + 1] = __float_as_uint(va.y);
                        xb2[k] = __float_as_uint(vb.x); xb2[k + 1] = __float_as_uint(vb.y);
                    }
                    cn_bp_core2<DC>(xa, xb2, oa, ob);
#pragma unroll
                    for (int k = 0; k < DC; k += 2) {
                        float2 va, vb;
                        va.x = oa[k]; va.y = oa[k + 1];
                        vb.x = ob[k]; vb.y = ob[k + 1];
                        *reinterpret_cast<float2*>(ba + 32 * k) = va;
                        *reinterpret_cast<float2*>(bb + 32 * k) = vb;
                    }
                }
            }
#endif
            for (; R < nR; R += wstep) {
                float* base = msg + 32 * DC * R + (PAIRED ? 2 * lane : lane);
                float x[DC], out[DC];
                if (PAIRED) {
#pragma unroll
                    for (int k = 0; k < DC; k += 2) {
                        const float2 v = *reinterpret_cast<const float2*>(base + 32 * k);
                        x[k] = v.x;
                        x[k + 1] = v.y;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < DC; k++) x[k] = base[32 * k];
                }
                if (MODE == 1) {
                    cn_ms_core<float, DC>(x, out, P.norm);
                } else {
                    uint32_t xb[DC];
#pragma unroll
                    for (int k = 0; k < DC; k++) xb[k] = __float_as_uint(x[k]);
                    cn_bp_core<DC, true, true>(xb, out, DC);
                }
                if (PAIRED) {
#pragma unroll
                    for (int k = 0; k < DC; k += 2) {
                        float2 v;
                        v.x = out[k];
                        v.y = out[k + 1];
                        *reinterpret_cast<float2*>(base + 32 * k) = v;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < DC; k++) base[32 * k] = out[k];
                }
            }
            sync();
            // 2. variable nodes + 3. hard decision, in position space
            const bool want_hard = Y.early_stop || it == Y.max_iter - 1;
            // (the pass exists in four compiled versions -- with / without hard decisions and the total-LLR
            // dump -- so that the common one carries no per-position tests; table words of four positions
            // are fetched ahead of the shared-memory accesses)
            auto var_pass = [&](auto hard_c, auto total_c) {
                constexpr bool HARD = decltype(hard_c)::value, TOTAL = decltype(total_c)::value;
                // three byte offsets into the message array, 16 bits each
                auto one = [&](int vb, unsigned long long pk, float chv = 0.0f, bool have = false) {
                    const int pi = vb + lane;
                    const uint32_t plo = (uint32_t)pk;
                    const pcl_saddr pa = msg_s + (plo & 0xffffu), pb = msg_s + (plo >> 16), pc = msg_s + (uint32_t)(pk >> 32);
                    const float ma = pcl_lds_f32(pa), mb = pcl_lds_f32(pb), mc = pcl_lds_f32(pc);
                    const float total = (have ? chv : sllr[pi]) + ((ma + mb) + mc);
                    pcl_sts_f32(pa, total - ma);
                    pcl_sts_f32(pb, total - mb);
                    pcl_sts_f32(pc, total - mc);
                    if (TOTAL) {
                        const int v = P.varof[pi];
                        if (v != 0xffff) P.total[f * n + v] = total * (MODE == 0 ? 0.6931471805599453f : 1.0f);
                    }
                    if (HARD) {
                        const unsigned bal = __ballot_sync(PCL_FULL_MASK, total <= 0.0f);
                        if (lane == 0) hard[vb >> 5] = bal;
                    }
                };
                const int step = 32 * wstep;
                int vb = 32 * w0;
                if constexpr (NPR > 0) {
#pragma unroll
                    for (int r4 = 0; r4 < NPR; r4 += 4) {
                        unsigned long long k4[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) k4[e] = P.bpack[32 * (w0 + wstep * (r4 + e)) + lane];
#pragma unroll
                        for (int e = 0; e < 4; e++) one(32 * (w0 + wstep * (r4 + e)), k4[e], cl[r4 + e], true);
                    }
                    return;
                }
                for (; vb + 3 * step < NP; vb += 4 * step) {
                    const unsigned long long k0 = P.bpack[vb + lane], k1 = P.bpack[vb + step + lane],
                                             k2 = P.bpack[vb + 2 * step + lane], k3 = P.bpack[vb + 3 * step + lane];
                    one(vb, k0);
                    one(vb + step, k1);
                    one(vb + 2 * step, k2);
                    one(vb + 3 * step, k3);
                }
                for (; vb < NP; vb += step) one(vb, P.bpack[vb + lane]);
            };
            if (P.total != nullptr) {
                if (want_hard) var_pass(pcl_true(), pcl_true());
                else var_pass(pcl_false(), pcl_true());
            } else {
                if (want_hard) var_pass(pcl_true(), pcl_false());
                else var_pass(pcl_false(), pcl_false());
            }
            sync();
            // 4. syndrome early stop
            if (Y.early_stop) {
                bool bad = false;
                for (int R = w0; R < nR; R += wstep) {
                    unsigned par = 0;
#pragma unroll
                    for (int k = 0; k < DC; k++) {
                        const int cp = P.cpos[PAIRED ? 32 * DC * R + 32 * (k & ~1) + 2 * lane + (k & 1) : 32 * (DC * R + k) + lane];
                        if (cp != 0xffff) par ^= hard[cp >> 5] >> (cp & 31);
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
        uint8_t* outb = P.bits + f * n;
        for (int v = tid; v < n; v += T) {
            const int pi = P.posof[v];
            outb[v] = (uint8_t)((hard[pi >> 5] >> (pi & 31)) & 1u);
        }
        if (P.iters != nullptr && tid == 0) P.iters[f] = iters;
        sync();
    }
}
