// tcgen05 / TMEM helpers shared by the tensor-core decode kernels.
//
// Shared-memory operands use the canonical K-major, no-swizzle UMMA layout: 8-row x 16-byte core
// matrices stored [K/8][rows][8 x 16-bit]; SBO (8-row group stride) = 128 B, LBO (K-chunk stride) =
// rows * 16 B.  TMEM operands / accumulators: lane = row; a 16-bit A operand packs k = 2j (low half)
// and k = 2j+1 (high half) into column j, one K block (16 elements) = 8 columns
// (layout pinned on hardware by tests/tc_probe/tmem_a_probe.cu).
#pragma once
#include <cuda/ptx>
#include <stdint.h>

namespace sw {
namespace ptx = cuda::ptx;

__device__ __forceinline__ uint64_t umma_desc(const void* smem_ptr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_ptr);
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);   // version = 1 (sm_100), base_offset = 0, layout_type = SWIZZLE_NONE
}

// Same descriptor, but broadcast from lane 0 with a shuffle: the compiler then KNOWS the value is warp-uniform, keeps
// it (and every "+ k-block offset" derived from it) in uniform registers and feeds UTCHMMA without the
// ELECT / R2UR.BROADCAST "waterfall" loop it otherwise emits per MMA.  Call from convergent code (all 32 lanes).
__device__ __forceinline__ uint64_t umma_desc_uniform(const void* smem_ptr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint64_t d = umma_desc(smem_ptr, lbo_bytes, sbo_bytes);
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)d, 0), hi = __shfl_sync(0xffffffffu, (uint32_t)(d >> 32), 0);
    return ((uint64_t)hi << 32) | lo;
}

// instruction descriptor: FP32 accumulate (c_format 1 @4), a/b format @7/@10 (0 = F16, 1 = BF16), K-major A and B,
// N >> 3 @17, M = 128 (>> 4) @24
__device__ __forceinline__ constexpr uint32_t umma_idesc(int n, uint32_t ab_format) {
    return (1u << 4) | (ab_format << 7) | (ab_format << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
}

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    // bounded spin: a lost MMA completion becomes a trap (reported as a CUDA error), never a hung GPU
    for (uint32_t spins = 0; !ptx::mbarrier_try_wait_parity(reinterpret_cast<uint64_t*>(bar), parity); ++spins)
        if (spins > (1u << 24)) __trap();
}

// TMA (bulk asynchronous copy engine), 1-D form: `bytes` (multiple of 16, both addresses 16-byte aligned) from global to this
// CTA's shared memory; completion is signalled on `bar` as transaction bytes.  One thread arms the barrier with
// mbar_expect_tx(total bytes) and issues the copies; consumers wait on the barrier's phase (mbar_wait).
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    ptx::mbarrier_arrive_expect_tx(ptx::sem_release, ptx::scope_cta, ptx::space_shared, reinterpret_cast<uint64_t*>(bar), bytes);
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* bar) {
    ptx::cp_async_bulk(ptx::space_cluster, ptx::space_global, smem_dst, gmem_src, bytes, reinterpret_cast<uint64_t*>(bar));
}

// 2-D tiled form through a tensor map (cuTensorMapEncodeTiled on the host, passed as a __grid_constant__ kernel parameter):
// box (c0 .. , c1 ..) -> shared memory, out-of-bounds elements zero-filled, optional 128-byte hardware swizzle (16-byte chunk
// index XOR (row & 7): row-per-lane reads of a [rows][128 B] box are then bank-conflict free).  Destination 1024-byte aligned.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tensor_map, int c0, int c1, unsigned long long* bar) {
    const int32_t coords[2] = {c0, c1};
    ptx::cp_async_bulk_tensor(ptx::space_cluster, ptx::space_global, smem_dst, tensor_map, coords, reinterpret_cast<uint64_t*>(bar));
}

// MMA / commit issue.  Called by ALL 32 lanes of the issuing warp (convergent code: descriptors stay in uniform
// registers); `leader` predicates the instruction itself so that exactly one lane issues it.
__device__ __forceinline__ void umma_commit(unsigned long long* bar, bool leader) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                 :: "r"(addr), "r"((uint32_t)leader) : "memory");
}

__device__ __forceinline__ void umma_issue_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              bool accumulate, bool leader) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate), "r"((uint32_t)leader)
                 : "memory");
}

__device__ __forceinline__ void umma_issue_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                              bool accumulate, bool leader) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate), "r"((uint32_t)leader)
                 : "memory");
}

// D[128 x N] (+)= A[128 x 16*KB] . B[N x 16*KB]^T, both operands 16-bit in shared memory.
// `a`: [K/8][128][8]; `b`: [K/8][B_ROWS][8] (the MMA uses N consecutive rows starting at `b`).
template <int B_ROWS, int N, int KB, typename T>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, const T* a, const T* b, uint32_t ab_format, bool accumulate_first,
                                        bool leader) {
    const uint32_t idesc = umma_idesc(N, ab_format);
    const uint64_t ad = umma_desc_uniform(a, 128 * 16, 128), bd = umma_desc_uniform(b, B_ROWS * 16, 128);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)      // start-address field is in 16-byte units: one K block = 2 chunks of rows*16 B
        umma_issue_ss(d_tmem, ad + (uint64_t)(kb * 2 * 128), bd + (uint64_t)(kb * 2 * B_ROWS), idesc, accumulate_first || kb > 0, leader);
}

// same with the A operand in TMEM (16-bit: one K block = 8 columns; consecutive K blocks A_STRIDE columns apart)
template <int B_ROWS, int N, int KB, int A_STRIDE = 8, typename T>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, const T* b, uint32_t ab_format, bool accumulate_first,
                                        bool leader) {
    const uint32_t idesc = umma_idesc(N, ab_format);
    const uint64_t bd = umma_desc_uniform(b, B_ROWS * 16, 128);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
        umma_issue_ts(d_tmem, a_tmem + kb * A_STRIDE, bd + (uint64_t)(kb * 2 * B_ROWS), idesc, accumulate_first || kb > 0, leader);
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair form (cta_group::2, a cluster of two CTAs on the two SMs of a TPC): ONE instruction, issued by a lane of the
// leader CTA (cluster rank 0), computes D[256 x N] = A[256 x K] . B[N x K]^T where each CTA supplies ITS 128 rows of A
// (same shared-memory offset / TMEM address in both CTAs), ITS N/2 rows of B (rank 0: rows [0, N/2), rank 1: [N/2, N)) and
// receives ITS 128 rows of D (all N columns) in its own TMEM.  Every B matrix is therefore stored once per pair --
// half per SM -- which is what lets two 128-row tiles per SM be in flight against one resident weight set.
// (operand placement and instruction form pinned on hardware by tests/tc_probe/pair_mma_probe.cu)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ constexpr uint32_t umma_idesc_pair(int n, uint32_t ab_format) {       // M = 256
    return (1u << 4) | (ab_format << 7) | (ab_format << 10) | ((uint32_t)(n >> 3) << 17) | (16u << 24);
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// (Remote arrivals -- mapa + mbarrier.arrive.shared::cluster, default .release.cta semantics as CUTLASS's
// ClusterBarrier::arrive(cta_id) -- are issued inline by the kernels: the cluster-scope form costs MEMBAR.ALL.GPU + ERRBAR per
// arrival and CCTL.IVALL per wait, 47 % of all stall samples of the first pair kernel.  What crosses the pair is consumed by
// the tensor core's async proxy (shared-memory operands, made visible by fence.proxy.async before the arrival) or lives in
// TMEM (tcgen05.wait::st + fence::before_thread_sync).)
// wait on a barrier of this CTA whose arrivals come from the peer CTA (operand-ready) or from a multicast tcgen05.commit
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long* bar, uint32_t parity) { mbar_wait(bar, parity); }
// completion of every MMA issued so far by this thread -> one arrival on `bar` in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(unsigned long long* bar, bool leader) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %2;\n\t}"
                 :: "r"(addr), "r"((uint32_t)leader), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_issue_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                   bool accumulate, bool leader) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate), "r"((uint32_t)leader)
                 : "memory");
}
__device__ __forceinline__ void umma_issue_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                   bool accumulate, bool leader) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate), "r"((uint32_t)leader)
                 : "memory");
}
// D[256 x N] (+)= A . B^T over KB K-blocks, A in shared memory ([K/8][128][8] per CTA), B [K/8][NL = N/2 local rows][8] per CTA
template <int NL, int KB, typename T>
__device__ __forceinline__ void pmma_ss(uint32_t d_tmem, const T* a, const T* b, uint32_t ab_format, bool accumulate_first, bool leader) {
    const uint32_t idesc = umma_idesc_pair(2 * NL, ab_format);
    const uint64_t ad = umma_desc_uniform(a, 128 * 16, 128), bd = umma_desc_uniform(b, NL * 16, 128);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
        umma_issue_ss_pair(d_tmem, ad + (uint64_t)(kb * 2 * 128), bd + (uint64_t)(kb * 2 * NL), idesc, accumulate_first || kb > 0, leader);
}
template <int NL, int KB, int A_STRIDE = 8, typename T>
__device__ __forceinline__ void pmma_ts(uint32_t d_tmem, uint32_t a_tmem, const T* b, uint32_t ab_format, bool accumulate_first, bool leader) {
    const uint32_t idesc = umma_idesc_pair(2 * NL, ab_format);
    const uint64_t bd = umma_desc_uniform(b, NL * 16, 128);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
        umma_issue_ts_pair(d_tmem, a_tmem + kb * A_STRIDE, bd + (uint64_t)(kb * 2 * NL), idesc, accumulate_first || kb > 0, leader);
}

// ---------------------------------------------------------------------------------------------------------------
// Single-thread issue forms: called INSIDE `if (elect_one()) { ... }` by the one elected lane -- no per-instruction
// predicate, so ptxas emits the MMAs back to back from uniform registers (the lane-predicated forms above compile to an
// ELECT / BRA.U.ANY loop around every UTCHMMA, ~15 instructions per MMA: enough to make ONE issuing warp the bottleneck of
// a kernel whose MMAs take 57-128 clk each).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// cta_group::1 forms
__device__ __forceinline__ void umma1_commit(unsigned long long* bar) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(addr) : "memory");
}
__device__ __forceinline__ void umma1_issue_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma1_issue_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
template <int B_ROWS, int N, int KB, typename T>
__device__ __forceinline__ void umma1_ss(uint32_t d_tmem, const T* a, const T* b, uint32_t ab_format, bool accumulate_first) {
    const uint32_t idesc = umma_idesc(N, ab_format);
    const uint64_t ad = umma_desc(a, 128 * 16, 128), bd = umma_desc(b, B_ROWS * 16, 128);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
        umma1_issue_ss(d_tmem, ad + (uint64_t)(kb * 2 * 128), bd + (uint64_t)(kb * 2 * B_ROWS), idesc, (accumulate_first || kb > 0) ? 1u : 0u);
}
template <int B_ROWS, int N, int KB, int A_STRIDE = 8, typename T>
__device__ __forceinline__ void umma1_ts(uint32_t d_tmem, uint32_t a_tmem, const T* b, uint32_t ab_format, bool accumulate_first) {
    const uint32_t idesc = umma_idesc(N, ab_format);
    const uint64_t bd = umma_desc(b, B_ROWS * 16, 128);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
        umma1_issue_ts(d_tmem, a_tmem + kb * A_STRIDE, bd + (uint64_t)(kb * 2 * B_ROWS), idesc, (accumulate_first || kb > 0) ? 1u : 0u);
}
// cta_group::2 forms
__device__ __forceinline__ void umma1_commit_pair(unsigned long long* bar) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(addr), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma1_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma1_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
template <int NL, int KB, typename T>
__device__ __forceinline__ void pmma1_ss(uint32_t d_tmem, const T* a, const T* b, uint32_t ab_format, bool accumulate_first) {
    const uint32_t idesc = umma_idesc_pair(2 * NL, ab_format);
    const uint64_t ad = umma_desc(a, 128 * 16, 128), bd = umma_desc(b, NL * 16, 128);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
        umma1_ss_pair(d_tmem, ad + (uint64_t)(kb * 2 * 128), bd + (uint64_t)(kb * 2 * NL), idesc, (accumulate_first || kb > 0) ? 1u : 0u);
}
template <int NL, int KB, int A_STRIDE = 8, typename T>
__device__ __forceinline__ void pmma1_ts(uint32_t d_tmem, uint32_t a_tmem, const T* b, uint32_t ab_format, bool accumulate_first) {
    const uint32_t idesc = umma_idesc_pair(2 * NL, ab_format);
    const uint64_t bd = umma_desc(b, NL * 16, 128);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
        umma1_ts_pair(d_tmem, a_tmem + kb * A_STRIDE, bd + (uint64_t)(kb * 2 * NL), idesc, (accumulate_first || kb > 0) ? 1u : 0u);
}

// TMEM load / store of NCOLS consecutive 32-bit columns of the calling thread's lane (32x32b shape),
// decomposed into the power-of-two instruction widths.
template <int NCOLS>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* v) {
    if constexpr (NCOLS >= 32) {
        ptx::tcgen05_ld_32x32b(*reinterpret_cast<uint32_t(*)[32]>(v), taddr);
        tmem_ld<NCOLS - 32>(taddr + 32, v + 32);
    } else if constexpr (NCOLS >= 16) {
        ptx::tcgen05_ld_32x32b(*reinterpret_cast<uint32_t(*)[16]>(v), taddr);
        tmem_ld<NCOLS - 16>(taddr + 16, v + 16);
    } else if constexpr (NCOLS >= 8) {
        ptx::tcgen05_ld_32x32b(*reinterpret_cast<uint32_t(*)[8]>(v), taddr);
        tmem_ld<NCOLS - 8>(taddr + 8, v + 8);
    } else if constexpr (NCOLS >= 4) {
        ptx::tcgen05_ld_32x32b(*reinterpret_cast<uint32_t(*)[4]>(v), taddr);
        tmem_ld<NCOLS - 4>(taddr + 4, v + 4);
    } else if constexpr (NCOLS >= 2) {
        ptx::tcgen05_ld_32x32b(*reinterpret_cast<uint32_t(*)[2]>(v), taddr);
        tmem_ld<NCOLS - 2>(taddr + 2, v + 2);
    } else if constexpr (NCOLS == 1) {
        ptx::tcgen05_ld_32x32b(*reinterpret_cast<uint32_t(*)[1]>(v), taddr);
    }
}

template <int NCOLS>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t* v) {
    if constexpr (NCOLS >= 32) {
        ptx::tcgen05_st_32x32b(taddr, *reinterpret_cast<const uint32_t(*)[32]>(v));
        tmem_st<NCOLS - 32>(taddr + 32, v + 32);
    } else if constexpr (NCOLS >= 16) {
        ptx::tcgen05_st_32x32b(taddr, *reinterpret_cast<const uint32_t(*)[16]>(v));
        tmem_st<NCOLS - 16>(taddr + 16, v + 16);
    } else if constexpr (NCOLS >= 8) {
        ptx::tcgen05_st_32x32b(taddr, *reinterpret_cast<const uint32_t(*)[8]>(v));
        tmem_st<NCOLS - 8>(taddr + 8, v + 8);
    } else if constexpr (NCOLS >= 4) {
        ptx::tcgen05_st_32x32b(taddr, *reinterpret_cast<const uint32_t(*)[4]>(v));
        tmem_st<NCOLS - 4>(taddr + 4, v + 4);
    } else if constexpr (NCOLS >= 2) {
        ptx::tcgen05_st_32x32b(taddr, *reinterpret_cast<const uint32_t(*)[2]>(v));
        tmem_st<NCOLS - 2>(taddr + 2, v + 2);
    } else if constexpr (NCOLS == 1) {
        ptx::tcgen05_st_32x32b(taddr, *reinterpret_cast<const uint32_t(*)[1]>(v));
    }
}

// exp2-based activations: abs error ~5e-7 (ex2.approx 2^-22 rel, rcp.approx 1 ulp) at ~5 instructions
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }


// e^{-x} for the logistic forms below, argument clamped so that products of two (1 + e) terms stay finite
__device__ __forceinline__ float expneg_clamped(float x) { return ex2_approx(fminf(-1.4426950408889634f * x, 60.0f)); }

// LSTM cell update of TWO hidden units with SHARED reciprocals: the gate epilogue is bound by the MUFU pipe (16 lanes/clk/SM),
// so the four logistic denominators of a unit share ONE reciprocal, 1/(A.B.C.D), from which 1/A .. 1/D follow by
// multiplications, and the two tanh(c) of the unit pair share another: 10 ex2 + 3 rcp per 2 units instead of 10 + 10.
// Exponents are clamped at 2^30 so that the 4-fold product stays finite (sigma(x) floors at 2^-30 ~ 1e-9: far below
// fp32 resolution of the cell update).  c updated in place.
// PRE-SCALED pre-activations: e = (-log2 e . x_i, -log2 e . x_f, -2 log2 e . x_g, -log2 e . x_o), i.e. the ex2
// arguments themselves (the scale lives in the packed gate weights, packing.pack_decoder_tcx) -- 4 multiplies per unit less.
// Reciprocal on the FMA pipe: magic-constant seed (12 % off) + three Newton steps (relative error 6e-8).  The cell update of
// the pair decode kernel sits at the balance point between the MUFU pipe (16 lanes/clk/SM) and instruction issue: with
// SW_RCP_NEWTON = n (defined by the including kernel file) the first n of the 3 reciprocals per unit pair leave the MUFU pipe
// (7 FMA-pipe instructions each).  Measured on B200, pair decode alone: n = 0: 7.98 ms, 1: 7.88, 2: 7.93, 3: 8.02; the
// one-tile kernel and the encoder are issue-bound and lose 3 % with n = 1, so only decode_fwd_pair.cu sets it.
__device__ __forceinline__ float rcp_newton(float x) {
    float y = __int_as_float(0x7EF311C7 - __float_as_int(x));
#pragma unroll
    for (int it = 0; it < 3; ++it) y = fmaf(y, fmaf(-x, y, 1.0f), y);
    return y;
}
#ifndef SW_RCP_NEWTON
#define SW_RCP_NEWTON 0
#endif
__device__ __forceinline__ void lstm_cell_pair_prescaled(const float (&ea)[4], const float (&eb)[4], float& ca, float& cb,
                                                         float& ha, float& hb) {
    float cn[2];
    const float* es[2] = {ea, eb};
    const float cs[2] = {ca, cb};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const float* e = es[q];
        const float ai = 1.0f + ex2_approx(fminf(e[0], 30.0f)), af = 1.0f + ex2_approx(fminf(e[1], 30.0f));
        const float ag = 1.0f + ex2_approx(fminf(e[2], 30.0f)), ao = 1.0f + ex2_approx(fminf(e[3], 30.0f));
        const float p_ig = ai * ag, p_fo = af * ao;
        const float r = (q < SW_RCP_NEWTON) ? rcp_newton(p_ig * p_fo) : rcp_approx(p_ig * p_fo);
        const float r_ig = r * p_fo, r_fo = r * p_ig;            // 1/(ai.ag), 1/(af.ao)
        const float sig_i = ag * r_ig, tanh_g = fmaf(2.0f * ai, r_ig, -1.0f);
        cn[q] = fmaf(ao * r_fo, cs[q], sig_i * tanh_g);          // sigma(f) = ao / (af.ao)
        if (q == 0) ha = af * r_fo; else hb = af * r_fo;         // sigma(o), multiplied by tanh(c) below
    }
    const float a0 = 1.0f + expneg_clamped(2.0f * cn[0]), a1 = 1.0f + expneg_clamped(2.0f * cn[1]);
    const float r = (SW_RCP_NEWTON > 2) ? rcp_newton(a0 * a1) : rcp_approx(a0 * a1);
    ha *= fmaf(2.0f * a1, r, -1.0f);
    hb *= fmaf(2.0f * a0, r, -1.0f);
    ca = cn[0];
    cb = cn[1];
}

// Packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2 -- one issue slot and one register-file access for two IEEE fp32
// operations on a 64-bit register pair) and the prescaled cell with the two units of the pair carried as the two halves of
// fp32x2 values: the same formulas and operation order per unit (bit-identical to lstm_cell_pair_prescaled), ~46 instead of
// ~74 instructions per pair.  Used by the pair decode kernel, which is bound by instruction issue in bursts and by the board's
// power cap when it runs continuously (scripts/sustain_ab.sh: 7.40 -> 7.27 ms burst, 8.28 -> 8.18 ms sustained).  (Measured
// 3 % SLOWER in an earlier version of that kernel -- before its step loop shrank -- and removed then; DESIGN.md.)
struct f32x2 { unsigned long long v; };
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk2(f32x2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ void lstm_cell_pair_prescaled_x2(const float (&ea)[4], const float (&eb)[4], float& ca, float& cb,
                                                            float& ha, float& hb) {
    const f32x2 one = pk2(1.0f, 1.0f), mone = pk2(-1.0f, -1.0f);
    const f32x2 ai = add2(one, pk2(ex2_approx(fminf(ea[0], 30.0f)), ex2_approx(fminf(eb[0], 30.0f))));
    const f32x2 af = add2(one, pk2(ex2_approx(fminf(ea[1], 30.0f)), ex2_approx(fminf(eb[1], 30.0f))));
    const f32x2 ag = add2(one, pk2(ex2_approx(fminf(ea[2], 30.0f)), ex2_approx(fminf(eb[2], 30.0f))));
    const f32x2 ao = add2(one, pk2(ex2_approx(fminf(ea[3], 30.0f)), ex2_approx(fminf(eb[3], 30.0f))));
    const f32x2 p_ig = mul2(ai, ag), p_fo = mul2(af, ao);
    const f32x2 pp = mul2(p_ig, p_fo);
    float ppa, ppb;
    unpk2(pp, ppa, ppb);
    const f32x2 r = pk2((SW_RCP_NEWTON > 0) ? rcp_newton(ppa) : rcp_approx(ppa), (SW_RCP_NEWTON > 1) ? rcp_newton(ppb) : rcp_approx(ppb));
    const f32x2 r_ig = mul2(r, p_fo), r_fo = mul2(r, p_ig);
    const f32x2 sig_i = mul2(ag, r_ig), tanh_g = fma2(add2(ai, ai), r_ig, mone);
    const f32x2 cn = fma2(mul2(ao, r_fo), pk2(ca, cb), mul2(sig_i, tanh_g));
    const f32x2 so = mul2(af, r_fo);
    // tanh(c) = 2 / (1 + 2^(-2 log2e c)) - 1, the pair sharing one reciprocal; -2 log2e . c == -log2e . (2 c) bit for bit
    // (doubling is exact), so the argument and the "1 +" are one packed operation each
    float eca, ecb;
    unpk2(mul2(cn, pk2(-2.8853900817779268f, -2.8853900817779268f)), eca, ecb);
    float a0, a1;
    unpk2(add2(one, pk2(ex2_approx(fminf(eca, 60.0f)), ex2_approx(fminf(ecb, 60.0f)))), a0, a1);
    const float rr = rcp_approx(a0 * a1);
    const f32x2 th = pk2(fmaf(2.0f * a1, rr, -1.0f), fmaf(2.0f * a0, rr, -1.0f));
    unpk2(mul2(so, th), ha, hb);
    unpk2(cn, ca, cb);
}

// bf16-mode cell: hardware tanh (MUFU.TANH, rel. error 2^-11, below the bf16 operand rounding), 5 MUFU per unit
__device__ __forceinline__ float tanh_hw(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void lstm_cell_hw(const float (&g)[4], float& c, float& h) {
    const float si = fmaf(0.5f, tanh_hw(0.5f * g[0]), 0.5f), sf = fmaf(0.5f, tanh_hw(0.5f * g[1]), 0.5f);
    const float so = fmaf(0.5f, tanh_hw(0.5f * g[3]), 0.5f);
    c = fmaf(sf, c, si * tanh_hw(g[2]));
    h = so * tanh_hw(c);
}

}  // namespace sw
