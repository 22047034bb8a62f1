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

__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    ptx::tcgen05_commit(ptx::cta_group_1, reinterpret_cast<uint64_t*>(bar));
}

// D[128 x n] (+)= A[128 x 16*kblocks] . B[n x 16*kblocks]^T, both operands 16-bit in shared memory.
// `a`: [K/8][128][8]; `b`: [K/8][b_rows][8] (the MMA uses n consecutive rows starting at `b`).
template <typename T>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, const T* a, const T* b, int b_rows, int n, int kblocks,
                                        uint32_t ab_format, bool accumulate_first) {
    const uint32_t idesc = umma_idesc(n, ab_format);
    for (int kb = 0; kb < kblocks; ++kb) {
        const uint64_t ad = umma_desc(a + (size_t)kb * 2 * 128 * 8, 128 * 16, 128);
        const uint64_t bd = umma_desc(b + (size_t)kb * 2 * b_rows * 8, b_rows * 16, 128);
        ptx::tcgen05_mma(ptx::kind_f16, ptx::cta_group_1, d_tmem, ad, bd, idesc, accumulate_first || kb > 0);
    }
}

// same with the A operand in TMEM (16-bit, 8 columns per K block)
template <typename T>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, const T* b, int b_rows, int n, int kblocks,
                                        uint32_t ab_format, bool accumulate_first) {
    const uint32_t idesc = umma_idesc(n, ab_format);
    for (int kb = 0; kb < kblocks; ++kb) {
        const uint64_t bd = umma_desc(b + (size_t)kb * 2 * b_rows * 8, b_rows * 16, 128);
        ptx::tcgen05_mma_tmem_a(ptx::kind_f16, ptx::cta_group_1, d_tmem, a_tmem + kb * 8, bd, idesc, accumulate_first || kb > 0);
    }
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
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.0f, rcp_approx(1.0f + ex2_approx(-2.8853900817779268f * x)), -1.0f); }

}  // namespace sw
