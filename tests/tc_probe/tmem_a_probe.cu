// Probe (bring-up aid, not product code): layout of a 16-bit A operand held in TMEM for
// tcgen05.mma [d], [a_tmem], b_desc.  B is an identity selector so D reveals A's element order.
#include <cuda/ptx>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
namespace ptx = cuda::ptx;

__device__ uint64_t umma_desc(const void* p, uint32_t lbo, uint32_t sbo) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__global__ void probe(float* out /*[128][16]*/) {
    __shared__ __align__(128) __half b[4 * 16 * 8];       // canonical [K/8 = 4][N = 16][8]
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 4 * 16 * 8; i += blockDim.x) {
        const int kc = i / 128, n = (i / 8) % 16, e = i % 8, k = kc * 8 + e;
        b[i] = __float2half((k - 16 == n) ? 1.0f : 0.0f);   // D[r][n] = A[r][16 + n]
    }
    if (warp == 0) { ptx::tcgen05_alloc(ptx::cta_group_1, &tmem_base, 64u); ptx::tcgen05_relinquish_alloc_permit(ptx::cta_group_1); }
    if (tid == 0) { ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&bar), 1); ptx::fence_mbarrier_init(ptx::sem_release, ptx::scope_cluster); }
    ptx::fence_proxy_async(ptx::space_shared);
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    ptx::tcgen05_fence_after_thread_sync();
    const uint32_t tm = tmem_base;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t tl = tm + ((uint32_t)((warp & 3) * 32) << 16);
    // A[r][k] = r + k/64 for k in [0, 32): 16 columns at [32, 48), column j = (k = 2j low half, k = 2j+1 high half)
    uint32_t v[16];
    for (int j = 0; j < 16; ++j) {
        const __half2 h2 = __floats2half2_rn((float)r + (2 * j) / 64.0f, (float)r + (2 * j + 1) / 64.0f);
        v[j] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    ptx::tcgen05_st_32x32b(tl + 32, v);
    ptx::tcgen05_wait_st();
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    if (tid == 0) {
        ptx::tcgen05_fence_after_thread_sync();
        const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);   // F16 x F16 -> F32
        for (int kb = 0; kb < 2; ++kb)
            ptx::tcgen05_mma_tmem_a(ptx::kind_f16, ptx::cta_group_1, tm + 0, tm + 32 + kb * 8, umma_desc(b + kb * 2 * 16 * 8, 16 * 16, 128),
                                    idesc, kb > 0);
        ptx::tcgen05_commit(ptx::cta_group_1, reinterpret_cast<uint64_t*>(&bar));
    }
    for (uint32_t s = 0; !ptx::mbarrier_try_wait_parity(reinterpret_cast<uint64_t*>(&bar), 0u); ++s) if (s > (1u << 22)) __trap();
    ptx::tcgen05_fence_after_thread_sync();
    uint32_t d[16];
    ptx::tcgen05_ld_32x32b(d, tl + 0);
    ptx::tcgen05_wait_ld();
    for (int n = 0; n < 16; ++n) out[r * 16 + n] = __uint_as_float(d[n]);
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) ptx::tcgen05_dealloc(ptx::cta_group_1, tm, 64u);
}

int main() {
    float* d; cudaMalloc(&d, 128 * 16 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    float h[128 * 16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int ok = 1;
    for (int r : {0, 1, 37, 127}) {
        printf("row %3d:", r);
        for (int n = 0; n < 16; ++n) { printf(" %.4f", h[r * 16 + n]); if (fabsf(h[r * 16 + n] - (r + (16 + n) / 64.0f)) > 1e-2f) ok = 0; }
        printf("\n");
    }
    printf("expected D[r][n] = r + (16+n)/64 : %s\n", ok ? "MATCH (k even in low half, +8 columns per K block)" : "MISMATCH");
    return 0;
}
