// Probe (bring-up aid, not product code): cost per tcgen05.mma when issued the way the product kernels issue them
// (sw_umma.cuh helpers: warp-uniform descriptors, predicated issue, fully unrolled K-block loops), single CTA and CTA pair.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I socialways_b200/csrc -o mma_rate_probe.bin mma_rate_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include "sw_umma.cuh"
using namespace sw;

struct Smem {
    __half a[8 * 128 * 8];        // K = 64
    __half b[20 * 256 * 8];       // up to K = 160, 256 rows
    unsigned long long full;
    uint32_t tmem_base;
};

// mode: 0 = SS N=128 K=64 (x3 products), 1 = SS N=160, 2 = SS N=256, 3 = TS N=80 K=160 (x3), 4 = TS N=160 K=160, 5 = TS N=96 K=160 (x3)
//       6 = SS N=128 but A-lo products use same A (tests A re-read), 7 = TS N=256 K=64 x3
template <int PAIR>
__global__ void __launch_bounds__(128, 1) rate(long long* out, int mode, int reps) {
    extern __shared__ __align__(1024) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t cta = 0;
    if (PAIR) cta = cluster_ctarank();
    for (int i = tid; i < 8 * 128 * 8; i += 128) s.a[i] = __float2half(0.01f);
    for (int i = tid; i < 20 * 256 * 8; i += 128) s.b[i] = __float2half(0.02f);
    if (warp == 0) {
        if (PAIR) { asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&s.tmem_base)), "r"(512u) : "memory");
                    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
        else { ptx::tcgen05_alloc(ptx::cta_group_1, &s.tmem_base, 512u); ptx::tcgen05_relinquish_alloc_permit(ptx::cta_group_1); }
    }
    if (tid == 0) { ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.full), 1); ptx::fence_mbarrier_init(ptx::sem_release, ptx::scope_cluster); }
    ptx::fence_proxy_async(ptx::space_shared);
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    ptx::tcgen05_fence_after_thread_sync();
    const uint32_t tm = __shfl_sync(0xffffffffu, s.tmem_base, 0);
    const bool leader = lane == 0;
    uint32_t ph = 0;
    for (int pass = 0; pass < 3; ++pass) {
        long long t0 = 0, t1 = 0;
        if (warp == 0 && cta == 0) {
            t0 = clock64();
            for (int i = 0; i < reps; ++i) {
                if (PAIR) {
                    switch (mode) {
                    case 0: for (int p = 0; p < 3; ++p) pmma_ss<64, 4>(tm, s.a, s.b, 0, p > 0, leader); break;
                    case 1: for (int p = 0; p < 3; ++p) pmma_ss<80, 4>(tm, s.a, s.b, 0, p > 0, leader); break;
                    case 2: for (int p = 0; p < 3; ++p) pmma_ss<128, 4>(tm, s.a, s.b, 0, p > 0, leader); break;
                    case 3: for (int p = 0; p < 3; ++p) pmma_ts<40, 10, 16>(tm + 256, tm, s.b, 0, p > 0, leader); break;
                    case 4: for (int p = 0; p < 3; ++p) pmma_ts<80, 10, 16>(tm + 256, tm, s.b, 0, p > 0, leader); break;
                    case 5: for (int p = 0; p < 3; ++p) pmma_ts<48, 10, 16>(tm + 256, tm, s.b, 0, p > 0, leader); break;
                    case 7: for (int p = 0; p < 3; ++p) pmma_ts<128, 4, 8>(tm + 256, tm, s.b, 0, p > 0, leader); break;
                    }
                } else {
                    switch (mode) {
                    case 0: for (int p = 0; p < 3; ++p) umma_ss<128, 128, 4>(tm, s.a, s.b, 0, p > 0, leader); break;
                    case 1: for (int p = 0; p < 3; ++p) umma_ss<160, 160, 4>(tm, s.a, s.b, 0, p > 0, leader); break;
                    case 2: for (int p = 0; p < 3; ++p) umma_ss<256, 256, 4>(tm, s.a, s.b, 0, p > 0, leader); break;
                    case 3: for (int p = 0; p < 3; ++p) umma_ts<80, 80, 10, 16>(tm + 256, tm, s.b, 0, p > 0, leader); break;
                    case 4: for (int p = 0; p < 3; ++p) umma_ts<160, 160, 10, 16>(tm + 256, tm, s.b, 0, p > 0, leader); break;
                    case 5: for (int p = 0; p < 3; ++p) umma_ts<96, 96, 10, 16>(tm + 256, tm, s.b, 0, p > 0, leader); break;
                    case 7: for (int p = 0; p < 3; ++p) umma_ts<256, 256, 4, 8>(tm + 256, tm, s.b, 0, p > 0, leader); break;
                    }
                }
            }
            t1 = clock64();
            if (PAIR) umma_commit_pair(&s.full, leader); else umma_commit(&s.full, leader);
        }
        mbar_wait_cluster(&s.full, ph); ph ^= 1;
        const long long t2 = clock64();
        if (warp == 0 && cta == 0 && lane == 0 && pass > 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        __syncthreads();
    }
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    if (warp == 0) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512u) : "memory");
        else ptx::tcgen05_dealloc(ptx::cta_group_1, tm, 512u);
    }
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    cudaFuncSetAttribute(rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    const char* names[] = {"SS N=128 K=64 x3 (12 MMAs)", "SS N=160 K=64 x3 (12)", "SS N=256 K=64 x3 (12)", "TS N=80 K=160 x3 (30)",
                           "TS N=160 K=160 x3 (30)", "TS N=96 K=160 x3 (30)", "", "TS N=256 K=64 x3 (12)"};
    const int nmma[] = {12, 12, 12, 30, 30, 30, 0, 12};
    for (int pair = 0; pair < 2; ++pair)
        for (int mode : {0, 1, 2, 3, 4, 5, 7}) {
            long long h4[2], h16[2];
            for (int reps : {4, 16}) {
                if (!pair) rate<0><<<1, 128, sizeof(Smem)>>>(d, mode, reps);
                else {
                    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = sizeof(Smem);
                    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
                    cfg.attrs = &at; cfg.numAttrs = 1;
                    cudaLaunchKernelEx(&cfg, rate<1>, d, mode, reps);
                }
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(reps == 4 ? h4 : h16, d, 16, cudaMemcpyDeviceToHost);
            }
            printf("%s %-28s: issue %.1f clk/MMA, complete %.1f clk/MMA (slope over reps 4 -> 16)\n", pair ? "pair  " : "single", names[mode],
                   (double)(h16[0] - h4[0]) / (12.0 * nmma[mode]), (double)(h16[1] - h4[1]) / (12.0 * nmma[mode]));
        }
    return 0;
}
