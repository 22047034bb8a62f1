// Probe (bring-up aid, not product code): tcgen05.mma with cta_group::2 (a CTA pair, UMMA M = 256).
//  part A  correctness: D = A.B^T for small-integer fp16 operands, A from shared memory (SS) and from TMEM (TS), operand-ready
//          signalling through remote mbarrier arrives on the leader CTA -- the protocol decode_fwd_pair uses
//  part B  cost: cycles per MMA of a back-to-back sequence, cta_group::1 and ::2, SS and TS, N = 64 .. 256
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pair_mma_probe.bin pair_mma_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(const void* p, uint32_t lbo, uint32_t sbo) {
    const uint32_t a = smem_u32(p);
    return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(void* bar, uint32_t cta) {   // arrive on `bar` of CTA `cta` of the cluster
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(ra) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(void* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spins > (1u << 22)) __trap();
    }
}
template <int CG>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, bool acc) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, bool acc) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(void* bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t cols) {
    if (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
    else         asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
#define FENCE_BEFORE() asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory")
#define FENCE_AFTER() asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")
#define FENCE_ASYNC_SMEM() asm volatile("fence.proxy.async.shared::cta;" ::: "memory")

__host__ __device__ inline float a_val(int r, int k) { return (float)((r * 3 + k * 5) % 7 - 3); }
__host__ __device__ inline float b_val(int n, int k) { return (float)((n * 2 + k) % 5 - 2); }

constexpr int KTOT = 32;        // 2 K blocks

struct ProbeSmem {
    __half a[KTOT / 8 * 128 * 8];        // [K/8][128][8]
    __half b[KTOT / 8 * 128 * 8];        // [K/8][N/2 <= 128][8]
    unsigned long long full, ready;
    uint32_t tmem_base;
};

// part A: one CTA pair, 128 threads per CTA.  mode 0 = SS, 1 = TS.  out [256][N]
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_correct(float* out, int mode) {
    extern __shared__ __align__(1024) unsigned char raw[];
    ProbeSmem& s = *reinterpret_cast<ProbeSmem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t cta = cluster_ctarank();
    constexpr int NH = N / 2;
    for (int i = tid; i < KTOT / 8 * 128 * 8; i += 128) {
        const int kc = i / (128 * 8), r = (i / 8) % 128, e = i % 8;
        s.a[i] = __float2half(a_val(cta * 128 + r, kc * 8 + e));
    }
    for (int i = tid; i < KTOT / 8 * NH * 8; i += 128) {
        const int kc = i / (NH * 8), n = (i / 8) % NH, e = i % 8;
        s.b[i] = __float2half(b_val(cta * NH + n, kc * 8 + e));
    }
    if (warp == 0) tmem_alloc<2>(&s.tmem_base, 512);
    if (tid == 0) {
        mbar_init(&s.full, 1);
        mbar_init(&s.ready, 2 * 4);            // one arrive per warp of both CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    FENCE_ASYNC_SMEM();
    FENCE_BEFORE();
    __syncthreads();
    cluster_sync();
    FENCE_AFTER();
    const uint32_t tm = s.tmem_base;
    const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
    if (mode == 1) {        // A rows of this CTA -> TMEM columns [256, 272): K block kb at +8 kb, column j = (k = 2j, 2j+1)
        for (int kb = 0; kb < 2; ++kb) {
            uint32_t v[8];
            for (int j = 0; j < 8; ++j) {
                const __half2 h2 = __floats2half2_rn(a_val(cta * 128 + warp * 32 + lane, kb * 16 + 2 * j), a_val(cta * 128 + warp * 32 + lane, kb * 16 + 2 * j + 1));
                v[j] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            tmem_st8(tl + 256 + kb * 8, v);
        }
    }
    FENCE_ASYNC_SMEM();
    FENCE_BEFORE();
    __syncwarp();
    if (lane == 0) mbar_arrive_remote(&s.ready, 0);
    if (cta == 0 && warp == 0) {
        mbar_wait_cluster(&s.ready, 0);
        FENCE_AFTER();
        if (lane == 0) {
            const uint32_t idesc = idesc_f16(256, N);
            for (int kb = 0; kb < 2; ++kb) {
                const uint64_t bd = umma_desc(s.b + kb * 2 * NH * 8, NH * 16, 128);
                if (mode == 0) mma_ss<2>(tm, umma_desc(s.a + kb * 2 * 128 * 8, 128 * 16, 128), bd, idesc, kb > 0);
                else           mma_ts<2>(tm, tm + 256 + kb * 8, bd, idesc, kb > 0);
            }
            commit<2>(&s.full);
        }
        __syncwarp();
    }
    mbar_wait_cluster(&s.full, 0);
    FENCE_AFTER();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t d[16];
        tmem_ld16(tl + c0, d);
        for (int j = 0; j < 16; ++j) out[(size_t)(cta * 128 + warp * 32 + lane) * N + c0 + j] = __uint_as_float(d[j]);
    }
    FENCE_BEFORE();
    __syncthreads();
    cluster_sync();
    if (warp == 0) tmem_dealloc<2>(tm, 512);
}

// part B: R back-to-back MMAs (K = 16 each, accumulating into the same D), cycles from first issue to completion
struct CostSmem {
    __half a[2 * 128 * 8];
    __half b[2 * 256 * 8];
    unsigned long long full;
    uint32_t tmem_base;
};
template <int CG>
__global__ void __launch_bounds__(128, 1) mma_cost(long long* cycles, int n, int ts, int reps, int d_stride) {
    extern __shared__ __align__(1024) unsigned char raw[];
    CostSmem& s = *reinterpret_cast<CostSmem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t cta = 0;
    if (CG == 2) cta = cluster_ctarank();
    for (int i = tid; i < 2 * 128 * 8; i += 128) s.a[i] = __float2half(0.5f);
    for (int i = tid; i < 2 * 256 * 8; i += 128) s.b[i] = __float2half(0.25f);
    if (warp == 0) tmem_alloc<CG>(&s.tmem_base, 512);
    if (tid == 0) { mbar_init(&s.full, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    FENCE_ASYNC_SMEM();
    FENCE_BEFORE();
    __syncthreads();
    if (CG == 2) cluster_sync();
    FENCE_AFTER();
    const uint32_t tm = s.tmem_base;
    const int nh = n / CG;
    for (int pass = 0; pass < 3; ++pass) {     // pass 0 warms up
        if (cta == 0 && tid == 0) {
            const uint32_t idesc = idesc_f16(128 * CG, n);
            const uint64_t ad = umma_desc(s.a, 128 * 16, 128), bd = umma_desc(s.b, nh * 16, 128);
            const long long t0 = clock64();
            for (int i = 0; i < reps; ++i) {
                const uint32_t d = tm + (uint32_t)((i & 1) * d_stride);      // d_stride = 0: one accumulator; 256: two alternating
                if (ts) mma_ts<CG>(d, tm + 480, bd, idesc, i > 1);
                else    mma_ss<CG>(d, ad, bd, idesc, i > 1);
            }
            const long long t1 = clock64();
            commit<CG>(&s.full);
            mbar_wait_cluster(&s.full, pass & 1);
            const long long t2 = clock64();
            if (pass > 0) { cycles[(pass - 1) * 2] = t1 - t0; cycles[(pass - 1) * 2 + 1] = t2 - t0; }
        } else {
            mbar_wait_cluster(&s.full, pass & 1);
        }
        __syncthreads();
    }
    FENCE_BEFORE();
    __syncthreads();
    if (CG == 2) cluster_sync();
    if (warp == 0) tmem_dealloc<CG>(tm, 512);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int N>
static int run_correct(int mode) {
    float* d_out;
    CK(cudaMalloc(&d_out, 256 * N * sizeof(float)));
    CK(cudaMemset(d_out, 0xff, 256 * N * sizeof(float)));
    CK(cudaFuncSetAttribute(pair_correct<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ProbeSmem)));
    pair_correct<N><<<2, 128, sizeof(ProbeSmem)>>>(d_out, mode);
    CK(cudaDeviceSynchronize());
    std::vector<float> h(256 * N);
    CK(cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r = 0; r < 256; ++r)
        for (int n = 0; n < N; ++n) {
            float want = 0;
            for (int k = 0; k < KTOT; ++k) want += a_val(r, k) * b_val(n, k);
            if (h[r * N + n] != want) { if (bad < 5) printf("  mismatch r=%d n=%d got %g want %g\n", r, n, h[r * N + n], want); ++bad; }
        }
    printf("pair MMA %s  M=256 N=%3d : %s (%d mismatches)\n", mode ? "TS" : "SS", N, bad ? "FAIL" : "ok", bad);
    cudaFree(d_out);
    return 0;
}

int main() {
    for (int mode = 0; mode < 2; ++mode) {
        if (run_correct<80>(mode)) return 1;
        if (run_correct<96>(mode)) return 1;
        if (run_correct<160>(mode)) return 1;
        if (run_correct<256>(mode)) return 1;
    }
    long long* d_cyc;
    CK(cudaMalloc(&d_cyc, 4 * sizeof(long long)));
    CK(cudaFuncSetAttribute(mma_cost<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CostSmem)));
    CK(cudaFuncSetAttribute(mma_cost<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CostSmem)));
    const int ns[] = {64, 80, 96, 128, 160, 256};
    printf("cycles per MMA (K = 16): issue-only | issue+complete, reps 8 and 64, one accumulator (d0) or two alternating (d256)\n");
    for (int cg = 1; cg <= 2; ++cg)
        for (int ts = 0; ts < 2; ++ts)
            for (int n : ns)
                for (int ds = 0; ds <= 256; ds += 256) {
                    if (ds && n > 240) continue;
                    long long h8[4], h64[4];
                    for (int reps : {8, 64}) {
                        if (cg == 1) mma_cost<1><<<1, 128, sizeof(CostSmem)>>>(d_cyc, n, ts, reps, ds);
                        else {
                            cudaLaunchConfig_t cfg = {};
                            cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = sizeof(CostSmem);
                            cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
                            cfg.attrs = &at; cfg.numAttrs = 1;
                            CK(cudaLaunchKernelEx(&cfg, mma_cost<2>, d_cyc, n, ts, reps, ds));
                        }
                        CK(cudaDeviceSynchronize());
                        CK(cudaMemcpy(reps == 8 ? h8 : h64, d_cyc, 4 * sizeof(long long), cudaMemcpyDeviceToHost));
                    }
                    printf("cta_group::%d %s N=%3d d%-3d : reps8 issue %5lld total %5lld | reps64 issue %6lld total %6lld -> %.1f clk/MMA (slope), floor N/2 = %d\n",
                           cg, ts ? "TS" : "SS", n, ds, h8[2], h8[3], h64[2], h64[3], (double)(h64[3] - h8[3]) / 56.0, n / 2);
                }
    cudaFree(d_cyc);
    return 0;
}
