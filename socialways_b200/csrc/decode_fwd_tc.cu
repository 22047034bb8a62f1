// Tensor-core (tcgen05 / TMEM) variant of the K-sample decode kernel (see decode_fwd.cu for the
// reference mapping: the loop of predict(), train.py:418-430, for every (sample, agent) row).
//
// One CTA = one 128-row tile at a time (UMMA M = 128, cta_group::1), 512 threads.  Every dense layer
// of a decode step is a tcgen05.mma with BF16 operands in shared memory and an FP32 accumulator in
// TMEM (one row per TMEM lane):
//     L1    [128 x 160] = [h ; S ; z] (K = 160) . W1^T          -> +b1, LeakyReLU -> A1 (bf16, smem)
//     L2    [128 x  80] = A1 (K = 160) . W2^T                    -> +b2, LeakyReLU -> A2 (bf16, smem)
//     L34   [128 x  16] = A2 (K = 80) . W34^T (2 real columns)   -> +b34 = velocity; integrate; emit
//     gates [128 x 256] = h (K = 64) . Whh^T                     -> + Wx.x4 (fp32 FMA) + b -> LSTM cell
// Operands use the canonical K-major, no-swizzle UMMA layout: 8-row x 16-byte core matrices,
// stored [K/8][rows][8 bf16]; consecutive 8-row groups are 128 B apart (SBO), consecutive K chunks
// rows*16 B apart (LBO).  With one thread per row the epilogue writes whole 16-byte K chunks, 32
// lanes contiguous -> conflict-free, and the positions (p, v) enter the LSTM in fp32 (never rounded
// to bf16).  The gates MMA of a step is issued together with L1 (both only need h), so it runs under
// the L1/L2/L34 epilogues.  Thread t: row = 32*(warp%4)+lane (its TMEM lane), column quarter = warp/4.
//
// Precision: bf16 operands, fp32 accumulate -- the "fast" mode (BASELINE.json config 3 asks for bf16);
// it does NOT meet the 1e-4 fp32 parity bar, which stays with the FFMA kernel (decode_fwd.cu).
#include <cuda_bf16.h>

#include "sw_common.cuh"
#include "sw_umma.cuh"

namespace sw {

constexpr int TC_ROWS = 128;
constexpr int TC_THREADS = 512;
// bf16 weight section (elements): canonical [K/8][N][8]
constexpr int TW_W1 = 0, TW_W2 = TW_W1 + 160 * 160, TW_W34 = TW_W2 + 160 * 80, TW_WHH = TW_W34 + 80 * 16,
              TW_TOTAL = TW_WHH + 64 * 256;
// fp32 section (floats): Wx[4][256] | bL[256] | b1[160] | b2[80] | b34[2] | pad
constexpr int TF_WX = 0, TF_BL = 1024, TF_B1 = 1280, TF_B2 = 1440, TF_B34 = 1520, TF_TOTAL = 1536;
// TMEM columns
constexpr uint32_t COL_G = 0, COL_L1 = 256, COL_L2 = 416, COL_V = 496;

struct TcSmem {
    __nv_bfloat16 w[TW_TOTAL];                 // 112 128 B
    __nv_bfloat16 al1[20 * TC_ROWS * 8];       // [h(8 chunks) ; S(8) ; z(4)] x 128 rows x 8
    __nv_bfloat16 a12[20 * TC_ROWS * 8];       // A1 (20 chunks); A2 re-uses the first 10 chunks
    float f32[TF_TOTAL];
    float x4[4 * TC_ROWS];
    unsigned long long bar[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
decode_fwd_tc_kernel(const __nv_bfloat16* __restrict__ w16, const float* __restrict__ wf32,
                     const float* __restrict__ h0, const float* __restrict__ c0, const float* __restrict__ pooled,
                     const float* __restrict__ noise, const float* __restrict__ x_last, float* __restrict__ out,
                     int n_agents, long long n_rows, int n_next, int n_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TcSmem& s = *reinterpret_cast<TcSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lq = warp & 3, cq = warp >> 2;
    const int r = lq * 32 + lane;                          // row of the tile == TMEM lane
    const bool leader = lane == 0;                         // lane of warp 0 that issues the MMAs

    // ---- one-time setup: weights -> smem, TMEM allocation, mbarriers ----
    for (int i = tid * 8; i < TW_TOTAL; i += TC_THREADS * 8)
        *reinterpret_cast<uint4*>(s.w + i) = __ldg(reinterpret_cast<const uint4*>(w16 + i));
    for (int i = tid; i < TF_TOTAL; i += TC_THREADS) s.f32[i] = __ldg(wf32 + i);
    if (warp == 0) {
        ptx::tcgen05_alloc(ptx::cta_group_1, &s.tmem_base, 512u);
        ptx::tcgen05_relinquish_alloc_permit(ptx::cta_group_1);
    }
    if (tid == 0) {
        ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.bar[0]), 1);
        ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.bar[1]), 1);
        ptx::fence_mbarrier_init(ptx::sem_release, ptx::scope_cluster);
    }
    ptx::fence_proxy_async(ptx::space_shared);
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    ptx::tcgen05_fence_after_thread_sync();
    const uint32_t tmem = s.tmem_base;
    const uint32_t tlane = tmem + ((uint32_t)(lq * 32) << 16);
    uint32_t ph0 = 0, ph1 = 0;
    const float* bL = s.f32 + TF_BL;
    const float* wx = s.f32 + TF_WX;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = (long long)tile * TC_ROWS;
        const bool valid = row0 + r < n_rows;
        const int agent = valid ? (int)((row0 + r) % n_agents) : 0;
        // ---- tile prologue: h0 / S / z -> bf16 operand chunks; c0, p -> registers ----
        {
            float v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 t = valid ? __ldg(reinterpret_cast<const float4*>(h0 + (size_t)agent * SW_H + cq * 16) + q)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
                v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j)
                *reinterpret_cast<uint4*>(s.al1 + ((size_t)(cq * 2 + j) * TC_ROWS + r) * 8) =
                    make_uint4(pack_bf16(v[j * 8], v[j * 8 + 1]), pack_bf16(v[j * 8 + 2], v[j * 8 + 3]),
                               pack_bf16(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16(v[j * 8 + 6], v[j * 8 + 7]));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 t = (valid && pooled) ? __ldg(reinterpret_cast<const float4*>(pooled + (size_t)agent * SW_H + cq * 16) + q)
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j)
                *reinterpret_cast<uint4*>(s.al1 + ((size_t)(8 + cq * 2 + j) * TC_ROWS + r) * 8) =
                    make_uint4(pack_bf16(v[j * 8], v[j * 8 + 1]), pack_bf16(v[j * 8 + 2], v[j * 8 + 3]),
                               pack_bf16(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16(v[j * 8 + 6], v[j * 8 + 7]));
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float4 t = valid ? __ldg(reinterpret_cast<const float4*>(noise + (size_t)(row0 + r) * SW_Z + cq * 8) + q)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
                v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
            }
            *reinterpret_cast<uint4*>(s.al1 + ((size_t)(16 + cq) * TC_ROWS + r) * 8) =
                make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
        float c[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 t = valid ? __ldg(reinterpret_cast<const float4*>(c0 + (size_t)agent * SW_H + cq * 16) + q)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            c[q * 4] = t.x; c[q * 4 + 1] = t.y; c[q * 4 + 2] = t.z; c[q * 4 + 3] = t.w;
        }
        float p0 = 0.f, p1 = 0.f;
        if (cq == 0 && valid) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(x_last + (size_t)agent * 4));
            p0 = t.x; p1 = t.y;
        }
        ptx::fence_proxy_async(ptx::space_shared);
        ptx::tcgen05_fence_before_thread_sync();
        __syncthreads();

        for (int t = 0; t < n_next; ++t) {
            const bool feed_back = t + 1 < n_next;
            if (warp == 0) {
                ptx::tcgen05_fence_after_thread_sync();
                umma_ss<160, 160, 10>(tmem + COL_L1, s.al1, s.w + TW_W1, 1u, false, leader);
                umma_commit(&s.bar[0], leader);
                if (feed_back) {
                    umma_ss<256, 256, 4>(tmem + COL_G, s.al1, s.w + TW_WHH, 1u, false, leader);
                    umma_commit(&s.bar[1], leader);
                }
            }
            // ---- L1 epilogue: +b1, LeakyReLU -> A1 (bf16) ----
            mbar_wait(&s.bar[0], ph0); ph0 ^= 1;
            ptx::tcgen05_fence_after_thread_sync();
            {
                uint32_t a[40];
                uint32_t (&a32)[32] = *reinterpret_cast<uint32_t(*)[32]>(&a[0]);
                uint32_t (&a8)[8] = *reinterpret_cast<uint32_t(*)[8]>(&a[32]);
                ptx::tcgen05_ld_32x32b(a32, tlane + COL_L1 + cq * 40);
                ptx::tcgen05_ld_32x32b(a8, tlane + COL_L1 + cq * 40 + 32);
                ptx::tcgen05_wait_ld();
                const float* b1 = s.f32 + TF_B1 + cq * 40;
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    float y[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] = lrelu02(__uint_as_float(a[j * 8 + e]) + b1[j * 8 + e]);
                    *reinterpret_cast<uint4*>(s.a12 + ((size_t)(cq * 5 + j) * TC_ROWS + r) * 8) =
                        make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
                }
            }
            ptx::fence_proxy_async(ptx::space_shared);
            ptx::tcgen05_fence_before_thread_sync();
            __syncthreads();
            if (warp == 0) {
                ptx::tcgen05_fence_after_thread_sync();
                umma_ss<80, 80, 10>(tmem + COL_L2, s.a12, s.w + TW_W2, 1u, false, leader);
                umma_commit(&s.bar[0], leader);
            }
            // ---- L2 epilogue: +b2, LeakyReLU -> A2 (bf16, over the first 10 chunks of A1) ----
            mbar_wait(&s.bar[0], ph0); ph0 ^= 1;
            ptx::tcgen05_fence_after_thread_sync();
            {
                const int n0 = (cq < 2) ? cq * 24 : 48 + (cq - 2) * 16;   // column split 24 | 24 | 16 | 16
                const int nch = (cq < 2) ? 3 : 2;
                uint32_t a[24];
                uint32_t (&a0)[8] = *reinterpret_cast<uint32_t(*)[8]>(&a[0]);
                uint32_t (&a1)[8] = *reinterpret_cast<uint32_t(*)[8]>(&a[8]);
                uint32_t (&a2)[8] = *reinterpret_cast<uint32_t(*)[8]>(&a[16]);
                ptx::tcgen05_ld_32x32b(a0, tlane + COL_L2 + n0);
                ptx::tcgen05_ld_32x32b(a1, tlane + COL_L2 + n0 + 8);
                if (nch == 3) ptx::tcgen05_ld_32x32b(a2, tlane + COL_L2 + n0 + 16);   // warp-uniform branch
                ptx::tcgen05_wait_ld();
                const float* b2 = s.f32 + TF_B2 + n0;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (j < nch) {
                        float y[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) y[e] = lrelu02(__uint_as_float(a[j * 8 + e]) + b2[j * 8 + e]);
                        *reinterpret_cast<uint4*>(s.a12 + ((size_t)(n0 / 8 + j) * TC_ROWS + r) * 8) =
                            make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
                    }
                }
            }
            ptx::fence_proxy_async(ptx::space_shared);
            ptx::tcgen05_fence_before_thread_sync();
            __syncthreads();
            if (warp == 0) {
                ptx::tcgen05_fence_after_thread_sync();
                umma_ss<16, 16, 5>(tmem + COL_V, s.a12, s.w + TW_W34, 1u, false, leader);
                umma_commit(&s.bar[0], leader);
            }
            // ---- velocity, integration, emit, fp32 feedback state ----
            mbar_wait(&s.bar[0], ph0); ph0 ^= 1;
            ptx::tcgen05_fence_after_thread_sync();
            if (cq == 0) {
                uint32_t a[2];
                ptx::tcgen05_ld_32x32b(a, tlane + COL_V);
                ptx::tcgen05_wait_ld();
                const float v0 = __uint_as_float(a[0]) + s.f32[TF_B34], v1 = __uint_as_float(a[1]) + s.f32[TF_B34 + 1];
                p0 += v0; p1 += v1;
                s.x4[r] = p0; s.x4[TC_ROWS + r] = p1; s.x4[2 * TC_ROWS + r] = v0; s.x4[3 * TC_ROWS + r] = v1;
                if (valid)
                    *reinterpret_cast<float4*>(out + ((size_t)(row0 + r) * n_next + t) * 4) = make_float4(p0, p1, v0, v1);
            }
            ptx::tcgen05_fence_before_thread_sync();
            __syncthreads();
            if (!feed_back) break;
            // ---- LSTM cell: gates = acc(h.Whh^T) + Wx.x4 + b ; new c (registers), new h -> bf16 operand ----
            mbar_wait(&s.bar[1], ph1); ph1 ^= 1;
            ptx::tcgen05_fence_after_thread_sync();
            {
                const float x0 = s.x4[r], x1 = s.x4[TC_ROWS + r], x2 = s.x4[2 * TC_ROWS + r], x3 = s.x4[3 * TC_ROWS + r];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t a[32];
                    ptx::tcgen05_ld_32x32b(a, tlane + COL_G + cq * 64 + half * 32);
                    ptx::tcgen05_wait_ld();
                    float hv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        float g[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int n = cq * 64 + half * 32 + u * 4 + q;
                            g[q] = __uint_as_float(a[u * 4 + q]) + bL[n] +
                                   fmaf(wx[n], x0, fmaf(wx[256 + n], x1, fmaf(wx[512 + n], x2, wx[768 + n] * x3)));
                        }
                        lstm_cell_hw(g, c[half * 8 + u], hv[u]);
                    }
                    *reinterpret_cast<uint4*>(s.al1 + ((size_t)(cq * 2 + half) * TC_ROWS + r) * 8) =
                        make_uint4(pack_bf16(hv[0], hv[1]), pack_bf16(hv[2], hv[3]), pack_bf16(hv[4], hv[5]), pack_bf16(hv[6], hv[7]));
                }
            }
            ptx::fence_proxy_async(ptx::space_shared);
            ptx::tcgen05_fence_before_thread_sync();
            __syncthreads();
        }
    }
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) ptx::tcgen05_dealloc(ptx::cta_group_1, tmem, 512u);
}

}  // namespace sw

extern "C" int sw_decode_fwd_tc(const void* tc_w16, const float* tc_f32, const float* h0, const float* c0,
                                const float* pooled, const float* noise, const float* x_last, float* out,
                                int n_agents, int n_samples, int n_next, int sm_count, void* stream) {
    if (!tc_w16 || !tc_f32 || !h0 || !c0 || !noise || !x_last || !out) return SW_ERR_ARG;
    if (n_agents <= 0 || n_samples <= 0 || n_next <= 0 || sm_count <= 0) return SW_ERR_ARG;
    const long long n_rows = (long long)n_agents * n_samples;
    const long long tiles = (n_rows + sw::TC_ROWS - 1) / sw::TC_ROWS;
    if (tiles > 0x7fffffffLL) return SW_ERR_UNSUPPORTED;
    const int smem = (int)sizeof(sw::TcSmem) + 128;
    SW_SET_MAX_SMEM(sw::decode_fwd_tc_kernel, smem);
    const int grid = (int)(tiles < sm_count ? tiles : sm_count);
    sw::decode_fwd_tc_kernel<<<grid, sw::TC_THREADS, smem, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)tc_w16, tc_f32, h0, c0, pooled, noise, x_last, out, n_agents, n_rows, n_next, (int)tiles);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}

extern "C" int sw_decode_tc_pack_sizes(int* n_bf16, int* n_f32) {
    if (!n_bf16 || !n_f32) return SW_ERR_ARG;
    *n_bf16 = sw::TW_TOTAL;
    *n_f32 = sw::TF_TOTAL;
    return SW_OK;
}
