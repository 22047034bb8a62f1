// Tensor-core observation encoder: LSTM over the observed sequence from a zero state (reference EncoderLstm
// observation pass, train.py:404 -> :262-269, with get_traj_4d :130-134 fused in) -- the inference-path variant of
// lstm_seq_fwd.cu (no backward stash), same fp16 hi/lo split scheme and the same operand tricks as the decode kernels:
//     gates[128 x 256] = h (K = 64, hi|lo fp16 in shared memory) . Whh^T          3 tcgen05.mma passes, fp32 accumulate in TMEM
//                      + [x_hi | x_lo | x_hi | 1 | 1 | 0 0] . [Wx_hi | Wx_hi | Wx_lo | b_hi | b_lo | 0 0]^T   ONE extra K block
// with the gate rows of Whh / Wx / b pre-scaled on the host by -log2(e) (i, f, o) and -2 log2(e) (g): the accumulator IS the ex2
// argument of the logistic forms, the epilogue is the cell update alone (the first version formed Wx.x + b and the scales in
// the epilogue: ~61 instead of ~37 instructions per hidden unit, 0.26 -> see DESIGN.md).  h_0 = 0: step 0 issues the x block only.
// One CTA = one 128-row tile, 256 threads (2 per row: unit halves), 108 KB shared memory and 256 TMEM columns, so TWO
// CTAs are co-resident per SM and the MMAs of one overlap the gate epilogue of the other.
#include <cuda_fp16.h>

#include "sw_common.cuh"
#include "sw_umma.cuh"

namespace sw {

constexpr int E_ROWS = 128, E_THREADS = 256;
constexpr int EW_HI = 0, EW_LO = 64 * 256, EW_XK = 2 * 64 * 256, EW_TOTAL = 2 * 64 * 256 + 256 * 16;   // Whh hi | lo [8][256][8], x block [2][256][8]

struct EncSmem {
    __half w[EW_TOTAL];                // 73 728 B
    __half h[2][8 * E_ROWS * 8];       // 32 768 B
    __half xk[2 * E_ROWS * 8];         //  4 096 B: x-feedback A operand, one K block [2 chunks][128][8]
    unsigned long long bar;
    uint32_t tmem_base;
};

__device__ __forceinline__ void enc_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h2 = __floats2half2_rn(a, b);
    const float2 back = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(a - back.x, b - back.y);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}

__global__ void __launch_bounds__(E_THREADS, 2)
lstm_seq_fwd_tcx_kernel(const __half* __restrict__ w16, const float* __restrict__ x, int in_dim,
                        int n_rows, int T, float* __restrict__ h_out, float* __restrict__ c_out, float* __restrict__ x_last,
                        int n_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    EncSmem& s = *reinterpret_cast<EncSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lq = warp & 3, cq = warp >> 2;                 // TMEM lane quarter, column half (units 32 cq .. 32 cq + 31)
    const int r = lq * 32 + lane;

    for (int i = tid * 8; i < EW_TOTAL; i += E_THREADS * 8)
        *reinterpret_cast<uint4*>(s.w + i) = __ldg(reinterpret_cast<const uint4*>(w16 + i));
    if (warp == 0) {
        ptx::tcgen05_alloc(ptx::cta_group_1, &s.tmem_base, 256u);
        ptx::tcgen05_relinquish_alloc_permit(ptx::cta_group_1);
    }
    if (tid == 0) {
        ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.bar), 1);
        ptx::fence_mbarrier_init(ptx::sem_release, ptx::scope_cluster);
    }
    ptx::fence_proxy_async(ptx::space_shared);
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    ptx::tcgen05_fence_after_thread_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, s.tmem_base, 0);
    const uint32_t tl = tmem + ((uint32_t)(lq * 32) << 16) + cq * 128;
    uint32_t ph = 0;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row = tile * E_ROWS + r;
        const bool valid = row < n_rows;
        const float* xr = x + (size_t)(valid ? row : 0) * T * in_dim;
        // the 4-d state (p_t, p_t - p_{t-1}), v_0 := v_1 (train.py:131-133)
        auto load_x = [&](int t) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                if (in_dim == 2) {
                    const int tv = (t == 0) ? 1 : t;
                    const float2 p = __ldg(reinterpret_cast<const float2*>(xr) + t);
                    const float2 a = __ldg(reinterpret_cast<const float2*>(xr) + tv), b = __ldg(reinterpret_cast<const float2*>(xr) + tv - 1);
                    v = make_float4(p.x, p.y, a.x - b.x, a.y - b.y);
                } else {
                    v = __ldg(reinterpret_cast<const float4*>(xr) + t);
                }
            }
            return v;
        };
        // x_t -> hi|lo x block of the gate MMA (the unit-half-0 thread of the row)
        auto put_x = [&](const float4& v) {
            uint32_t hp, lp, hv, lv;
            enc_split2(v.x, v.y, hp, lp);
            enc_split2(v.z, v.w, hv, lv);
            *reinterpret_cast<uint4*>(s.xk + (size_t)r * 8) = make_uint4(hp, hv, lp, lv);                      // k 0..7
            *reinterpret_cast<uint4*>(s.xk + (size_t)(E_ROWS + r) * 8) = make_uint4(hp, hv, 0x3C003C00u, 0u);  // k 8..15
        };
        float c[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) c[u] = 0.0f;
        float4 xt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cq == 0) { xt = load_x(0); put_x(xt); }
        ptx::fence_proxy_async(ptx::space_shared);
        ptx::tcgen05_fence_before_thread_sync();
        __syncthreads();

        for (int t = 0; t < T; ++t) {
            if (warp == 0) {    // one elected lane issues the MMAs back to back (sw_umma.cuh: single-thread issue forms)
                ptx::tcgen05_fence_after_thread_sync();
                if (elect_one()) {
                    if (t > 0) {                                  // h_0 = 0: no recurrent term at step 0
                        umma1_ss<256, 256, 4>(tmem, s.h[0], s.w + EW_HI, 0u, false);
                        umma1_ss<256, 256, 4>(tmem, s.h[0], s.w + EW_LO, 0u, true);
                        umma1_ss<256, 256, 4>(tmem, s.h[1], s.w + EW_HI, 0u, true);
                    }
                    umma1_ss<256, 256, 1>(tmem, s.xk, s.w + EW_XK, 0u, t > 0);
                    umma1_commit(&s.bar);
                }
                __syncwarp();
            }
            if (cq == 0) {      // x_last = x_{T-1}; the next step's state is fetched while the MMAs run
                if (t == T - 1) { if (valid && x_last) *reinterpret_cast<float4*>(x_last + (size_t)row * 4) = xt; }
                else xt = load_x(t + 1);
            }
            mbar_wait(&s.bar, ph); ph ^= 1;
            ptx::tcgen05_fence_after_thread_sync();
#pragma unroll
            for (int part = 0; part < 4; ++part) {               // 32 gate columns = 8 units at a time
                uint32_t a[32];
                tmem_ld<32>(tl + part * 32, a);
                ptx::tcgen05_wait_ld();
                float hv[8];
#pragma unroll
                for (int u = 0; u < 8; u += 2) {
                    float g[2][4];
#pragma unroll
                    for (int w2 = 0; w2 < 2; ++w2)
#pragma unroll
                        for (int q = 0; q < 4; ++q) g[w2][q] = __uint_as_float(a[(u + w2) * 4 + q]);
                    lstm_cell_pair_prescaled_x2(g[0], g[1], c[part * 8 + u], c[part * 8 + u + 1], hv[u], hv[u + 1]);   // packed fp32x2 form (bit-identical)
                }
                if (t == T - 1) {
                    if (valid) {
                        float* ho = h_out + (size_t)row * SW_H + cq * 32 + part * 8;
                        *reinterpret_cast<float4*>(ho) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                        *reinterpret_cast<float4*>(ho + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
                    }
                } else {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) enc_split2(hv[2 * e], hv[2 * e + 1], hi[e], lo[e]);
                    const size_t off = ((size_t)(cq * 4 + part) * E_ROWS + r) * 8;
                    *reinterpret_cast<uint4*>(s.h[0] + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(s.h[1] + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            if (cq == 0 && t + 1 < T) put_x(xt);                  // (this step's MMAs, the only readers of the x block, have completed)
            ptx::fence_proxy_async(ptx::space_shared);
            ptx::tcgen05_fence_before_thread_sync();
            __syncthreads();
        }
        if (valid) {
            float* co = c_out + (size_t)row * SW_H + cq * 32;
#pragma unroll
            for (int u = 0; u < 32; u += 4) *reinterpret_cast<float4*>(co + u) = make_float4(c[u], c[u + 1], c[u + 2], c[u + 3]);
        }
    }
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) ptx::tcgen05_dealloc(ptx::cta_group_1, tmem, 256u);
}

}  // namespace sw

extern "C" int sw_lstm_seq_fwd_tcx(const void* enc_w16, const float* x, int in_dim, int n_rows,
                                   int n_steps, float* h_out, float* c_out, float* x_last, int sm_count, void* stream) {
    if (!enc_w16 || !x || !h_out || !c_out) return SW_ERR_ARG;
    if (n_rows <= 0 || sm_count <= 0 || (in_dim != 2 && in_dim != 4)) return SW_ERR_ARG;
    if (n_steps < (in_dim == 2 ? 2 : 1)) return SW_ERR_UNSUPPORTED;
    const int tiles = (n_rows + sw::E_ROWS - 1) / sw::E_ROWS;
    const int smem = (int)sizeof(sw::EncSmem);
    SW_SET_MAX_SMEM(sw::lstm_seq_fwd_tcx_kernel, smem);
    const int grid = tiles < 2 * sm_count ? tiles : 2 * sm_count;
    sw::lstm_seq_fwd_tcx_kernel<<<grid, sw::E_THREADS, smem, (cudaStream_t)stream>>>(
        (const __half*)enc_w16, x, in_dim, n_rows, n_steps, h_out, c_out, x_last, tiles);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
