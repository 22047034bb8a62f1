// Backward of sw_lstm_seq_fwd (zero initial state): back-propagation through time of the observation
// LSTM of the generator's encoder (reference train.py:404 -> autograd of nn.LSTM, train.py:254,268)
// and of Discriminator.obsv_encoder_lstm (train.py:299), as one kernel.
//
// The kernel walks the T steps in reverse on 32-row tiles: gate gradients from the forward stash
// (elementwise, registers), then dL/dh_{t-1} = dG_t . Whh^T as an FFMA register-tile contraction
// (K = 256 gate columns -> 64).  The pre-activation gate gradients dG_t are written out in the
// tile-image layout [T][tiles][256][32]; the WEIGHT gradient is then one plain GEMM over all
// (step, row) pairs,  d(lstm_pack)[0:68] = XH^T . dG  (stash_xh is the forward operand image), done
// by the host with cuBLAS (socialways_b200/autograd_path.py) -- it has no sequential dependency.
#include "sw_common.cuh"

namespace sw {

constexpr int WT_LD = 68;   // transposed pack row: [Wx(4) | Whh(64)] per gate column; 68 % 32 == 4

struct SeqBwdSmem {
    float wt[SW_G * WT_LD];          // pack^T [256][68]
    float dg[SW_G * SW_ROWS];        // gate gradients, k-major
    float dh[SW_H * SW_ROWS];        // dL/dh_t, k-major
};

__global__ void __launch_bounds__(SW_THREADS, 2)
lstm_seq_bwd_kernel(const float* __restrict__ pack_t, const float* __restrict__ stash_gates,
                    const float* __restrict__ dh_last, const float* __restrict__ dc_last,
                    float* __restrict__ g_gates, float* __restrict__ dx /*[N][T][4] or null*/,
                    int n_rows, int T, int n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SeqBwdSmem& s = *reinterpret_cast<SeqBwdSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    copy_f4(s.wt, pack_t, SW_G * WT_LD);
    const LaneMap<1> lmG;   // gate-gradient threads: same ownership as the forward step
    const LaneMap<4> lmH;   // dh contraction: 8 rg x 8 cg x 4 ks, TN = 8

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * SW_ROWS;
        const int rows_valid = min(SW_ROWS, n_rows - row0);
        __syncthreads();
        // dL/dh_T -> shared (k-major), dL/dc_T -> registers
        if (dh_last) {
            load_rows_kmajor(s.dh, s.dg, dh_last, SW_H, rows_valid, [&](int r) { return row0 + r; });
        } else {
            for (int i = tid; i < SW_H * SW_ROWS; i += SW_THREADS) s.dh[i] = 0.0f;
        }
        float dc[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = lmG.rg * 4 + i;
                dc[i][u] = (dc_last && r < rows_valid) ? __ldg(dc_last + (size_t)(row0 + r) * SW_H + lmG.cg * 2 + u) : 0.0f;
            }
        __syncthreads();

        for (int t = T - 1; t >= 0; --t) {
            const float* st = stash_gates + ((size_t)t * n_tiles + tile) * SW_GATE_STASH_FLOATS;
            const float* cp = (t > 0) ? stash_gates + ((size_t)(t - 1) * n_tiles + tile) * SW_GATE_STASH_FLOATS + 4 * SW_H * SW_ROWS
                                      : nullptr;
            lstm_tile_bwd_gates(st, cp, s.dh, s.dg, dc, lmG, rows_valid);
            __syncthreads();
            store_image(g_gates + ((size_t)t * n_tiles + tile) * (SW_G * SW_ROWS), s.dg, SW_G * SW_ROWS);
            if (dx) {   // dL/dx4_t = dG_t . Wx^T : warp owns 4 rows, lane = ks(8) + 8*rl
                const int ks = lane & 7, r = warp * 4 + (lane >> 3);
                float a[4] = {0.f, 0.f, 0.f, 0.f};
                for (int k = ks; k < SW_G; k += 8) {
                    const float g = s.dg[k * SW_ROWS + r];
                    const float4 w = *reinterpret_cast<const float4*>(s.wt + k * WT_LD);
                    a[0] = fmaf(g, w.x, a[0]); a[1] = fmaf(g, w.y, a[1]); a[2] = fmaf(g, w.z, a[2]); a[3] = fmaf(g, w.w, a[3]);
                }
#pragma unroll
                for (int off = 1; off < 8; off <<= 1)
#pragma unroll
                    for (int q = 0; q < 4; ++q) a[q] += __shfl_xor_sync(0xffffffffu, a[q], off);
                if (ks == 0 && r < rows_valid)
                    *reinterpret_cast<float4*>(dx + ((size_t)(row0 + r) * T + t) * 4) = make_float4(a[0], a[1], a[2], a[3]);
            }
            if (t > 0) {
                float acc[4][8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
                fma_tile<8, 4>(acc, s.dg, s.wt + 4, WT_LD, SW_G, lmH);
                ksplit_reduce<8, 4>(acc);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if ((j & 3) == lmH.ks)
                        *reinterpret_cast<float4*>(s.dh + (lmH.cg * 8 + j) * SW_ROWS + lmH.rg * 4) =
                            make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
            }
            __syncthreads();
        }
    }
}

}  // namespace sw

extern "C" int sw_lstm_seq_bwd(const float* lstm_pack_t, const float* stash_gates, const float* dh_last,
                               const float* dc_last, float* g_gates, float* dx, int n_rows, int n_steps,
                               int sm_count, void* stream) {
    if (!lstm_pack_t || !stash_gates || !g_gates) return SW_ERR_ARG;
    if (n_rows <= 0 || n_steps <= 0 || sm_count <= 0) return SW_ERR_ARG;
    const int tiles = (n_rows + SW_ROWS - 1) / SW_ROWS;
    const int smem = (int)sizeof(sw::SeqBwdSmem);
    SW_SET_MAX_SMEM(sw::lstm_seq_bwd_kernel, smem);
    const int grid = tiles < 2 * sm_count ? tiles : 2 * sm_count;
    sw::lstm_seq_bwd_kernel<<<grid, SW_THREADS, smem, (cudaStream_t)stream>>>(lstm_pack_t, stash_gates, dh_last, dc_last,
                                                                              g_gates, dx, n_rows, n_steps, tiles);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
