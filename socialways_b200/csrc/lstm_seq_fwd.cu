// LSTM over an observed sequence from a zero state -- the observation pass of EncoderLstm
// (reference train.py:262-269 called at :404, with get_traj_4d :130-134 fused in when the input is
// positions) and of Discriminator.obsv_encoder_lstm (train.py:296-299).
//
// in_dim == 2: x is positions [N][T][2]; the 4-d state (p_t, p_t - p_{t-1}), v_0 := v_1, is formed
//              on the fly (train.py:131-133).
// in_dim == 4: x is already [N][T][4].
// The initial state is zero (predict(), train.py:399-401; Discriminator, :296-297) unless h_in/c_in
// are given (the module-level EncoderLstm.forward carries its state between calls, train.py:268).
// Outputs: h, c after the last step [N][64]; optionally every h_t as y [N][T][64] (the value
// EncoderLstm.forward returns), the last 4-d state [N][4] (what predict() integrates from,
// train.py:416) and, for the backward pass, the per-step stash.
//
// One CTA = 32 rows, weights resident in shared memory, cell state in registers (sw_common.cuh).
#include "sw_common.cuh"

namespace sw {

constexpr int SEQ_T_MAX = 32;

struct SeqSmem {
    float wl[SW_LSTM_PACK_FLOATS];
    float xb[2][SW_LSTM_K * SW_ROWS];
    float xin[SW_ROWS * SEQ_T_MAX * 4];   // the tile's raw input rows
};

template <bool STASH>
__global__ void __launch_bounds__(SW_THREADS, 2)
lstm_seq_fwd_kernel(const float* __restrict__ lstm_pack, const float* __restrict__ x, int in_dim, int n_rows,
                    int T, const float* __restrict__ h_in, const float* __restrict__ c_in,
                    float* __restrict__ y_out, float* __restrict__ h_out, float* __restrict__ c_out, float* __restrict__ x_last,
                    float* __restrict__ stash_gates /*[T][tiles][5][64][32]*/,
                    float* __restrict__ stash_xh /*[T][tiles][68][32]: the {x4 ; h_{t-1}} operand of every step*/,
                    int n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SeqSmem& s = *reinterpret_cast<SeqSmem*>(smem_raw);
    const int tid = threadIdx.x;
    copy_f4(s.wl, lstm_pack, SW_LSTM_PACK_FLOATS);
    const LaneMap<1> lm;
    const int row_floats = T * in_dim;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * SW_ROWS;
        const int rows_valid = min(SW_ROWS, n_rows - row0);
        __syncthreads();
        for (int i = tid; i < SW_ROWS * row_floats; i += SW_THREADS)
            s.xin[i] = (i < rows_valid * row_floats) ? __ldg(x + (size_t)row0 * row_floats + i) : 0.0f;
        float c[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
        if (h_in) {
            load_rows_kmajor(s.xb[0] + 4 * SW_ROWS, s.xb[1], h_in, SW_H, rows_valid, [&](int r) { return row0 + r; });
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = lm.rg * 4 + i;
                if (r < rows_valid) {
                    const float2 cv = __ldg(reinterpret_cast<const float2*>(c_in + (size_t)(row0 + r) * SW_H + lm.cg * 2));
                    c[i][0] = cv.x; c[i][1] = cv.y;
                }
            }
        } else {
            for (int i = tid; i < SW_H * SW_ROWS; i += SW_THREADS) s.xb[0][4 * SW_ROWS + i] = 0.0f;
        }
        __syncthreads();

        for (int t = 0; t < T; ++t) {
            float* X = s.xb[t & 1];
            float* Hn = s.xb[(t + 1) & 1] + 4 * SW_ROWS;
            if (tid < SW_ROWS) {
                const float* xr = s.xin + tid * row_floats;
                float4 st;
                if (in_dim == 2) {
                    const int tv = (t == 0) ? 1 : t;   // v_0 := v_1 (train.py:132)
                    st = make_float4(xr[t * 2], xr[t * 2 + 1], xr[tv * 2] - xr[(tv - 1) * 2],
                                     xr[tv * 2 + 1] - xr[(tv - 1) * 2 + 1]);
                } else {
                    st = make_float4(xr[t * 4], xr[t * 4 + 1], xr[t * 4 + 2], xr[t * 4 + 3]);
                }
                X[0 * SW_ROWS + tid] = st.x; X[1 * SW_ROWS + tid] = st.y;
                X[2 * SW_ROWS + tid] = st.z; X[3 * SW_ROWS + tid] = st.w;
                if (tid < rows_valid && t == T - 1 && x_last)
                    *reinterpret_cast<float4*>(x_last + (size_t)(row0 + tid) * 4) = st;
            }
            __syncthreads();
            if (STASH) store_image(stash_xh + ((size_t)t * n_tiles + tile) * (SW_LSTM_K * SW_ROWS), X, SW_LSTM_K * SW_ROWS);
            lstm_tile_step<STASH>(s.wl, X, Hn, c, lm,
                                  STASH ? stash_gates + ((size_t)t * n_tiles + tile) * SW_GATE_STASH_FLOATS : nullptr);
            __syncthreads();
            if (y_out) {   // h_t row-major: the module's return value (train.py:268-269)
                for (int i = tid; i < SW_ROWS * SW_H; i += SW_THREADS) {
                    const int r = i >> 6, k = i & 63;
                    if (r < rows_valid) y_out[((size_t)(row0 + r) * T + t) * SW_H + k] = Hn[k * SW_ROWS + r];
                }
            }
        }
        const float* Hf = s.xb[T & 1] + 4 * SW_ROWS;
        for (int i = tid; i < SW_ROWS * SW_H; i += SW_THREADS) {
            const int r = i >> 6, k = i & 63;
            if (r < rows_valid) h_out[(size_t)(row0 + r) * SW_H + k] = Hf[k * SW_ROWS + r];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = lm.rg * 4 + i;
            if (r < rows_valid)
                *reinterpret_cast<float2*>(c_out + (size_t)(row0 + r) * SW_H + lm.cg * 2) = make_float2(c[i][0], c[i][1]);
        }
    }
}

}  // namespace sw

extern "C" int sw_lstm_seq_fwd(const float* lstm_pack, const float* x, int in_dim, int n_rows, int n_steps,
                               const float* h_in, const float* c_in, float* y_out, float* h_out, float* c_out,
                               float* x_last, float* stash_gates, float* stash_xh, int sm_count, void* stream) {
    if (!lstm_pack || !x || !h_out || !c_out) return SW_ERR_ARG;
    if (n_rows <= 0 || sm_count <= 0 || (in_dim != 2 && in_dim != 4)) return SW_ERR_ARG;
    if (n_steps < (in_dim == 2 ? 2 : 1) || n_steps > sw::SEQ_T_MAX) return SW_ERR_UNSUPPORTED;
    const bool stash = stash_gates != nullptr;
    if (stash && !stash_xh) return SW_ERR_ARG;
    if ((h_in == nullptr) != (c_in == nullptr)) return SW_ERR_ARG;
    const int tiles = (n_rows + SW_ROWS - 1) / SW_ROWS;
    const int smem = (int)sizeof(sw::SeqSmem);
    auto kern = stash ? sw::lstm_seq_fwd_kernel<true> : sw::lstm_seq_fwd_kernel<false>;
    SW_SET_MAX_SMEM(sw::lstm_seq_fwd_kernel<true>, smem);     // one static cache per call site: set both instantiations
    SW_SET_MAX_SMEM(sw::lstm_seq_fwd_kernel<false>, smem);
    // two CTAs fit per SM (104 KB each): let short grids spread over more SMs' worth of slots
    const int grid = tiles < 2 * sm_count ? tiles : 2 * sm_count;
    kern<<<grid, SW_THREADS, smem, (cudaStream_t)stream>>>(lstm_pack, x, in_dim, n_rows, n_steps, h_in, c_in, y_out, h_out, c_out, x_last,
                                                          stash_gates, stash_xh, tiles);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
