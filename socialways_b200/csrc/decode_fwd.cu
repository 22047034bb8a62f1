// K-sample autoregressive decode: the loop of predict() (reference train.py:418-430) for every
// (sample k, agent n) row in ONE launch.
//
// Per row and step:  v = DecoderFC(h, S, z) (train.py:320-335);  p += v;  (p, v) is emitted and fed
// back through one EncoderLstm step (train.py:262-269).  After pooling the rows are independent
// (SURVEY.md §3.2), so K samples x N agents are folded into the row dimension: row = k*N + n.
//
// Algebra done once on the host (socialways_b200/packing.py), exact up to fp32 re-association:
//   * embed Linear(4,64) folded into the LSTM input projection  (gates = Wx.x4 + Whh.h + b, K = 68)
//   * the S and z columns of DecoderFC layer 1 are constant over the 12 steps -> hoisted: one
//     96 x 160 contraction per row, kept in registers as the accumulator seed of every step
//   * the last two Linear layers of DecoderFC have no activation between them -> one 80 x 2 layer
//   * the encoder step after the last prediction (train.py:430) is never observed -> skipped
//
// Layout: persistent CTAs, one per SM; all per-step weights resident in shared memory (210 KB);
// 32-row tiles; activations k-major in shared memory; fp32 FFMA register tiles (sw_common.cuh).
#include "sw_common.cuh"

namespace sw {

constexpr int W1H_LD = 164;  // 160 cols + 4: consecutive k rows land 4 banks apart (KS = 2)
constexpr int W2_LD = 84;    // 80 cols + 4 (KS = 4)
constexpr int XB = SW_LSTM_K * SW_ROWS;  // one {x4 ; h} operand buffer

struct DecodeSmem {
    float wl[SW_LSTM_PACK_FLOATS];   // LSTM pack
    float w1h[64 * W1H_LD];          // DecoderFC layer 1, h rows
    float w2[160 * W2_LD];           // DecoderFC layer 2
    float tail[256];                 // b2[80] | W34[80][2] | b34[2]
    float xb[2][XB];                 // ping-pong {x4 ; h}
    float a1[160 * SW_ROWS];
    float a2[80 * SW_ROWS];
};

// dec_pack layout (floats): W1[160][160] k-major rows {h 0..63, S 64..127, z 128..159} | b1[160] |
//                           W2[160][80] | b2[80] | W34[80][2] | b34[2]
constexpr int DP_W1 = 0, DP_B1 = 160 * 160, DP_W2 = DP_B1 + 160, DP_B2 = DP_W2 + 160 * 80,
              DP_W34 = DP_B2 + 80, DP_B34 = DP_W34 + 160, DP_TOTAL = DP_B34 + 2;

template <bool STASH>
__global__ void __launch_bounds__(SW_THREADS, 1)
decode_fwd_kernel(const float* __restrict__ lstm_pack, const float* __restrict__ dec_pack,
                  const float* __restrict__ h0, const float* __restrict__ c0,
                  const float* __restrict__ pooled, const float* __restrict__ noise,
                  const float* __restrict__ x_last, float* __restrict__ out,
                  float* __restrict__ stash_xh /*[T][tiles][68][32]*/, float* __restrict__ stash_gates /*[T-1][tiles][5][64][32]*/,
                  float* __restrict__ stash_a1 /*[T][tiles][160][32]*/, float* __restrict__ stash_a2 /*[T][tiles][80][32]*/,
                  float* __restrict__ stash_sz /*[tiles][96][32]: the hoisted operand [S ; z], or null*/,
                  int n_agents, long long n_rows, int n_next, int n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DecodeSmem& s = *reinterpret_cast<DecodeSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- weights -> shared, once per CTA ----
    copy_f4(s.wl, lstm_pack, SW_LSTM_PACK_FLOATS);
    for (int i = tid; i < 64 * 160; i += SW_THREADS) s.w1h[(i / 160) * W1H_LD + (i % 160)] = __ldg(dec_pack + DP_W1 + i);
    for (int i = tid; i < 160 * 80; i += SW_THREADS) s.w2[(i / 80) * W2_LD + (i % 80)] = __ldg(dec_pack + DP_W2 + i);
    for (int i = tid; i < 242; i += SW_THREADS) s.tail[i] = __ldg(dec_pack + DP_B2 + i);
    __syncthreads();

    const LaneMap<1> lmL;   // LSTM:    8 rg x 32 cg, TN = 8
    const LaneMap<2> lm1;   // layer 1: 8 rg x 16 cg x 2 ks, TN = 10
    const LaneMap<4> lm2;   // layer 2: 8 rg x  8 cg x 4 ks, TN = 10
    // final 80 -> 2 layer: warp owns 4 rows; lane = ks + 4*o + 8*rl
    const int f_ks = lane & 3, f_o = (lane >> 2) & 1, f_r = warp * 4 + (lane >> 3);
    const float* b2 = s.tail;
    const float* w34 = s.tail + 80;
    const float* b34 = s.tail + 240;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = (long long)tile * SW_ROWS;
        const int rows_valid = (int)min((long long)SW_ROWS, n_rows - row0);
        auto agent_of = [&](int r) { return (int)((row0 + r) % n_agents); };

        // ---- hoisted layer-1 constant: c1 = b1 + W1[S rows].S + W1[z rows].z ----
        float c1[4][10];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 10; ++j) c1[i][j] = 0.0f;
        if (pooled != nullptr) {
            load_rows_kmajor(s.a1, s.a2, pooled, 64, rows_valid, agent_of);
            if (STASH && stash_sz) store_image(stash_sz + (size_t)tile * (96 * SW_ROWS), s.a1, 64 * SW_ROWS);
            fma_tile<10, 2>(c1, s.a1, dec_pack + DP_W1 + 64 * 160, 160, 64, lm1);
            __syncthreads();
        } else if (STASH && stash_sz) {
            for (int i = tid * 4; i < 64 * SW_ROWS; i += SW_THREADS * 4)
                *reinterpret_cast<float4*>(stash_sz + (size_t)tile * (96 * SW_ROWS) + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        load_rows_kmajor(s.a1, s.a2, noise, SW_Z, rows_valid, [&](int r) { return row0 + r; });
        if (STASH && stash_sz) store_image(stash_sz + (size_t)tile * (96 * SW_ROWS) + 64 * SW_ROWS, s.a1, SW_Z * SW_ROWS);
        fma_tile<10, 2>(c1, s.a1, dec_pack + DP_W1 + 128 * 160, 160, SW_Z, lm1);
        ksplit_reduce<10, 2>(c1);
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            const float b = __ldg(dec_pack + DP_B1 + lm1.cg * 10 + j);
#pragma unroll
            for (int i = 0; i < 4; ++i) c1[i][j] = (lm1.ks == 0) ? c1[i][j] + b : 0.0f;   // seed lives in the ks==0 lane
        }
        __syncthreads();

        // ---- initial state: h0 -> xb[0] rows 4..67, c0 -> registers, x_last -> xb[0] rows 0..3 ----
        load_rows_kmajor(s.xb[0] + 4 * SW_ROWS, s.a2, h0, SW_H, rows_valid, agent_of);
        float c[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = lmL.rg * 4 + i;
                c[i][u] = (r < rows_valid) ? __ldg(c0 + (size_t)agent_of(r) * SW_H + lmL.cg * 2 + u) : 0.0f;
            }
        float p_cur = 0.0f;   // position component f_o of row f_r (meaningful in the f_ks == 0 lanes)
        if (f_r < rows_valid) p_cur = __ldg(x_last + (size_t)agent_of(f_r) * 4 + f_o);

        for (int t = 0; t < n_next; ++t) {
            float* X = s.xb[t & 1];
            // ---- layer 1: 64 (+hoisted 96) -> 160, LeakyReLU(0.2) ----
            {
                float acc[4][10];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 10; ++j) acc[i][j] = c1[i][j];
                fma_tile<10, 2>(acc, X + 4 * SW_ROWS, s.w1h, W1H_LD, SW_H, lm1);
                ksplit_reduce<10, 2>(acc);
#pragma unroll
                for (int j = 0; j < 10; ++j)
                    if ((j & 1) == lm1.ks)
                        *reinterpret_cast<float4*>(s.a1 + (lm1.cg * 10 + j) * SW_ROWS + lm1.rg * 4) =
                            make_float4(lrelu02(acc[0][j]), lrelu02(acc[1][j]), lrelu02(acc[2][j]), lrelu02(acc[3][j]));
            }
            __syncthreads();
            if (STASH) store_image(stash_a1 + ((size_t)t * n_tiles + tile) * (160 * SW_ROWS), s.a1, 160 * SW_ROWS);
            // ---- layer 2: 160 -> 80, LeakyReLU(0.2) ----
            {
                float acc[4][10];
#pragma unroll
                for (int j = 0; j < 10; ++j) {
                    const float b = (lm2.ks == 0) ? b2[lm2.cg * 10 + j] : 0.0f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i][j] = b;
                }
                fma_tile<10, 4>(acc, s.a1, s.w2, W2_LD, 160, lm2);
                ksplit_reduce<10, 4>(acc);
#pragma unroll
                for (int j = 0; j < 10; ++j)
                    if ((j & 3) == lm2.ks)
                        *reinterpret_cast<float4*>(s.a2 + (lm2.cg * 10 + j) * SW_ROWS + lm2.rg * 4) =
                            make_float4(lrelu02(acc[0][j]), lrelu02(acc[1][j]), lrelu02(acc[2][j]), lrelu02(acc[3][j]));
            }
            __syncthreads();
            if (STASH) store_image(stash_a2 + ((size_t)t * n_tiles + tile) * (80 * SW_ROWS), s.a2, 80 * SW_ROWS);
            // ---- folded layers 3+4: 80 -> 2 velocity; integrate; emit (p, v); feed back as x4 ----
            {
                float v = 0.0f;
#pragma unroll 5
                for (int k = f_ks; k < SW_DEC_H2; k += 4) v = fmaf(s.a2[k * SW_ROWS + f_r], w34[k * 2 + f_o], v);
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += b34[f_o];
                p_cur += v;
                const float p_other = __shfl_xor_sync(0xffffffffu, p_cur, 4);
                const float v_other = __shfl_xor_sync(0xffffffffu, v, 4);
                if (f_ks == 0) {
                    X[f_o * SW_ROWS + f_r] = p_cur;
                    X[(2 + f_o) * SW_ROWS + f_r] = v;
                    if (f_o == 0 && f_r < rows_valid)
                        *reinterpret_cast<float4*>(out + ((size_t)(row0 + f_r) * n_next + t) * 4) =
                            make_float4(p_cur, p_other, v, v_other);
                }
            }
            if (!STASH && t + 1 == n_next) break;
            __syncthreads();
            if (STASH) store_image(stash_xh + ((size_t)t * n_tiles + tile) * XB, X, XB);
            if (t + 1 == n_next) break;
            // ---- encoder LSTM step on (p, v): {x4 ; h} -> h', c' ----
            lstm_tile_step<STASH>(s.wl, X, s.xb[(t + 1) & 1] + 4 * SW_ROWS, c, lmL,
                                  STASH ? stash_gates + ((size_t)t * n_tiles + tile) * SW_GATE_STASH_FLOATS : nullptr);
            __syncthreads();
        }
        __syncthreads();
    }
}

}  // namespace sw

extern "C" int sw_decode_fwd(const float* lstm_pack, const float* dec_pack, const float* h0, const float* c0,
                             const float* pooled, const float* noise, const float* x_last, float* out,
                             float* stash_xh, float* stash_gates, float* stash_a1, float* stash_a2, float* stash_sz,
                             int n_agents, int n_samples, int n_next, int sm_count, void* stream) {
    if (!lstm_pack || !dec_pack || !h0 || !c0 || !noise || !x_last || !out) return SW_ERR_ARG;
    if (n_agents <= 0 || n_samples <= 0 || n_next <= 0 || sm_count <= 0) return SW_ERR_ARG;
    const long long n_rows = (long long)n_agents * n_samples;
    const long long tiles = (n_rows + SW_ROWS - 1) / SW_ROWS;
    if (tiles > 0x7fffffffLL) return SW_ERR_UNSUPPORTED;
    const bool stash = stash_xh != nullptr;
    if (stash && (!stash_a1 || !stash_a2 || (n_next > 1 && !stash_gates))) return SW_ERR_ARG;
    const int smem = (int)sizeof(sw::DecodeSmem);
    auto kern = stash ? sw::decode_fwd_kernel<true> : sw::decode_fwd_kernel<false>;
    SW_SET_MAX_SMEM(sw::decode_fwd_kernel<true>, smem);     // one static cache per call site: set both instantiations
    SW_SET_MAX_SMEM(sw::decode_fwd_kernel<false>, smem);
    const int grid = (int)(tiles < sm_count ? tiles : sm_count);
    kern<<<grid, SW_THREADS, smem, (cudaStream_t)stream>>>(lstm_pack, dec_pack, h0, c0, pooled, noise, x_last, out, stash_xh,
                                                           stash_gates, stash_a1, stash_a2, stash_sz, n_agents, n_rows, n_next,
                                                           (int)tiles);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}

extern "C" int sw_decode_pack_floats(void) { return sw::DP_TOTAL; }
