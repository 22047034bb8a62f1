// Backward of sw_decode_fwd: back-propagation through the 12-step decode loop of predict()
// (reference train.py:418-430; in the reference this is the autograd graph of 12 x {DecoderFC,
// integrate, one EncoderLstm step} that g_loss.backward() walks, train.py:538).
//
// One launch, 32-row tiles, reverse time.  Per step t (state entering the step: h_t, c_t, p_{t-1}):
//     forward was   a1 = lrelu(W1h h_t + c1), a2 = lrelu(W2 a1 + b2), v_t = W34 a2 + b34,
//                   p_t = p_{t-1} + v_t, (h_{t+1}, c_{t+1}) = LSTM((p_t, v_t), h_t, c_t)  [t < T-1]
//     backward:     dG_t       = LSTM gate gradients from (dh_{t+1}, dc_{t+1})      [t < T-1]
//                   dh_t       = dG_t . Whh^T  +  da1pre_t . W1h^T
//                   d(p_t,v_t) = dOut_t + dG_t . Wx^T ; dp_t += dp_{t+1} ; dv_t += dp_t
//                   da2pre_t   = (W34^T dv_t) * lrelu'(a2) ; da1pre_t = (da2pre_t . W2^T) * lrelu'(a1)
// The data-gradient chain (the part with a sequential dependency) runs here as FFMA register-tile
// contractions with transposed weights resident in shared memory.  dG, da1pre, da2pre, dv are
// written out as tile images; every WEIGHT gradient is then a plain GEMM of a forward stash image
// against one of those (cuBLAS on the host side, autograd_path.py), as is dS = (sum_t da1pre) . W1s^T.
#include "sw_common.cuh"

namespace sw {

constexpr int WT_LD = 68;     // pack^T rows [Wx(4) | Whh(64)], KS = 4
constexpr int W2T_LD = 164;   // W2^T [80][160], KS = 2
constexpr int W1T_LD = 68;    // W1h^T [160][64] (+4 pad), KS = 4

struct DecodeBwdSmem {
    float wt[SW_G * WT_LD];
    float w2t[80 * W2T_LD];
    float w1t[160 * W1T_LD];
    float w34[256];                   // W34[80][2] | pad
    float dg[SW_G * SW_ROWS];         // gate gradients; later in the step re-used: da1pre [160][32] | da2pre [80][32]
    float dh[SW_H * SW_ROWS];
    float c0img[SW_H * SW_ROWS];      // cell state entering step 0 (the encoder's post-observation c), k-major
    float dv[2 * SW_ROWS];
};

// dec_pack_t layout (floats): W1h^T [160][64] | W2^T [80][160] | W34 [80][2]
constexpr int DT_W1 = 0, DT_W2 = 160 * 64, DT_W34 = DT_W2 + 80 * 160, DT_TOTAL = DT_W34 + 160;

__global__ void __launch_bounds__(SW_THREADS, 1)
decode_bwd_kernel(const float* __restrict__ pack_t, const float* __restrict__ dec_pack_t,
                  const float* __restrict__ c0, const float* __restrict__ stash_gates,
                  const float* __restrict__ stash_a1, const float* __restrict__ stash_a2,
                  const float* __restrict__ d_out /*[rows][T][4]*/,
                  float* __restrict__ g_gates /*[T-1][tiles][256][32]*/, float* __restrict__ g_a1 /*[T][tiles][160][32]*/,
                  float* __restrict__ g_a2 /*[T][tiles][80][32]*/, float* __restrict__ g_v /*[T][tiles][2][32]*/,
                  float* __restrict__ dh0 /*[rows][64]*/, float* __restrict__ dc0 /*[rows][64]*/,
                  float* __restrict__ g_a1sum /*[tiles][160][32] = sum_t da1pre, or null*/,
                  const float* __restrict__ w1_sz /*DecoderFC.fc1.0.weight [160][160] (torch layout), or null*/,
                  float* __restrict__ d_pooled /*[rows][64] = (sum_t da1pre) . W1[:, 64:128], or null*/,
                  int n_agents, long long n_rows, int T, int n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DecodeBwdSmem& s = *reinterpret_cast<DecodeBwdSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    copy_f4(s.wt, pack_t, SW_G * WT_LD);
    for (int i = tid; i < 80 * 160; i += SW_THREADS) s.w2t[(i / 160) * W2T_LD + (i % 160)] = __ldg(dec_pack_t + DT_W2 + i);
    for (int i = tid; i < 160 * 64; i += SW_THREADS) s.w1t[(i / 64) * W1T_LD + (i % 64)] = __ldg(dec_pack_t + DT_W1 + i);
    for (int i = tid; i < 160; i += SW_THREADS) s.w34[i] = __ldg(dec_pack_t + DT_W34 + i);
    float* da1 = s.dg;                     // [160][32]
    float* da2 = s.dg + 160 * SW_ROWS;     // [80][32]
    const LaneMap<1> lmG;
    const LaneMap<2> lm2;   // da1 contraction: K = 80 -> 160, TN = 10
    const LaneMap<4> lmH;   // dh contractions: K = 256 / 160 -> 64, TN = 8
    const int x_ks = lane & 7, x_r = warp * 4 + (lane >> 3);   // x4-gradient threads: warp owns 4 rows

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = (long long)tile * SW_ROWS;
        const int rows_valid = (int)min((long long)SW_ROWS, n_rows - row0);
        __syncthreads();
        for (int i = tid; i < SW_H * SW_ROWS; i += SW_THREADS) {
            const int r = i >> 6, k = i & 63;
            s.dh[i] = 0.0f;
            s.c0img[k * SW_ROWS + r] = (r < rows_valid) ? __ldg(c0 + (size_t)((row0 + r) % n_agents) * SW_H + k) : 0.0f;
        }
        float dc[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
        float carry0 = 0.0f, carry1 = 0.0f;          // dL/dp_t flowing to p_{t-1}  (x_ks == 0 lanes)
        float sum_a1[5][4];                          // sum_t da1pre of this thread's columns (the hoisted [S ; z] operand's gradient)
#pragma unroll
        for (int q = 0; q < 5; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) sum_a1[q][i] = 0.0f;
        __syncthreads();

        for (int t = T - 1; t >= 0; --t) {
            float acc_h[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc_h[i][j] = 0.0f;
            float a[4] = {0.f, 0.f, 0.f, 0.f};       // dG_t . Wx^T partials
            if (t < T - 1) {
                const float* st = stash_gates + ((size_t)t * n_tiles + tile) * SW_GATE_STASH_FLOATS;
                const float* cp = (t > 0) ? stash_gates + ((size_t)(t - 1) * n_tiles + tile) * SW_GATE_STASH_FLOATS + 4 * SW_H * SW_ROWS
                                          : s.c0img;
                lstm_tile_bwd_gates(st, cp, s.dh, s.dg, dc, lmG, rows_valid);
                __syncthreads();
                store_image(g_gates + ((size_t)t * n_tiles + tile) * (SW_G * SW_ROWS), s.dg, SW_G * SW_ROWS);
                fma_tile<8, 4>(acc_h, s.dg, s.wt + 4, WT_LD, SW_G, lmH);
                for (int k = x_ks; k < SW_G; k += 8) {
                    const float g = s.dg[k * SW_ROWS + x_r];
                    const float4 w = *reinterpret_cast<const float4*>(s.wt + k * WT_LD);
                    a[0] = fmaf(g, w.x, a[0]); a[1] = fmaf(g, w.y, a[1]); a[2] = fmaf(g, w.z, a[2]); a[3] = fmaf(g, w.w, a[3]);
                }
#pragma unroll
                for (int off = 1; off < 8; off <<= 1)
#pragma unroll
                    for (int q = 0; q < 4; ++q) a[q] += __shfl_xor_sync(0xffffffffu, a[q], off);
            }
            // ---- d(p_t, v_t): loss gradient + LSTM-input gradient + position carry ----
            if (x_ks == 0) {
                float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
                if (x_r < rows_valid) go = __ldg(reinterpret_cast<const float4*>(d_out) + (size_t)(row0 + x_r) * T + t);
                const float dp0 = go.x + a[0] + carry0, dp1 = go.y + a[1] + carry1;
                const float dv0 = go.z + a[2] + dp0, dv1 = go.w + a[3] + dp1;
                carry0 = dp0; carry1 = dp1;
                s.dv[x_r] = dv0; s.dv[SW_ROWS + x_r] = dv1;
                float* gv = g_v + ((size_t)t * n_tiles + tile) * (2 * SW_ROWS);
                gv[x_r] = dv0; gv[SW_ROWS + x_r] = dv1;
            }
            __syncthreads();   // dg fully consumed, dv visible
            // ---- da2pre = (W34^T dv) * lrelu'(a2) ----
            {
                const float* a2 = stash_a2 + ((size_t)t * n_tiles + tile) * (80 * SW_ROWS);
                float* ga2 = g_a2 + ((size_t)t * n_tiles + tile) * (80 * SW_ROWS);
                for (int i = tid; i < 80 * SW_ROWS; i += SW_THREADS) {
                    const int j = i >> 5, r = i & 31;
                    const float d = fmaf(s.w34[j * 2], s.dv[r], s.w34[j * 2 + 1] * s.dv[SW_ROWS + r]);
                    const float g = (__ldg(a2 + i) > 0.0f) ? d : 0.2f * d;
                    da2[i] = g;
                    ga2[i] = g;
                }
            }
            __syncthreads();
            // ---- da1pre = (da2pre . W2^T) * lrelu'(a1) ----
            {
                float acc[4][10];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 10; ++j) acc[i][j] = 0.0f;
                fma_tile<10, 2>(acc, da2, s.w2t, W2T_LD, 80, lm2);
                ksplit_reduce<10, 2>(acc);
                const float* a1 = stash_a1 + ((size_t)t * n_tiles + tile) * (160 * SW_ROWS);
                float* ga1 = g_a1 + ((size_t)t * n_tiles + tile) * (160 * SW_ROWS);
#pragma unroll
                for (int j = 0; j < 10; ++j)
                    if ((j & 1) == lm2.ks) {
                        const int off = (lm2.cg * 10 + j) * SW_ROWS + lm2.rg * 4;
                        const float4 av = __ldg(reinterpret_cast<const float4*>(a1 + off));
                        const float4 g = make_float4(av.x > 0.f ? acc[0][j] : 0.2f * acc[0][j], av.y > 0.f ? acc[1][j] : 0.2f * acc[1][j],
                                                     av.z > 0.f ? acc[2][j] : 0.2f * acc[2][j], av.w > 0.f ? acc[3][j] : 0.2f * acc[3][j]);
                        *reinterpret_cast<float4*>(da1 + off) = g;
                        *reinterpret_cast<float4*>(ga1 + off) = g;
                        sum_a1[j >> 1][0] += g.x; sum_a1[j >> 1][1] += g.y; sum_a1[j >> 1][2] += g.z; sum_a1[j >> 1][3] += g.w;
                    }
            }
            __syncthreads();
            // ---- dh_t = dG_t . Whh^T (already in acc_h) + da1pre . W1h^T ----
            fma_tile<8, 4>(acc_h, da1, s.w1t, W1T_LD, 160, lmH);
            ksplit_reduce<8, 4>(acc_h);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if ((j & 3) == lmH.ks)
                    *reinterpret_cast<float4*>(s.dh + (lmH.cg * 8 + j) * SW_ROWS + lmH.rg * 4) =
                        make_float4(acc_h[0][j], acc_h[1][j], acc_h[2][j], acc_h[3][j]);
            __syncthreads();   // dh visible; da1/da2 (aliasing dg) consumed before the next step's gates overwrite them
        }
        // ---- gradients w.r.t. the initial state (h_0, c_0) ----
        for (int i = tid; i < SW_ROWS * SW_H; i += SW_THREADS) {
            const int r = i >> 6, k = i & 63;
            if (r < rows_valid) dh0[(size_t)(row0 + r) * SW_H + k] = s.dh[k * SW_ROWS + r];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = lmG.rg * 4 + i;
            if (r < rows_valid)
                *reinterpret_cast<float2*>(dc0 + (size_t)(row0 + r) * SW_H + lmG.cg * 2) = make_float2(dc[i][0], dc[i][1]);
        }
        if (g_a1sum) {
            // ---- the step-invariant part of layer 1: its gradient operand is sum_t da1pre (image, for the W1[S,z] / b1
            //      weight gradients) and dL/dS = (sum_t da1pre) . W1[:, S columns] ----
#pragma unroll
            for (int j = 0; j < 10; ++j)
                if ((j & 1) == lm2.ks)
                    *reinterpret_cast<float4*>(da1 + (lm2.cg * 10 + j) * SW_ROWS + lm2.rg * 4) =
                        make_float4(sum_a1[j >> 1][0], sum_a1[j >> 1][1], sum_a1[j >> 1][2], sum_a1[j >> 1][3]);
            __syncthreads();     // also: the dh0 writer above is done with s.dh
            store_image(g_a1sum + (size_t)tile * (160 * SW_ROWS), da1, 160 * SW_ROWS);
            if (d_pooled) {
                float acc[4][8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
                fma_tile<8, 4>(acc, da1, w1_sz + 64, 160, 160, lmH);       // W[n][64 + k], n = contraction index
                ksplit_reduce<8, 4>(acc);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if ((j & 3) == lmH.ks)
                        *reinterpret_cast<float4*>(s.dh + (lmH.cg * 8 + j) * SW_ROWS + lmH.rg * 4) =
                            make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
                __syncthreads();
                for (int i = tid; i < SW_ROWS * SW_H; i += SW_THREADS) {
                    const int r = i >> 6, k = i & 63;
                    if (r < rows_valid) d_pooled[(size_t)(row0 + r) * SW_H + k] = s.dh[k * SW_ROWS + r];
                }
            }
        }
    }
}

}  // namespace sw

extern "C" int sw_decode_bwd(const float* lstm_pack_t, const float* dec_pack_t, const float* c0,
                             const float* stash_gates, const float* stash_a1, const float* stash_a2, const float* d_out,
                             float* g_gates, float* g_a1, float* g_a2, float* g_v, float* dh0, float* dc0,
                             float* g_a1sum, const float* w1_torch, float* d_pooled,
                             int n_agents, int n_samples, int n_next, int sm_count, void* stream) {
    if (!lstm_pack_t || !dec_pack_t || !c0 || !stash_a1 || !stash_a2 || !d_out || !g_a1 || !g_a2 || !g_v || !dh0 || !dc0)
        return SW_ERR_ARG;
    if (n_next > 1 && (!stash_gates || !g_gates)) return SW_ERR_ARG;
    if (n_agents <= 0 || n_samples <= 0 || n_next <= 0 || sm_count <= 0) return SW_ERR_ARG;
    if (d_pooled && (!g_a1sum || !w1_torch)) return SW_ERR_ARG;
    const long long n_rows = (long long)n_agents * n_samples;
    const long long tiles = (n_rows + SW_ROWS - 1) / SW_ROWS;
    if (tiles > 0x7fffffffLL) return SW_ERR_UNSUPPORTED;
    const int smem = (int)sizeof(sw::DecodeBwdSmem);
    SW_SET_MAX_SMEM(sw::decode_bwd_kernel, smem);
    const int grid = (int)(tiles < sm_count ? tiles : sm_count);
    sw::decode_bwd_kernel<<<grid, SW_THREADS, smem, (cudaStream_t)stream>>>(
        lstm_pack_t, dec_pack_t, c0, stash_gates, stash_a1, stash_a2, d_out, g_gates, g_a1, g_a2, g_v, dh0, dc0, g_a1sum,
        w1_torch, d_pooled, n_agents, n_rows, n_next, (int)tiles);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}

extern "C" int sw_decode_pack_t_floats(void) { return sw::DT_TOTAL; }
