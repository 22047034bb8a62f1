// Parameter folding / packing for the training step, as kernels (one launch each) instead of ~100 tiny tensor ops per
// iteration, together with the adjoint of the folds (pack gradients -> parameter gradients).
//
// The algebra is socialways_b200/packing.py's (exact up to fp32 re-association):
//   * EncoderLstm.embed (Linear 4 -> 64, reference train.py:251) folded into the LSTM input projection;
//   * DecoderFC.fc1 (train.py:324-328): layer weights k-major, the last two Linear layers (no activation between) as
//     one 80 -> 2 layer;
//   * EmbedSocialFeatures.fc.4 and AttentionPooling.W (train.py:158,185) folded into the per-agent map (u | beta) = h.M + m0;
//   * Discriminator (train.py:278-292): LSTM(4, 64) gate-interleaved, the 8 Linear layers of the heads in the layouts
//     csrc/disc_step.cu reads (k-major for the forward pass, [out][in] for the backward pass).
// Generator parameter order = Generator.optimizer_parameters() (train.py:379-380: attention, feature_embedder, encoder,
// decoder), discriminator order = Discriminator.parameters(); see include/socialways_b200.h.
#include "sw_common.cuh"
#include "disc_layout.cuh"

namespace sw {

struct GenParams { const float* p[22]; };
struct GenGrads { float* g[22]; };
struct DiscParams { const float* p[20]; };

enum { G_ATT_W, G_ATT_B, G_FC0_W, G_FC0_B, G_FC2_W, G_FC2_B, G_FC4_W, G_FC4_B, G_EMB_W, G_EMB_B, G_WIH, G_WHH, G_BIH, G_BHH,
       G_W1, G_B1, G_W2, G_B2, G_W3, G_B3, G_W4, G_B4 };

// dec_pack (csrc/decode_fwd.cu): W1t [160][160] | b1 | W2t [160][80] | b2 | W34 [80][2] | b34
constexpr int GP_W1 = 0, GP_B1 = 25600, GP_W2 = 25760, GP_B2 = 38560, GP_W34 = 38640, GP_B34 = 38800, GP_DEC = 38802;
// dec_pack_t (csrc/decode_bwd.cu): W1h^T [160][64] | W2^T [80][160] | W34 [80][2]
constexpr int GT_W1 = 0, GT_W2 = 10240, GT_W34 = 23040, GT_DEC = 23200;
constexpr int GP_ENC = 69 * 256, GP_ENC_T = 256 * 68, GP_POOL = 2240, GP_M = 65 * 65 /* M [64][65] | m0 [65] */, GP_MT = 65 * 64;
constexpr int GEN_PACK_TOTAL = GP_ENC + GP_ENC_T + GP_DEC + GT_DEC + GP_POOL + GP_M + GP_MT;

__device__ __forceinline__ float dotn(const float* __restrict__ a, int sa, const float* __restrict__ b, int sb, int n) {
    float s = 0.0f;
    for (int i = 0; i < n; ++i) s = fmaf(__ldg(a + (size_t)i * sa), __ldg(b + (size_t)i * sb), s);
    return s;
}

// value of enc_pack[k][n'] (n' = 4*unit + gate; torch row R = gate*64 + unit)
__device__ __forceinline__ float enc_pack_value(const GenParams& P, int k, int np) {
    const int R = (np & 3) * 64 + (np >> 2);
    if (k < 4) return dotn(P.p[G_WIH] + R * 64, 1, P.p[G_EMB_W] + k, 4, 64);              // (W_ih W_e)[R][k]
    if (k < 68) return __ldg(P.p[G_WHH] + R * 64 + (k - 4));
    return dotn(P.p[G_WIH] + R * 64, 1, P.p[G_EMB_B], 1, 64) + __ldg(P.p[G_BIH] + R) + __ldg(P.p[G_BHH] + R);
}

__device__ __forceinline__ float pool_tail(const GenParams& P, int m, int n) {      // [fc4_w | fc4_b] [64][65]
    return n < 64 ? __ldg(P.p[G_FC4_W] + m * 64 + n) : __ldg(P.p[G_FC4_B] + m);
}

__global__ void __launch_bounds__(256)
gen_pack_kernel(const __grid_constant__ GenParams P, float* __restrict__ enc_pack, float* __restrict__ enc_pack_t,
                float* __restrict__ dec_pack, float* __restrict__ dec_pack_t, float* __restrict__ pool_pack,
                float* __restrict__ pool_m /*[64][65] | m0[65]*/, float* __restrict__ pool_mt /*[65][64]*/) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < GEN_PACK_TOTAL; i += gridDim.x * blockDim.x) {
        int e = i;
        if (e < GP_ENC) { enc_pack[e] = enc_pack_value(P, e >> 8, e & 255); continue; }
        e -= GP_ENC;
        if (e < GP_ENC_T) { enc_pack_t[e] = enc_pack_value(P, e % 68, e / 68); continue; }
        e -= GP_ENC_T;
        if (e < GP_DEC) {
            float v;
            if (e < GP_B1) v = __ldg(P.p[G_W1] + (e % 160) * 160 + e / 160);                        // W1t[k][n] = W1[n][k]
            else if (e < GP_W2) v = __ldg(P.p[G_B1] + e - GP_B1);
            else if (e < GP_B2) { const int q = e - GP_W2; v = __ldg(P.p[G_W2] + (q % 80) * 160 + q / 80); }
            else if (e < GP_W34) v = __ldg(P.p[G_B2] + e - GP_B2);
            else if (e < GP_B34) { const int q = e - GP_W34; v = dotn(P.p[G_W4] + (q & 1) * 40, 1, P.p[G_W3] + (q >> 1), 80, 40); }
            else { const int o = e - GP_B34; v = dotn(P.p[G_W4] + o * 40, 1, P.p[G_B3], 1, 40) + __ldg(P.p[G_B4] + o); }
            dec_pack[e] = v;
            continue;
        }
        e -= GP_DEC;
        if (e < GT_DEC) {
            float v;
            if (e < GT_W2) v = __ldg(P.p[G_W1] + (e >> 6) * 160 + (e & 63));                         // W1h^T[n][k] = W1[n][k], k < 64
            else if (e < GT_W34) v = __ldg(P.p[G_W2] + e - GT_W2);                                   // W2^T = torch layout
            else { const int q = e - GT_W34; v = dotn(P.p[G_W4] + (q & 1) * 40, 1, P.p[G_W3] + (q >> 1), 80, 40); }
            dec_pack_t[e] = v;
            continue;
        }
        e -= GT_DEC;
        if (e < GP_POOL) {
            float v;
            if (e < 128) v = (e & 3) < 3 ? __ldg(P.p[G_FC0_W] + (e >> 2) * 3 + (e & 3)) : __ldg(P.p[G_FC0_B] + (e >> 2));
            else if (e < 128 + 2048) v = __ldg(P.p[G_FC2_W] + e - 128);
            else v = __ldg(P.p[G_FC2_B] + e - 2176);
            pool_pack[e] = v;
            continue;
        }
        e -= GP_POOL;
        if (e < GP_M) {
            const int k = e / 65, n = e % 65;
            float s = 0.0f;
            if (k < 64) { for (int m = 0; m < 64; ++m) s = fmaf(__ldg(P.p[G_ATT_W] + m * 64 + k), pool_tail(P, m, n), s); }
            else        { for (int m = 0; m < 64; ++m) s = fmaf(__ldg(P.p[G_ATT_B] + m), pool_tail(P, m, n), s); }
            pool_m[e] = s;
            continue;
        }
        e -= GP_M;
        {
            const int n = e >> 6, k = e & 63;                                                       // Mt[n][k] = M[k][n]
            float s = 0.0f;
            for (int m = 0; m < 64; ++m) s = fmaf(__ldg(P.p[G_ATT_W] + m * 64 + k), pool_tail(P, m, n), s);
            pool_mt[e] = s;
        }
    }
}

// Adjoint of the folds.  d_enc [69][256] (pack layout), d_w34 [80][2] | d_b34 [2], d_m [64][65] | d_m0 [65] come from the
// contraction kernel; every other generator gradient is written by sw_contract directly in parameter layout.
// Writes the gradients of: attention.W.{weight,bias}, fc.4.{weight,bias}, embed.{weight,bias}, lstm.{weight_ih,weight_hh,
// bias_ih,bias_hh}, fc1.4.{weight,bias}, fc1.5.{weight,bias}.  have_pool == 0 writes zeros to the four pooling tensors.
constexpr int GB_ATT_W = 0, GB_ATT_B = 4096, GB_FC4_W = 4160, GB_FC4_B = 8256, GB_EMB_W = 8320, GB_EMB_B = 8576, GB_WIH = 8640,
              GB_WHH = 25024, GB_BI = 41408, GB_W3 = 41664, GB_B3 = 44864, GB_W4 = 44904, GB_B4 = 44984, GB_TOTAL = 44986;

__global__ void __launch_bounds__(256)
gen_pack_bwd_kernel(const __grid_constant__ GenParams P, const __grid_constant__ GenGrads Gd, const float* __restrict__ d_enc,
                    const float* __restrict__ d_w34, const float* __restrict__ d_m, int have_pool) {
    const float* d_b34 = d_w34 + 160;
    const float* d_m0 = d_m + 64 * 65;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < GB_TOTAL; i += gridDim.x * blockDim.x) {
        if (i < GB_ATT_B) {                 // d att_w[m][k] = sum_n tail[m][n] dM[k][n]
            const int m = i >> 6, k = i & 63;
            float s = 0.0f;
            if (have_pool) for (int n = 0; n < 65; ++n) s = fmaf(pool_tail(P, m, n), __ldg(d_m + k * 65 + n), s);
            Gd.g[G_ATT_W][i] = s;
        } else if (i < GB_FC4_W) {          // d att_b[m] = sum_n tail[m][n] dm0[n]
            const int m = i - GB_ATT_B;
            float s = 0.0f;
            if (have_pool) for (int n = 0; n < 65; ++n) s = fmaf(pool_tail(P, m, n), __ldg(d_m0 + n), s);
            Gd.g[G_ATT_B][m] = s;
        } else if (i < GB_EMB_W) {          // d tail[m][n] = sum_k att_w[m][k] dM[k][n] + att_b[m] dm0[n]
            const int e = i - GB_FC4_W;
            const int m = e < 4096 ? e >> 6 : e - 4096, n = e < 4096 ? e & 63 : 64;
            float s = 0.0f;
            if (have_pool) {
                for (int k = 0; k < 64; ++k) s = fmaf(__ldg(P.p[G_ATT_W] + m * 64 + k), __ldg(d_m + k * 65 + n), s);
                s = fmaf(__ldg(P.p[G_ATT_B] + m), __ldg(d_m0 + n), s);
            }
            if (n < 64) Gd.g[G_FC4_W][m * 64 + n] = s; else Gd.g[G_FC4_B][m] = s;
        } else if (i < GB_EMB_B) {          // d W_e[m][k] = sum_R W_ih[R][m] dWx[k][n'(R)]
            const int e = i - GB_EMB_W, m = e >> 2, k = e & 3;
            float s = 0.0f;
            for (int R = 0; R < 256; ++R) s = fmaf(__ldg(P.p[G_WIH] + R * 64 + m), __ldg(d_enc + k * 256 + (R & 63) * 4 + (R >> 6)), s);
            Gd.g[G_EMB_W][e] = s;
        } else if (i < GB_WIH) {            // d b_e[m] = sum_R W_ih[R][m] d68[n'(R)]
            const int m = i - GB_EMB_B;
            float s = 0.0f;
            for (int R = 0; R < 256; ++R) s = fmaf(__ldg(P.p[G_WIH] + R * 64 + m), __ldg(d_enc + 68 * 256 + (R & 63) * 4 + (R >> 6)), s);
            Gd.g[G_EMB_B][m] = s;
        } else if (i < GB_WHH) {            // d W_ih[R][m] = sum_k dWx[k][n'] W_e[m][k] + d68[n'] b_e[m]
            const int e = i - GB_WIH, R = e >> 6, m = e & 63, np = (R & 63) * 4 + (R >> 6);
            float s = __ldg(d_enc + 68 * 256 + np) * __ldg(P.p[G_EMB_B] + m);
#pragma unroll
            for (int k = 0; k < 4; ++k) s = fmaf(__ldg(d_enc + k * 256 + np), __ldg(P.p[G_EMB_W] + m * 4 + k), s);
            Gd.g[G_WIH][e] = s;
        } else if (i < GB_BI) {             // d W_hh[R][k] = d_enc[4 + k][n']
            const int e = i - GB_WHH, R = e >> 6, k = e & 63;
            Gd.g[G_WHH][e] = __ldg(d_enc + (4 + k) * 256 + (R & 63) * 4 + (R >> 6));
        } else if (i < GB_W3) {             // d b_ih[R] = d b_hh[R] = d68[n']
            const int R = i - GB_BI;
            const float v = __ldg(d_enc + 68 * 256 + (R & 63) * 4 + (R >> 6));
            Gd.g[G_BIH][R] = v;
            Gd.g[G_BHH][R] = v;
        } else if (i < GB_B3) {             // d W3[m][k] = sum_o W4[o][m] dW34[k][o]
            const int e = i - GB_W3, m = e / 80, k = e % 80;
            Gd.g[G_W3][e] = fmaf(__ldg(P.p[G_W4] + m), __ldg(d_w34 + k * 2), __ldg(P.p[G_W4] + 40 + m) * __ldg(d_w34 + k * 2 + 1));
        } else if (i < GB_W4) {             // d b3[m] = sum_o W4[o][m] db34[o]
            const int m = i - GB_B3;
            Gd.g[G_B3][m] = fmaf(__ldg(P.p[G_W4] + m), __ldg(d_b34), __ldg(P.p[G_W4] + 40 + m) * __ldg(d_b34 + 1));
        } else if (i < GB_B4) {             // d W4[o][m] = sum_k dW34[k][o] W3[m][k] + db34[o] b3[m]
            const int e = i - GB_W4, o = e / 40, m = e % 40;
            float s = __ldg(d_b34 + o) * __ldg(P.p[G_B3] + m);
            for (int k = 0; k < 80; ++k) s = fmaf(__ldg(d_w34 + k * 2 + o), __ldg(P.p[G_W3] + m * 80 + k), s);
            Gd.g[G_W4][e] = s;
        } else {
            Gd.g[G_B4][i - GB_B4] = __ldg(d_b34 + (i - GB_B4));
        }
    }
}

// ---- discriminator ----
__global__ void __launch_bounds__(256)
disc_pack_kernel(const __grid_constant__ DiscParams P, int pred_dim, float* __restrict__ lstm_pack, float* __restrict__ lstm_pack_t,
                 float* __restrict__ heads_work) {
    const HeadsWork L(pred_dim);
    const int total = GP_ENC + GP_ENC_T + L.total;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int e = i;
        if (e < GP_ENC + GP_ENC_T) {
            int k, np;
            if (e < GP_ENC) { k = e >> 8; np = e & 255; } else { const int q = e - GP_ENC; k = q % 68; np = q / 68; }
            const int R = (np & 3) * 64 + (np >> 2);
            float v;
            if (k < 4) v = __ldg(P.p[0] + R * 4 + k);
            else if (k < 68) v = __ldg(P.p[1] + R * 64 + (k - 4));
            else v = __ldg(P.p[2] + R) + __ldg(P.p[3] + R);
            if (e < GP_ENC) lstm_pack[e] = v; else lstm_pack_t[e - GP_ENC] = v;
            continue;
        }
        e -= GP_ENC + GP_ENC_T;
        heads_work[e] = L.value(P.p + 4, e);
    }
}

}  // namespace sw

static int check_ptrs(const float* const* p, int n) {
    if (!p) return 0;
    for (int i = 0; i < n; ++i)
        if (!p[i]) return 0;
    return 1;
}

extern "C" int sw_gen_pack_sizes(int* enc_pack, int* enc_pack_t, int* dec_pack, int* dec_pack_t, int* pool_pack, int* pool_m,
                                 int* pool_mt) {
    if (!enc_pack || !enc_pack_t || !dec_pack || !dec_pack_t || !pool_pack || !pool_m || !pool_mt) return SW_ERR_ARG;
    *enc_pack = sw::GP_ENC; *enc_pack_t = sw::GP_ENC_T; *dec_pack = sw::GP_DEC; *dec_pack_t = sw::GT_DEC;
    *pool_pack = sw::GP_POOL; *pool_m = sw::GP_M; *pool_mt = sw::GP_MT;
    return SW_OK;
}

extern "C" int sw_gen_pack(const float* const* params22, float* enc_pack, float* enc_pack_t, float* dec_pack,
                           float* dec_pack_t, float* pool_pack, float* pool_m, float* pool_mt, void* stream) {
    if (!check_ptrs(params22, 22) || !enc_pack || !enc_pack_t || !dec_pack || !dec_pack_t || !pool_pack || !pool_m || !pool_mt)
        return SW_ERR_ARG;
    sw::GenParams P;
    for (int i = 0; i < 22; ++i) P.p[i] = params22[i];
    sw::gen_pack_kernel<<<(sw::GEN_PACK_TOTAL + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, enc_pack, enc_pack_t, dec_pack,
                                                                                           dec_pack_t, pool_pack, pool_m, pool_mt);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}

extern "C" int sw_gen_pack_bwd(const float* const* params22, float* const* grads22, const float* d_enc, const float* d_w34,
                               const float* d_m, int have_pool, void* stream) {
    if (!check_ptrs(params22, 22) || !grads22 || !d_enc || !d_w34 || (have_pool && !d_m)) return SW_ERR_ARG;
    sw::GenParams P;
    sw::GenGrads G;
    for (int i = 0; i < 22; ++i) {
        if (!grads22[i]) return SW_ERR_ARG;
        P.p[i] = params22[i];
        G.g[i] = grads22[i];
    }
    sw::gen_pack_bwd_kernel<<<(sw::GB_TOTAL + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, G, d_enc, d_w34, d_m, have_pool);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}

extern "C" int sw_disc_pack_sizes(int pred_dim, int* lstm_pack, int* lstm_pack_t, int* heads_work) {
    if (!lstm_pack || !lstm_pack_t || !heads_work || pred_dim <= 0 || pred_dim > SW_DISC_PMAX || (pred_dim & 1)) return SW_ERR_ARG;
    *lstm_pack = sw::GP_ENC; *lstm_pack_t = sw::GP_ENC_T; *heads_work = sw::HeadsWork(pred_dim).total;
    return SW_OK;
}

extern "C" int sw_disc_pack(const float* const* params20, int pred_dim, float* lstm_pack, float* lstm_pack_t,
                            float* heads_work, void* stream) {
    if (!check_ptrs(params20, 20) || !lstm_pack || !lstm_pack_t || !heads_work) return SW_ERR_ARG;
    if (pred_dim <= 0 || pred_dim > SW_DISC_PMAX || (pred_dim & 1)) return SW_ERR_ARG;
    sw::DiscParams P;
    for (int i = 0; i < 20; ++i) P.p[i] = params20[i];
    const int total = sw::GP_ENC + sw::GP_ENC_T + sw::HeadsWork(pred_dim).total;
    sw::disc_pack_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, pred_dim, lstm_pack, lstm_pack_t, heads_work);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
