// Discriminator FC heads: forward + LSGAN / InfoGAN losses + backward in ONE launch (training step).
//
// Reference: Discriminator.forward train.py:300-309 (obsv_encoder_fc 64->32->32, pred_encoder (n_next*4)->32->32,
// classifier 64->32->1, latent_decoder 64->32->2, LeakyReLU(0.2) inside each block, train.py:281-292), the three
// nn.MSELoss terms of the D step (train.py:484-493) or the two of the G step (train.py:514-521) and the part of
// d_loss.backward() / g_loss.backward() (train.py:495,538) that runs through these layers.
//
// mode 0 (D step): a 32-row tile = 16 agents x {fake row (pred_hat_4d), real row (pred_4d formed on the fly from the
//   ground-truth positions, get_traj_4d train.py:135-137)}; both rows share the agent's observation code, so
//   dL/d(obsv code) of the two branches is summed inside the tile.  Losses: fake -> (label - zeros)^2, info; real -> (label - ones)^2.
// mode 1 (G step): 32 agents, fake rows only, target `ones`; emits dL/d(pred_hat_4d) for the generator's backward pass.
// The discriminator's own gradients from g_loss.backward() are never used by the reference (D.zero_grad() precedes every
// D update, train.py:478,503), so mode 1 writes no parameter-gradient records.
//
// Every dense layer is an FFMA register-tile contraction on k-major shared-memory operands (sw_common.cuh); the weights
// arrive in the working layout of csrc/disc_layout.cuh.  The activation / gradient records leave as TILE IMAGES
// (literally the shared-memory buffers), which sw_contract turns into all 16 parameter gradients.
// Loss terms leave as per-tile partial sums (fixed order; summed by sw_train_stats).
#include "sw_common.cuh"
#include "disc_layout.cuh"

namespace sw {

// out[n][r] = act(bias[n] + sum_k X[k][r] Wt[k][n]),  n < 32
__device__ __forceinline__ void dense32(float* __restrict__ out, const float* __restrict__ X, const float* __restrict__ Wt, int K,
                                        const float* __restrict__ bias, bool act, const LaneMap<2>& lm) {
    float acc[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    fma_tile<2, 2>(acc, X, Wt, HW_LD32, K, lm);
    ksplit_reduce<2, 2>(acc);
    const int j = lm.ks, n = lm.cg * 2 + j;
    const float b = bias ? bias[n] : 0.0f;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = (j == 0 ? acc[i][0] : acc[i][1]) + b;
        if (act) v[i] = lrelu02(v[i]);
    }
    *reinterpret_cast<float4*>(out + n * SW_ROWS + lm.rg * 4) = make_float4(v[0], v[1], v[2], v[3]);
}

// out[k][r] = (sum_j W[j][k] D[j][r]) * lrelu'(A[k][r]),  k < 32  (A == nullptr: no activation derivative)
__device__ __forceinline__ void dense32_t(float* __restrict__ out, const float* __restrict__ D, const float* __restrict__ W, int J,
                                          const float* __restrict__ A, const LaneMap<2>& lm) {
    float acc[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    fma_tile<2, 2>(acc, D, W, HW_LD32, J, lm);
    ksplit_reduce<2, 2>(acc);
    const int j = lm.ks, k = lm.cg * 2 + j;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (j == 0 ? acc[i][0] : acc[i][1]);
    if (A) {
        const float4 a = *reinterpret_cast<const float4*>(A + k * SW_ROWS + lm.rg * 4);
        v[0] = a.x > 0.f ? v[0] : 0.2f * v[0]; v[1] = a.y > 0.f ? v[1] : 0.2f * v[1];
        v[2] = a.z > 0.f ? v[2] : 0.2f * v[2]; v[3] = a.w > 0.f ? v[3] : 0.2f * v[3];
    }
    *reinterpret_cast<float4*>(out + k * SW_ROWS + lm.rg * 4) = make_float4(v[0], v[1], v[2], v[3]);
}

// out[k][r] = sum_j W[j][k] D[j][r],  k < 64, W rows of HW_LD64 floats
__device__ __forceinline__ void dense64_t(float* __restrict__ out, const float* __restrict__ D, const float* __restrict__ W, int J,
                                          const LaneMap<2>& lm) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = 0.0f;
    fma_tile<4, 2>(acc, D, W, HW_LD64, J, lm);
    ksplit_reduce<4, 2>(acc);
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if ((q & 1) == lm.ks)
            *reinterpret_cast<float4*>(out + (lm.cg * 4 + q) * SW_ROWS + lm.rg * 4) = make_float4(acc[0][q], acc[1][q], acc[2][q], acc[3][q]);
}

__global__ void __launch_bounds__(SW_THREADS, 1)
disc_step_kernel(const float* __restrict__ heads_work, int P, int mode, const float* __restrict__ obsv_h,
                 const float* __restrict__ pred_fake /*[N][P]*/, const float* __restrict__ pred_pos /*[N][P/4][2]*/,
                 const float* __restrict__ obsv_pos /*[N][n_past][2]*/, int n_past, const float* __restrict__ noise, int noise_ld,
                 const float* __restrict__ targets /*[2]: zeros value, ones value*/, float inv_n, float info_w,
                 float* __restrict__ d_h, float* __restrict__ d_pred, float* __restrict__ x_img, float* __restrict__ g_img,
                 float* __restrict__ loss_part, float* __restrict__ label_out, float* __restrict__ code_out, int n_agents,
                 int n_tiles) {
    extern __shared__ __align__(16) float sm[];
    const HeadsWork L(P);
    float* W = sm;
    float* sX = W + L.total;                         // X image: (256 + P) rows
    float* sG = sX + hx_rows(P) * SW_ROWS;           // G image: 196 rows
    float* sD = sG + HG_ROWS * SW_ROWS;              // dboth [64][32] | dh [64][32] | pad; scratch of the transposing loads
    float* sOut = sD + 4352;                         // label [32] | code [2][32] | targets etc.
    const int tid = threadIdx.x;
    copy_f4(W, heads_work, L.total);
    const LaneMap<2> lm;
    const int T = P >> 2;
    const int agents_per_tile = mode == 0 ? 16 : 32;
    const float t_zero = __ldg(targets), t_one = __ldg(targets + 1);

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int a0 = tile * agents_per_tile;
        auto agent_of = [&](int r) { return a0 + (mode == 0 ? (r & 15) : r); };
        auto row_valid = [&](int r) { return agent_of(r) < n_agents; };
        __syncthreads();
        // ---- inputs, k-major: h -> X rows [0,64); pred -> X rows [96, 96+P) ----
        {
            float* scratch = sD;                                           // [32][65]
            for (int i = tid; i < SW_ROWS * 64; i += SW_THREADS) {
                const int r = i >> 6, k = i & 63;
                scratch[r * 65 + k] = row_valid(r) ? __ldg(obsv_h + (size_t)agent_of(r) * 64 + k) : 0.0f;
            }
            __syncthreads();
            for (int i = tid; i < SW_ROWS * 64; i += SW_THREADS) {
                const int k = i >> 5, r = i & 31;
                sX[(HX_H + k) * SW_ROWS + r] = scratch[r * 65 + k];
            }
            __syncthreads();
            const int ldp = P + 1;                                         // [32][P + 1]
            for (int i = tid; i < SW_ROWS * P; i += SW_THREADS) {
                const int r = i / P, p = i - r * P;
                float v = 0.0f;
                if (row_valid(r)) {
                    const int ag = agent_of(r);
                    if (mode == 1 || r < 16) {
                        v = __ldg(pred_fake + (size_t)ag * P + p);
                    } else {                                               // real branch: (p_t, p_t - p_{t-1}), p_{-1} = last observation
                        const int t = p >> 2, c = p & 3;
                        const float cur = __ldg(pred_pos + ((size_t)ag * T + t) * 2 + (c & 1));
                        if (c < 2) v = cur;
                        else {
                            const float prev = t > 0 ? __ldg(pred_pos + ((size_t)ag * T + t - 1) * 2 + (c & 1))
                                                     : __ldg(obsv_pos + ((size_t)ag * n_past + n_past - 1) * 2 + (c & 1));
                            v = cur - prev;
                        }
                    }
                }
                scratch[r * ldp + p] = v;
            }
            __syncthreads();
            for (int i = tid; i < SW_ROWS * P; i += SW_THREADS) {
                const int p = i >> 5, r = i & 31;
                sX[(HX_PRED + p) * SW_ROWS + r] = scratch[r * ldp + p];
            }
            __syncthreads();
        }
        float* xh = sX + HX_H * SW_ROWS;
        float* xo1 = sX + HX_O1 * SW_ROWS;
        float* xpred = sX + HX_PRED * SW_ROWS;
        float* xp1 = sX + hx_p1(P) * SW_ROWS;
        float* xboth = sX + hx_both(P) * SW_ROWS;
        float* xc1 = sX + hx_c1(P) * SW_ROWS;
        float* xl1 = sX + hx_l1(P) * SW_ROWS;
        // ---- forward ----
        dense32(xo1, xh, W + L.f_wo1t, 64, W + L.v_bo1, true, lm);
        dense32(xp1, xpred, W + L.f_wp1t, P, W + L.v_bp1, true, lm);
        __syncthreads();
        dense32(xboth, xo1, W + L.f_wo2t, 32, W + L.v_bo2, false, lm);
        dense32(xboth + 32 * SW_ROWS, xp1, W + L.f_wp2t, 32, W + L.v_bp2, false, lm);
        __syncthreads();
        dense32(xc1, xboth, W + L.f_wc1t, 64, W + L.v_bc1, true, lm);
        dense32(xl1, xboth, W + L.f_wl1t, 64, W + L.v_bl1, true, lm);
        __syncthreads();
        // ---- output layer, losses, output gradients: thread = (row, output o in {label, code0, code1}) ----
        if (tid < 96) {
            const int r = tid & 31, o = tid >> 5;
            const float* act = o == 0 ? xc1 : xl1;
            const float* w = o == 0 ? W + L.v_wc2 : W + L.v_wl2 + (o - 1) * 32;
            float y = o == 0 ? W[L.v_bc2] : W[L.v_bl2 + o - 1];
#pragma unroll 8
            for (int k = 0; k < 32; ++k) y = fmaf(w[k], act[k * SW_ROWS + r], y);
            const bool valid = row_valid(r);
            const bool real = mode == 0 && r >= 16;
            float err = 0.0f, grad = 0.0f;
            if (valid) {
                const int ag = agent_of(r);
                if (o == 0) {
                    err = y - ((mode == 0 && !real) ? t_zero : t_one);
                    grad = 2.0f * err * inv_n;
                    if (label_out) label_out[(size_t)(mode == 0 ? (real ? n_agents + ag : ag) : ag)] = y;
                } else if (!real) {
                    err = y - __ldg(noise + (size_t)ag * noise_ld + (o - 1));
                    grad = info_w * err * inv_n;                           // mean over N x 2 elements: 2 err / (2 N)
                    if (code_out) code_out[(size_t)ag * 2 + (o - 1)] = y;
                }
            }
            sG[(o == 0 ? HG_DLABEL : HG_DCODE + o - 1) * SW_ROWS + r] = grad;
            // squared-error partial sums over the 32 rows (fixed butterfly order): slots fake | real | info
            float sq_a = (o == 0 && !real) ? err * err : 0.0f;             // fake (or fooling) label term
            float sq_b = (o == 0 && real) ? err * err : 0.0f;              // real label term
            float sq_c = o > 0 ? err * err : 0.0f;                         // info term of this code component
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                sq_a += __shfl_xor_sync(0xffffffffu, sq_a, off);
                sq_b += __shfl_xor_sync(0xffffffffu, sq_b, off);
                sq_c += __shfl_xor_sync(0xffffffffu, sq_c, off);
            }
            if (r == 0) {
                if (o == 0) { sOut[0] = sq_a; sOut[1] = sq_b; } else sOut[1 + o] = sq_c;
            }
        }
        if (tid >= 96 && tid < 128) sG[(HG_ROWS - 1) * SW_ROWS + (tid - 96)] = 0.0f;     // padding row of the G image
        __syncthreads();
        if (tid == 0 && loss_part) {
            float* lp = loss_part + (size_t)tile * 4;
            lp[0] = sOut[0]; lp[1] = sOut[1]; lp[2] = sOut[2] + sOut[3]; lp[3] = 0.0f;
        }
        // ---- backward: pre-activation gradients of c1 / l1 (elementwise), then the transposed contractions ----
        for (int i = tid; i < 32 * SW_ROWS; i += SW_THREADS) {
            const int j = i >> 5, r = i & 31;
            const float dl = sG[HG_DLABEL * SW_ROWS + r];
            const float a = W[L.v_wc2 + j] * dl;
            sG[(HG_DC1 + j) * SW_ROWS + r] = xc1[i] > 0.0f ? a : 0.2f * a;
            const float b = fmaf(W[L.v_wl2 + j], sG[HG_DCODE * SW_ROWS + r], W[L.v_wl2 + 32 + j] * sG[(HG_DCODE + 1) * SW_ROWS + r]);
            sG[(HG_DL1 + j) * SW_ROWS + r] = xl1[i] > 0.0f ? b : 0.2f * b;
        }
        __syncthreads();
        // dboth = Wc1^T dc1 + Wl1^T dl1: stage [dc1 ; dl1] contiguously (the G image has dlabel between them)
        float* dcl = sD + 2048 + 64;            // [64][32] scratch inside sD (dh region, written later)
        for (int i = tid; i < 32 * SW_ROWS; i += SW_THREADS) {
            dcl[i] = sG[HG_DC1 * SW_ROWS + i];
            dcl[32 * SW_ROWS + i] = sG[HG_DL1 * SW_ROWS + i];
        }
        __syncthreads();
        // dboth is not contiguous in the G image (doc at rows [32,64), dpc at [96,128)): compute it in sD, then split
        dense64_t(sD, dcl, W + L.b_wcl, 64, lm);                           // sD[0:64][32] = dboth
        __syncthreads();
        for (int i = tid; i < 32 * SW_ROWS; i += SW_THREADS) {
            sG[HG_DOC * SW_ROWS + i] = sD[i];
            sG[HG_DPC * SW_ROWS + i] = sD[32 * SW_ROWS + i];
        }
        __syncthreads();
        dense32_t(sG + HG_DO1 * SW_ROWS, sG + HG_DOC * SW_ROWS, W + L.b_wo2, 32, xo1, lm);
        dense32_t(sG + HG_DP1 * SW_ROWS, sG + HG_DPC * SW_ROWS, W + L.b_wp2, 32, xp1, lm);
        __syncthreads();
        if (mode == 0) {
            // dL/d(obsv code) = Wo1^T do1, fake + real rows of the same agent summed
            float* sdh = sD;                    // dboth is consumed
            dense64_t(sdh, sG + HG_DO1 * SW_ROWS, W + L.b_wo1, 32, lm);
            __syncthreads();
            if (d_h)
                for (int i = tid; i < 16 * 64; i += SW_THREADS) {
                    const int a = i & 15, k = i >> 4;
                    if (a0 + a < n_agents) d_h[(size_t)(a0 + a) * 64 + k] = sdh[k * SW_ROWS + a] + sdh[k * SW_ROWS + a + 16];
                }
            if (x_img) store_image(x_img + (size_t)tile * hx_rows(P) * SW_ROWS, sX, hx_rows(P) * SW_ROWS);
            if (g_img) store_image(g_img + (size_t)tile * HG_ROWS * SW_ROWS, sG, HG_ROWS * SW_ROWS);
        } else if (d_pred) {
            const float* wp1 = W + L.b_wp1;
            const float* dp1 = sG + HG_DP1 * SW_ROWS;
            for (int i = tid; i < P * SW_ROWS; i += SW_THREADS) {
                const int p = i >> 5, r = i & 31;
                float s = 0.0f;
#pragma unroll 8
                for (int j = 0; j < 32; ++j) s = fmaf(wp1[j * L.P4 + p], dp1[j * SW_ROWS + r], s);
                if (a0 + r < n_agents) d_pred[(size_t)(a0 + r) * P + p] = s;
            }
        }
    }
}

__host__ int disc_step_smem_floats(int P) { return HeadsWork(P).total + (hx_rows(P) + HG_ROWS) * SW_ROWS + 4352 + 64; }

}  // namespace sw

extern "C" int sw_disc_step_image_rows(int pred_dim, int* x_rows, int* g_rows) {
    if (!x_rows || !g_rows || pred_dim <= 0 || pred_dim > SW_DISC_PMAX || (pred_dim & 3)) return SW_ERR_ARG;
    *x_rows = sw::hx_rows(pred_dim);
    *g_rows = sw::HG_ROWS;
    return SW_OK;
}

extern "C" int sw_disc_step(const float* heads_work, int pred_dim, int mode, const float* obsv_h, const float* pred_fake,
                            const float* pred_pos, const float* obsv_pos, int n_past, const float* noise, int noise_ld,
                            const float* targets, float inv_n, float info_w, float* d_h, float* d_pred, float* x_img,
                            float* g_img, float* loss_part, float* label_out, float* code_out, int n_agents, int sm_count,
                            void* stream) {
    if (!heads_work || !obsv_h || !pred_fake || !noise || !targets) return SW_ERR_ARG;
    if (mode != 0 && mode != 1) return SW_ERR_ARG;
    if (mode == 0 && (!pred_pos || !obsv_pos || n_past <= 0)) return SW_ERR_ARG;
    if (n_agents <= 0 || sm_count <= 0 || noise_ld < 2) return SW_ERR_ARG;
    if (pred_dim <= 0 || pred_dim > SW_DISC_PMAX || (pred_dim & 3)) return SW_ERR_UNSUPPORTED;
    const int per = mode == 0 ? 16 : 32;
    const int tiles = (n_agents + per - 1) / per;
    const int smem = sw::disc_step_smem_floats(pred_dim) * 4;
    SW_SET_MAX_SMEM(sw::disc_step_kernel, smem);
    const int grid = tiles < sm_count ? tiles : sm_count;
    sw::disc_step_kernel<<<grid, SW_THREADS, smem, (cudaStream_t)stream>>>(
        heads_work, pred_dim, mode, obsv_h, pred_fake, pred_pos, obsv_pos, n_past, noise, noise_ld, targets, inv_n, info_w, d_h,
        d_pred, x_img, g_img, loss_part, label_out, code_out, n_agents, tiles);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
