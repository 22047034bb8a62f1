// Fused pairwise social features + embedding MLP + attention pooling, per scene:
//   SocialFeatures / BearingMTX / DCA_MTX   reference train.py:208-241
//   EmbedSocialFeatures.fc (3->32->64->64)   reference train.py:178-189
//   AttentionPooling.forward                 reference train.py:153-175
// The reference materialises [N,N,3] and [N,N,64] over the WHOLE mini-batch (cross-scene pairs
// included, then discarded) and loops over agents in Python.  Here nothing pairwise ever reaches
// HBM: per ordered pair (i, j) of one scene the 3 features, the first two MLP layers and the
// attention score are evaluated in registers; the third MLP layer is folded into the score
// (packing.py):   sigma_ij = fc4(a2_ij) . Wh_j = a2_ij . u_j + beta_j,
//                 u_j = W3^T (W h_j + b_W),  beta_j = b3 . (W h_j + b_W)
// u, beta come from one [N,64]x[64,65] GEMM on the encoder output (host side, packing.py).
// Then sigma_ii = -1000 (train.py:170), softmax over the scene, S_i = sum_j a_ij h_j (RAW h, :173);
// scenes with one agent give S = 0 (train.py:165).
//
// Work unit = 32 consecutive agent rows; the agents of every scene those rows belong to (the
// "span") are staged in shared memory: last state x[.,4], h[.,64], u[.,64], beta, with coalesced
// float4 loads.  A group of G lanes (G = 8/16/32 by the largest scene) owns one row i and strides
// over j.  Algorithmic HBM traffic: 16 + 256 + 256 + 4 B read, 256 B written per agent.
#include "sw_common.cuh"

namespace sw {

constexpr int POOL_LD = 65;          // padded row stride of h / u in shared memory: lane j, column n -> bank (j+n)%32
constexpr int POOL_SPAN_MAX = 320;   // agents staged per unit (166 KB); larger spans read h/u through L1/L2
constexpr int POOL_A_MAX = 2048;     // largest scene the score buffers are sized for
// pool_pack layout: P1[32][4] = (w_dist, w_bearing, w_dca, bias) per hidden unit | P2[64][32] | b2[64]
constexpr int PP_P1 = 0, PP_P2 = 128, PP_B2 = PP_P2 + 64 * 32, PP_TOTAL = PP_B2 + 64;

__device__ __forceinline__ float pair_score(const float4 xi, const float4 xj, const float* __restrict__ W,
                                            const float* __restrict__ uj /*stride 1*/, float betaj) {
    // D[i][j] = x_i - x_j (train.py:232-234)
    const float dpx = xi.x - xj.x, dpy = xi.y - xj.y, dvx = xi.z - xj.z, dvy = xi.w - xj.w;
    const float dist = sqrtf(dpx * dpx + dpy * dpy);
    const float vnorm = sqrtf(xi.z * xi.z + xi.w * xi.w);
    const float bearing = (dpx * xi.z + dpy * xi.w) / (dist * vnorm + 1e-6f);          // :224-225
    const float ttca = -((dpx * dvx + dpy * dvy) / (dvx * dvx + dvy * dvy + 1e-6f));   // :211-213
    const float cx = dpx + ttca * dvx, cy = dpy + ttca * dvy;
    const float dca = sqrtf(cx * cx + cy * cy);                                          // :214-217
    float a1[32];
#pragma unroll
    for (int n = 0; n < 32; ++n) {
        const float4 w = *reinterpret_cast<const float4*>(W + PP_P1 + n * 4);
        a1[n] = fmaxf(fmaf(w.x, dist, fmaf(w.y, bearing, fmaf(w.z, dca, w.w))), 0.0f);
    }
    float sigma = betaj;
#pragma unroll 2
    for (int n = 0; n < 64; ++n) {
        const float* w2 = W + PP_P2 + n * 32;
        float a = W[PP_B2 + n];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 w = *reinterpret_cast<const float4*>(w2 + q * 4);
            a = fmaf(w.x, a1[q * 4 + 0], a); a = fmaf(w.y, a1[q * 4 + 1], a);
            a = fmaf(w.z, a1[q * 4 + 2], a); a = fmaf(w.w, a1[q * 4 + 3], a);
        }
        sigma = fmaf(fmaxf(a, 0.0f), uj[n], sigma);
    }
    return sigma;
}

template <int G>
__global__ void __launch_bounds__(SW_THREADS)
pool_fwd_kernel(const float* __restrict__ pool_pack, const float* __restrict__ x_last, const float* __restrict__ h,
                const float* __restrict__ ub /*[N][65]: u | beta*/, const int* __restrict__ scene_offsets,
                const int* __restrict__ agent_scene, float* __restrict__ pooled, float* __restrict__ attn_out,
                int n_agents, int a_cap /*score buffer length per slot*/, int span_cap /*0: do not stage*/) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* W = reinterpret_cast<float*>(smem_raw);                 // PP_TOTAL
    constexpr int SLOTS = (SW_THREADS / 32) * (32 / G);
    float* sig = W + PP_TOTAL;                                      // [SLOTS][a_cap]
    float* sx = sig + SLOTS * a_cap;                                // [span][4]   (16 B aligned: a_cap % 4 == 0)
    float* sh = sx + span_cap * 4;                                  // [span][65]
    float* su = sh + span_cap * POOL_LD;                            // [span][65]  (col 64 = beta)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    copy_f4(W, pool_pack, PP_TOTAL);

    const int row0 = blockIdx.x * SW_ROWS;
    const int row1 = min(row0 + SW_ROWS, n_agents);
    const int span0 = scene_offsets[agent_scene[row0]];
    const int span1 = scene_offsets[agent_scene[row1 - 1] + 1];
    const int span = span1 - span0;
    const bool staged = span <= span_cap;
    const float* hp; const float* up; const float* xp; int ld, ldu;
    if (staged) {
        for (int i = tid; i < span; i += SW_THREADS)
            *reinterpret_cast<float4*>(sx + i * 4) = __ldg(reinterpret_cast<const float4*>(x_last) + span0 + i);
        for (int i = tid; i < span * 16; i += SW_THREADS) {         // 16 float4 per agent row of h
            const int a = i >> 4, q = i & 15;
            const float4 v = __ldg(reinterpret_cast<const float4*>(h) + (size_t)(span0 + a) * 16 + q);
            float* d = sh + a * POOL_LD + q * 4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        for (int i = tid; i < span * 65; i += SW_THREADS) su[i] = __ldg(ub + (size_t)span0 * 65 + i);
        hp = sh; up = su; xp = sx; ld = POOL_LD; ldu = POOL_LD;
    } else {
        hp = h + (size_t)span0 * SW_H; up = ub + (size_t)span0 * 65; xp = x_last + (size_t)span0 * 4; ld = SW_H; ldu = 65;
    }
    __syncthreads();

    const int gl = lane % G;                          // lane inside the group
    const int slot = warp * (32 / G) + lane / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
    float* my_sig = sig + slot * a_cap;

    for (int i = row0 + slot; i < row1; i += SLOTS) {
        const int sc = agent_scene[i];
        const int a = scene_offsets[sc], b = scene_offsets[sc + 1];
        const int A = b - a;
        if (A == 1) {                                  // train.py:165
            for (int n = gl; n < SW_H; n += G) pooled[(size_t)i * SW_H + n] = 0.0f;
            continue;
        }
        const float4 xi = *reinterpret_cast<const float4*>(xp + (size_t)(i - span0) * 4);
        float mx = -3.0e38f;
        for (int j = a + gl; j < b; j += G) {
            const int jj = j - span0;
            const float4 xj = *reinterpret_cast<const float4*>(xp + (size_t)jj * 4);
            float sg = pair_score(xi, xj, W, up + (size_t)jj * ldu, up[(size_t)jj * ldu + 64]);
            if (j == i) sg = -1000.0f;                 // train.py:170
            my_sig[j - a] = sg;
            mx = fmaxf(mx, sg);
        }
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, off));
        float sum = 0.0f;
        for (int j = gl; j < A; j += G) {
            const float e = expf(my_sig[j] - mx);
            my_sig[j] = e;
            sum += e;
        }
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(gmask, sum, off);
        __syncwarp(gmask);
        for (int j = gl; j < A; j += G) my_sig[j] = my_sig[j] / sum;      // attention weights (softmax, :172)
        __syncwarp(gmask);
        if (attn_out)                                                     // saved for the backward pass
            for (int j = gl; j < A; j += G) attn_out[(size_t)i * a_cap + j] = my_sig[j];
        for (int n = gl; n < SW_H; n += G) {
            float acc = 0.0f;
            for (int j = 0; j < A; ++j) acc = fmaf(my_sig[j], hp[(size_t)(a - span0 + j) * ld + n], acc);
            pooled[(size_t)i * SW_H + n] = acc;
        }
        __syncwarp(gmask);
    }
}

}  // namespace sw

// scene_offsets: [n_scenes + 1] ascending agent offsets (the reference's `sub_batches` [start,end)
// pairs flattened, train.py:461); agent_scene: [N] scene index of every agent; ub: [N][65] = u | beta.
// attn (optional, may be null): [N][max_scene] softmax weights, row stride = max_scene rounded up to 4.
extern "C" int sw_pool_fwd(const float* pool_pack, const float* x_last, const float* h, const float* ub,
                           const int* scene_offsets, const int* agent_scene, float* pooled, float* attn,
                           int n_agents, int max_scene, void* stream) {
    if (!pool_pack || !x_last || !h || !ub || !scene_offsets || !agent_scene || !pooled) return SW_ERR_ARG;
    if (n_agents <= 0 || max_scene <= 0) return SW_ERR_ARG;
    if (max_scene > sw::POOL_A_MAX) return SW_ERR_UNSUPPORTED;
    const int a_cap = (max_scene + 3) & ~3;
    const int G = max_scene <= 8 ? 8 : (max_scene <= 16 ? 16 : 32);
    const int slots = (SW_THREADS / 32) * (32 / G);
    // a unit of 32 rows touches at most (max_scene - 1) extra agents on each side: stage that worst
    // case when it fits, else up to POOL_SPAN_MAX (re-checked per unit in the kernel), else nothing
    int span_cap = SW_ROWS + 2 * (max_scene - 1);
    if (span_cap > sw::POOL_SPAN_MAX) span_cap = (max_scene <= sw::POOL_SPAN_MAX) ? sw::POOL_SPAN_MAX : 0;
    const size_t smem = (size_t)(sw::PP_TOTAL + slots * a_cap + span_cap * (4 + 2 * sw::POOL_LD)) * 4;
    const int grid = (n_agents + SW_ROWS - 1) / SW_ROWS;
    cudaStream_t st = (cudaStream_t)stream;
#define SW_POOL_LAUNCH(GG)                                                                                         \
    do {                                                                                                           \
        SW_SET_MAX_SMEM(sw::pool_fwd_kernel<GG>, \
                                         (int)smem);                                                              \
        sw::pool_fwd_kernel<GG><<<grid, SW_THREADS, smem, st>>>(pool_pack, x_last, h, ub, scene_offsets,            \
                                                                 agent_scene, pooled, attn, n_agents, a_cap, span_cap); \
    } while (0)
    if (G == 8) SW_POOL_LAUNCH(8);
    else if (G == 16) SW_POOL_LAUNCH(16);
    else SW_POOL_LAUNCH(32);
#undef SW_POOL_LAUNCH
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}

extern "C" int sw_pool_pack_floats(void) { return sw::PP_TOTAL; }
