// Backward of sw_pool_fwd (reference: the autograd graph of AttentionPooling.forward train.py:160-175
// through EmbedSocialFeatures.fc train.py:183-188; the social features themselves carry no gradient --
// they are a function of the observed data only).
//
// Given dL/dS [N][64]:
//   softmax:   dsigma_ij = a_ij * dS_i . (h_j - S_i)          (row sum folded: sum_j a_ij dS_i.h_j = dS_i.S_i)
//   pooling:   dh_j (direct) = sum_i a_ij dS_i
//   score:     sigma_ij = relu(a2pre_ij) . u_j + beta_j
//              du_j = sum_i dsigma_ij a2_ij ; dbeta_j = sum_i dsigma_ij
//              da2pre_ij = dsigma_ij u_j * [a2pre > 0] ; da1pre_ij = (da2pre_ij . P2) * [a1pre > 0]
// Everything indexed by j is a reduction over the rows i of the scene, so the kernel is organised
// by COLUMN: a group of G lanes owns agent j and strides over i (deterministic reduction order:
// fixed lane partition + shuffle butterfly, no atomics).  The pair activations are recomputed (never
// stored by the forward pass).  The per-pair vectors needed for the MLP weight gradients
// (a1, da1pre, da2pre, features) are written to HBM once and contracted by plain GEMMs on the host:
//   dP2 = G2^T . A1, db2 = sum G2, dP1 = G1^T . F    (autograd_path.py)
#include "sw_common.cuh"

namespace sw {

constexpr int PB_LD = 65;
constexpr int PB_SPAN_MAX = 640;    // staged agents per unit: x[4] + dS[65] floats each
constexpr int PBP_P1 = 0, PBP_P2 = 128, PBP_B2 = PBP_P2 + 64 * 32;

template <int G>
__global__ void __launch_bounds__(SW_THREADS, 1)
pool_bwd_kernel(const float* __restrict__ pool_pack, const float* __restrict__ x_last, const float* __restrict__ h,
                const float* __restrict__ ub, const float* __restrict__ dS, const float* __restrict__ tdot /*[N] dS_i.S_i, or null*/,
                const float* __restrict__ pooled /*[N][64] S (used when tdot is null)*/,
                const float* __restrict__ attn, const int* __restrict__ scene_offsets, const int* __restrict__ agent_scene,
                const long long* __restrict__ pair_offsets, float* __restrict__ dub /*[N][65]*/,
                float* __restrict__ dh_direct /*[N][64]*/, float* __restrict__ stA1 /*[P][32]*/, float* __restrict__ stG2 /*[P][64]*/,
                float* __restrict__ stG1 /*[P][32]*/, float* __restrict__ stF /*[P][4]*/, int n_agents, int a_cap, int span_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* W = reinterpret_cast<float*>(smem_raw);
    float* sx = W + 2240;                        // [span][4]
    float* sd = sx + span_cap * 4;               // [span][65]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    copy_f4(W, pool_pack, 2240);
    const int row0 = blockIdx.x * SW_ROWS;
    const int row1 = min(row0 + SW_ROWS, n_agents);
    const int span0 = scene_offsets[agent_scene[row0]];
    const int span1 = scene_offsets[agent_scene[row1 - 1] + 1];
    const int span = span1 - span0;
    const bool staged = span <= span_cap;
    const float* xp; const float* dp; int ldd;
    if (staged) {
        for (int i = tid; i < span; i += SW_THREADS)
            *reinterpret_cast<float4*>(sx + i * 4) = __ldg(reinterpret_cast<const float4*>(x_last) + span0 + i);
        for (int i = tid; i < span * 16; i += SW_THREADS) {
            const int a = i >> 4, q = i & 15;
            const float4 v = __ldg(reinterpret_cast<const float4*>(dS) + (size_t)(span0 + a) * 16 + q);
            float* d = sd + a * PB_LD + q * 4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        xp = sx; dp = sd; ldd = PB_LD;
    } else {
        xp = x_last + (size_t)span0 * 4; dp = dS + (size_t)span0 * SW_H; ldd = SW_H;
    }
    __syncthreads();

    constexpr int SLOTS = (SW_THREADS / 32) * (32 / G);
    const int gl = lane % G;
    const int slot = warp * (32 / G) + lane / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));

    for (int j = row0 + slot; j < row1; j += SLOTS) {
        const int sc = agent_scene[j];
        const int a = scene_offsets[sc], b = scene_offsets[sc + 1];
        const int A = b - a;
        if (A == 1) {
            for (int n = gl; n < 65; n += G) dub[(size_t)j * 65 + n] = 0.0f;
            for (int n = gl; n < SW_H; n += G) dh_direct[(size_t)j * SW_H + n] = 0.0f;
            continue;
        }
        const float4 xj = *reinterpret_cast<const float4*>(xp + (size_t)(j - span0) * 4);
        const float* uj = ub + (size_t)j * 65;
        const float* hj = h + (size_t)j * SW_H;
        const long long pbase = pair_offsets[sc] + (long long)(j - a) * A;
        float du[64];
#pragma unroll
        for (int n = 0; n < 64; ++n) du[n] = 0.0f;
        float dbeta = 0.0f;
        for (int i = a + gl; i < b; i += G) {
            const int ii = i - span0;
            const float4 xi = *reinterpret_cast<const float4*>(xp + (size_t)ii * 4);
            // ---- forward recompute of pair (i, j): D = x_i - x_j (train.py:232-238) ----
            const float dpx = xi.x - xj.x, dpy = xi.y - xj.y, dvx = xi.z - xj.z, dvy = xi.w - xj.w;
            const float dist = sqrtf(dpx * dpx + dpy * dpy);
            const float vnorm = sqrtf(xi.z * xi.z + xi.w * xi.w);
            const float bearing = (dpx * xi.z + dpy * xi.w) / (dist * vnorm + 1e-6f);
            const float ttca = -((dpx * dvx + dpy * dvy) / (dvx * dvx + dvy * dvy + 1e-6f));
            const float cx = dpx + ttca * dvx, cy = dpy + ttca * dvy;
            const float dca = sqrtf(cx * cx + cy * cy);
            float a1[32];
#pragma unroll
            for (int n = 0; n < 32; ++n) {
                const float4 w = *reinterpret_cast<const float4*>(W + PBP_P1 + n * 4);
                a1[n] = fmaxf(fmaf(w.x, dist, fmaf(w.y, bearing, fmaf(w.z, dca, w.w))), 0.0f);
            }
            // ---- dsigma_ij = a_ij * (dS_i . h_j - dS_i . S_i) ----
            const float* dsi = dp + (size_t)ii * ldd;
            float da = 0.0f;
            if (tdot) {
                da = -__ldg(tdot + i);
            } else {                       // dS_i . S_i evaluated here (64 FMAs against ~4 500 of the pair): no extra launch
                const float* si = pooled + (size_t)i * SW_H;
#pragma unroll 8
                for (int n = 0; n < SW_H; ++n) da = fmaf(-dsi[n], __ldg(si + n), da);
            }
#pragma unroll 8
            for (int n = 0; n < SW_H; ++n) da = fmaf(dsi[n], __ldg(hj + n), da);
            const float g = __ldg(attn + (size_t)i * a_cap + (j - a)) * da;
            dbeta += g;
            const long long pid = pbase + (i - a);
            float da1[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) da1[k] = 0.0f;
            float* g2row = stG2 + pid * 64;
#pragma unroll
            for (int n4 = 0; n4 < 16; ++n4) {
                float d2v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int n = n4 * 4 + e;
                    const float* w2 = W + PBP_P2 + n * 32;
                    float s2 = W[PBP_B2 + n];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 w = *reinterpret_cast<const float4*>(w2 + q * 4);
                        s2 = fmaf(w.x, a1[q * 4 + 0], s2); s2 = fmaf(w.y, a1[q * 4 + 1], s2);
                        s2 = fmaf(w.z, a1[q * 4 + 2], s2); s2 = fmaf(w.w, a1[q * 4 + 3], s2);
                    }
                    du[n] = fmaf(g, fmaxf(s2, 0.0f), du[n]);
                    const float d2 = (s2 > 0.0f) ? g * __ldg(uj + n) : 0.0f;
                    d2v[e] = d2;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 w = *reinterpret_cast<const float4*>(w2 + q * 4);
                        da1[q * 4 + 0] = fmaf(d2, w.x, da1[q * 4 + 0]); da1[q * 4 + 1] = fmaf(d2, w.y, da1[q * 4 + 1]);
                        da1[q * 4 + 2] = fmaf(d2, w.z, da1[q * 4 + 2]); da1[q * 4 + 3] = fmaf(d2, w.w, da1[q * 4 + 3]);
                    }
                }
                *reinterpret_cast<float4*>(g2row + n4 * 4) = make_float4(d2v[0], d2v[1], d2v[2], d2v[3]);
            }
            float* a1row = stA1 + pid * 32;
            float* g1row = stG1 + pid * 32;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                *reinterpret_cast<float4*>(a1row + q * 4) = make_float4(a1[q * 4], a1[q * 4 + 1], a1[q * 4 + 2], a1[q * 4 + 3]);
                *reinterpret_cast<float4*>(g1row + q * 4) =
                    make_float4(a1[q * 4] > 0.f ? da1[q * 4] : 0.f, a1[q * 4 + 1] > 0.f ? da1[q * 4 + 1] : 0.f,
                                a1[q * 4 + 2] > 0.f ? da1[q * 4 + 2] : 0.f, a1[q * 4 + 3] > 0.f ? da1[q * 4 + 3] : 0.f);
            }
            *reinterpret_cast<float4*>(stF + pid * 4) = make_float4(dist, bearing, dca, 1.0f);
        }
        // ---- reduce over the group's lanes (fixed order) ----
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) {
            dbeta += __shfl_xor_sync(gmask, dbeta, off);
#pragma unroll
            for (int n = 0; n < 64; ++n) du[n] += __shfl_xor_sync(gmask, du[n], off);
        }
#pragma unroll
        for (int n = 0; n < 64; ++n)
            if ((n % G) == gl) dub[(size_t)j * 65 + n] = du[n];
        if (gl == 0) dub[(size_t)j * 65 + 64] = dbeta;
        // ---- direct term: dh_j = sum_i a_ij dS_i ----
        for (int n = gl; n < SW_H; n += G) {
            float acc = 0.0f;
            for (int i = a; i < b; ++i)
                acc = fmaf(__ldg(attn + (size_t)i * a_cap + (j - a)), dp[(size_t)(i - span0) * ldd + n], acc);
            dh_direct[(size_t)j * SW_H + n] = acc;
        }
    }
}

}  // namespace sw

extern "C" int sw_pool_bwd(const float* pool_pack, const float* x_last, const float* h, const float* ub,
                           const float* dS, const float* tdot, const float* pooled, const float* attn, const int* scene_offsets,
                           const int* agent_scene, const long long* pair_offsets, float* dub, float* dh_direct,
                           float* st_a1, float* st_g2, float* st_g1, float* st_f, int n_agents, int max_scene,
                           void* stream) {
    if (!pool_pack || !x_last || !h || !ub || !dS || (!tdot && !pooled) || !attn || !scene_offsets || !agent_scene || !pair_offsets ||
        !dub || !dh_direct || !st_a1 || !st_g2 || !st_g1 || !st_f)
        return SW_ERR_ARG;
    if (n_agents <= 0 || max_scene <= 0) return SW_ERR_ARG;
    const int a_cap = (max_scene + 3) & ~3;
    const int G = max_scene <= 8 ? 8 : (max_scene <= 16 ? 16 : 32);
    int span_cap = SW_ROWS + 2 * (max_scene - 1);
    if (span_cap > sw::PB_SPAN_MAX) span_cap = (max_scene <= sw::PB_SPAN_MAX) ? sw::PB_SPAN_MAX : 0;
    const size_t smem = (size_t)(2240 + span_cap * (4 + sw::PB_LD)) * 4;
    const int grid = (n_agents + SW_ROWS - 1) / SW_ROWS;
    cudaStream_t st = (cudaStream_t)stream;
#define SW_POOLB_LAUNCH(GG)                                                                                          \
    do {                                                                                                             \
        SW_SET_MAX_SMEM(sw::pool_bwd_kernel<GG>, \
                                         (int)smem);                                                                \
        sw::pool_bwd_kernel<GG><<<grid, SW_THREADS, smem, st>>>(pool_pack, x_last, h, ub, dS, tdot, pooled, attn,     \
                                                                 scene_offsets, agent_scene, pair_offsets, dub,      \
                                                                 dh_direct, st_a1, st_g2, st_g1, st_f, n_agents,     \
                                                                 a_cap, span_cap);                                   \
    } while (0)
    if (G == 8) SW_POOLB_LAUNCH(8);
    else if (G == 16) SW_POOLB_LAUNCH(16);
    else SW_POOLB_LAUNCH(32);
#undef SW_POOLB_LAUNCH
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
