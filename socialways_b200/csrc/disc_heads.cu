// Discriminator FC heads, forward and backward, one thread per trajectory
// (reference Discriminator.forward train.py:300-309 after the observation LSTM:
//   obsv_encoder_fc 64->32->32, pred_encoder (n_next*4)->32->32, classifier 64->32->1, latent_decoder 64->32->n_latent,
//   LeakyReLU(0.2) between the two Linear layers of each block, train.py:281-292).
//
// Forward: all 8 Linear layers + the concatenation in one launch; every weight row is read from shared
// memory as a warp-wide broadcast (lanes = rows of the batch).  It also writes, per row, the activation
// record  X = [h | o1 | pred | p1 | both | c1 | l1 | 1]  that the backward pass needs.
// Backward: the data-gradient chain per row (d_h for the LSTM BPTT kernel, d_pred for the generator) and
// the per-row gradient record  G = [d_o1 | d_oc | d_p1 | d_pc | d_c1 | d_label | d_l1 | d_code].
// ALL 16 parameter gradients are then blocks of ONE plain GEMM  X^T . G  done by the host (cuBLAS):
// the trailing 1 of X yields the bias gradients.  No atomics, fixed reduction order.
#include "sw_common.cuh"

namespace sw {

constexpr int DH_H = 64, DH_M = 32, DH_PMAX = 128;

// parameter pack (floats), torch layouts [out][in]:
//   Wo1[32][64] bo1[32] Wo2[32][32] bo2[32] Wp1[32][P] bp1[32] Wp2[32][32] bp2[32]
//   Wc1[32][64] bc1[32] Wc2[1][32] bc2[1] Wl1[32][64] bl1[32] Wl2[L][32] bl2[L]
struct HeadOffsets {
    int wo1, bo1, wo2, bo2, wp1, bp1, wp2, bp2, wc1, bc1, wc2, bc2, wl1, bl1, wl2, bl2, total;
    __host__ __device__ HeadOffsets(int P, int L) {
        int o = 0;
        wo1 = o; o += 32 * 64; bo1 = o; o += 32; wo2 = o; o += 32 * 32; bo2 = o; o += 32;
        wp1 = o; o += 32 * P;  bp1 = o; o += 32; wp2 = o; o += 32 * 32; bp2 = o; o += 32;
        wc1 = o; o += 32 * 64; bc1 = o; o += 32; wc2 = o; o += 32;      bc2 = o; o += 1;
        wl1 = o; o += 32 * 64; bl1 = o; o += 32; wl2 = o; o += L * 32;  bl2 = o; o += L;
        total = o;
    }
};

// y[j] = b[j] + sum_k W[j][k] x[k]   (W rows in shared memory, broadcast across the warp)
template <int NOUT, int KMAX>
__device__ __forceinline__ void dense(const float* __restrict__ W, const float* __restrict__ b, const float (&x)[KMAX], int k_in,
                                      float (&y)[NOUT]) {
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
        float a = b[j];
        const float* w = W + j * k_in;
        if (KMAX <= 64) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) a = fmaf(w[k], x[k], a);
        } else {
#pragma unroll 8
            for (int k = 0; k < KMAX; ++k)
                if (k < k_in) a = fmaf(w[k], x[k], a);
        }
        y[j] = a;
    }
}

// x_grad[k] += sum_j W[j][k] d[j]
template <int NOUT, int KMAX>
__device__ __forceinline__ void dense_t(const float* __restrict__ W, const float (&d)[NOUT], int k_in, float (&xg)[KMAX]) {
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
        const float* w = W + j * k_in;
        if (KMAX <= 64) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) xg[k] = fmaf(w[k], d[j], xg[k]);
        } else {
#pragma unroll 8
            for (int k = 0; k < KMAX; ++k)
                if (k < k_in) xg[k] = fmaf(w[k], d[j], xg[k]);
        }
    }
}

template <int L>
__global__ void __launch_bounds__(128)
disc_heads_fwd_kernel(const float* __restrict__ pack, const float* __restrict__ h, const float* __restrict__ pred, int P,
                      int n_rows, float* __restrict__ label, float* __restrict__ code, float* __restrict__ xrec, int x_ld) {
    extern __shared__ __align__(16) float sp[];
    const HeadOffsets o(P, L);
    for (int i = threadIdx.x; i < o.total; i += blockDim.x) sp[i] = __ldg(pack + i);
    __syncthreads();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    float* xr = xrec ? xrec + (size_t)row * x_ld : nullptr;
    float both[64];
    {   // observation branch: h -> o1 -> oc
        float x[64], o1[32], oc[32];
#pragma unroll
        for (int k = 0; k < 64; k += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(h + (size_t)row * 64 + k));
            x[k] = t.x; x[k + 1] = t.y; x[k + 2] = t.z; x[k + 3] = t.w;
        }
        dense<32, 64>(sp + o.wo1, sp + o.bo1, x, 64, o1);
#pragma unroll
        for (int j = 0; j < 32; ++j) o1[j] = lrelu02(o1[j]);
        dense<32, 32>(sp + o.wo2, sp + o.bo2, o1, 32, oc);
#pragma unroll
        for (int j = 0; j < 32; ++j) both[j] = oc[j];
        if (xr) {
#pragma unroll
            for (int k = 0; k < 64; ++k) xr[k] = x[k];
#pragma unroll
            for (int j = 0; j < 32; ++j) xr[64 + j] = o1[j];
        }
    }
    {   // prediction branch: pred (P = n_next*4 values) -> p1 -> pc ; P is small and runtime: stream it
        float p1[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) p1[j] = sp[o.bp1 + j];
        for (int k = 0; k < P; ++k) {
            const float v = __ldg(pred + (size_t)row * P + k);
            if (xr) xr[96 + k] = v;
#pragma unroll
            for (int j = 0; j < 32; ++j) p1[j] = fmaf(sp[o.wp1 + j * P + k], v, p1[j]);
        }
        float pc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) p1[j] = lrelu02(p1[j]);
        dense<32, 32>(sp + o.wp2, sp + o.bp2, p1, 32, pc);
#pragma unroll
        for (int j = 0; j < 32; ++j) both[32 + j] = pc[j];
        if (xr)
#pragma unroll
            for (int j = 0; j < 32; ++j) xr[96 + P + j] = p1[j];
    }
    float c1[32], l1[32];
    dense<32, 64>(sp + o.wc1, sp + o.bc1, both, 64, c1);
    dense<32, 64>(sp + o.wl1, sp + o.bl1, both, 64, l1);
    float lab = sp[o.bc2];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        c1[j] = lrelu02(c1[j]);
        l1[j] = lrelu02(l1[j]);
        lab = fmaf(sp[o.wc2 + j], c1[j], lab);
    }
    label[row] = lab;
#pragma unroll
    for (int q = 0; q < L; ++q) {
        float a = sp[o.bl2 + q];
#pragma unroll
        for (int j = 0; j < 32; ++j) a = fmaf(sp[o.wl2 + q * 32 + j], l1[j], a);
        code[(size_t)row * L + q] = a;
    }
    if (xr) {
        float* xb = xr + 128 + P;
#pragma unroll
        for (int k = 0; k < 64; ++k) xb[k] = both[k];
#pragma unroll
        for (int j = 0; j < 32; ++j) { xb[64 + j] = c1[j]; xb[96 + j] = l1[j]; }
        xb[128] = 1.0f;
    }
}

template <int L>
__global__ void __launch_bounds__(128)
disc_heads_bwd_kernel(const float* __restrict__ pack, const float* __restrict__ xrec, int x_ld, int P, int n_rows,
                      const float* __restrict__ d_label, const float* __restrict__ d_code, float* __restrict__ d_h,
                      float* __restrict__ d_pred, float* __restrict__ grec, int g_ld) {
    extern __shared__ __align__(16) float sp[];
    const HeadOffsets o(P, L);
    for (int i = threadIdx.x; i < o.total; i += blockDim.x) sp[i] = __ldg(pack + i);
    __syncthreads();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const float* xr = xrec + (size_t)row * x_ld;
    const float* xb = xr + 128 + P;                 // both[64] | c1[32] | l1[32] | 1
    float* g = grec + (size_t)row * g_ld;           // d_o1 | d_oc | d_p1 | d_pc | d_c1 | d_label | d_l1 | d_code
    const float dl = d_label ? d_label[row] : 0.0f;
    float dboth[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) dboth[k] = 0.0f;
    {
        float dc1[32], dl1[32], dcode[L];
#pragma unroll
        for (int q = 0; q < L; ++q) dcode[q] = d_code ? d_code[(size_t)row * L + q] : 0.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float a = sp[o.wc2 + j] * dl;
            dc1[j] = (xb[64 + j] > 0.0f) ? a : 0.2f * a;
            float b = 0.0f;
#pragma unroll
            for (int q = 0; q < L; ++q) b = fmaf(sp[o.wl2 + q * 32 + j], dcode[q], b);
            dl1[j] = (xb[96 + j] > 0.0f) ? b : 0.2f * b;
            g[128 + j] = dc1[j];
            g[161 + j] = dl1[j];
        }
        g[160] = dl;
#pragma unroll
        for (int q = 0; q < L; ++q) g[193 + q] = dcode[q];
        dense_t<32, 64>(sp + o.wc1, dc1, 64, dboth);
        dense_t<32, 64>(sp + o.wl1, dl1, 64, dboth);
    }
    {   // observation branch
        float doc[32], do1[32], dh[64];
#pragma unroll
        for (int j = 0; j < 32; ++j) { doc[j] = dboth[j]; do1[j] = 0.0f; g[32 + j] = doc[j]; }
        dense_t<32, 32>(sp + o.wo2, doc, 32, do1);
#pragma unroll
        for (int j = 0; j < 32; ++j) { do1[j] = (xr[64 + j] > 0.0f) ? do1[j] : 0.2f * do1[j]; g[j] = do1[j]; }
        if (d_h) {
#pragma unroll
            for (int k = 0; k < 64; ++k) dh[k] = 0.0f;
            dense_t<32, 64>(sp + o.wo1, do1, 64, dh);
#pragma unroll
            for (int k = 0; k < 64; k += 4)
                *reinterpret_cast<float4*>(d_h + (size_t)row * 64 + k) = make_float4(dh[k], dh[k + 1], dh[k + 2], dh[k + 3]);
        }
    }
    {   // prediction branch
        float dpc[32], dp1[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { dpc[j] = dboth[32 + j]; dp1[j] = 0.0f; g[96 + j] = dpc[j]; }
        dense_t<32, 32>(sp + o.wp2, dpc, 32, dp1);
#pragma unroll
        for (int j = 0; j < 32; ++j) { dp1[j] = (xr[96 + P + j] > 0.0f) ? dp1[j] : 0.2f * dp1[j]; g[64 + j] = dp1[j]; }
        if (d_pred)
            for (int k = 0; k < P; ++k) {
                float a = 0.0f;
#pragma unroll
                for (int j = 0; j < 32; ++j) a = fmaf(sp[o.wp1 + j * P + k], dp1[j], a);
                d_pred[(size_t)row * P + k] = a;
            }
    }
}

}  // namespace sw

extern "C" int sw_disc_heads_pack_floats(int pred_dim, int n_latent) { return sw::HeadOffsets(pred_dim, n_latent).total; }

// xrec (optional, for backward): [N][x_ld] with x_ld >= 129 + 128 + P ... see sw_disc_heads_record_dims
extern "C" int sw_disc_heads_record_dims(int pred_dim, int n_latent, int* x_dim, int* g_dim) {
    if (!x_dim || !g_dim) return SW_ERR_ARG;
    *x_dim = 64 + 32 + pred_dim + 32 + 64 + 32 + 32 + 1;
    *g_dim = 32 * 5 + 1 + 32 + n_latent;
    return SW_OK;
}

extern "C" int sw_disc_heads_fwd(const float* pack, const float* h, const float* pred, int pred_dim, int n_latent,
                                 int n_rows, float* label, float* code, float* xrec, void* stream) {
    if (!pack || !h || !pred || !label || !code) return SW_ERR_ARG;
    if (n_rows <= 0 || pred_dim <= 0 || pred_dim > sw::DH_PMAX) return SW_ERR_ARG;
    if (n_latent != 2) return SW_ERR_UNSUPPORTED;             // n_latent_codes = 2 (train.py:65)
    const sw::HeadOffsets o(pred_dim, n_latent);
    const int x_ld = 257 + pred_dim;
    const int block = 128, grid = (n_rows + block - 1) / block;
    const size_t smem = (size_t)o.total * 4;
    SW_SET_MAX_SMEM(sw::disc_heads_fwd_kernel<2>, (int)smem);
    sw::disc_heads_fwd_kernel<2><<<grid, block, smem, (cudaStream_t)stream>>>(pack, h, pred, pred_dim, n_rows, label, code, xrec, x_ld);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}

extern "C" int sw_disc_heads_bwd(const float* pack, const float* xrec, int pred_dim, int n_latent, int n_rows,
                                 const float* d_label, const float* d_code, float* d_h, float* d_pred, float* grec,
                                 void* stream) {
    if (!pack || !xrec || !grec) return SW_ERR_ARG;
    if (n_rows <= 0 || pred_dim <= 0 || pred_dim > sw::DH_PMAX) return SW_ERR_ARG;
    if (n_latent != 2) return SW_ERR_UNSUPPORTED;
    const sw::HeadOffsets o(pred_dim, n_latent);
    const int x_ld = 257 + pred_dim, g_ld = 193 + n_latent;
    const int block = 128, grid = (n_rows + block - 1) / block;
    const size_t smem = (size_t)o.total * 4;
    SW_SET_MAX_SMEM(sw::disc_heads_bwd_kernel<2>, (int)smem);
    sw::disc_heads_bwd_kernel<2><<<grid, block, smem, (cudaStream_t)stream>>>(pack, xrec, x_ld, pred_dim, n_rows, d_label, d_code,
                                                                             d_h, d_pred, grec, g_ld);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
