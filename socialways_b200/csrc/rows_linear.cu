// Small per-row linear maps of the training step that sit between the big kernels:
//     out[row][n] = bias[n] + add1[row][n] + add2[row][n] + sum_k X[row][k] * W[k][n]        (K, n <= 80)
// Uses: (u | beta) = h . M + m0, the per-agent operand of the pooling kernel (AttentionPooling.W and
// EmbedSocialFeatures.fc.4 folded, reference train.py:158,167-169,185 -- packing.pool_agent_matrix), and its adjoint
// dL/dh = dh_decode + dh_pool + d(u | beta) . M^T inside g_loss.backward() (train.py:538).
// One CTA = 32 rows (128 from 8192 rows up: 4 rows per thread); X tile and W in shared memory; thread = (row(s), column group of
// 8-strided columns).
#include "sw_common.cuh"

namespace sw {

constexpr int RL_MAX = 80;

// RPT rows per thread (rows r, r + 32, ...: the same shared-memory bank pattern as one row): every weight read from shared
// memory feeds RPT FMAs.  The sum over k of each output runs in the same order for every RPT: results are bit-identical.
template <int RPT>
__global__ void __launch_bounds__(SW_THREADS)
rows_linear_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W /*[K][n_out]*/, const float* __restrict__ bias,
                   const float* __restrict__ add1, const float* __restrict__ add2, float* __restrict__ out, int ldo,
                   int n_rows, int K, int n_out) {
    extern __shared__ __align__(16) float sm[];
    constexpr int ROWS = SW_ROWS * RPT;
    float* sw_ = sm;                       // [K][n_out]
    float* sx = sm + K * n_out;            // [ROWS][K + 1]
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * ROWS;
    for (int i = tid; i < K * n_out; i += SW_THREADS) sw_[i] = __ldg(W + i);
    for (int i = tid; i < ROWS * K; i += SW_THREADS) {
        const int r = i / K, k = i - r * K;
        sx[r * (K + 1) + k] = (row0 + r < n_rows) ? __ldg(X + (size_t)(row0 + r) * ldx + k) : 0.0f;
    }
    __syncthreads();
    const int r = tid >> 3, cg = tid & 7;
    if (row0 + r >= n_rows) return;
    float acc[RPT][RL_MAX / 8];
#pragma unroll
    for (int j = 0; j < RPT; ++j)
#pragma unroll
        for (int q = 0; q < RL_MAX / 8; ++q) acc[j][q] = 0.0f;
    const float* xr = sx + r * (K + 1);
    for (int k = 0; k < K; ++k) {
        float x[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) x[j] = xr[j * SW_ROWS * (K + 1) + k];
        const float* w = sw_ + k * n_out + cg;
#pragma unroll
        for (int q = 0; q < RL_MAX / 8; ++q)
            if (cg + 8 * q < n_out) {
                const float wv = w[8 * q];
#pragma unroll
                for (int j = 0; j < RPT; ++j) acc[j][q] = fmaf(x[j], wv, acc[j][q]);
            }
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int row = row0 + r + j * SW_ROWS;
        if (row >= n_rows) break;
#pragma unroll
        for (int q = 0; q < RL_MAX / 8; ++q) {
            const int n = cg + 8 * q;
            if (n < n_out) {
                float v = acc[j][q];
                if (bias) v += __ldg(bias + n);
                if (add1) v += __ldg(add1 + (size_t)row * ldo + n);
                if (add2) v += __ldg(add2 + (size_t)row * ldo + n);
                out[(size_t)row * ldo + n] = v;
            }
        }
    }
}

}  // namespace sw

extern "C" int sw_rows_linear(const float* x, int ldx, const float* w, const float* bias, const float* add1,
                              const float* add2, float* out, int ldo, int n_rows, int k_in, int n_out, void* stream) {
    if (!x || !w || !out) return SW_ERR_ARG;
    if (n_rows <= 0 || k_in <= 0 || n_out <= 0 || ldx < k_in || ldo < n_out) return SW_ERR_ARG;
    if (k_in > sw::RL_MAX || n_out > sw::RL_MAX) return SW_ERR_UNSUPPORTED;
    if (n_rows >= 8192) {       // large batches (inference): 4 rows per thread, 128 rows per CTA
        constexpr int RPT = 4;
        const size_t smem = (size_t)(k_in * n_out + SW_ROWS * RPT * (k_in + 1)) * 4;
        SW_SET_MAX_SMEM(sw::rows_linear_kernel<RPT>, (int)smem);
        const int grid = (n_rows + SW_ROWS * RPT - 1) / (SW_ROWS * RPT);
        sw::rows_linear_kernel<RPT><<<grid, SW_THREADS, smem, (cudaStream_t)stream>>>(x, ldx, w, bias, add1, add2, out, ldo, n_rows,
                                                                                      k_in, n_out);
    } else {
        const size_t smem = (size_t)(k_in * n_out + SW_ROWS * (k_in + 1)) * 4;
        SW_SET_MAX_SMEM(sw::rows_linear_kernel<1>, (int)smem);
        const int grid = (n_rows + SW_ROWS - 1) / SW_ROWS;
        sw::rows_linear_kernel<1><<<grid, SW_THREADS, smem, (cudaStream_t)stream>>>(x, ldx, w, bias, add1, add2, out, ldo, n_rows,
                                                                                    k_in, n_out);
    }
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
