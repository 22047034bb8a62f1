// Shared device helpers for the socialways_b200 sm_100a kernels.
//
// Tile convention used by every dense contraction on the path (LSTM gates, DecoderFC layers,
// discriminator heads): a CTA owns SW_ROWS = 32 independent rows (agents, or (agent, sample)
// pairs).  Activations live in shared memory "k-major": X[k][row], 32 rows contiguous (128 B, one
// shared-memory wavefront).  Weights live in shared memory k-major as well, W[k][col-group][TNP],
// a thread's TN output columns contiguous (LDS.128 when TN % 4 == 0, else LDS.64).
// A thread accumulates a 4 x TN register tile; KS lanes of the same warp split the k range and
// are summed with a shuffle butterfly.  Lane layout inside a warp:
//     lane = rg + 8 * (ks + KS * cgl),  rg = row group (4 rows), ks = k split, cgl = local col group.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SW_ROWS 32
#define SW_THREADS 256
#define SW_H 64          // LSTM hidden size of the path (train.py:43-45 default, BASELINE configs)
#define SW_G 256         // 4 gates x H
#define SW_Z 32          // noise_len = H / 2 (train.py:81)
#define SW_DEC_IN 160    // H + F + Z (train.py:375)
#define SW_DEC_H2 80

// error codes of the C-ABI (include/socialways_b200.h)
#define SW_OK 0
#define SW_ERR_ARG -1
#define SW_ERR_CUDA -2
#define SW_ERR_UNSUPPORTED -3

#define SW_CUDA_TRY(expr)                                 \
    do {                                                  \
        cudaError_t _e = (expr);                          \
        if (_e != cudaSuccess) { sw_set_last_cuda_error((int)_e); return SW_ERR_CUDA; } \
    } while (0)

void sw_set_last_cuda_error(int e);

namespace sw {

__device__ __forceinline__ float sigmoidf_acc(float x) { return __fdividef(1.0f, 1.0f + expf(-x)); }
// 1 - 2/(1+e^{2x}): absolute error ~1e-7 everywhere (what matters for the state update), saturates cleanly
__device__ __forceinline__ float tanhf_acc(float x) { return 1.0f - __fdividef(2.0f, 1.0f + expf(2.0f * x)); }
__device__ __forceinline__ float lrelu02(float x) { return x > 0.0f ? x : 0.2f * x; }

template <int KS>
struct LaneMap {
    int rg, ks, cg;  // row group 0..7, k split 0..KS-1, global col group
    __device__ __forceinline__ LaneMap() {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        rg = lane & 7;
        ks = (lane >> 3) % KS;
        cg = warp * (4 / KS) + (lane >> 3) / KS;
    }
};

// acc[i][j] += sum over this lane's k of X[k][4*rg + i] * W[k][cg*TN + j]
//   X: shared, row stride SW_ROWS floats.  W: shared or global, row stride ldw floats (even; TN % 4 == 0
//   needs ldw % 4 == 0).  The lane's k run over ks, ks + KS, ... < kcount (kcount % KS == 0): the KS
//   lanes of a (rg, cg) group hit neighbouring weight rows, which with ldw % 32 == 4 are different banks.
template <int TN, int KS>
__device__ __forceinline__ void fma_tile(float (&acc)[4][TN], const float* __restrict__ X,
                                         const float* __restrict__ W, int ldw, int kcount,
                                         const LaneMap<KS>& lm) {
    static_assert(TN % 2 == 0, "tile width must be even");
    const float* xp = X + lm.ks * SW_ROWS + lm.rg * 4;
    const float* wp = W + lm.ks * ldw + lm.cg * TN;
#pragma unroll 4
    for (int k = 0; k < kcount; k += KS) {
        const float4 xv = *reinterpret_cast<const float4*>(xp + k * SW_ROWS);
        float w[TN];
        if (TN % 4 == 0) {
#pragma unroll
            for (int q = 0; q < TN / 4; ++q) {
                const float4 t = *reinterpret_cast<const float4*>(wp + k * ldw + q * 4);
                w[q * 4 + 0] = t.x; w[q * 4 + 1] = t.y; w[q * 4 + 2] = t.z; w[q * 4 + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < TN / 2; ++q) {
                const float2 t = *reinterpret_cast<const float2*>(wp + k * ldw + q * 2);
                w[q * 2 + 0] = t.x; w[q * 2 + 1] = t.y;
            }
        }
        const float x[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(x[i], w[j], acc[i][j]);
    }
}

// sum the KS partial tiles held by the KS lanes that share (rg, cg); every lane ends with the total
template <int TN, int KS>
__device__ __forceinline__ void ksplit_reduce(float (&acc)[4][TN]) {
#pragma unroll
    for (int off = 8; off < 8 * KS; off <<= 1)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], off);
}

// cooperative copy of n floats (n % 4 == 0, both 16 B aligned) global -> shared
__device__ __forceinline__ void copy_f4(float* __restrict__ dst, const float* __restrict__ src, int n) {
    for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4)
        *reinterpret_cast<float4*>(dst + i) = __ldg(reinterpret_cast<const float4*>(src + i));
}

// Transposing load: src row-major [rows][K] gathered through row_of(r) -> dst k-major [K][SW_ROWS].
// `scratch` is a [SW_ROWS][K+1] shared staging area (may alias any buffer free at that moment).
template <typename RowFn>
__device__ __forceinline__ void load_rows_kmajor(float* __restrict__ dst, float* __restrict__ scratch,
                                                 const float* __restrict__ src, int K, int nrows_valid,
                                                 RowFn row_of) {
    const int ldp = K + 1;
    for (int i = threadIdx.x; i < SW_ROWS * K; i += blockDim.x) {
        const int r = i / K, k = i - r * K;
        scratch[r * ldp + k] = (r < nrows_valid) ? __ldg(src + (size_t)row_of(r) * K + k) : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SW_ROWS * K; i += blockDim.x) {
        const int k = i >> 5, r = i & 31;
        dst[k * SW_ROWS + r] = scratch[r * ldp + k];
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// LSTM layer on a 32-row tile.  Packed weights (socialways_b200/packing.py, `pack_lstm`):
//   row 0..3   Wx[k][n']   input projection of the 4-d state (for the encoder: W_ih . W_embed folded)
//   row 4..67  Whh[k][n']
//   row 68     bias[n']    (b_ih + b_hh [+ W_ih . b_embed])
// with gate-interleaved columns n' = 4*unit + gate, gate order i,f,g,o (torch.nn.LSTM).
// X points at a [68][32] k-major operand {x4 ; h}.  Thread (rg, cg) owns units 2cg, 2cg+1 of rows
// 4rg..4rg+3: the cell state c of those 8 (row, unit) pairs stays in registers across steps.
// ------------------------------------------------------------------------------------------
#define SW_LSTM_K 68
#define SW_LSTM_PACK_FLOATS (69 * 256)

struct LstmGates {  // post-activation gates of one (row, unit): what the backward pass needs
    float i, f, g, o;
};

template <bool STASH>
__device__ __forceinline__ void lstm_tile_step(const float* __restrict__ Wl /*smem [69][256]*/,
                                               const float* __restrict__ X /*smem [68][32]*/,
                                               float* __restrict__ Hout /*smem [64][32]*/, float (&c)[4][2],
                                               const LaneMap<1>& lm, float* __restrict__ stash_row0, int stash_ld,
                                               int rows_valid) {
    float acc[4][8];
    const float* b = Wl + 68 * SW_G + lm.cg * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = b[j];
    fma_tile<8, 1>(acc, X, Wl, SW_G, SW_LSTM_K, lm);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        float hv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float gi = sigmoidf_acc(acc[i][u * 4 + 0]);
            const float gf = sigmoidf_acc(acc[i][u * 4 + 1]);
            const float gg = tanhf_acc(acc[i][u * 4 + 2]);
            const float go = sigmoidf_acc(acc[i][u * 4 + 3]);
            c[i][u] = fmaf(gf, c[i][u], gi * gg);
            hv[i] = go * tanhf_acc(c[i][u]);
            if (STASH) {
                const int r = lm.rg * 4 + i;
                if (r < rows_valid) {
                    // stash layout per row: [i f g o c] x 64 units, unit-major: 5 floats per unit
                    float* s = stash_row0 + (size_t)r * stash_ld + (lm.cg * 2 + u) * 5;
                    s[0] = gi; s[1] = gf; s[2] = gg; s[3] = go; s[4] = c[i][u];
                }
            }
        }
        *reinterpret_cast<float4*>(Hout + (lm.cg * 2 + u) * SW_ROWS + lm.rg * 4) = make_float4(hv[0], hv[1], hv[2], hv[3]);
    }
}

}  // namespace sw
