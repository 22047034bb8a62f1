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

// Raise a kernel's dynamic shared-memory limit, once per call site, device and size (a per-device function
// attribute, not a stream operation; doing it only when the requested size grows keeps steady-state launches free of
// driver calls, which also makes them capturable into CUDA graphs).
#define SW_MAX_DEVICES 64
#define SW_SET_MAX_SMEM(kernel, bytes)                                                                       \
    do {                                                                                                     \
        static int _sw_smem_set[SW_MAX_DEVICES];                                                             \
        int _sw_dev = 0;                                                                                     \
        SW_CUDA_TRY(cudaGetDevice(&_sw_dev));                                                                \
        const int _sw_want = (int)(bytes);                                                                   \
        if (_sw_dev < 0 || _sw_dev >= SW_MAX_DEVICES || _sw_want > _sw_smem_set[_sw_dev]) {                  \
            SW_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, _sw_want)); \
            if (_sw_dev >= 0 && _sw_dev < SW_MAX_DEVICES) _sw_smem_set[_sw_dev] = _sw_want;                  \
        }                                                                                                    \
    } while (0)

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

// Backward-pass stash of one LSTM tile step ("tile k-major", the shared-memory image):
//   gates [5][64][32] floats: post-activation i, f, g, o and the new cell state c, per unit, 32 rows contiguous.
#define SW_GATE_STASH_FLOATS (5 * SW_H * SW_ROWS)

template <bool STASH>
__device__ __forceinline__ void lstm_tile_step(const float* __restrict__ Wl /*smem [69][256]*/,
                                               const float* __restrict__ X /*smem [68][32]*/,
                                               float* __restrict__ Hout /*smem [64][32]*/, float (&c)[4][2],
                                               const LaneMap<1>& lm, float* __restrict__ stash /*global, tile image*/) {
    float acc[4][8];
    const float* b = Wl + 68 * SW_G + lm.cg * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = b[j];
    fma_tile<8, 1>(acc, X, Wl, SW_G, SW_LSTM_K, lm);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        float gi[4], gf[4], gg[4], go[4], hv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            gi[i] = sigmoidf_acc(acc[i][u * 4 + 0]);
            gf[i] = sigmoidf_acc(acc[i][u * 4 + 1]);
            gg[i] = tanhf_acc(acc[i][u * 4 + 2]);
            go[i] = sigmoidf_acc(acc[i][u * 4 + 3]);
            c[i][u] = fmaf(gf[i], c[i][u], gi[i] * gg[i]);
            hv[i] = go[i] * tanhf_acc(c[i][u]);
        }
        const int off = (lm.cg * 2 + u) * SW_ROWS + lm.rg * 4;
        *reinterpret_cast<float4*>(Hout + off) = make_float4(hv[0], hv[1], hv[2], hv[3]);
        if (STASH) {
            float* s = stash + off;
            *reinterpret_cast<float4*>(s + 0 * SW_H * SW_ROWS) = make_float4(gi[0], gi[1], gi[2], gi[3]);
            *reinterpret_cast<float4*>(s + 1 * SW_H * SW_ROWS) = make_float4(gf[0], gf[1], gf[2], gf[3]);
            *reinterpret_cast<float4*>(s + 2 * SW_H * SW_ROWS) = make_float4(gg[0], gg[1], gg[2], gg[3]);
            *reinterpret_cast<float4*>(s + 3 * SW_H * SW_ROWS) = make_float4(go[0], go[1], go[2], go[3]);
            *reinterpret_cast<float4*>(s + 4 * SW_H * SW_ROWS) = make_float4(c[0][u], c[1][u], c[2][u], c[3][u]);
        }
    }
}

// Reverse of lstm_tile_step for thread (rg, cg): from dL/dh_t (shared DH[64][32]) and the running dL/dc_t
// (registers) to the pre-activation gate gradients DG[n'][32] (shared, gate-interleaved n' = 4*unit+gate,
// the layout of the pack columns) and dL/dc_{t-1} (registers).  `st` = this step's gate stash image,
// `c_prev` = previous step's cell image (nullptr: zero initial state).  Rows >= rows_valid produce zeros.
__device__ __forceinline__ void lstm_tile_bwd_gates(const float* __restrict__ st, const float* __restrict__ c_prev,
                                                    const float* __restrict__ DH, float* __restrict__ DG,
                                                    float (&dc)[4][2], const LaneMap<1>& lm, int rows_valid) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int unit = lm.cg * 2 + u;
        const int off = unit * SW_ROWS + lm.rg * 4;
        const float4 vi = __ldg(reinterpret_cast<const float4*>(st + 0 * SW_H * SW_ROWS + off));
        const float4 vf = __ldg(reinterpret_cast<const float4*>(st + 1 * SW_H * SW_ROWS + off));
        const float4 vg = __ldg(reinterpret_cast<const float4*>(st + 2 * SW_H * SW_ROWS + off));
        const float4 vo = __ldg(reinterpret_cast<const float4*>(st + 3 * SW_H * SW_ROWS + off));
        const float4 vc = __ldg(reinterpret_cast<const float4*>(st + 4 * SW_H * SW_ROWS + off));
        float4 vp = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c_prev) vp = *reinterpret_cast<const float4*>(c_prev + off);   // global stash or shared image
        const float4 vdh = *reinterpret_cast<const float4*>(DH + off);
        const float gi[4] = {vi.x, vi.y, vi.z, vi.w}, gf[4] = {vf.x, vf.y, vf.z, vf.w}, gg[4] = {vg.x, vg.y, vg.z, vg.w},
                    go[4] = {vo.x, vo.y, vo.z, vo.w}, cc[4] = {vc.x, vc.y, vc.z, vc.w}, cp[4] = {vp.x, vp.y, vp.z, vp.w},
                    dh[4] = {vdh.x, vdh.y, vdh.z, vdh.w};
        float dai[4], daf[4], dag[4], dao[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool valid = (lm.rg * 4 + i) < rows_valid;
            const float tc = tanhf_acc(cc[i]);
            const float d_o = dh[i] * tc;
            const float dct = fmaf(dh[i] * go[i], 1.0f - tc * tc, dc[i][u]);
            dai[i] = valid ? dct * gg[i] * gi[i] * (1.0f - gi[i]) : 0.0f;
            daf[i] = valid ? dct * cp[i] * gf[i] * (1.0f - gf[i]) : 0.0f;
            dag[i] = valid ? dct * gi[i] * (1.0f - gg[i] * gg[i]) : 0.0f;
            dao[i] = valid ? d_o * go[i] * (1.0f - go[i]) : 0.0f;
            dc[i][u] = valid ? dct * gf[i] : 0.0f;
        }
        float* g = DG + (unit * 4) * SW_ROWS + lm.rg * 4;
        *reinterpret_cast<float4*>(g + 0 * SW_ROWS) = make_float4(dai[0], dai[1], dai[2], dai[3]);
        *reinterpret_cast<float4*>(g + 1 * SW_ROWS) = make_float4(daf[0], daf[1], daf[2], daf[3]);
        *reinterpret_cast<float4*>(g + 2 * SW_ROWS) = make_float4(dag[0], dag[1], dag[2], dag[3]);
        *reinterpret_cast<float4*>(g + 3 * SW_ROWS) = make_float4(dao[0], dao[1], dao[2], dao[3]);
    }
}

// cooperative copy of a tile image (n floats, n % 4 == 0) shared <-> global
__device__ __forceinline__ void store_image(float* __restrict__ gdst, const float* __restrict__ ssrc, int n) {
    for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4)
        *reinterpret_cast<float4*>(gdst + i) = *reinterpret_cast<const float4*>(ssrc + i);
}

}  // namespace sw
