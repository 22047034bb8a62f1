// Per-iteration scalars of train() in one launch: the running ADE / FDE terms (reference train.py:546-551) and the six
// nn.MSELoss values of the iteration (train.py:484-493, 514-521) from the per-tile partial sums sw_disc_step wrote.
//   stats[0] = sum_rows sum_t ||p_hat - p|| / ss / n_next     (the reference's `e`)
//   stats[1] = sum_rows ||p_hat_T - p_T|| / ss
//   stats[2..7] = d_loss, d_loss_fake, d_loss_real, d_loss_info, g_loss_fooling, g_loss_info
// Fixed summation order (thread-strided rows, shared-memory tree, per-CTA partials added in CTA order by the CTA that
// finishes last), no host synchronisation: the host reads `stats` whenever it wants (once per epoch in train_native()).
#include "sw_common.cuh"

namespace sw {

__global__ void __launch_bounds__(256)
train_stats_kernel(const float* __restrict__ pred_hat /*[N][T][4]*/, const float* __restrict__ pred /*[N][T][2]*/, int n_rows, int T,
                   float ss, const float* __restrict__ d_parts, int d_tiles, const float* __restrict__ g_parts, int g_tiles,
                   float inv_n, float info_w, float* __restrict__ partial /*[grid][2]*/, unsigned* __restrict__ counter,
                   float* __restrict__ stats /*[8]*/) {
    __shared__ float s_a[256], s_f[256];
    __shared__ bool last_s;
    const int tid = threadIdx.x;
    float ade = 0.0f, fde = 0.0f;
    for (int r = blockIdx.x * 256 + tid; r < n_rows; r += gridDim.x * 256) {
        const float4* ph = reinterpret_cast<const float4*>(pred_hat) + (size_t)r * T;
        const float2* pg = reinterpret_cast<const float2*>(pred) + (size_t)r * T;
        float e = 0.0f;
        for (int t = 0; t < T; ++t) {
            const float4 a = __ldg(ph + t);
            const float2 b = __ldg(pg + t);
            const float dx = (a.x - b.x) / ss, dy = (a.y - b.y) / ss;
            e = sqrtf(dx * dx + dy * dy);
            ade += e;
        }
        fde += e;
    }
    s_a[tid] = ade; s_f[tid] = fde;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (tid < off) { s_a[tid] += s_a[tid + off]; s_f[tid] += s_f[tid + off]; }
        __syncthreads();
    }
    if (tid == 0) {
        partial[blockIdx.x * 2] = s_a[0];
        partial[blockIdx.x * 2 + 1] = s_f[0];
        __threadfence();
        last_s = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last_s) return;
    __threadfence();
    // warp 0: ADE/FDE partials; warps 1, 2: loss partials of the D and the G pass
    const int warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        float a = 0.0f, f = 0.0f;
        for (int i = lane; i < (int)gridDim.x; i += 32) { a += __ldcg(partial + i * 2); f += __ldcg(partial + i * 2 + 1); }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, off); f += __shfl_xor_sync(0xffffffffu, f, off); }
        if (lane == 0) { stats[0] = a / (float)T; stats[1] = f; *counter = 0u; }
    } else if (warp == 1 || warp == 2) {
        const float* parts = warp == 1 ? d_parts : g_parts;
        const int tiles = warp == 1 ? d_tiles : g_tiles;
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
        if (parts)
            for (int i = lane; i < tiles; i += 32) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(parts) + i);
                s0 += v.x; s1 += v.y; s2 += v.z;
            }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, off);
            s1 += __shfl_xor_sync(0xffffffffu, s1, off);
            s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        }
        if (lane == 0) {
            if (warp == 1) {
                const float fake = s0 * inv_n, real = s1 * inv_n, info = 0.5f * s2 * inv_n;
                stats[2] = fake + real + info_w * info; stats[3] = fake; stats[4] = real; stats[5] = info;
            } else {
                stats[6] = s0 * inv_n; stats[7] = 0.5f * s2 * inv_n;
            }
        }
    }
}

}  // namespace sw

extern "C" int sw_train_stats(const float* pred_hat, const float* pred, int n_rows, int n_next, float ss, const float* d_parts,
                              int d_tiles, const float* g_parts, int g_tiles, float inv_n, float info_w, float* partial,
                              unsigned* counter, float* stats, int sm_count, void* stream) {
    if (!pred_hat || !pred || !partial || !counter || !stats) return SW_ERR_ARG;
    if (n_rows <= 0 || n_next <= 0 || sm_count <= 0 || !(ss > 0.0f)) return SW_ERR_ARG;
    int grid = (n_rows + 255) / 256;
    if (grid > sm_count) grid = sm_count;
    sw::train_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred_hat, pred, n_rows, n_next, ss, d_parts, d_tiles,
                                                                   g_parts, g_tiles, inv_n, info_w, partial, counter, stats);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
