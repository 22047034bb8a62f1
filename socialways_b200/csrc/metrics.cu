// Best-of-K error metrics of test() (reference train.py:587, :602-607), one thread per agent:
//   err[k][t] = || (pred_hat[k,n,t,:2] - gt[n,t]) / ss ||_2
//   out[n] = ( mean_k mean_t err,  mean_k err[T-1],  min_k mean_t err,  min_k err[T-1] )
// The reference sums these four per scene and divides by the number of test agents (:611-614); the
// host does that sum over out[.] (socialways_b200/api.py).
#include "sw_common.cuh"

namespace sw {
__global__ void bestofk_kernel(const float* __restrict__ pred /*[K][N][T][4]*/, const float* __restrict__ gt /*[N][T][2]*/,
                               float inv_ss, int n_agents, int n_samples, int n_next, float* __restrict__ out /*[N][4]*/) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_agents) return;
    float ade_sum = 0.f, fde_sum = 0.f, ade_min = 3.0e38f, fde_min = 3.0e38f;
    for (int k = 0; k < n_samples; ++k) {
        const float4* p = reinterpret_cast<const float4*>(pred) + ((size_t)k * n_agents + n) * n_next;
        const float2* g = reinterpret_cast<const float2*>(gt) + (size_t)n * n_next;
        float s = 0.f, e = 0.f;
        for (int t = 0; t < n_next; ++t) {
            const float4 a = __ldg(p + t);
            const float2 b = __ldg(g + t);
            const float dx = (a.x - b.x) * inv_ss, dy = (a.y - b.y) * inv_ss;
            e = sqrtf(dx * dx + dy * dy);
            s += e;
        }
        s /= (float)n_next;
        ade_sum += s; fde_sum += e;
        ade_min = fminf(ade_min, s); fde_min = fminf(fde_min, e);
    }
    *reinterpret_cast<float4*>(out + (size_t)n * 4) =
        make_float4(ade_sum / (float)n_samples, fde_sum / (float)n_samples, ade_min, fde_min);
}

// The same sums in the same order for even horizons up to 16 steps with 32-byte aligned rows: the thread reads its rows in
// 32-byte pieces (LDG.256 -- each lane touches a different cache line, so the number of requests is what the pass costs) and
// keeps the ground truth of its agent in registers instead of re-reading it for every sample.
__device__ __forceinline__ void ldg256f(const float* p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__global__ void bestofk_wide_kernel(const float* __restrict__ pred /*[K][N][T][4]*/, const float* __restrict__ gt /*[N][T][2]*/,
                                    float inv_ss, int n_agents, int n_samples, int n_next, float* __restrict__ out /*[N][4]*/) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_agents) return;
    float g[32];                                               // (x, y) of up to 16 steps
    const float* gr = gt + (size_t)n * n_next * 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (j * 4 < n_next) ldg256f(gr + j * 8, v);            // 4 steps per piece (n_next is a multiple of 4 here)
#pragma unroll
        for (int i = 0; i < 8; ++i) g[j * 8 + i] = v[i];
    }
    float ade_sum = 0.f, fde_sum = 0.f, ade_min = 3.0e38f, fde_min = 3.0e38f;
    for (int k = 0; k < n_samples; ++k) {
        const float* p = pred + ((size_t)k * n_agents + n) * n_next * 4;
        float s = 0.f, e = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {                          // 2 steps per piece
            if (2 * j < n_next) {
                float a[8];
                ldg256f(p + j * 8, a);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float dx = (a[4 * h] - g[(2 * j + h) * 2]) * inv_ss, dy = (a[4 * h + 1] - g[(2 * j + h) * 2 + 1]) * inv_ss;
                    e = sqrtf(dx * dx + dy * dy);
                    s += e;
                }
            }
        }
        s /= (float)n_next;
        ade_sum += s; fde_sum += e;
        ade_min = fminf(ade_min, s); fde_min = fminf(fde_min, e);
    }
    *reinterpret_cast<float4*>(out + (size_t)n * 4) =
        make_float4(ade_sum / (float)n_samples, fde_sum / (float)n_samples, ade_min, fde_min);
}
}  // namespace sw

extern "C" int sw_bestofk_metrics(const float* pred, const float* gt, float ss, int n_agents, int n_samples,
                                  int n_next, float* out, void* stream) {
    if (!pred || !gt || !out) return SW_ERR_ARG;
    if (n_agents <= 0 || n_samples <= 0 || n_next <= 0 || !(ss > 0.f)) return SW_ERR_ARG;
    const int block = 128, grid = (n_agents + block - 1) / block;
    if (n_next % 4 == 0 && n_next <= 16 && (((uintptr_t)pred | (uintptr_t)gt) & 31u) == 0)
        sw::bestofk_wide_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(pred, gt, 1.0f / ss, n_agents, n_samples, n_next, out);
    else
        sw::bestofk_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(pred, gt, 1.0f / ss, n_agents, n_samples, n_next, out);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
