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
}  // namespace sw

extern "C" int sw_bestofk_metrics(const float* pred, const float* gt, float ss, int n_agents, int n_samples,
                                  int n_next, float* out, void* stream) {
    if (!pred || !gt || !out) return SW_ERR_ARG;
    if (n_agents <= 0 || n_samples <= 0 || n_next <= 0 || !(ss > 0.f)) return SW_ERR_ARG;
    const int block = 128, grid = (n_agents + block - 1) / block;
    sw::bestofk_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(pred, gt, 1.0f / ss, n_agents, n_samples, n_next, out);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
