// Device-side latent noise for the K-sample inference path: uniform [0, 1) fp32 from Philox4x32-10.
//
// The reference draws the noise on the HOST (torch.rand on the CPU generator, train.py:584 / :473) and uploads it:
// 128 B per predicted trajectory over PCIe, 94 % of the host->device bytes of an inference step.  A caller that does not
// need the reference's exact noise stream can let the GPU draw it (Generator.predict_k(noise=None, seed=...)): same
// distribution (24-bit uniform grid, like torch.rand for float32), a DIFFERENT stream -- parity tests keep host noise.
//
// Stream definition (so that any implementation can reproduce it): element e of the output belongs to group
// g = first_group + e / 4; the four values of group g are Philox4x32-10(counter = (g_lo, g_hi, offset_lo, offset_hi),
// key = (seed_lo, seed_hi)), each mapped to float by (x >> 8) * 2^-24.  `first_group` lets a rank draw exactly its rows of a
// larger logical tensor (sharded training: same noise whatever the number of GPUs).  tests/test_gpu_noise.py checks the
// stream against a numpy restatement.
#include "sw_common.cuh"

namespace sw {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

__global__ void __launch_bounds__(256)
noise_uniform_kernel(float* __restrict__ out, long long n, unsigned long long seed, unsigned long long offset,
                     unsigned long long first_group) {
    const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
    const long long groups = (n + 3) >> 2;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
        const unsigned long long gg = first_group + (unsigned long long)g;
        const uint4 r = philox4x32_10(make_uint4((unsigned)gg, (unsigned)(gg >> 32), (unsigned)offset, (unsigned)(offset >> 32)), key);
        const float4 v = make_float4((r.x >> 8) * 5.9604644775390625e-8f, (r.y >> 8) * 5.9604644775390625e-8f,
                                     (r.z >> 8) * 5.9604644775390625e-8f, (r.w >> 8) * 5.9604644775390625e-8f);
        if (4 * g + 3 < n) {
            *reinterpret_cast<float4*>(out + 4 * g) = v;
        } else {
            const float t[4] = {v.x, v.y, v.z, v.w};
            for (int q = 0; 4 * g + q < n; ++q) out[4 * g + q] = t[q];
        }
    }
}

}  // namespace sw

extern "C" int sw_noise_uniform(float* out, long long n, unsigned long long seed, unsigned long long offset,
                                unsigned long long first_group, int sm_count, void* stream) {
    if (!out || n <= 0 || sm_count <= 0) return SW_ERR_ARG;
    if (((uintptr_t)out & 15u) != 0) return SW_ERR_ARG;
    long long blocks = ((n + 3) / 4 + 255) / 256;
    if (blocks > (long long)sm_count * 8) blocks = (long long)sm_count * 8;
    sw::noise_uniform_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(out, n, seed, offset, first_group);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
