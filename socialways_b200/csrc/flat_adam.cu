// Adam over ONE flat fp32 parameter buffer, optionally fused with the gradient all-reduce of the sharded training step
// (SURVEY.md §8e, §8f-3).  Replaces, per optimiser step of train() (reference train.py:495-496, :538-539 with the
// optimisers of :381,:385): [distributed only: flat-buffer NCCL all-reduce] + torch.optim.Adam's ~12 foreach kernels.
//
// Update rule = torch.optim.Adam (betas, eps, no weight decay, no amsgrad), in its operation order:
//     m <- m + (g - m)(1 - b1) ;  v <- v b2 + ((1 - b2) g) g ;  p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The step count t lives in device memory and is advanced by the kernel, so a captured CUDA graph replays correctly.
//
// sw_allreduce_adam: ONE kernel per optimiser step and rank.  Every rank's gradient buffer lives in symmetric (peer-
// mapped) memory; the kernel (1) publishes "my gradients are written" to every peer's flag array through NVLink,
// (2) waits for all peers, (3) reads ALL ranks' gradients with peer loads in rank order 0..W-1 -- the same order on
// every rank, so the replicated parameters stay bit-identical without any broadcast -- and applies Adam, (4) publishes
// "I have finished reading" and waits for the peers' same message before returning, so nobody overwrites a gradient
// buffer that is still being read.  The payload is 27 939 (D) / 86 122 (G) floats: latency-bound, so a handful of CTAs
// with every peer load of a round in flight; plain kernels, so the whole sharded iteration can sit in a CUDA graph (NCCL
// capture hung in this stack).  Waits are bounded in TIME (%globaltimer; sw_set_peer_wait_timeout_ms, default 60 s): a peer
// that never shows up makes the kernel give up WITHOUT touching the parameters and raise the status word seq[2] -- the
// context stays alive (no trap) and the host reads the word when it next synchronises (FlatAdam.check_status()).
#include "sw_common.cuh"

namespace sw {

// hyper-parameters arrive as the python doubles torch sees; the float casts below are the ones its kernels make
struct AdamHyper {
    double lr, b1, b2;
    float b2f, omb1, omb2, eps;      // (float)b2, (float)(1 - b1), (float)(1 - b2), (float)eps
};

static AdamHyper make_hyper(double lr, double b1, double b2, double eps) {
    return AdamHyper{lr, b1, b2, (float)b2, (float)(1.0 - b1), (float)(1.0 - b2), (float)eps};
}

__device__ __forceinline__ void adam_update(float g, float& p, float& m, float& v, const AdamHyper h, float step_size,
                                            float inv_bc2_sqrt) {
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), h.omb1));
    v = __fadd_rn(__fmul_rn(v, h.b2f), __fmul_rn(__fmul_rn(h.omb2, g), g));
    const float denom = __fadd_rn(__fmul_rn(__fsqrt_rn(v), inv_bc2_sqrt), h.eps);
    p = __fadd_rn(p, __fmul_rn(-step_size, __fdiv_rn(m, denom)));
}

// bias corrections of step t (t >= 1), evaluated in double like the python scalars of torch's default (non-capturable) path
__device__ __forceinline__ void adam_scalars(float t, const AdamHyper h, float& step_size, float& inv_bc2_sqrt) {
    const double bc1 = 1.0 - pow(h.b1, (double)t), bc2 = 1.0 - pow(h.b2, (double)t);
    step_size = (float)(h.lr / bc1);
    inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
}

// step[0] = step count t (float, as torch keeps it), step[1] = finished-CTA counter (uint32 bits, zero between launches).
// Every CTA reads t when it starts; the CTA that finishes LAST advances it -- one launch per optimiser step, and no CTA
// can observe the advanced value.
__global__ void adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                 float* __restrict__ step, int n, AdamHyper h) {
    __shared__ float t_s;
    if (threadIdx.x == 0) t_s = *reinterpret_cast<volatile float*>(step) + 1.0f;
    __syncthreads();
    const float t = t_s;
    float step_size, inv_bc2_sqrt;
    adam_scalars(t, h, step_size, inv_bc2_sqrt);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float pi = p[i], mi = m[i], vi = v[i];
        adam_update(g[i], pi, mi, vi, h, step_size, inv_bc2_sqrt);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* counter = reinterpret_cast<unsigned*>(step + 1);
        __threadfence();
        if (atomicAdd(counter, 1u) == gridDim.x - 1) { *counter = 0u; *step = t; }
    }
}

__device__ __forceinline__ void st_release_sys(unsigned* addr, unsigned val) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(addr), "r"(val) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* addr) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// wait until *flag has reached `seq` (wrap-safe); false after timeout_ns
__device__ __forceinline__ bool wait_flag(const unsigned* flag, unsigned seq, unsigned long long timeout_ns) {
    const unsigned long long t0 = global_ns();
    for (unsigned spins = 0; (int)(ld_acquire_sys(flag) - seq) < 0; ++spins)
        if ((spins & 1023u) == 1023u && global_ns() - t0 > timeout_ns) return false;
    return true;
}
__device__ __forceinline__ float4 ld_volatile_f32x4(const float* addr) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
    return v;
}

// peer_bufs[r] = base of rank r's symmetric buffer: [n_pad floats of gradient | flags: ready[W] | done[W]] (uint32).
// p, m, v are n_pad floats long as well (zero padding), so every access is a float4.
// Grid = a few CTAs (latency-bound: what matters is the number of peer loads in flight).  No CTA waits for another CTA of
// its own grid: CTA 0 publishes "ready", every CTA polls the LOCAL flag array, and the CTA that finishes last (device-scope
// counter) publishes "done" and waits for the peers' "done".
constexpr int AR_THREADS = 512;
constexpr int AR_MAX_WORLD = 16;
__global__ void __launch_bounds__(AR_THREADS, 1)
allreduce_adam_kernel(const unsigned long long* __restrict__ peer_bufs, int rank, int world, int n_pad,
                      float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, float* __restrict__ step,
                      unsigned* __restrict__ seq_ptr /* [0] sequence number, [1] finished-CTA counter, [2] status */,
                      AdamHyper h, unsigned long long timeout_ns) {
    __shared__ unsigned seq_s;
    __shared__ float t_s;
    __shared__ bool last_s;
    __shared__ int timed_out_s;
    __shared__ unsigned long long bufs[AR_MAX_WORLD];
    const int tid = threadIdx.x;
    if (tid == 0) { seq_s = *seq_ptr + 1u; t_s = *step + 1.0f; timed_out_s = 0; }
    if (tid < world) bufs[tid] = peer_bufs[tid];
    __syncthreads();
    const unsigned seq = seq_s;
    unsigned* my_flags = reinterpret_cast<unsigned*>(reinterpret_cast<float*>(bufs[rank]) + n_pad);
    // (1) my gradients (written by earlier kernels of this stream) are complete -> tell every peer; (2) wait for all peers
    if (tid < world) {
        if (blockIdx.x == 0) {
            unsigned* peer_flags = reinterpret_cast<unsigned*>(reinterpret_cast<float*>(bufs[tid]) + n_pad);
            __threadfence_system();
            st_release_sys(peer_flags + rank, seq);
        }
        if (!wait_flag(my_flags + tid, seq, timeout_ns)) atomicExch(&timed_out_s, 1);
    }
    __syncthreads();
    if (timed_out_s) {            // a peer never published its gradients: leave the parameters alone, tell the host
        if (tid == 0) atomicExch(seq_ptr + 2, 1u);
        return;
    }
    // (3) sum in rank order + Adam, one float4 per thread and round, all peers' loads of a round in flight together
    float step_size, inv_bc2_sqrt;
    adam_scalars(t_s, h, step_size, inv_bc2_sqrt);
    for (int i = (blockIdx.x * AR_THREADS + tid) * 4; i < n_pad; i += gridDim.x * AR_THREADS * 4) {
        float4 gr[AR_MAX_WORLD];
#pragma unroll
        for (int r = 0; r < AR_MAX_WORLD; ++r)
            if (r < world) gr[r] = ld_volatile_f32x4(reinterpret_cast<const float*>(bufs[r]) + i);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < AR_MAX_WORLD; ++r)
            if (r < world) {
                g.x = __fadd_rn(g.x, gr[r].x); g.y = __fadd_rn(g.y, gr[r].y);
                g.z = __fadd_rn(g.z, gr[r].z); g.w = __fadd_rn(g.w, gr[r].w);
            }
        float4 pi = *reinterpret_cast<float4*>(p + i), mi = *reinterpret_cast<float4*>(m + i), vi = *reinterpret_cast<float4*>(v + i);
        adam_update(g.x, pi.x, mi.x, vi.x, h, step_size, inv_bc2_sqrt);
        adam_update(g.y, pi.y, mi.y, vi.y, h, step_size, inv_bc2_sqrt);
        adam_update(g.z, pi.z, mi.z, vi.z, h, step_size, inv_bc2_sqrt);
        adam_update(g.w, pi.w, mi.w, vi.w, h, step_size, inv_bc2_sqrt);
        *reinterpret_cast<float4*>(p + i) = pi;
        *reinterpret_cast<float4*>(m + i) = mi;
        *reinterpret_cast<float4*>(v + i) = vi;
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        last_s = atomicAdd(seq_ptr + 1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last_s) return;
    // (4) every CTA of this rank has finished reading: tell every peer, wait until every peer has finished reading mine
    if (tid < world) {
        unsigned* peer_flags = reinterpret_cast<unsigned*>(reinterpret_cast<float*>(bufs[tid]) + n_pad);
        st_release_sys(peer_flags + world + rank, seq);
        if (!wait_flag(my_flags + world + tid, seq, timeout_ns)) atomicExch(seq_ptr + 2, 2u);
    }
    __syncthreads();
    if (tid == 0) { seq_ptr[1] = 0u; *seq_ptr = seq; *step = t_s; }
}

}  // namespace sw

static bool adam_args_ok(const void* p, const void* m, const void* v, const void* step, int n, double lr, double b1, double b2, double eps) {
    return p && m && v && step && n > 0 && lr >= 0. && b1 >= 0. && b1 < 1. && b2 >= 0. && b2 < 1. && eps >= 0.;
}

extern "C" int sw_adam_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* step, int n, double lr,
                            double beta1, double beta2, double eps, int sm_count, void* stream) {
    if (!adam_args_ok(params, exp_avg, exp_avg_sq, step, n, lr, beta1, beta2, eps) || !grads || sm_count <= 0) return SW_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const sw::AdamHyper h = sw::make_hyper(lr, beta1, beta2, eps);
    const int block = 256;
    int grid = (n + block - 1) / block;
    if (grid > sm_count * 4) grid = sm_count * 4;
    sw::adam_flat_kernel<<<grid, block, 0, st>>>(params, grads, exp_avg, exp_avg_sq, step, n, h);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}

static unsigned long long g_peer_wait_timeout_ns = 60ull * 1000000000ull;
extern "C" int sw_set_peer_wait_timeout_ms(int ms) {
    if (ms <= 0) return SW_ERR_ARG;
    g_peer_wait_timeout_ns = (unsigned long long)ms * 1000000ull;
    return SW_OK;
}

extern "C" int sw_allreduce_adam(const void* peer_bufs_dev, int rank, int world, int n, int n_pad, float* params, float* exp_avg,
                                 float* exp_avg_sq, float* step, unsigned* seq, double lr, double beta1, double beta2, double eps,
                                 void* stream) {
    if (!adam_args_ok(params, exp_avg, exp_avg_sq, step, n, lr, beta1, beta2, eps) || !peer_bufs_dev || !seq) return SW_ERR_ARG;
    if (world < 1 || world > sw::AR_MAX_WORLD || rank < 0 || rank >= world || n_pad < n || (n_pad & 31)) return SW_ERR_ARG;
    const sw::AdamHyper h = sw::make_hyper(lr, beta1, beta2, eps);
    int grid = (n_pad / 4 + sw::AR_THREADS - 1) / sw::AR_THREADS;
    if (grid > 32) grid = 32;
    sw::allreduce_adam_kernel<<<grid, sw::AR_THREADS, 0, (cudaStream_t)stream>>>(
        (const unsigned long long*)peer_bufs_dev, rank, world, n_pad, params, exp_avg, exp_avg_sq, step, seq, h,
        g_peer_wait_timeout_ns);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
