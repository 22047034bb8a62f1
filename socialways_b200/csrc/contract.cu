// Weight-gradient contractions of the training step, batched: every parameter gradient of train()'s two backward
// passes (reference train.py:495 d_loss.backward(), :538 g_loss.backward(); in the reference these are the
// `grad_weight = grad_output^T . input` halves of autograd's Linear / LSTM nodes) is a sum over batch rows
//        out[k][n] = sum_rows A[row][k] * B[row][n]
// of a forward stash A against a gradient record B that the data-gradient kernels wrote.  ONE launch evaluates a whole
// list of such jobs (all of D's, or all of G's): no library GEMM, no reduction kernels, bias gradients as a K = 1 job
// against an all-ones operand.
//
// Operand layouts: "tile image" [image][k][32 rows] (the shared-memory operand of the FFMA kernels, written with
// coalesced float4 stores; padding rows of the gradient images are zero) or row-major records [row][ld].
// Work item = (job, 64 x 64 output tile, chunk of images); 256 threads, 4 x 4 outputs per thread.  Reduction order is
// FIXED: a chunk sums its images in order into registers, writes its partial tile to the workspace, and the CTA that
// arrives last at the tile's counter adds the partials in chunk order -- the result does not depend on scheduling.
#include "sw_common.cuh"
#include "sw_contract.h"

namespace sw {

constexpr int CT_TILE = 64;        // output tile edge
constexpr int CT_LD = 36;          // padded shared row: 32 rows + 4 (a float4 per lane, lane stride 36 -> conflict-free)
constexpr int CT_MAX_JOBS = SW_CONTRACT_MAX_JOBS;

struct ContractParams {
    sw_contract_job job[CT_MAX_JOBS];
    int first_cta[CT_MAX_JOBS + 1];   // prefix sum of CTAs per job
    int chunks[CT_MAX_JOBS];          // image chunks per output tile
    int ipc[CT_MAX_JOBS];             // images per chunk
    int first_tile[CT_MAX_JOBS];      // index of the job's first output tile (counter / workspace slot base)
    long long ws_off[CT_MAX_JOBS];    // workspace offset (floats) of the job's partial tiles
    int n_jobs;
};

__device__ __forceinline__ int gate_perm(int n) { return (n & 3) * 64 + (n >> 2); }   // n' = 4*unit + gate -> gate*64 + unit

// stage rows [r0, r0 + 64) of one operand image into shared memory s[64][CT_LD] (zero beyond `rows` / `n_rows`)
__device__ __forceinline__ void stage_operand(float* __restrict__ s, const float* __restrict__ base, long long stride, int r0,
                                              int rows /*rows of the job from r0 on*/, int image, int kind, int total_rows) {
    const int tid = threadIdx.x;
    if (kind == SW_CONTRACT_IMAGE) {
        const float* img = base + (size_t)image * stride + (size_t)r0 * 32;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = tid + q * 256, row = i >> 3, piece = i & 7;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < rows) v = __ldg(reinterpret_cast<const float4*>(img + row * 32) + piece);
            *reinterpret_cast<float4*>(s + row * CT_LD + piece * 4) = v;
        }
    } else if (kind == SW_CONTRACT_ROWS) {
        // records [row][stride]: element (k, r) = base[(image*32 + r) * stride + r0 + k]; read k-fastest (coalesced)
        for (int i = tid; i < CT_TILE * 32; i += 256) {
            const int r = i >> 6, k = i & 63;
            const long long grow = (long long)image * 32 + r;
            float v = 0.0f;
            if (k < rows && grow < total_rows) v = __ldg(base + (size_t)grow * stride + r0 + k);
            s[k * CT_LD + r] = v;
        }
    } else {   // SW_CONTRACT_ONES: k = 0 is the all-ones row (bias gradients); rows past the batch are zero
        for (int i = tid; i < CT_TILE * 32; i += 256) {
            const int r = i & 31, k = i >> 5;
            const long long grow = (long long)image * 32 + r;
            s[k * CT_LD + r] = (k == 0 && grow < total_rows) ? 1.0f : 0.0f;
        }
    }
}

__global__ void __launch_bounds__(256, 3)
contract_kernel(const __grid_constant__ ContractParams P, float* __restrict__ ws, unsigned* __restrict__ counters) {
    __shared__ __align__(16) float sa[CT_TILE * CT_LD];
    __shared__ __align__(16) float sb[CT_TILE * CT_LD];
    __shared__ bool last_s;
    int j = 0;
    while (j + 1 < P.n_jobs && (int)blockIdx.x >= P.first_cta[j + 1]) ++j;
    const sw_contract_job& J = P.job[j];
    const int local = blockIdx.x - P.first_cta[j];
    const int tiles_n = (J.N + CT_TILE - 1) / CT_TILE;
    const int tiles_k = (J.K + CT_TILE - 1) / CT_TILE;
    const int C = P.chunks[j];
    const int chunk = local % C, tile = local / C;
    const int kt = tile / tiles_n, nt = tile % tiles_n;
    (void)tiles_k;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int img0 = chunk * P.ipc[j], img1 = min(img0 + P.ipc[j], J.n_images);
    const int krows = min(CT_TILE, J.K - kt * CT_TILE), nrows = min(CT_TILE, J.N - nt * CT_TILE);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = 0.0f;

    for (int img = img0; img < img1; ++img) {
        __syncthreads();
        stage_operand(sa, J.a, J.a_stride, J.a_k0 + kt * CT_TILE, krows, img, J.a_kind, J.n_rows);
        stage_operand(sb, J.b, J.b_stride, J.b_n0 + nt * CT_TILE, nrows, img, J.b_kind, J.n_rows);
        __syncthreads();
#pragma unroll
        for (int r4 = 0; r4 < 8; ++r4) {
            float4 av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(sa + (ty + 16 * i) * CT_LD + r4 * 4);
#pragma unroll
            for (int q = 0; q < 4; ++q) bv[q] = *reinterpret_cast<const float4*>(sb + (tx + 16 * q) * CT_LD + r4 * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc[i][q] = fmaf(av[i].x, bv[q].x, acc[i][q]);
                    acc[i][q] = fmaf(av[i].y, bv[q].y, acc[i][q]);
                    acc[i][q] = fmaf(av[i].z, bv[q].z, acc[i][q]);
                    acc[i][q] = fmaf(av[i].w, bv[q].w, acc[i][q]);
                }
        }
    }

    // element (i, q) of this thread: k = kt*64 + ty + 16 i, n = nt*64 + tx + 16 q
    auto store_out = [&](const float (&v)[4][4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = ty + 16 * i, n = tx + 16 * q;
                if (k < krows && n < nrows) {
                    int gn = nt * CT_TILE + n;
                    if (J.n_perm == SW_CONTRACT_PERM_GATES) gn = gate_perm(gn);
                    float* o = J.out + (size_t)(kt * CT_TILE + k) * J.out_sk + (size_t)gn * J.out_sn;
                    const float val = v[i][q] * J.scale;
                    if (J.out2) J.out2[(size_t)(kt * CT_TILE + k) * J.out_sk + (size_t)gn * J.out_sn] = val;
                    *o = J.accumulate ? *o + val : val;
                }
            }
    };
    if (C == 1) { store_out(acc); return; }

    // partial tile -> workspace [tile][chunk][256 threads][16], then the last arrival reduces in chunk order
    const int tile_id = P.first_tile[j] + tile;
    float* part = ws + P.ws_off[j] + ((size_t)tile * C + chunk) * (CT_TILE * CT_TILE);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(part + (i * 256 + tid) * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __threadfence();
    __syncthreads();
    if (tid == 0) last_s = atomicAdd(counters + tile_id, 1u) == (unsigned)(C - 1);
    __syncthreads();
    if (!last_s) return;
    __threadfence();
    float sum[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) sum[i][q] = 0.0f;
    const float* base = ws + P.ws_off[j] + (size_t)tile * C * (CT_TILE * CT_TILE);
    for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(base + (size_t)c * (CT_TILE * CT_TILE) + (i * 256 + tid) * 4));
            sum[i][0] += v.x; sum[i][1] += v.y; sum[i][2] += v.z; sum[i][3] += v.w;
        }
    }
    store_out(sum);
    if (tid == 0) counters[tile_id] = 0u;     // ready for the next launch (CUDA-graph replay included)
}

struct ContractPlan {
    ContractParams p;
    long long ws_floats;
    int n_tiles, n_ctas;
};

static int make_plan(const sw_contract_job* jobs, int n_jobs, int sm_count, ContractPlan& plan) {
    if (!jobs || n_jobs <= 0 || n_jobs > CT_MAX_JOBS || sm_count <= 0) return SW_ERR_ARG;
    long long image_tiles = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const sw_contract_job& J = jobs[j];
        if (!J.b || !J.out || J.K <= 0 || J.N <= 0 || J.n_images <= 0 || J.n_rows <= 0) return SW_ERR_ARG;
        if (J.a_kind != SW_CONTRACT_ONES && !J.a) return SW_ERR_ARG;
        if (J.a_kind < 0 || J.a_kind > SW_CONTRACT_ONES || J.b_kind < 0 || J.b_kind > SW_CONTRACT_ROWS) return SW_ERR_ARG;
        if (J.n_perm == SW_CONTRACT_PERM_GATES && J.N != 256) return SW_ERR_ARG;
        const int tiles = ((J.K + CT_TILE - 1) / CT_TILE) * ((J.N + CT_TILE - 1) / CT_TILE);
        image_tiles += (long long)tiles * J.n_images;
    }
    // images per chunk: about 3 CTAs per SM in flight, at least 4 images per CTA (amortises the partial-tile round trip)
    long long ipc = (image_tiles + 3LL * sm_count - 1) / (3LL * sm_count);
    if (ipc < 4) ipc = 4;
    plan.p.n_jobs = n_jobs;
    plan.ws_floats = 0;
    int cta = 0, tile0 = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const sw_contract_job& J = jobs[j];
        const int tiles = ((J.K + CT_TILE - 1) / CT_TILE) * ((J.N + CT_TILE - 1) / CT_TILE);
        const int chunks = (int)((J.n_images + ipc - 1) / ipc);
        plan.p.job[j] = J;
        plan.p.first_cta[j] = cta;
        plan.p.chunks[j] = chunks;
        plan.p.ipc[j] = (int)ipc;
        plan.p.first_tile[j] = tile0;
        plan.p.ws_off[j] = plan.ws_floats;
        if (chunks > 1) plan.ws_floats += (long long)tiles * chunks * CT_TILE * CT_TILE;
        cta += tiles * chunks;
        tile0 += tiles;
    }
    plan.p.first_cta[n_jobs] = cta;
    plan.n_tiles = tile0;
    plan.n_ctas = cta;
    return SW_OK;
}

}  // namespace sw

extern "C" int sw_contract_plan(const sw_contract_job* jobs, int n_jobs, int sm_count, long long* workspace_floats,
                                int* n_counters) {
    if (!workspace_floats || !n_counters) return SW_ERR_ARG;
    sw::ContractPlan plan;
    const int rc = sw::make_plan(jobs, n_jobs, sm_count, plan);
    if (rc != SW_OK) return rc;
    *workspace_floats = plan.ws_floats;
    *n_counters = plan.n_tiles;
    return SW_OK;
}

extern "C" int sw_contract(const sw_contract_job* jobs, int n_jobs, float* workspace, long long workspace_floats,
                           unsigned* counters, int n_counters, int sm_count, void* stream) {
    sw::ContractPlan plan;
    const int rc = sw::make_plan(jobs, n_jobs, sm_count, plan);
    if (rc != SW_OK) return rc;
    if (plan.ws_floats > workspace_floats || plan.n_tiles > n_counters) return SW_ERR_ARG;
    if ((plan.ws_floats > 0 && !workspace) || !counters) return SW_ERR_ARG;
    sw::contract_kernel<<<plan.n_ctas, 256, 0, (cudaStream_t)stream>>>(plan.p, workspace, counters);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
