// Weight-gradient contractions of the training step, batched: every parameter gradient of train()'s two backward
// passes (reference train.py:495 d_loss.backward(), :538 g_loss.backward(); in the reference these are the
// `grad_weight = grad_output^T . input` / grad_bias halves of autograd's Linear / LSTM nodes) is a sum over batch rows
//        G[k][n] = sum_rows A[row][k] * B[row][n]
// of a forward stash A against a gradient record B that the data-gradient kernels wrote.  ONE launch evaluates a whole
// list of such jobs (all of D's, or all of G's): no library GEMM, no reduction kernels; bias gradients ride along as an
// all-ones row appended to A; a job scatters its result into up to three parameter tensors (sw_contract.h).
//
// Two kernels behind the same job list:
//   sw_contract_tc  tcgen05.mma kind::tf32 on fp32 operands SPLIT into two tf32 terms (x = hi + lo, hi = x with the low 13
//                   mantissa bits cleared, lo = rna_tf32(x - hi)): A.B ~= Ahi.Bhi + Ahi.Blo + Alo.Bhi, fp32 accumulation in
//                   TMEM over all images of a chunk -- fp32 exponent range (gradient records sit at 1e-3 .. 1e-9, far below
//                   fp16's range, which rules out the fp16 split of the inference kernels), ~2^-21 per product.  One work
//                   item = (job, 128 x 128 slab of G, chunk of images); the batch rows are the MMA's K dimension (32 per
//                   image = 4 MMAs of K = 8 per product).  Operands are staged fp32 -> (hi, lo) by all 256 threads into the
//                   canonical K-major no-swizzle layout [K/4][rows][4] with a PADDED chunk stride (rows*16 + 16 bytes: the 8
//                   chunk-consecutive lanes of a quarter-warp hit 8 different bank groups).  The loads of image i + 1 are
//                   issued into registers right after the MMAs of image i, so they fly under those MMAs and under the wait
//                   for them; 66 KB of shared memory and 128 TMEM columns per CTA -> 2 CTAs per SM interleave.  The kernel is
//                   bound by the operand read (each image is read once per slab), not by the tensor pipe.
//   sw_contract     fp32 FFMA register tiles (64 x 64 of G per CTA, 4 x 4 per thread), the round-2 first version; kept as
//                   the arithmetic reference of the tensor-core kernel (tests) and for A/B timing.
// Reduction order is FIXED in both: a chunk sums its images in order, writes its partial slab to the workspace, and the
// CTA that arrives last at the slab's counter adds the partials in chunk order -- results do not depend on scheduling.
#include "sw_common.cuh"
#include "sw_contract.h"
#include "sw_umma.cuh"

namespace sw {

constexpr int CT_MAX_JOBS = SW_CONTRACT_MAX_JOBS;

struct ContractParams {
    sw_contract_job job[CT_MAX_JOBS];
    int first_cta[CT_MAX_JOBS + 1];   // prefix sum of CTAs per job
    int chunks[CT_MAX_JOBS];          // image chunks per output tile
    int ipc[CT_MAX_JOBS];             // images per chunk
    int first_tile[CT_MAX_JOBS];      // index of the job's first output tile (counter slot base)
    long long ws_off[CT_MAX_JOBS];    // workspace offset (floats) of the job's partial tiles
    int n_jobs;
};

__device__ __forceinline__ int gate_perm(int n) { return (n & 3) * 64 + (n >> 2); }   // n' = 4*unit + gate -> gate*64 + unit

// G[k][n] -> the segment that owns row k
__device__ __forceinline__ void scatter(const sw_contract_job& J, int k, int n, float val) {
#pragma unroll
    for (int s = 0; s < SW_CONTRACT_MAX_SEGS; ++s) {
        if (s >= J.n_segs) break;
        const sw_contract_seg& S = J.seg[s];
        const int kk = k - S.k_begin;
        if (kk >= 0 && kk < S.k_count) {
            const int gn = J.n_perm == SW_CONTRACT_PERM_GATES ? gate_perm(n) : n;
            const size_t idx = (size_t)kk * S.out_sk + (size_t)gn * S.out_sn;
            S.out[idx] = val;
            if (S.out2) S.out2[idx] = val;
        }
    }
}

// =====================================================================================================================
// FFMA kernel: 64 x 64 tiles of G
// =====================================================================================================================
constexpr int CT_TILE = 64;
constexpr int CT_LD = 36;          // padded shared row: 32 rows + 4 (a float4 per lane, lane stride 36 -> conflict-free)

// stage rows [r0, r0 + 64) of one operand image into shared memory s[64][CT_LD].  `rows` = operand rows of the job
// (rows >= `rows`: the ones row at index `rows` when `ones`, zeros beyond).
__device__ __forceinline__ void stage_operand(float* __restrict__ s, const float* __restrict__ base, long long stride, int k0,
                                              int r0, int rows, bool ones, int image, int kind, int total_rows) {
    const int tid = threadIdx.x;
    if (kind == SW_CONTRACT_IMAGE) {
        const float* img = base + (size_t)image * stride + (size_t)(k0 + r0) * 32;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = tid + q * 256, row = i >> 3, piece = i & 7;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + row < rows) v = __ldg(reinterpret_cast<const float4*>(img + row * 32) + piece);
            else if (ones && r0 + row == rows) {
                const long long g0 = (long long)image * 32 + piece * 4;
                v = make_float4(g0 < total_rows ? 1.f : 0.f, g0 + 1 < total_rows ? 1.f : 0.f, g0 + 2 < total_rows ? 1.f : 0.f,
                                g0 + 3 < total_rows ? 1.f : 0.f);
            }
            *reinterpret_cast<float4*>(s + row * CT_LD + piece * 4) = v;
        }
    } else {
        // records [row][stride]: element (k, r) = base[(image*32 + r) * stride + k0 + r0 + k]; read k-fastest (coalesced)
        for (int i = tid; i < CT_TILE * 32; i += 256) {
            const int r = i >> 6, k = i & 63;
            const long long grow = (long long)image * 32 + r;
            float v = 0.0f;
            if (grow < total_rows) {
                if (r0 + k < rows) v = __ldg(base + (size_t)grow * stride + k0 + r0 + k);
                else if (ones && r0 + k == rows) v = 1.0f;
            }
            s[k * CT_LD + r] = v;
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
    const int rows_total = J.K + (J.ones_row ? 1 : 0);
    const int tiles_n = (J.N + CT_TILE - 1) / CT_TILE;
    const int C = P.chunks[j];
    const int chunk = local % C, tile = local / C;
    const int kt = tile / tiles_n, nt = tile % tiles_n;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int img0 = chunk * P.ipc[j], img1 = min(img0 + P.ipc[j], J.n_images);
    const int krows = min(CT_TILE, rows_total - kt * CT_TILE), nrows = min(CT_TILE, J.N - nt * CT_TILE);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = 0.0f;

    for (int img = img0; img < img1; ++img) {
        __syncthreads();
        stage_operand(sa, J.a, J.a_stride, J.a_k0, kt * CT_TILE, J.K, J.ones_row != 0, img, J.a_kind, J.n_rows);
        stage_operand(sb, J.b, J.b_stride, J.b_n0, nt * CT_TILE, J.N, false, img, J.b_kind, J.n_rows);
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
                if (k < krows && n < nrows) scatter(J, kt * CT_TILE + k, nt * CT_TILE + n, v[i][q]);
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

// =====================================================================================================================
// tcgen05 kernel: 128 x 128 slabs of G, tf32 split operands
// =====================================================================================================================
constexpr int TC_M = 128, TC_N = 128, TC_THREADS = 256;
constexpr int TC_CHUNK = 128 * 4 + 4;           // floats per K-chunk (4 r's x 128 rows + 16 B pad)
constexpr uint32_t FMT_TF32 = 2;

struct TcSmem {
    float a_hi[8 * TC_CHUNK], a_lo[8 * TC_CHUNK];
    float b_hi[8 * TC_CHUNK], b_lo[8 * TC_CHUNK];
    unsigned long long bar;
    uint32_t tmem_base;
    int last;
};

// issued by ONE elected lane (inside `if (elect_one())`, sw_umma.cuh)
__device__ __forceinline__ void umma1_issue_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

// x -> (hi, lo): hi = x with the 13 low mantissa bits cleared (exactly a tf32 number), lo = (x - hi) rounded to tf32
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    const float d = x - hi;
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(d));
    lo = __uint_as_float(r);
}

// One operand slab of one image = up to 128 rows x 32 batch rows = 1024 float4 items (row, chunk c: r = 4c .. 4c+3), four per
// thread.  fetch_operand pulls them into REGISTERS (so that the loads of image i + 1 are in flight while the MMAs of image i
// run and while the CTA waits for them); put_operand splits them into (hi, lo) and writes the canonical K-major layout
// [8 chunks][TC_CHUNK].  Row index == rows -> the all-ones row when `ones`, larger -> zeros.
struct OperandRegs { float4 v[4]; };

__device__ __forceinline__ void item_of(int i, int count, int kind, int& row, int& c) {
    if (kind == SW_CONTRACT_IMAGE) { row = i >> 3; c = i & 7; }      // chunk fastest: a warp reads whole 128-byte rows
    else { c = i / count; row = i - c * count; }                      // row (= record column) fastest: coalesced along k
}

__device__ __forceinline__ void fetch_operand(OperandRegs& R, const float* __restrict__ base, long long stride, int k0, int row0,
                                              int count, int rows, bool ones, int image, int kind, int total_rows) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = threadIdx.x + q * TC_THREADS;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < count * 8) {
            int row, c;
            item_of(i, count, kind, row, c);
            const long long g0 = (long long)image * 32 + c * 4;
            if (row0 + row < rows) {
                if (kind == SW_CONTRACT_IMAGE) {
                    v = __ldg(reinterpret_cast<const float4*>(base + (size_t)image * stride + (size_t)(k0 + row0 + row) * 32) + c);
                } else {
                    const float* p = base + (size_t)g0 * stride + k0 + row0 + row;
                    if (g0 < total_rows) v.x = __ldg(p);
                    if (g0 + 1 < total_rows) v.y = __ldg(p + stride);
                    if (g0 + 2 < total_rows) v.z = __ldg(p + 2 * stride);
                    if (g0 + 3 < total_rows) v.w = __ldg(p + 3 * stride);
                }
            } else if (ones && row0 + row == rows) {
                v = make_float4(g0 < total_rows ? 1.f : 0.f, g0 + 1 < total_rows ? 1.f : 0.f, g0 + 2 < total_rows ? 1.f : 0.f,
                                g0 + 3 < total_rows ? 1.f : 0.f);
            }
        }
        R.v[q] = v;
    }
}

__device__ __forceinline__ void put_operand(const OperandRegs& R, float* __restrict__ hi, float* __restrict__ lo, int count, int kind) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = threadIdx.x + q * TC_THREADS;
        if (i < count * 8) {
            int row, c;
            item_of(i, count, kind, row, c);
            float4 h, l;
            split_tf32(R.v[q].x, h.x, l.x); split_tf32(R.v[q].y, h.y, l.y);
            split_tf32(R.v[q].z, h.z, l.z); split_tf32(R.v[q].w, h.w, l.w);
            *reinterpret_cast<float4*>(hi + c * TC_CHUNK + row * 4) = h;
            *reinterpret_cast<float4*>(lo + c * TC_CHUNK + row * 4) = l;
        }
    }
}

__global__ void __launch_bounds__(TC_THREADS, 2)
contract_tc_kernel(const __grid_constant__ ContractParams P, float* __restrict__ ws, unsigned* __restrict__ counters) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TcSmem& s = *reinterpret_cast<TcSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int j = 0;
    while (j + 1 < P.n_jobs && (int)blockIdx.x >= P.first_cta[j + 1]) ++j;
    const sw_contract_job& J = P.job[j];
    const int local = blockIdx.x - P.first_cta[j];
    const int rows_total = J.K + (J.ones_row ? 1 : 0);
    const int tiles_n = (J.N + TC_N - 1) / TC_N;
    const int C = P.chunks[j];
    const int chunk = local % C, tile = local / C;
    const int mt = tile / tiles_n, nt = tile % tiles_n;
    const int img0 = chunk * P.ipc[j], img1 = min(img0 + P.ipc[j], J.n_images);
    const int mrows = min(TC_M, rows_total - mt * TC_M);
    const int ncols = min(TC_N, J.N - nt * TC_N);
    const int n_pad = (ncols + 15) & ~15;               // UMMA N: multiple of 16 for M = 128
    const bool a_ones = J.ones_row != 0;

    OperandRegs ra, rb;                                  // the first image's loads fly under the TMEM / barrier set-up
    fetch_operand(ra, J.a, J.a_stride, J.a_k0, mt * TC_M, mrows, J.K, a_ones, img0, J.a_kind, J.n_rows);
    fetch_operand(rb, J.b, J.b_stride, J.b_n0, nt * TC_N, ncols, J.N, false, img0, J.b_kind, J.n_rows);

    if (warp == 0) {
        ptx::tcgen05_alloc(ptx::cta_group_1, &s.tmem_base, 128u);
        ptx::tcgen05_relinquish_alloc_permit(ptx::cta_group_1);
    }
    if (tid == 0) {
        ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.bar), 1);
        ptx::fence_mbarrier_init(ptx::sem_release, ptx::scope_cluster);
    }
    // rows of the B operand between ncols and n_pad are never written by the staging: clear them once
    for (int i = tid; i < (n_pad - ncols) * 8; i += TC_THREADS) {
        const int c = i & 7, row = ncols + (i >> 3);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(s.b_hi + c * TC_CHUNK + row * 4) = z;
        *reinterpret_cast<float4*>(s.b_lo + c * TC_CHUNK + row * 4) = z;
    }
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    ptx::tcgen05_fence_after_thread_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, s.tmem_base, 0);
    const uint32_t idesc = umma_idesc(n_pad, FMT_TF32);
    uint32_t phase = 0u;

    for (int img = img0; img < img1; ++img) {
        if (img > img0) { mbar_wait(&s.bar, phase); phase ^= 1u; }    // the MMAs of the previous image have read the operands
        put_operand(ra, s.a_hi, s.a_lo, mrows, J.a_kind);
        put_operand(rb, s.b_hi, s.b_lo, ncols, J.b_kind);
        ptx::fence_proxy_async(ptx::space_shared);
        ptx::tcgen05_fence_before_thread_sync();
        __syncthreads();
        if (warp == 0) {
            ptx::tcgen05_fence_after_thread_sync();
            // canonical K-major, no swizzle: core matrix = 8 rows x 16 B; SBO (8-row group stride) = 128 B,
            // LBO (K-chunk stride) = the padded chunk size; one tf32 MMA (K = 8) spans two chunks
            const uint64_t ah = umma_desc_uniform(s.a_hi, TC_CHUNK * 4, 128), al = umma_desc_uniform(s.a_lo, TC_CHUNK * 4, 128);
            const uint64_t bh = umma_desc_uniform(s.b_hi, TC_CHUNK * 4, 128), bl = umma_desc_uniform(s.b_lo, TC_CHUNK * 4, 128);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t off = (uint64_t)(ks * 2 * TC_CHUNK * 4 / 16);
                    umma1_issue_ss_tf32(tmem, ah + off, bh + off, idesc, img > img0 || ks > 0);
                    umma1_issue_ss_tf32(tmem, ah + off, bl + off, idesc, true);
                    umma1_issue_ss_tf32(tmem, al + off, bh + off, idesc, true);
                }
                umma1_commit(&s.bar);
            }
            __syncwarp();
        }
        if (img + 1 < img1) {                                         // next image's loads: in flight during the MMAs + the wait
            fetch_operand(ra, J.a, J.a_stride, J.a_k0, mt * TC_M, mrows, J.K, a_ones, img + 1, J.a_kind, J.n_rows);
            fetch_operand(rb, J.b, J.b_stride, J.b_n0, nt * TC_N, ncols, J.N, false, img + 1, J.b_kind, J.n_rows);
        }
    }
    mbar_wait(&s.bar, phase);
    ptx::tcgen05_fence_after_thread_sync();

    // ---- epilogue: thread = (G row m = TMEM lane, column half of the slab) ----
    const int m = (warp & 3) * 32 + lane, half = warp >> 2;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int tile_id = P.first_tile[j] + tile;
    float* part = ws + P.ws_off[j] + ((size_t)tile * C + chunk) * (size_t)(TC_M * TC_N);
    for (int c0 = half * 64; c0 < min(n_pad, half * 64 + 64); c0 += 8) {
        uint32_t v[8];
        tmem_ld<8>(taddr + c0, v);
        ptx::tcgen05_wait_ld();
        if (C == 1) {
            if (m < mrows)
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (c0 + q < ncols) scatter(J, mt * TC_M + m, nt * TC_N + c0 + q, __uint_as_float(v[q]));
        } else {
            float* dst = part + (size_t)m * TC_N + c0;
            *reinterpret_cast<uint4*>(dst) = make_uint4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<uint4*>(dst + 4) = make_uint4(v[4], v[5], v[6], v[7]);
        }
    }
    ptx::tcgen05_fence_before_thread_sync();
    if (C > 1) __threadfence();
    __syncthreads();
    if (warp == 0) ptx::tcgen05_dealloc(ptx::cta_group_1, tmem, 128u);
    if (C == 1) return;
    if (tid == 0) s.last = atomicAdd(counters + tile_id, 1u) == (unsigned)(C - 1);
    __syncthreads();
    if (!s.last) return;
    __threadfence();
    // last arrival: add the partial slabs in chunk order (fixed summation order), four columns per thread and step
    const float* base = ws + P.ws_off[j] + (size_t)tile * C * (size_t)(TC_M * TC_N);
    const int quads = n_pad >> 2;
    for (int e = tid; e < mrows * quads; e += TC_THREADS) {
        const int mm = e / quads, n4 = (e - mm * quads) * 4;
        const float* src = base + (size_t)mm * TC_N + n4;
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        int c = 0;
        for (; c + 4 <= C; c += 4) {                                  // four independent loads in flight
            const float4 v0 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(c + 0) * (TC_M * TC_N)));
            const float4 v1 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(c + 1) * (TC_M * TC_N)));
            const float4 v2 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(c + 2) * (TC_M * TC_N)));
            const float4 v3 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(c + 3) * (TC_M * TC_N)));
            sum.x = ((sum.x + v0.x) + v1.x) + v2.x + v3.x; sum.y = ((sum.y + v0.y) + v1.y) + v2.y + v3.y;
            sum.z = ((sum.z + v0.z) + v1.z) + v2.z + v3.z; sum.w = ((sum.w + v0.w) + v1.w) + v2.w + v3.w;
        }
        for (; c < C; ++c) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (size_t)c * (TC_M * TC_N)));
            sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        }
        const float t[4] = {sum.x, sum.y, sum.z, sum.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (n4 + q < ncols) scatter(J, mt * TC_M + mm, nt * TC_N + n4 + q, t[q]);
    }
    if (tid == 0) counters[tile_id] = 0u;
}

// =====================================================================================================================
struct ContractPlan {
    ContractParams p;
    long long ws_floats;
    int n_tiles, n_ctas;
};

static int make_plan(const sw_contract_job* jobs, int n_jobs, int sm_count, bool tc, ContractPlan& plan) {
    if (!jobs || n_jobs <= 0 || n_jobs > CT_MAX_JOBS || sm_count <= 0) return SW_ERR_ARG;
    auto tiles_of = [&](const sw_contract_job& J) {
        const int rows = J.K + (J.ones_row ? 1 : 0);
        return tc ? ((rows + TC_M - 1) / TC_M) * ((J.N + TC_N - 1) / TC_N)
                  : ((rows + CT_TILE - 1) / CT_TILE) * ((J.N + CT_TILE - 1) / CT_TILE);
    };
    auto tile_floats = [&](const sw_contract_job&) { return tc ? (long long)TC_M * TC_N : (long long)CT_TILE * CT_TILE; };
    long long image_tiles = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const sw_contract_job& J = jobs[j];
        if (!J.a || !J.b || J.K <= 0 || J.N <= 0 || J.N > SW_CONTRACT_MAX_N || J.n_images <= 0 || J.n_rows <= 0) return SW_ERR_ARG;
        if (J.a_kind < 0 || J.a_kind > SW_CONTRACT_ROWS || J.b_kind < 0 || J.b_kind > SW_CONTRACT_ROWS) return SW_ERR_ARG;
        if (J.n_perm == SW_CONTRACT_PERM_GATES && J.N != 256) return SW_ERR_ARG;
        if (J.n_segs <= 0 || J.n_segs > SW_CONTRACT_MAX_SEGS) return SW_ERR_ARG;
        for (int s = 0; s < J.n_segs; ++s)
            if (!J.seg[s].out || J.seg[s].k_begin < 0 || J.seg[s].k_count <= 0) return SW_ERR_ARG;
        image_tiles += (long long)tiles_of(J) * J.n_images;
    }
    // images per chunk: a few CTAs per SM in flight, at least 8 images per CTA (amortises the set-up and the partial-slab
    // round trip; small problems stay unsplit and store directly)
    const long long target = (tc ? 4LL : 3LL) * sm_count;
    long long ipc = (image_tiles + target - 1) / target;
    if (ipc < 8) ipc = 8;
    plan.p.n_jobs = n_jobs;
    plan.ws_floats = 0;
    int cta = 0, tile0 = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const sw_contract_job& J = jobs[j];
        const int tiles = tiles_of(J);
        const int chunks = (int)((J.n_images + ipc - 1) / ipc);
        plan.p.job[j] = J;
        plan.p.first_cta[j] = cta;
        plan.p.chunks[j] = chunks;
        plan.p.ipc[j] = (int)ipc;
        plan.p.first_tile[j] = tile0;
        plan.p.ws_off[j] = plan.ws_floats;
        if (chunks > 1) plan.ws_floats += (long long)tiles * chunks * tile_floats(J);
        cta += tiles * chunks;
        tile0 += tiles;
    }
    plan.p.first_cta[n_jobs] = cta;
    plan.n_tiles = tile0;
    plan.n_ctas = cta;
    return SW_OK;
}

}  // namespace sw

extern "C" int sw_contract_plan(const sw_contract_job* jobs, int n_jobs, int sm_count, int tensor_cores,
                                long long* workspace_floats, int* n_counters) {
    if (!workspace_floats || !n_counters) return SW_ERR_ARG;
    sw::ContractPlan plan;
    const int rc = sw::make_plan(jobs, n_jobs, sm_count, tensor_cores != 0, plan);
    if (rc != SW_OK) return rc;
    *workspace_floats = plan.ws_floats;
    *n_counters = plan.n_tiles;
    return SW_OK;
}

static int contract_launch(const sw_contract_job* jobs, int n_jobs, float* workspace, long long workspace_floats,
                           unsigned* counters, int n_counters, int sm_count, bool tc, void* stream) {
    sw::ContractPlan plan;
    const int rc = sw::make_plan(jobs, n_jobs, sm_count, tc, plan);
    if (rc != SW_OK) return rc;
    if (plan.ws_floats > workspace_floats || plan.n_tiles > n_counters) return SW_ERR_ARG;
    if ((plan.ws_floats > 0 && !workspace) || !counters) return SW_ERR_ARG;
    if (tc) {
        const int smem = (int)sizeof(sw::TcSmem);
        SW_SET_MAX_SMEM(sw::contract_tc_kernel, smem);
        sw::contract_tc_kernel<<<plan.n_ctas, sw::TC_THREADS, smem, (cudaStream_t)stream>>>(plan.p, workspace, counters);
    } else {
        sw::contract_kernel<<<plan.n_ctas, 256, 0, (cudaStream_t)stream>>>(plan.p, workspace, counters);
    }
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}

extern "C" int sw_contract(const sw_contract_job* jobs, int n_jobs, float* workspace, long long workspace_floats,
                           unsigned* counters, int n_counters, int sm_count, void* stream) {
    return contract_launch(jobs, n_jobs, workspace, workspace_floats, counters, n_counters, sm_count, false, stream);
}

extern "C" int sw_contract_tc(const sw_contract_job* jobs, int n_jobs, float* workspace, long long workspace_floats,
                              unsigned* counters, int n_counters, int sm_count, void* stream) {
    return contract_launch(jobs, n_jobs, workspace, workspace_floats, counters, n_counters, sm_count, true, stream);
}
