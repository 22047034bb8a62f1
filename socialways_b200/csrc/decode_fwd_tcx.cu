// fp32-faithful tensor-core decode kernel: tcgen05.mma on FP16 hi/lo SPLIT operands.
//
// Same computation as decode_fwd.cu (the loop of predict(), reference train.py:418-430, for every
// (sample, agent) row), but every dense layer runs on the 5th-gen tensor cores with each fp32 operand x
// represented as two fp16 numbers  x = hi + lo,  hi = fp16(x), lo = fp16(x - hi)  (22 significant bits; the
// operand ranges of this network -- |x| < ~1e2 -- sit well inside fp16), and each product evaluated as
//     A.B  ~=  A_hi.B_hi + A_hi.B_lo + A_lo.B_hi        (3 MMAs, fp32 accumulation in TMEM)
// which keeps the result within ~1e-6 of the fp32 FFMA kernel: the 1e-4 ADE/FDE parity bar holds.
//
// One CTA = one 128-row tile at a time (UMMA M = 128, cta_group::1), 512 threads, thread t owns row
// 32*(warp%4)+lane (its TMEM lane) and column quarter warp/4.  Operand placement:
//   weights  W1h, W2, W34, Whh as hi|lo fp16 in shared memory (163 KB), canonical K-major no-swizzle layout
//   h        hi|lo fp16 in shared memory (32 KB), rewritten by the LSTM epilogue every step
//   a1, a2   hi|lo fp16 written IN PLACE over their own fp32 accumulators in TMEM (tcgen05.st): the 16 accumulator
//            columns of a K block become 8 hi + 8 lo columns, written by the very thread that read them (no
//            barrier in between), and are consumed as the TMEM A operand of the next layer -- they never touch
//            shared memory.  (Issuing L1 / L2 as two N groups to overlap epilogue and MMA was measured SLOWER: every
//            tcgen05.mma carries a fixed cost of the order of 60 clk, so fewer, wider MMAs win.)  The folded 80 -> 2
//            output layer runs as fp32 FMAs inside the layer-2 epilogue (no MMA round trip for 160 MACs per row).
//   c1       the step-invariant part of layer 1, [S ; z] . W1[S,z rows] (hoisted), computed
//            once per tile by MMAs whose weights are streamed through a 20 KB staging buffer, and
//            kept in 160 TMEM columns for the 12 steps
//   (p, v)   fed back to the LSTM as ONE extra K block of the gate MMA: A row = [x_hi(4) | x_lo(4) | x_hi(4) | 1 | 1 | 0 0],
//            B row n' = [Wx_hi | Wx_hi | Wx_lo | b_hi | b_lo | 0 0], i.e. the same three-product split plus the bias, so the
//            gate accumulator IS the pre-activation (the gate epilogue was issue-bound: 37 % of its instructions were
//            this 4-term projection + bias on CUDA cores).  The gate rows of Whh / Wx / b are pre-scaled on the host by
//            -log2(e) (i, f, o) and -2 log2(e) (g), so the accumulator is directly the ex2 argument of the logistic forms.
// TMEM columns: [0,160) c1 | [160,320) L1 acc -> a1 hi|lo ; later gates half 1 | [320,480) L2 acc (hi | lo weight halves) ;
//               later gates half 0.  MMAs execute in issue order, so the gates MMAs are queued right behind the
//               last reader of the region they overwrite and run under the epilogues.
#include <cuda.h>
#include <cuda_fp16.h>

#include "sw_common.cuh"
#include "sw_umma.cuh"

namespace sw {

constexpr int X_ROWS = 128;
constexpr int X_THREADS = 512;
// fp16 weight section (elements), every matrix canonical [K/8][N][8], hi block then lo block
constexpr int XW_W1H_HI = 0, XW_W1H_LO = 10240, XW_W2_CAT = 20480 /* [20][160 = hi|lo][8] */, XW_WHH_HI = 46080,
              XW_WHH_LO = 62464, XW_WXK = 78848 /* x-feedback K block [2][256][8] */, XW_TOTAL = 82944;
// hoist weights in global memory: 3 chunks of K = 32 rows of W1[S,z]: [chunk][hi|lo][4][160][8]
constexpr int XW_SZ_CHUNK = 2 * 4 * 160 * 8;      // 10240 halves = 20480 B
// fp32 section: b1[160] | b2[80] | b34[2] | pad | W34[80][2]
constexpr int XF_B1 = 0, XF_B2 = 160, XF_B34 = 240, XF_W34 = 256, XF_TOTAL = 256 + 160;
constexpr uint32_t XC_C1 = 0, XC_R1 = 160, XC_RG = 320;
constexpr uint32_t FMT_F16 = 0;

struct TcxSmem {
    __half w[XW_TOTAL];                    // 162 816 B
    __half h[2][8 * X_ROWS * 8];           // h hi | lo : [8 chunks][128][8]  (32 768 B)
    __half stage[XW_SZ_CHUNK];             // hoist weight chunk (20 480 B)
    __half xk[2 * X_ROWS * 8];             // x-feedback A operand, one K block: [2 chunks][128][8]  (4 096 B)
    float f32[XF_TOTAL];
    float vpart[8 * X_ROWS];                // partial velocities [quarter][component][row]
    unsigned long long bar[3];
    unsigned long long bar_w;               // TMA: weights (once per CTA)
    unsigned long long bar_z;               // TMA: the tile's noise block (prefetched one tile ahead)
    uint32_t tmem_base;
};

// x -> (hi, lo) fp16 pair for two values at once: returns packed hi2, lo2
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h2 = __floats2half2_rn(a, b);
    const float2 back = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(a - back.x, b - back.y);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}

// three-pass product with both operands in shared memory (called by the whole issuing warp)
// (all MMA forms below are the single-thread ones of sw_umma.cuh: called by ONE elected lane of warp 0, back to back)
template <int B_ROWS, int N, int KB>
__device__ __forceinline__ void mma3_ss(uint32_t d, const __half* a_hi, const __half* a_lo, const __half* b_hi,
                                        const __half* b_lo, bool accumulate_first = false) {
    umma1_ss<B_ROWS, N, KB>(d, a_hi, b_hi, FMT_F16, accumulate_first);
    umma1_ss<B_ROWS, N, KB>(d, a_hi, b_lo, FMT_F16, true);
    umma1_ss<B_ROWS, N, KB>(d, a_lo, b_hi, FMT_F16, true);
}

// three-pass product with the A operand (hi / lo column blocks) in TMEM
template <int B_ROWS, int N, int KB, int A_STRIDE = 8>
__device__ __forceinline__ void mma3_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, const __half* b_hi, const __half* b_lo,
                                        bool accumulate_first) {
    umma1_ts<B_ROWS, N, KB, A_STRIDE>(d, a_hi, b_hi, FMT_F16, accumulate_first);
    umma1_ts<B_ROWS, N, KB, A_STRIDE>(d, a_hi, b_lo, FMT_F16, true);
    umma1_ts<B_ROWS, N, KB, A_STRIDE>(d, a_lo, b_hi, FMT_F16, true);
}

// Epilogue of layer 1, NKB K-blocks (16 output features each) of this thread: y = lrelu(acc + c1) (c1 carries the bias),
// then the hi|lo fp16 pieces of each block are written IN PLACE over the 16 accumulator columns that produced them
// (hi -> columns +0..7, lo -> +8..15), i.e. only over columns this thread itself has just read: no barrier between
// the read and the write, and the next layer addresses K block kb at column 16*kb (hi) / 16*kb + 8 (lo).
template <int NKB>
__device__ __forceinline__ void hidden_epilogue(uint32_t t_acc, uint32_t t_c1) {
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb) {
        uint32_t acc[16], c1v[16], pc[16];
        tmem_ld<16>(t_acc + kb * 16, acc);
        tmem_ld<16>(t_c1 + kb * 16, c1v);
        ptx::tcgen05_wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float y0 = __uint_as_float(acc[2 * e]) + __uint_as_float(c1v[2 * e]);
            const float y1 = __uint_as_float(acc[2 * e + 1]) + __uint_as_float(c1v[2 * e + 1]);
            split2(lrelu02(y0), lrelu02(y1), pc[e], pc[8 + e]);
        }
        tmem_st<16>(t_acc + kb * 16, pc);
    }
    ptx::tcgen05_wait_st();
}

// c1 += b1 for this thread's NKB K-blocks, once per tile (instead of one bias add per value per step)
template <int NKB>
__device__ __forceinline__ void fold_bias_into_c1(uint32_t t_c1, const float* __restrict__ bias) {
    uint32_t v[NKB * 16];
    tmem_ld<NKB * 16>(t_c1, v);
    ptx::tcgen05_wait_ld();
#pragma unroll
    for (int j = 0; j < NKB * 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + bias[j]);
    tmem_st<NKB * 16>(t_c1, v);
    ptx::tcgen05_wait_st();
}

// Epilogue of layer 2 fused with the folded 80 -> 2 output layer: y = lrelu(acc + b2) stays in registers and is
// contracted with W34 in fp32 FMAs (2 per column); the per-quarter partial velocities meet in shared memory.
// No a2 operand, no extra MMA round trip for a 160-MAC-per-row layer.
template <int NC>
__device__ __forceinline__ void output_epilogue(uint32_t t_acc, const float* __restrict__ bias, const float2* __restrict__ w34,
                                                float& v0, float& v1) {
    v0 = 0.0f; v1 = 0.0f;
    uint32_t acc[NC], acc2[NC];
    tmem_ld<NC>(t_acc, acc);                     // a1_hi.W2_hi + a1_lo.W2_hi
    tmem_ld<NC>(t_acc + 80, acc2);               // a1_hi.W2_lo
    ptx::tcgen05_wait_ld();
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        const float y = lrelu02(__uint_as_float(acc[j]) + __uint_as_float(acc2[j]) + bias[j]);
        const float2 w = w34[j];
        v0 = fmaf(y, w.x, v0);
        v1 = fmaf(y, w.y, v1);
    }
}

__global__ void __launch_bounds__(X_THREADS, 1)
decode_fwd_tcx_kernel(const __grid_constant__ CUtensorMap noise_map /* [n_rows][32] fp32, box 128 x 32, 128-byte swizzle */,
                      const __half* __restrict__ w16, const __half* __restrict__ wsz16, const float* __restrict__ wf32,
                      const float* __restrict__ h0, const float* __restrict__ c0, const float* __restrict__ pooled,
                      const float* __restrict__ noise, const float* __restrict__ x_last, float* __restrict__ out,
                      int* __restrict__ status, int n_agents, long long n_rows, int n_next, int n_tiles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcxSmem& s = *reinterpret_cast<TcxSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lq = warp & 3, cq = warp >> 2;
    const int r = lq * 32 + lane;

    if (warp == 0) {
        ptx::tcgen05_alloc(ptx::cta_group_1, &s.tmem_base, 512u);
        ptx::tcgen05_relinquish_alloc_permit(ptx::cta_group_1);
    }
    // noise block of a tile: [128 rows][32] fp32 = 16 KB -> ONE tensor-map TMA copy into the hoist staging buffer (free from
    // the end of a tile's hoist to the next tile's prologue), issued one tile ahead.  128-byte swizzle: the 16-byte piece q of
    // row r lands at piece q ^ (r & 7), so the row-per-lane reads below are conflict free; rows past the batch arrive as zeros.
    auto prefetch_noise = [&](int tile) {
        mbar_expect_tx(&s.bar_z, X_ROWS * SW_Z * 4);
        tma_load_2d(s.stage, &noise_map, 0, tile * X_ROWS, &s.bar_z);
    };
    if (tid == 0) {
        ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.bar[0]), 1);
        ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.bar[1]), 1);
        ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.bar[2]), 1);
        ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.bar_w), 1);
        ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.bar_z), 1);
        ptx::fence_mbarrier_init(ptx::sem_release, ptx::scope_cluster);
        // all step weights (hi | lo fp16, 162 KB, already in the UMMA operand layout) + the fp32 tail: TMA, once per CTA
        constexpr uint32_t W_BYTES = XW_TOTAL * 2, W_PIECE = W_BYTES / 4, F_BYTES = XF_TOTAL * 4;
        static_assert(W_PIECE % 16 == 0 && F_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
        mbar_expect_tx(&s.bar_w, W_BYTES + F_BYTES);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            tma_load_1d(reinterpret_cast<unsigned char*>(s.w) + q * W_PIECE, reinterpret_cast<const unsigned char*>(w16) + q * W_PIECE,
                        W_PIECE, &s.bar_w);
        tma_load_1d(s.f32, wf32, F_BYTES, &s.bar_w);
        if ((int)blockIdx.x < n_tiles) prefetch_noise(blockIdx.x);
    }
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    ptx::tcgen05_fence_after_thread_sync();
    mbar_wait(&s.bar_w, 0u);                // weights have landed (async proxy -> visible to every thread and to the MMAs)
    const uint32_t tmem = __shfl_sync(0xffffffffu, s.tmem_base, 0);   // warp-uniform for the compiler (uniform datapath)
    const uint32_t tl = tmem + ((uint32_t)(lq * 32) << 16);       // this thread's lane, column 0
    uint32_t ph0 = 0, ph1 = 0, phz = 0;     // parities: bar0 (hoist, L1, L2) | bar1 / bar2 (gate halves) | bar_z (noise block)

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = (long long)tile * X_ROWS;
        const bool valid = row0 + r < n_rows;
        const int abase = (int)(row0 % n_agents);               // agent of tile row j = (abase + j) % n_agents (32-bit)
        const int agent = valid ? (abase + r) % n_agents : 0;
        // ---------------- tile prologue.  Every global load of the tile is issued up front, and every load is COALESCED:
        //   a warp instruction reads whole 128-byte lines (the row-per-lane loads of the first version cost 32 L1 tag
        //   lookups per instruction -- 10 K of the 18 K clk of this prologue).  [S ; z] goes through shared memory (S in the
        //   not-yet-used h operand region, z in the hoist staging buffer, 16-byte pieces XOR-swizzled by row so that both the
        //   row-contiguous writes and the row-per-lane reads are conflict-free) to reach the thread that owns the row's
        //   TMEM lane; h0 is loaded as (row, 8-column chunk) items, 8 rows x 128 B per instruction.  c0 is loaded last and
        //   first used at the end of step 0: its latency and tag traffic hide under the hoist and step-0 MMAs. --------------
        uint4 wreg[3];
        const uint4* wsz4 = reinterpret_cast<const uint4*>(wsz16);
        constexpr int CHUNK_U4 = XW_SZ_CHUNK / 8;                         // 1280 uint4 per chunk: 2.5 per thread
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (tid + q * X_THREADS < CHUNK_U4) wreg[q] = __ldg(wsz4 + tid + q * X_THREADS);
        float4 sreg[4], hreg[2][2];
        float2 xl = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {                                     // S tile [128][16 pieces]: piece g = tid + 512 i
            const int g = tid + i * X_THREADS, row = g >> 4, piece = g & 15;
            sreg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pooled && row0 + row < n_rows)
                sreg[i] = __ldg(reinterpret_cast<const float4*>(pooled + (size_t)((abase + row) % n_agents) * SW_H) + piece);
        }
        // (the z tile [128][8 pieces] arrives by TMA in the staging buffer: prefetched during the previous tile)
        const int hrow = warp * 8 + (lane & 7);                           // h0 items: (row, chunk = (lane >> 3) + 4 i)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            hreg[i][0] = hreg[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + hrow < n_rows) {
                const float4* src = reinterpret_cast<const float4*>(h0 + (size_t)((abase + hrow) % n_agents) * SW_H) + ((lane >> 3) + 4 * i) * 2;
                hreg[i][0] = __ldg(src);
                hreg[i][1] = __ldg(src + 1);
            }
        }
        if (cq == 1 && valid) xl = __ldg(reinterpret_cast<const float2*>(x_last + (size_t)agent * 4));
        {
            float4* sS = reinterpret_cast<float4*>(s.h);                  // [128 rows][16 pieces], piece' = piece ^ (row & 7)
            const float4* sZ = reinterpret_cast<const float4*>(s.stage);  // [128 rows][8 pieces], piece' = piece ^ (row & 7) (TMA swizzle)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int g = tid + i * X_THREADS, row = g >> 4, piece = g & 15;
                sS[row * 16 + (piece ^ (row & 7))] = sreg[i];
            }
            mbar_wait(&s.bar_z, phz); phz ^= 1;                           // this tile's noise block has landed
            __syncthreads();
            // [S ; z] (K = 96): this thread owns K 24cq .. 24cq+23 of its row = pieces 6cq .. 6cq+5 (S: 0..15, z: 16..23)
            // -> hi|lo TMEM A operand in R1: hi columns [160,208), lo [208,256)
            uint32_t hi[12], lo[12];
#pragma unroll
            for (int e = 0; e < 6; ++e) {
                const int piece = cq * 6 + e;
                const float4 v = piece < 16 ? sS[r * 16 + (piece ^ (r & 7))] : sZ[r * 8 + ((piece - 16) ^ (r & 7))];
                split2(v.x, v.y, hi[2 * e], lo[2 * e]);
                split2(v.z, v.w, hi[2 * e + 1], lo[2 * e + 1]);
            }
            tmem_st<12>(tl + XC_R1 + cq * 12, hi);
            tmem_st<12>(tl + XC_R1 + 48 + cq * 12, lo);
            ptx::tcgen05_wait_st();
            __syncthreads();                                              // the h region / staging buffer get their real contents now
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {                                     // h0 -> hi|lo operand chunks [chunk][row][8]
            uint32_t hi[4], lo[4];
            split2(hreg[i][0].x, hreg[i][0].y, hi[0], lo[0]);
            split2(hreg[i][0].z, hreg[i][0].w, hi[1], lo[1]);
            split2(hreg[i][1].x, hreg[i][1].y, hi[2], lo[2]);
            split2(hreg[i][1].z, hreg[i][1].w, hi[3], lo[3]);
            const size_t off = ((size_t)((lane >> 3) + 4 * i) * X_ROWS + hrow) * 8;
            *reinterpret_cast<uint4*>(s.h[0] + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(s.h[1] + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        float4 cq4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)     // cell state of this thread's units: c[0..7] = units 32 + 8cq .. (round 0), c[8..15] = units 8cq ..
            cq4[q] = valid ? __ldg(reinterpret_cast<const float4*>(c0 + (size_t)agent * SW_H + (q < 2 ? 32 : 0) + cq * 8) + (q & 1))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
        float c[16] = {cq4[0].x, cq4[0].y, cq4[0].z, cq4[0].w, cq4[1].x, cq4[1].y, cq4[1].z, cq4[1].w,
                       cq4[2].x, cq4[2].y, cq4[2].z, cq4[2].w, cq4[3].x, cq4[3].y, cq4[3].z, cq4[3].w};
        float p0 = xl.x, p1 = xl.y;
        bool out_of_range = false;          // fp16 range guard (see the row finish below), one predicate per thread
        // c1 = [S ; z] . W1[S,z rows]^T accumulated into TMEM [0,160): three K = 32 chunks of streamed weights
        for (int ch = 0; ch < 3; ++ch) {
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (tid + q * X_THREADS < CHUNK_U4) reinterpret_cast<uint4*>(s.stage)[tid + q * X_THREADS] = wreg[q];
            ptx::fence_proxy_async(ptx::space_shared);
            ptx::tcgen05_fence_before_thread_sync();
            __syncthreads();
            if (ch < 2) {
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    if (tid + q * X_THREADS < CHUNK_U4) wreg[q] = __ldg(wsz4 + (size_t)(ch + 1) * CHUNK_U4 + tid + q * X_THREADS);
            }
            if (warp == 0) {
                ptx::tcgen05_fence_after_thread_sync();
                if (elect_one()) {
                    mma3_ts<160, 160, 2>(tmem + XC_C1, tmem + XC_R1 + ch * 16, tmem + XC_R1 + 48 + ch * 16, s.stage,
                                         s.stage + XW_SZ_CHUNK / 2, ch > 0);
                    umma1_commit(&s.bar[0]);
                }
                __syncwarp();
            }
            mbar_wait(&s.bar[0], ph0); ph0 ^= 1;       // staging buffer is free again / c1 complete
            ptx::tcgen05_fence_after_thread_sync();
        }
        // the staging buffer is idle until the next tile's prologue: fetch that tile's noise block now (12 steps ahead of its use)
        if (tid == 0 && tile + (int)gridDim.x < n_tiles) prefetch_noise(tile + gridDim.x);
        {   // b1 joins c1 (the columns this thread reads back in the layer-1 epilogue)
            const int col0 = (cq < 2) ? cq * 48 : 96 + (cq - 2) * 32;
            if (cq < 2) fold_bias_into_c1<3>(tl + XC_C1 + col0, s.f32 + XF_B1 + col0);
            else        fold_bias_into_c1<2>(tl + XC_C1 + col0, s.f32 + XF_B1 + col0);
        }
        ptx::tcgen05_fence_before_thread_sync();
        __syncthreads();
        if (warp == 0) {    // layer 1 of step 0 (the later steps' layer-1 MMAs are issued from inside the previous gate epilogue)
            ptx::tcgen05_fence_after_thread_sync();
            if (elect_one()) {
                mma3_ss<160, 160, 4>(tmem + XC_R1, s.h[0], s.h[1], s.w + XW_W1H_HI, s.w + XW_W1H_LO);
                umma1_commit(&s.bar[0]);
            }
            __syncwarp();
        }

        for (int t = 0; t < n_next; ++t) {
            const bool feed_back = t + 1 < n_next;
            // ---------------- layer 1: h (K = 64, smem) -> 160 in R1; column quarters own 3, 3, 2, 2 K-blocks of a1 ----------------
            mbar_wait(&s.bar[0], ph0); ph0 ^= 1;
            ptx::tcgen05_fence_after_thread_sync();
            {
                const int col0 = (cq < 2) ? cq * 48 : 96 + (cq - 2) * 32;
                if (cq < 2) hidden_epilogue<3>(tl + XC_R1 + col0, tl + XC_C1 + col0);
                else        hidden_epilogue<2>(tl + XC_R1 + col0, tl + XC_C1 + col0);
            }
            ptx::tcgen05_fence_before_thread_sync();
            __syncthreads();
            // ---------------- layer 2: a1 (K = 160, TMEM, K block kb at column 16 kb) -> 80 in [320,400) ----------------
            // (Issuing these MMAs per K block from inside the layer-1 epilogue, so that layer 2 runs under it, was measured: no
            //  gain -- 0.812 vs 0.806 ms per step -- the TMEM-operand MMAs and the epilogue's tcgen05.ld/st share the TMEM ports.)
            if (warp == 0) {
                ptx::tcgen05_fence_after_thread_sync();
                // W2 hi and lo rows are stacked along N ([20 chunks][hi 80 | lo 80 rows][8]): one N = 160 pass yields a1_hi.W2_hi
                // in columns [0,80) and a1_hi.W2_lo in [80,160); a1_lo.W2_hi accumulates into [0,80).  20 MMAs instead of 30
                // (every tcgen05.mma carries a fixed cost); the epilogue adds the two column halves.
                if (elect_one()) {
                    umma1_ts<160, 160, 10, 16>(tmem + XC_RG, tmem + XC_R1, s.w + XW_W2_CAT, FMT_F16, false);
                    umma1_ts<160, 80, 10, 16>(tmem + XC_RG, tmem + XC_R1 + 8, s.w + XW_W2_CAT, FMT_F16, true);
                    umma1_commit(&s.bar[0]);
                    if (feed_back)     // gates, N half 1, h part -> [160,288): runs under the L2 / L34 epilogues (its x block and
                                       // commit follow once the velocity exists)
                        mma3_ss<256, 128, 4>(tmem + XC_R1, s.h[0], s.h[1], s.w + XW_WHH_HI + 128 * 8, s.w + XW_WHH_LO + 128 * 8);
                }
                __syncwarp();
            }
            mbar_wait(&s.bar[0], ph0); ph0 ^= 1;
            ptx::tcgen05_fence_after_thread_sync();
            {   // layer-2 epilogue + folded layers 3+4 (80 -> 2) on CUDA cores: partial velocity of this column quarter (20 columns)
                const int col0 = cq * 20;
                const float2* w34 = reinterpret_cast<const float2*>(s.f32 + XF_W34) + col0;
                float v0, v1;
                output_epilogue<20>(tl + XC_RG + col0, s.f32 + XF_B2 + col0, w34, v0, v1);
                s.vpart[(cq * 2 + 0) * X_ROWS + r] = v0;
                s.vpart[(cq * 2 + 1) * X_ROWS + r] = v1;
            }
            ptx::tcgen05_fence_before_thread_sync();
            __syncthreads();
            // ---------------- velocity, integration, emit; (p, v) -> hi|lo x block of the gate MMA ----------------
            if (cq == 1) {      // quarter 1 finishes the row
                const float v0 = s.vpart[0 * X_ROWS + r] + s.vpart[2 * X_ROWS + r] + s.vpart[4 * X_ROWS + r] + s.vpart[6 * X_ROWS + r] + s.f32[XF_B34];
                const float v1 = s.vpart[1 * X_ROWS + r] + s.vpart[3 * X_ROWS + r] + s.vpart[5 * X_ROWS + r] + s.vpart[7 * X_ROWS + r] + s.f32[XF_B34 + 1];
                p0 += v0; p1 += v1;
                // fp16 range guard: the operands of the split (x here, a1 upstream) live in fp16's exponent range; a layer-1
                // activation beyond 65 504 turns into inf -> NaN and reaches this velocity, a state beyond it would overflow
                // the next step's x block.  One predicate update per row and step (no branch in the step loop); the word is
                // raised once per tile, the host reads it when it next synchronises.
                out_of_range |= !(fmaxf(fmaxf(fabsf(p0), fabsf(p1)), fmaxf(fabsf(v0), fabsf(v1))) <= 6.0e4f);
                if (feed_back) {
                    uint32_t hp, lp, hv, lv;
                    split2(p0, p1, hp, lp);
                    split2(v0, v1, hv, lv);
                    *reinterpret_cast<uint4*>(s.xk + (size_t)r * 8) = make_uint4(hp, hv, lp, lv);                      // k 0..7
                    *reinterpret_cast<uint4*>(s.xk + (size_t)(X_ROWS + r) * 8) = make_uint4(hp, hv, 0x3C003C00u, 0u);  // k 8..15
                }
                if (valid)
                    *reinterpret_cast<float4*>(out + ((size_t)(row0 + r) * n_next + t) * 4) = make_float4(p0, p1, v0, v1);
            }
            if (!feed_back) { __syncthreads(); break; }
            ptx::fence_proxy_async(ptx::space_shared);
            __syncthreads();
            // ---------------- gates: x block of half 1 (its h part ran under the epilogues) -> bar2; half 0 (h part + x block)
            //                  -> [320,448) (the layer-2 accumulator is consumed) -> bar1.  Quarters 2,3 start their cell update
            //                  while the tensor pipe works on half 0. ----------------
            if (warp == 0) {
                ptx::tcgen05_fence_after_thread_sync();
                if (elect_one()) {
                    umma1_ss<256, 128, 1>(tmem + XC_R1, s.xk, s.w + XW_WXK + 128 * 8, FMT_F16, true);
                    umma1_commit(&s.bar[2]);
                    mma3_ss<256, 128, 4>(tmem + XC_RG, s.h[0], s.h[1], s.w + XW_WHH_HI, s.w + XW_WHH_LO);
                    umma1_ss<256, 128, 1>(tmem + XC_RG, s.xk, s.w + XW_WXK, FMT_F16, true);
                    umma1_commit(&s.bar[1]);
                }
                __syncwarp();
            }
            // ---------------- LSTM cell.  EVERY quarter first updates 8 units of gate half 1 (units 32 + 8cq .., ready early:
            //                  its h part ran under the epilogues) while the tensor pipe works on half 0, then 8 units of half 0
            //                  (units 8cq ..).  After each round the finished half of h (K blocks 2,3 / 0,1) goes straight into the
            //                  NEXT step's layer-1 MMAs, so half of layer 1 runs under the second round. ----------------
#pragma unroll
            for (int round = 0; round < 2; ++round) {
                if (round == 0) mbar_wait(&s.bar[2], ph1); else mbar_wait(&s.bar[1], ph1);
                ptx::tcgen05_fence_after_thread_sync();
                uint32_t a[32];
                tmem_ld<32>(tl + (round == 0 ? XC_R1 : XC_RG) + cq * 32, a);
                ptx::tcgen05_wait_ld();
                float hv[8];
#pragma unroll
                for (int u = 0; u < 8; u += 2) {
                    float g[2][4];
#pragma unroll
                    for (int w2 = 0; w2 < 2; ++w2)
#pragma unroll
                        for (int q = 0; q < 4; ++q) g[w2][q] = __uint_as_float(a[(u + w2) * 4 + q]);
                    lstm_cell_pair_prescaled_x2(g[0], g[1], c[round * 8 + u], c[round * 8 + u + 1], hv[u], hv[u + 1]);
                }
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split2(hv[2 * e], hv[2 * e + 1], hi[e], lo[e]);
                // h is an operand of the half-0 gate MMAs: nothing may overwrite it before they have completed (bar1; done
                // long before round 0 has finished its arithmetic -- the wait closes the race, it does not cost time)
                if (round == 0) mbar_wait(&s.bar[1], ph1);
                const size_t off = ((size_t)((round == 0 ? 4 : 0) + cq) * X_ROWS + r) * 8;
                *reinterpret_cast<uint4*>(s.h[0] + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(s.h[1] + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                ptx::fence_proxy_async(ptx::space_shared);
                ptx::tcgen05_fence_before_thread_sync();
                __syncthreads();
                if (warp == 0) {   // next step's layer 1, the two K blocks whose h chunks are complete
                    ptx::tcgen05_fence_after_thread_sync();
                    const int kb0 = round == 0 ? 2 : 0;
                    if (elect_one()) {
                        mma3_ss<160, 160, 2>(tmem + XC_R1, s.h[0] + kb0 * 2 * X_ROWS * 8, s.h[1] + kb0 * 2 * X_ROWS * 8,
                                             s.w + XW_W1H_HI + kb0 * 2 * 160 * 8, s.w + XW_W1H_LO + kb0 * 2 * 160 * 8, round == 1);
                        if (round == 1) umma1_commit(&s.bar[0]);
                    }
                    __syncwarp();
                }
            }
            ph1 ^= 1;
        }
        if (out_of_range && valid && status) atomicOr(status, 1);
    }
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) ptx::tcgen05_dealloc(ptx::cta_group_1, tmem, 512u);
}

}  // namespace sw

// cuTensorMapEncodeTiled through the runtime's driver-entry-point lookup (no link-time dependency on libcuda)
static int encode_noise_map(CUtensorMap* map, const float* noise, long long n_rows) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        SW_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !fn) return SW_ERR_UNSUPPORTED;
        encode = (EncodeFn)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)SW_Z, (cuuint64_t)n_rows};         // innermost first: 32 floats per row
    const cuuint64_t strides[1] = {(cuuint64_t)SW_Z * 4};                       // bytes between rows
    const cuuint32_t box[2] = {(cuuint32_t)SW_Z, (cuuint32_t)sw::X_ROWS}, elem[2] = {1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(noise), dims, strides, box, elem,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? SW_OK : SW_ERR_ARG;
}

extern "C" int sw_decode_fwd_tcx(const void* tcx_w16, const void* tcx_wsz16, const float* tcx_f32, const float* h0,
                                 const float* c0, const float* pooled, const float* noise, const float* x_last, float* out,
                                 int* status, int n_agents, int n_samples, int n_next, int sm_count, void* stream) {
    if (!tcx_w16 || !tcx_wsz16 || !tcx_f32 || !h0 || !c0 || !noise || !x_last || !out) return SW_ERR_ARG;
    if (n_agents <= 0 || n_samples <= 0 || n_next <= 0 || sm_count <= 0) return SW_ERR_ARG;
    const long long n_rows = (long long)n_agents * n_samples;
    const long long tiles = (n_rows + sw::X_ROWS - 1) / sw::X_ROWS;
    if (tiles > 0x7fffffffLL) return SW_ERR_UNSUPPORTED;
    if (((uintptr_t)noise & 15u) != 0) return SW_ERR_ARG;         // TMA: 16-byte aligned global address
    CUtensorMap noise_map;
    const int rc = encode_noise_map(&noise_map, noise, n_rows);
    if (rc != SW_OK) return rc;
    const int smem = (int)sizeof(sw::TcxSmem);
    SW_SET_MAX_SMEM(sw::decode_fwd_tcx_kernel, smem);
    const int grid = (int)(tiles < sm_count ? tiles : sm_count);
    sw::decode_fwd_tcx_kernel<<<grid, sw::X_THREADS, smem, (cudaStream_t)stream>>>(
        noise_map, (const __half*)tcx_w16, (const __half*)tcx_wsz16, tcx_f32, h0, c0, pooled, noise, x_last, out, status, n_agents, n_rows,
        n_next, (int)tiles);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}

extern "C" int sw_decode_tcx_pack_sizes(int* n_w16, int* n_wsz16, int* n_f32) {
    if (!n_w16 || !n_wsz16 || !n_f32) return SW_ERR_ARG;
    *n_w16 = sw::XW_TOTAL;
    *n_wsz16 = 3 * sw::XW_SZ_CHUNK;
    *n_f32 = sw::XF_TOTAL;
    return SW_OK;
}
