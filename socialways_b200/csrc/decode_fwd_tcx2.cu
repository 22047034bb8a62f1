// fp32-faithful tensor-core decode kernel, TWO 128-row tiles in flight per SM (CTA pairs, tcgen05 cta_group::2).
//
// Same computation and the same arithmetic as decode_fwd_tcx.cu (the loop of predict(), reference train.py:418-430, on fp16
// hi/lo split operands, 3 MMAs per product, fp32 accumulation in TMEM) -- re-organised so that the tensor pipe and the CUDA
// cores work on DIFFERENT tiles at the same time.  decode_fwd_tcx runs one tile per SM: one tile fills TMEM (480 of 512
// columns) and shared memory (224 KB, 162 KB of it weights), so its ~3.9 K clk of MMA time and ~7.5 K clk of epilogue time per
// step are serial (ncu: tensor pipe 35 %).  Here
//   * two CTAs on the two SMs of a TPC form a pair and issue every MMA as ONE cta_group::2 instruction (UMMA M = 256): each
//     CTA holds only HALF of every weight matrix (N/2 rows of B) -- 111 KB instead of 192 KB incl. the hoist weights, which
//     are now resident too;
//   * each CTA runs TWO tile slots (warps 0-7 / 8-15, thread = (TMEM lane = row, column half)); a slot owns 256 TMEM
//     columns, its own h / x operand buffers and its own barriers, and is a strictly serial MMA -> epilogue chain; the two
//     slots of an SM interleave by themselves (MMAs execute in issue order, whichever slot is in an epilogue leaves the
//     tensor pipe to the other);
//   * the step-invariant layer-1 term c1 (160 fp32 per row) no longer lives in TMEM: it is computed once per tile by MMAs,
//     parked in a per-slot scratch buffer in global memory (80 KB per slot, L2 resident: 23.7 MB for the chip) and re-read
//     by the thread that wrote it, coalesced, one K block ahead of its use.
// TMEM columns of a slot: [0,160) layer-1 accumulator -> a1 hi|lo in place | [160,240) layer-2 accumulator | gates half 0 ->
// [0,128) (queued behind the layer-2 MMAs, runs under the layer-2 epilogue), gates half 1 -> [128,256).  Tile prologue:
// [160,256) holds the [S ; z] hi|lo A operand of the hoist, [0,160) its result.
// Synchronisation: the threads of both CTAs signal "operands written" per warp on the LEADER CTA's `ready` mbarrier (remote
// arrive); the slot's ISSUING WARP -- warps 16 / 17 of the leader CTA, one per slot, which do nothing else -- waits for it,
// one elected lane issues the slot's MMAs back to back and commits with a multicast arrive on the `full` mbarriers of both CTAs.
// (First version: warp 0 of each slot issued, between its own epilogue work, through lane-predicated tcgen05.mma -- ptxas
// wraps each of those in an ELECT / BRA.U.ANY loop, ~15 instructions per MMA, ~1.5 K instructions per step on ONE warp while
// the slot's other 7 warps waited: 28 % of all warp samples sat on the MMA barriers.)  The issuing warps' warpgroup gives
// its registers to the epilogue warps (setmaxnreg: 32 / 112).
#include "decode_pair.cuh"

namespace sw {

struct Tcx2Smem {
    float zst[2][P_ROWS * SW_Z];             // noise block of each slot's tile (TMA, 128-byte swizzle; 1024-byte aligned)
    __half w[PW_TOTAL];                      // this rank's half of every weight matrix (113 664 B)
    __half h[2][2][8 * P_ROWS * 8];          // [slot][hi|lo][8 chunks][128][8]
    __half xk[2][2 * P_ROWS * 8];            // [slot] x-feedback A operand, one K block
    float f32[PF_TOTAL];
    float vpart[2][2 * P_ROWS];              // [slot][component][row]: partial velocity of column half 1
    unsigned long long ready[2];             // operands written (16 warp arrivals: 8 warps x 2 CTAs; used in the leader CTA)
    unsigned long long ready_x[2];           // x block written (8 arrivals: the 4 row-finishing warps x 2 CTAs)
    unsigned long long full[2][3];           // MMA completion: hoist / L1 / L2 | gates half 0 | gates half 1
    unsigned long long bar_z[2];             // TMA: the slot's noise block
    unsigned long long bar_w;                // TMA: weights
    uint32_t tmem_base;
};

constexpr int P_THREADS_ALL = P_THREADS + 128;   // + the issuing warps' warpgroup
__device__ __forceinline__ void slot_sync(int slot) { asm volatile("bar.sync %0, %1;" :: "r"(slot + 1), "n"(P_SLOT_THREADS) : "memory"); }
// partial-velocity exchange (named barrier 3 + slot): column half 1 only arrives, column half 0 (which finishes the rows) waits
__device__ __forceinline__ void vel_arrive2(int slot) { asm volatile("bar.arrive %0, %1;" :: "r"(slot + 3), "n"(P_SLOT_THREADS) : "memory"); }
__device__ __forceinline__ void vel_sync2(int slot) { asm volatile("bar.sync %0, %1;" :: "r"(slot + 3), "n"(P_SLOT_THREADS) : "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS_ALL, 1)
decode_fwd_tcx2_kernel(const __grid_constant__ CUtensorMap noise_map /* [n_rows][32] fp32, box 128 x 32, 128-byte swizzle */,
                       const __half* __restrict__ w16 /* [2 ranks][PW_TOTAL] */, const float* __restrict__ wf32,
                       const float* __restrict__ h0, const float* __restrict__ c0, const float* __restrict__ pooled,
                       const float* __restrict__ x_last, float* __restrict__ out, float4* __restrict__ scratch,
                       int* __restrict__ status, int n_agents, long long n_rows, int n_next, int n_tiles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Tcx2Smem& s = *reinterpret_cast<Tcx2Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int w8 = warp & 7, st = tid & (P_SLOT_THREADS - 1);
    const int lq = w8 & 3;                          // TMEM lane quarter (= warp % 4)
    int lane = tid & 31, slot = (warp >> 3) & 1, hf = w8 >> 2;      // tile slot, column half
    int r = lq * 32 + lane;
    // opaque: kept in registers (ptxas otherwise re-derives them from S2R tid, a ~25 clk special-register read, at every use)
    asm volatile("" : "+r"(lane), "+r"(slot), "+r"(hf), "+r"(r));
    const uint32_t cta = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_units = (n_tiles + 1) >> 1;         // work unit = two consecutive tiles, one per CTA of the pair

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(&s.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    auto prefetch_noise = [&](int sl, int tile) {
        mbar_expect_tx(&s.bar_z[sl], P_ROWS * SW_Z * 4);
        tma_load_2d(s.zst[sl], &noise_map, 0, tile * P_ROWS, &s.bar_z[sl]);
    };
    if (tid == 0) {
        for (int sl = 0; sl < 2; ++sl) {
            ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.ready[sl]), 16);
            ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.ready_x[sl]), 8);
            for (int j = 0; j < 3; ++j) ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.full[sl][j]), 1);
            ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.bar_z[sl]), 1);
        }
        ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.bar_w), 1);
        ptx::fence_mbarrier_init(ptx::sem_release, ptx::scope_cluster);
        constexpr uint32_t W_BYTES = PW_TOTAL * 2, W_PIECE = W_BYTES / 4, F_BYTES = PF_TOTAL * 4;
        static_assert(W_PIECE % 16 == 0 && F_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
        mbar_expect_tx(&s.bar_w, W_BYTES + F_BYTES);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(w16 + (size_t)cta * PW_TOTAL);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            tma_load_1d(reinterpret_cast<unsigned char*>(s.w) + q * W_PIECE, src + q * W_PIECE, W_PIECE, &s.bar_w);
        tma_load_1d(s.f32, wf32, F_BYTES, &s.bar_w);
        for (int sl = 0; sl < 2; ++sl) {
            const int u = sl + 2 * pair, tile = 2 * u + (int)cta;
            if (u < n_units && tile < n_tiles) prefetch_noise(sl, tile);
        }
    }
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    cluster_sync_all();                     // the peer's barriers exist before anything arrives on them
    ptx::tcgen05_fence_after_thread_sync();
    mbar_wait(&s.bar_w, 0u);
    const uint32_t tmem = __shfl_sync(0xffffffffu, s.tmem_base, 0);
    if (warp >= 16) {
        // =========================== the issuing warps (leader CTA: warp 16 -> slot 0, warp 17 -> slot 1) ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (cta == 0 && warp < 18) {
            const int sl = warp - 16;
            const uint32_t ts = tmem + (uint32_t)(sl * 256);
            uint32_t ph_r = 0, ph_x = 0;
            auto wait_ready = [&]() {
                mbar_wait(&s.ready[sl], ph_r); ph_r ^= 1;
                ptx::tcgen05_fence_after_thread_sync();
            };
            const __half* const h_hi = s.h[sl][0];
            const __half* const h_lo = s.h[sl][1];
            const __half* const xk = s.xk[sl];
            for (int u = sl + 2 * pair; u < n_units; u += 2 * n_pairs) {
                wait_ready();           // c1 = [S ; z] . W1[S,z rows]^T -> [0,160)
                if (elect_one()) {
                    pmma3_ts1<80, 6, 8>(ts + PC_R1, ts + PC_AHI, ts + PC_ALO, s.w + PW_WSZ_HI, s.w + PW_WSZ_LO);
                    umma1_commit_pair(&s.full[sl][0]);
                }
                __syncwarp();
                wait_ready();           // layer 1 of step 0
                if (elect_one()) {
                    pmma3_ss1<80, 4>(ts + PC_R1, h_hi, h_lo, s.w + PW_W1H_HI, s.w + PW_W1H_LO);
                    umma1_commit_pair(&s.full[sl][0]);
                }
                __syncwarp();
#pragma unroll 1
                for (int t = 0; t < n_next; ++t) {
                    const bool feed_back = t + 1 < n_next;
                    wait_ready();       // layer 2: a1 (K = 160, TMEM) -> [160,240); the h part of gates half 0 queued behind it
                    if (elect_one()) {
                        pmma3_ts1<P_L2NL, 10, 16>(ts + PC_R2, ts + PC_R1, ts + PC_R1 + 8, s.w + PW_W2_HI, s.w + PW_W2_LO);
                        umma1_commit_pair(&s.full[sl][0]);
                        if (feed_back) pmma3_ss1<64, 4>(ts, h_hi, h_lo, s.w + PW_WHH, s.w + PW_WHH + 4096);
                    }
                    __syncwarp();
                    if (!feed_back) break;
                    mbar_wait(&s.ready_x[sl], ph_x); ph_x ^= 1;
                    ptx::tcgen05_fence_after_thread_sync();
                    if (elect_one()) {  // gates: x block of half 0 -> full[1]; half 1 (h part + x block) -> [128,256) -> full[2]
                        pmma1_ss<64, 1>(ts, xk, s.w + PW_WXK, PFMT, true);
                        umma1_commit_pair(&s.full[sl][1]);
                        pmma3_ss1<64, 4>(ts + 128, h_hi, h_lo, s.w + PW_WHH + 8192, s.w + PW_WHH + 8192 + 4096);
                        pmma1_ss<64, 1>(ts + 128, xk, s.w + PW_WXK + 1024, PFMT, true);
                        umma1_commit_pair(&s.full[sl][2]);
                    }
                    __syncwarp();
                    wait_ready();       // next step's layer 1
                    if (elect_one()) {
                        pmma3_ss1<80, 4>(ts + PC_R1, h_hi, h_lo, s.w + PW_W1H_HI, s.w + PW_W1H_LO);
                        umma1_commit_pair(&s.full[sl][0]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
    // =========================== the epilogue warps: 8 per slot ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const uint32_t ts = tmem + (uint32_t)(slot * 256);                 // this slot's columns, lane 0
    const uint32_t tl = ts + ((uint32_t)(lq * 32) << 16);              // this thread's lane
    uint32_t ph_l = 0, ph_g = 0, ph_z = 0;
    unsigned long long* const bar_l = &s.full[slot][0];
    unsigned long long* const bar_g0 = &s.full[slot][1];
    unsigned long long* const bar_g1 = &s.full[slot][2];
    __half* const h_hi = s.h[slot][0];
    __half* const h_lo = s.h[slot][1];
    __half* const xk = s.xk[slot];
    float4* const my_scratch = scratch + ((size_t)blockIdx.x * 2 + slot) * P_SCRATCH_F4_PER_SLOT + r;
    // address of this slot's `ready` barrier in the LEADER CTA's shared memory (ready_x[slot] follows 16 bytes later)
    uint32_t ready_leader;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ready_leader) : "r"((uint32_t)__cvta_generic_to_shared(&s.ready[slot])), "r"(0));

    // every warp: "my operand writes are done" -> one arrival on the leader's barrier
    auto arrive_ready = [&]() {
        ptx::tcgen05_fence_before_thread_sync();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(ready_leader) : "memory");
    };
    auto wait_full = [&](unsigned long long* bar, uint32_t parity) {
        mbar_wait_cluster(bar, parity);
        ptx::tcgen05_fence_after_thread_sync();
    };

    for (int u = slot + 2 * pair; u < n_units; u += 2 * n_pairs) {
        const int tile = 2 * u + (int)cta;
        const bool has_tile = tile < n_tiles;                            // the odd last unit has one tile only
        const long long row0 = (long long)tile * P_ROWS;
        const bool valid = has_tile && row0 + r < n_rows;
        const int abase = (int)(row0 % n_agents);
        const int agent = valid ? (abase + r) % n_agents : 0;
        // ---------------- tile prologue: every global load coalesced and issued up front ----------------
        float4 sreg[8], hreg[4][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) {                                     // S tile [128][16 pieces]: piece g = st + 256 i
            const int g = st + i * P_SLOT_THREADS, row = g >> 4, piece = g & 15;
            sreg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pooled && has_tile && row0 + row < n_rows)
                sreg[i] = __ldg(reinterpret_cast<const float4*>(pooled + (size_t)((abase + row) % n_agents) * SW_H) + piece);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {                                 // h0 items: (row, 8-column chunk), 8 rows x 128 B per instruction
                const int hrow = w8 * 8 + (lane & 7) + 64 * j, chunk = (lane >> 3) + 4 * i;
                hreg[j * 2 + i][0] = hreg[j * 2 + i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (has_tile && row0 + hrow < n_rows) {
                    const float4* src = reinterpret_cast<const float4*>(h0 + (size_t)((abase + hrow) % n_agents) * SW_H) + chunk * 2;
                    hreg[j * 2 + i][0] = __ldg(src);
                    hreg[j * 2 + i][1] = __ldg(src + 1);
                }
            }
        float2 xl = make_float2(0.f, 0.f);
        if (hf == 0 && valid) xl = __ldg(reinterpret_cast<const float2*>(x_last + (size_t)agent * 4));
        {
            float4* sS = reinterpret_cast<float4*>(h_hi);                 // [128 rows][16 pieces], piece' = piece ^ (row & 7); 32 KB = h hi|lo
            const float4* sZ = reinterpret_cast<const float4*>(s.zst[slot]);   // [128 rows][8 pieces], TMA swizzle: piece' = piece ^ (row & 7)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int g = st + i * P_SLOT_THREADS, row = g >> 4, piece = g & 15;
                sS[row * 16 + (piece ^ (row & 7))] = sreg[i];
            }
            if (has_tile) { mbar_wait(&s.bar_z[slot], ph_z); ph_z ^= 1; }
            slot_sync(slot);
            // [S ; z] (K = 96 = 24 pieces): this thread owns pieces 12 hf .. 12 hf + 11 of its row = K blocks 3 hf .. 3 hf + 2
            uint32_t hi[24], lo[24];
#pragma unroll
            for (int e = 0; e < 12; ++e) {
                const int piece = hf * 12 + e;
                float4 v;
                if (piece < 16) v = sS[r * 16 + (piece ^ (r & 7))];
                else            v = has_tile ? sZ[r * 8 + ((piece - 16) ^ (r & 7))] : make_float4(0.f, 0.f, 0.f, 0.f);
                psplit2(v.x, v.y, hi[2 * e], lo[2 * e]);
                psplit2(v.z, v.w, hi[2 * e + 1], lo[2 * e + 1]);
            }
            tmem_st<24>(tl + PC_AHI + hf * 24, hi);
            tmem_st<24>(tl + PC_ALO + hf * 24, lo);
            ptx::tcgen05_wait_st();
            slot_sync(slot);                                              // staging consumed: h region and noise buffer are free
        }
        {   // next tile's noise block: 12 steps ahead of its use
            const int un = u + 2 * n_pairs, tn = 2 * un + (int)cta;
            if (st == 0 && un < n_units && tn < n_tiles) prefetch_noise(slot, tn);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {                                     // h0 -> hi|lo operand chunks [chunk][row][8]
            const int hrow = w8 * 8 + (lane & 7) + 64 * (q >> 1), chunk = (lane >> 3) + 4 * (q & 1);
            uint32_t hi[4], lo[4];
            psplit2(hreg[q][0].x, hreg[q][0].y, hi[0], lo[0]);
            psplit2(hreg[q][0].z, hreg[q][0].w, hi[1], lo[1]);
            psplit2(hreg[q][1].x, hreg[q][1].y, hi[2], lo[2]);
            psplit2(hreg[q][1].z, hreg[q][1].w, hi[3], lo[3]);
            const size_t off = ((size_t)chunk * P_ROWS + hrow) * 8;
            *reinterpret_cast<uint4*>(h_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(h_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        ptx::fence_proxy_async(ptx::space_shared);
        arrive_ready();
        // cell state of this thread's units: c[0..15] = units 16 hf .. (gates half 0), c[16..31] = units 32 + 16 hf .. (half 1)
        float c[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 v = valid ? __ldg(reinterpret_cast<const float4*>(c0 + (size_t)agent * SW_H + (q >> 2) * 32 + hf * 16) + (q & 3))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
        }
        float p0 = xl.x, p1 = xl.y;
        bool out_of_range = false;
        wait_full(bar_l, ph_l); ph_l ^= 1;
#pragma unroll
        for (int kb = 0; kb < 5; ++kb) {       // c1 + b1 -> scratch (this thread's 5 K blocks of 16 columns)
            const int col0 = (hf * 5 + kb) * 16;
            uint32_t v[16];
            tmem_ld<16>(tl + PC_R1 + col0, v);
            ptx::tcgen05_wait_ld();
#pragma unroll
            for (int q = 0; q < 4; ++q)
                __stcg(my_scratch + ((hf * 5 + kb) * 4 + q) * P_ROWS,
                       make_float4(__uint_as_float(v[4 * q]) + s.f32[PF_B1 + col0 + 4 * q], __uint_as_float(v[4 * q + 1]) + s.f32[PF_B1 + col0 + 4 * q + 1],
                                   __uint_as_float(v[4 * q + 2]) + s.f32[PF_B1 + col0 + 4 * q + 2], __uint_as_float(v[4 * q + 3]) + s.f32[PF_B1 + col0 + 4 * q + 3]));
        }
        arrive_ready();

        for (int t = 0; t < n_next; ++t) {
            const bool feed_back = t + 1 < n_next;
            // ---------------- layer 1 epilogue: a1 = lrelu(acc + c1) -> hi|lo in place; c1 read one K block ahead ----------------
            float4 cn[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) cn[q] = __ldcg(my_scratch + ((hf * 5) * 4 + q) * P_ROWS);
            wait_full(bar_l, ph_l); ph_l ^= 1;
#pragma unroll
            for (int kb = 0; kb < 5; ++kb) {
                const float4 cc[4] = {cn[0], cn[1], cn[2], cn[3]};
                if (kb < 4) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) cn[q] = __ldcg(my_scratch + ((hf * 5 + kb + 1) * 4 + q) * P_ROWS);
                }
                uint32_t acc[16], pc[16];
                const uint32_t ta = tl + PC_R1 + (hf * 5 + kb) * 16;
                tmem_ld<16>(ta, acc);
                ptx::tcgen05_wait_ld();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float y0 = __uint_as_float(acc[4 * q]) + cc[q].x, y1 = __uint_as_float(acc[4 * q + 1]) + cc[q].y;
                    const float y2 = __uint_as_float(acc[4 * q + 2]) + cc[q].z, y3 = __uint_as_float(acc[4 * q + 3]) + cc[q].w;
                    psplit2(lrelu02(y0), lrelu02(y1), pc[2 * q], pc[8 + 2 * q]);
                    psplit2(lrelu02(y2), lrelu02(y3), pc[2 * q + 1], pc[8 + 2 * q + 1]);
                }
                tmem_st<16>(ta, pc);
            }
            ptx::tcgen05_wait_st();
            arrive_ready();
            // ---------------- layer 2: a1 (K = 160, TMEM, K block kb at column 16 kb) -> 80 columns in [160,240); the h part of
            //                  gates half 0 is queued right behind it (overwrites a1 once layer 2 has consumed it) ----------------
            wait_full(bar_l, ph_l); ph_l ^= 1;
            float v0 = 0.0f, v1 = 0.0f;
            {   // layer-2 epilogue + folded layers 3+4 (80 -> 2): partial velocity over this thread's 40 columns
                uint32_t acc[40];
                tmem_ld<40>(tl + PC_R2 + hf * 40, acc);
                ptx::tcgen05_wait_ld();
                const float* b2 = s.f32 + PF_B2 + hf * 40;
                const float2* w34 = reinterpret_cast<const float2*>(s.f32 + PF_W34) + hf * 40;
#pragma unroll
                for (int j = 0; j < 40; ++j) {
                    const float y = lrelu02(__uint_as_float(acc[j]) + b2[j]);
                    const float2 w = w34[j];
                    v0 = fmaf(y, w.x, v0);
                    v1 = fmaf(y, w.y, v1);
                }
            }
            ptx::tcgen05_fence_before_thread_sync();
            if (hf == 1) { s.vpart[slot][r] = v0; s.vpart[slot][P_ROWS + r] = v1; vel_arrive2(slot); }
            else {              // column half 0 finishes the row: velocity, integration, emit; (p, v) -> hi|lo x block of the gate MMA
                vel_sync2(slot);
                v0 += s.vpart[slot][r] + s.f32[PF_B34];
                v1 += s.vpart[slot][P_ROWS + r] + s.f32[PF_B34 + 1];
                p0 += v0; p1 += v1;
                out_of_range |= !(fmaxf(fmaxf(fabsf(p0), fabsf(p1)), fmaxf(fabsf(v0), fabsf(v1))) <= 6.0e4f);   // fp16 range guard
                if (feed_back) {
                    uint32_t hp, lp, hv, lv;
                    psplit2(p0, p1, hp, lp);
                    psplit2(v0, v1, hv, lv);
                    *reinterpret_cast<uint4*>(xk + (size_t)r * 8) = make_uint4(hp, hv, lp, lv);                      // k 0..7
                    *reinterpret_cast<uint4*>(xk + (size_t)(P_ROWS + r) * 8) = make_uint4(hp, hv, 0x3C003C00u, 0u);  // k 8..15
                    ptx::fence_proxy_async(ptx::space_shared);
                    ptx::tcgen05_fence_before_thread_sync();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(ready_leader + 16u) : "memory");
                }
                if (valid)
                    *reinterpret_cast<float4*>(out + ((size_t)(row0 + r) * n_next + t) * 4) = make_float4(p0, p1, v0, v1);
            }
            if (!feed_back) break;
            // ---------------- LSTM cell: 16 units of half 0 (units 16 hf ..), then 16 units of half 1 (units 32 + 16 hf ..) ----------------
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (half == 0) wait_full(bar_g0, ph_g); else ptx::tcgen05_fence_after_thread_sync();
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    uint32_t a[32];
                    tmem_ld<32>(tl + half * 128 + hf * 64 + ch * 32, a);
                    ptx::tcgen05_wait_ld();
                    float hv[8];
#pragma unroll
                    for (int uu = 0; uu < 8; uu += 2) {
                        float g[2][4];
#pragma unroll
                        for (int w2 = 0; w2 < 2; ++w2)
#pragma unroll
                            for (int q = 0; q < 4; ++q) g[w2][q] = __uint_as_float(a[(uu + w2) * 4 + q]);
                        lstm_cell_pair_prescaled(g[0], g[1], c[half * 16 + ch * 8 + uu], c[half * 16 + ch * 8 + uu + 1], hv[uu], hv[uu + 1]);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) psplit2(hv[2 * e], hv[2 * e + 1], hi[ch * 4 + e], lo[ch * 4 + e]);
                }
                // h is an operand of the half-1 gate MMAs: nothing may overwrite it before they have completed
                if (half == 0) wait_full(bar_g1, ph_g);
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    const size_t off = ((size_t)(half * 4 + hf * 2 + ch) * P_ROWS + r) * 8;
                    *reinterpret_cast<uint4*>(h_hi + off) = make_uint4(hi[ch * 4], hi[ch * 4 + 1], hi[ch * 4 + 2], hi[ch * 4 + 3]);
                    *reinterpret_cast<uint4*>(h_lo + off) = make_uint4(lo[ch * 4], lo[ch * 4 + 1], lo[ch * 4 + 2], lo[ch * 4 + 3]);
                }
            }
            ph_g ^= 1;
            ptx::fence_proxy_async(ptx::space_shared);
            arrive_ready();
        }
        if (out_of_range && valid && status) atomicOr(status, 1);
    }
    }   // epilogue warps
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

}  // namespace sw

static int tcx2_grid(long long tiles, int sm_count) {
    const long long units = (tiles + 1) / 2;            // two tiles (one per CTA of a pair) per unit, two unit slots per pair
    long long pairs = (units + 1) / 2;
    if (pairs > sm_count / 2) pairs = sm_count / 2;
    if (pairs < 1) pairs = 1;
    return (int)(2 * pairs);
}

extern "C" long long sw_decode_tcx2_scratch_bytes(int sm_count) {
    if (sm_count < 2) return 0;
    return (long long)(sm_count / 2) * 2 * 2 * sw::P_SCRATCH_F4_PER_SLOT * 16;
}

extern "C" int sw_decode_fwd_tcx2(const void* tcx2_w16, const float* tcx2_f32, const float* h0, const float* c0,
                                  const float* pooled, const float* noise, const float* x_last, float* out, void* scratch,
                                  long long scratch_bytes, int* status, int n_agents, int n_samples, int n_next, int sm_count,
                                  void* stream) {
    if (!tcx2_w16 || !tcx2_f32 || !h0 || !c0 || !noise || !x_last || !out || !scratch) return SW_ERR_ARG;
    if (n_agents <= 0 || n_samples <= 0 || n_next <= 0 || sm_count < 2) return SW_ERR_ARG;
    if (scratch_bytes < sw_decode_tcx2_scratch_bytes(sm_count) || ((uintptr_t)scratch & 15u) != 0) return SW_ERR_ARG;
    const long long n_rows = (long long)n_agents * n_samples;
    const long long tiles = (n_rows + sw::P_ROWS - 1) / sw::P_ROWS;
    if (tiles > 0x3fffffffLL) return SW_ERR_UNSUPPORTED;
    if (((uintptr_t)noise & 15u) != 0) return SW_ERR_ARG;
    CUtensorMap noise_map;
    const int rc = encode_noise_map2(&noise_map, noise, n_rows);
    if (rc != SW_OK) return rc;
    const int smem = (int)sizeof(sw::Tcx2Smem);
    SW_SET_MAX_SMEM(sw::decode_fwd_tcx2_kernel, smem);
    const int grid = tcx2_grid(tiles, sm_count);
    sw::decode_fwd_tcx2_kernel<<<grid, sw::P_THREADS_ALL, smem, (cudaStream_t)stream>>>(
        noise_map, (const __half*)tcx2_w16, tcx2_f32, h0, c0, pooled, x_last, out, (float4*)scratch, status, n_agents, n_rows, n_next,
        (int)tiles);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}

extern "C" int sw_decode_tcx2_pack_sizes(int* n_w16, int* n_f32) {
    if (!n_w16 || !n_f32) return SW_ERR_ARG;
    *n_w16 = 2 * sw::PW_TOTAL;
    *n_f32 = sw::PF_TOTAL;
    return SW_OK;
}
