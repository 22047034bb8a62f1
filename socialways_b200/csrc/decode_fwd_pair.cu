// fp32-faithful tensor-core decode kernel, TWO 128-row tiles in flight per SM: CTA pairs (tcgen05 cta_group::2), the tensor
// pipe and the CUDA cores in PING-PONG over two tile slots.  The default decode kernel of the inference path.
//
// Same computation and arithmetic as decode_fwd_tcx.cu (the loop of predict(), reference train.py:418-430, on fp16 hi/lo split
// operands, 3 MMAs per product, fp32 accumulation in TMEM).  decode_fwd_tcx runs one tile per SM: one tile fills TMEM (480 of
// 512 columns) and shared memory (224 KB, 162 KB of it weights), so its MMA time and its epilogue time per step are serial.  Here
//   * two CTAs on the two SMs of a TPC form a pair and issue every MMA as ONE cta_group::2 instruction (UMMA M = 256): each CTA
//     holds only HALF of every weight matrix (N/2 rows of B) -- 111 KB, hoist weights included -- which leaves room for TWO
//     tile slots per CTA (a slot = 256 TMEM columns, its own h / x operand buffers and barriers);
//   * the step-invariant layer-1 term c1 (160 fp32 per row) does not live in TMEM: it is computed once per tile by MMAs, parked in a
//     per-slot scratch buffer in global memory (80 KB per slot, L2 resident: 23.7 MB for the chip) and re-read by the thread that
//     wrote it, coalesced, ahead of its use;
//   * ALL 16 epilogue warps (thread = (TMEM lane = row, column quarter)) serve BOTH slots, half a step apart:
//         L1 epilogue (0,t) | cell update (1,t-1) | L2 epilogue + row finish (0,t) | L1 epilogue (1,t) | cell (0,t) | L2 (1,t)
//     so that each MMA group of one slot runs under an epilogue phase of the other that is longer than it (layer 2 + the h part
//     of gates half 0, ~2.5 K clk, under the ~4.5 K clk cell update; layer 1 and the gate x blocks under the L2 / L1 epilogues);
//   * a 17th warp of the leader CTA does nothing but issue: it walks the same static schedule, waits for the "operands written"
//     arrivals of a slot (16 warps x 2 CTAs, remote mbarrier arrives), ONE elected lane issues the group's MMAs back to back and
//     commits with a multicast arrive on the slot's `full` barriers in both CTAs.  Its warpgroup gives its registers to the
//     epilogue warps (setmaxnreg 32 / 112: a 17th warp alone would cap every thread at 96);
//   * quarter -> work assignment is mirrored between the slots (layer-1 K blocks 3,3,2,2 in slot 0 and 2,2,3,3 in slot 1, the
//     row finish on quarter 2 / quarter 1), so every warp carries the same load over a slot pair.
// TMEM columns of a slot: [0,160) layer-1 accumulator -> a1 hi|lo in place | [160,240) layer-2 accumulator | gates half 0 ->
// [0,128) (queued behind the layer-2 MMAs), gates half 1 -> [128,256).  Tile prologue: [160,256) holds the [S ; z] hi|lo A
// operand of the hoist, [0,160) its result.
// Measured history (B200, 2.62 M trajectories x 12 steps; decode_fwd_tcx = 10.03 ms): slot-private warps with lane-predicated
// issue from an epilogue warp 9.64 ms (ncu: 28 % of warp samples on MMA barriers -- ptxas wraps a lane-predicated tcgen05.mma in
// an ELECT / BRA.U.ANY loop, ~15 instructions per MMA, ~1.5 K instructions per step on ONE warp); the same with dedicated issuing
// warps 9.15 ms; ping-pong in lockstep order 9.78 ms (layer 2 of one slot does not fit under the 1.1 K clk L1 epilogue of the
// other); half-step order with the lane-predicated issue 10.37 ms (the issuing warp never caught up: 175 clk per MMA);
// single-lane issue 8.71 ms; non-blocking velocity exchange + indices kept in registers 8.24 ms.
#include <type_traits>

#include "decode_pair.cuh"

#if SW_PAIR_BF16        // the bf16 build of this file (decode_fwd_pair_bf16.cu): own kernel and entry point, shared host helpers
#define PAIR_KERNEL decode_fwd_pair_bf16_kernel
#define PAIR_ENTRY sw_decode_fwd_pair_bf16
#else
#define PAIR_KERNEL decode_fwd_pair_kernel
#define PAIR_ENTRY sw_decode_fwd_pair
#endif

namespace sw {

#ifdef SW_PAIR_TRACE      // timeline of one tile pair (scripts/pair_trace.py): clock64 of thread 0 of CTA 0 at the marked points
__device__ long long g_pair_trace[64];
#define SW_TR(i) do { if (tr_on) g_pair_trace[i] = clock64(); } while (0)
#else
#define SW_TR(i) do { } while (0)
#endif

constexpr int Q_EPI = 512;            // epilogue threads (warps 0..15)
constexpr int Q_NB = 3, Q_NS = 2;     // layer-1 K blocks of a "big" / "small" quarter (2 big + 2 small = 10 blocks; 4 / 1 measured equal)
constexpr int Q_THREADS = 640;        // + the issuing warp's warpgroup (register file = 4 x 16 K: a 17th warp alone would cap
                                      //   every thread at 96 registers; setmaxnreg moves the idle group's registers over)

struct PairSmem {
    float zst[2][P_ROWS * SW_Z];             // noise block of each slot's tile (TMA, 128-byte swizzle; 1024-byte aligned)
    __half w[PW_TOTAL];                      // this rank's half of every weight matrix (113 664 B)
    __half h[2][2][8 * P_ROWS * 8];          // [slot][hi|lo][8 chunks][128][8]
    __half xk[2][2 * P_ROWS * 8];            // [slot] x-feedback A operand, one K block
    float f32[PF_TOTAL];
    float vpart[2][8 * P_ROWS];              // [slot][quarter * 2 + component][row]: partial velocities
    unsigned long long ready[2];             // operands written: 32 warp arrivals (16 warps x 2 CTAs), used in the leader CTA
    unsigned long long ready_x[2];           // x block written: 8 arrivals (the 4 finishing warps x 2 CTAs)
    unsigned long long full[2][3];           // MMA completion: hoist / L1 / L2 | gates half 0 | gates half 1
    unsigned long long vfree[2];             // partial velocities consumed by the finishing warps (4 warp arrivals per step)
    unsigned long long bar_z[2];             // TMA: the slot's noise block
    unsigned long long bar_w;                // TMA: weights
    uint32_t tmem_base;
};
static_assert(sizeof(PairSmem) <= 227 * 1024, "shared memory of one CTA");

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" :: "n"(Q_EPI) : "memory"); }
// partial-velocity exchange of a slot (named barrier 2 + slot): the 12 contributing warps only ARRIVE (they never block), the 4
// finishing warps wait for all 16
__device__ __forceinline__ void vel_arrive(int sl) { asm volatile("bar.arrive %0, %1;" :: "r"(2 + sl), "n"(Q_EPI) : "memory"); }
__device__ __forceinline__ void vel_sync(int sl) { asm volatile("bar.sync %0, %1;" :: "r"(2 + sl), "n"(Q_EPI) : "memory"); }
// ... and the way back (mbarrier vfree[slot], 4 warp arrivals): the finishing warps announce "partials consumed", a contributing
// warp checks it before it overwrites the buffer a step later.  The arrival is a whole step old by then (the next write sits
// behind this slot's cell update, layer-1 MMAs and layer-1 epilogue, all of which wait for the finishing warps), so the check
// passes at once; it makes the write-after-read order explicit instead of implied by the mbarrier / MMA chain.  (A named
// barrier in its place -- bar.sync for the 12 contributing warps -- made those warps wait for EACH OTHER: 2 % of all samples.)
// 32 bytes of a row in ONE request (LDG.256): the per-row loads of the tile prologue touch a different cache line per lane, and
// the tag stage of L1 serves ~1 line per clock -- with 16-byte loads the c0 rows of a tile pair took ~7 K clk to ISSUE
__device__ __forceinline__ void ldg256(const float* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

__device__ __forceinline__ void wait_full3(unsigned long long* bar, uint32_t parity) {
    mbar_wait(bar, parity);
    ptx::tcgen05_fence_after_thread_sync();
}

// layer-1 epilogue of NKB K blocks: a1 = lrelu(acc + c1) -> hi|lo fp16 written in place (hi -> columns +0..7, lo -> +8..15 of
// the 16 accumulator columns the thread has just read); c1 (+ b1) comes from this thread's scratch lines, one block ahead
// `third`: the thread owns three blocks (a "big" quarter), else two.  ONE copy of the block code serves both cases (the step
// loop is ~50 KB of SASS and sensitive to its size: a second copy of the step costs 18 %).
__device__ __forceinline__ void l1_epilogue(uint32_t t_acc, const float4* sc, unsigned long long* bar, uint32_t parity, float4 (&cn)[4],
                                            bool third) {
    // cn = c1 of the first block, loaded a whole phase earlier (c1_prefetch): the scratch lives in L2 (~700 clk away; the L1 is
    // all shared memory here), and a load issued just before the barrier wait was exposed at every phase start
    wait_full3(bar, parity);
#pragma unroll
    for (int kb = 0; kb < Q_NB; ++kb) {
        if (kb == Q_NB - 1 && !third) break;
        const float4 cc[4] = {cn[0], cn[1], cn[2], cn[3]};
        if (kb + 1 < Q_NS || (kb + 1 < Q_NB && third)) {
#pragma unroll
            for (int q = 0; q < 4; ++q) cn[q] = __ldcg(sc + ((kb + 1) * 4 + q) * P_ROWS);
        }
        uint32_t acc[16], pc[16];
        tmem_ld<16>(t_acc + kb * 16, acc);
        ptx::tcgen05_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            l1_pair(acc[4 * q], acc[4 * q + 1], cc[q].x, cc[q].y, pc[2 * q], pc[8 + 2 * q]);
            l1_pair(acc[4 * q + 2], acc[4 * q + 3], cc[q].z, cc[q].w, pc[2 * q + 1], pc[8 + 2 * q + 1]);
        }
#if SW_PAIR_BF16
        tmem_st<8>(t_acc + kb * 16, pc);                  // hi only
#else
        tmem_st<16>(t_acc + kb * 16, pc);
#endif
    }
    ptx::tcgen05_wait_st();
}

// c1 + b1 of this thread's K blocks (2, or 3 with `third`): TMEM -> its scratch lines (once per tile); all loads in flight before the one wait
__device__ __forceinline__ void c1_to_scratch(uint32_t t_acc, float4* sc, const float* __restrict__ b1, bool third) {
    uint32_t v[Q_NB][16];
#pragma unroll
    for (int kb = 0; kb < Q_NB; ++kb)
        if (kb < Q_NS || third) tmem_ld<16>(t_acc + kb * 16, v[kb]);
    ptx::tcgen05_wait_ld();
#pragma unroll
    for (int kb = 0; kb < Q_NB; ++kb) {
        if (kb >= Q_NS && !third) break;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 b = *reinterpret_cast<const float4*>(b1 + kb * 16 + 4 * q);
            __stcg(sc + (kb * 4 + q) * P_ROWS, make_float4(__uint_as_float(v[kb][4 * q]) + b.x, __uint_as_float(v[kb][4 * q + 1]) + b.y,
                                                          __uint_as_float(v[kb][4 * q + 2]) + b.z, __uint_as_float(v[kb][4 * q + 3]) + b.w));
        }
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Q_THREADS, 1)
PAIR_KERNEL(const __grid_constant__ CUtensorMap noise_map /* [n_rows][32] fp32, box 128 x 32, 128-byte swizzle */,
                       const __half* __restrict__ w16 /* [2 ranks][PW_TOTAL] */, const float* __restrict__ wf32,
                       const float* __restrict__ h0, const float* __restrict__ c0, const float* __restrict__ pooled,
                       const float* __restrict__ x_last, float* __restrict__ out, float4* __restrict__ scratch,
                       int* __restrict__ status, int n_agents, long long n_rows, int n_next, int n_tiles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    PairSmem& s = *reinterpret_cast<PairSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int lq = warp & 3;                        // TMEM lane quarter (= warp % 4)
    int lane = tid & 31, cq = (warp >> 2) & 3;      // column quarter
    int r = lq * 32 + lane;
    // opaque: kept in registers (ptxas otherwise re-derives them from S2R tid -- a ~25 clk special-register read -- at every use
    // inside the step loop: 13 % of the layer-1 epilogue's stall samples)
    asm volatile("" : "+r"(cq), "+r"(lane), "+r"(r));
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
            ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.ready[sl]), 32);
            ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.ready_x[sl]), 8);
            for (int j = 0; j < 3; ++j) ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.full[sl][j]), 1);
            ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.bar_z[sl]), 1);
            ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.vfree[sl]), 4);
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
            const int u = 2 * pair + sl, tile = 2 * u + (int)cta;
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
        // =========================== the issuing warp (leader CTA only): the static MMA schedule ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (cta == 0 && warp == 16) {
            uint32_t ph_r = 0, ph_x = 0;            // bit sl = parity of the slot's ready / ready_x barrier
            auto wait_ready = [&](int sl) {
                mbar_wait(&s.ready[sl], (ph_r >> sl) & 1u); ph_r ^= 1u << sl;
                ptx::tcgen05_fence_after_thread_sync();
            };
            // Every group: the whole warp waits for the arrivals, ONE elected lane issues the MMAs and the commit back to back
            // (sw_umma.cuh: the lane-predicated forms cost ~15 instructions per MMA, which made this warp the bottleneck).
            auto issue_hoist = [&](int sl) {
                const uint32_t ts = tmem + (uint32_t)(sl * 256);
                wait_ready(sl);         // c1 = [S ; z] . W1[S,z rows]^T -> [0,160)
                if (elect_one()) {
                    pmma3_ts1<80, 6, 8>(ts + PC_R1, ts + PC_AHI, ts + PC_ALO, s.w + PW_WSZ_HI, s.w + PW_WSZ_LO);
                    umma1_commit_pair(&s.full[sl][0]);
                }
                __syncwarp();
            };
            auto issue_l1 = [&](int sl) {
                const uint32_t ts = tmem + (uint32_t)(sl * 256);
                wait_ready(sl);         // layer 1: h (K = 64, smem) -> [0,160)
                if (elect_one()) {
                    pmma3_ss1<80, 4>(ts + PC_R1, s.h[sl][0], s.h[sl][1], s.w + PW_W1H_HI, s.w + PW_W1H_LO);
                    umma1_commit_pair(&s.full[sl][0]);
                }
                __syncwarp();
            };
            auto issue_l2 = [&](int sl, bool feed_back) {
                const uint32_t ts = tmem + (uint32_t)(sl * 256);
                wait_ready(sl);         // layer 2: a1 (K = 160, TMEM) -> [160,240); the h part of gates half 0 queued behind it
                if (elect_one()) {
                    pmma3_ts1<P_L2NL, 10, 16>(ts + PC_R2, ts + PC_R1, ts + PC_R1 + 8, s.w + PW_W2_HI, s.w + PW_W2_LO);
                    umma1_commit_pair(&s.full[sl][0]);
                    if (feed_back) pmma3_ss1<64, 4>(ts, s.h[sl][0], s.h[sl][1], s.w + PW_WHH, s.w + PW_WHH + 4096);
                }
                __syncwarp();
            };
            auto issue_x = [&](int sl) {
                const uint32_t ts = tmem + (uint32_t)(sl * 256);
                mbar_wait(&s.ready_x[sl], (ph_x >> sl) & 1u); ph_x ^= 1u << sl;
                ptx::tcgen05_fence_after_thread_sync();
                if (elect_one()) {
                    pmma1_ss<64, 1>(ts, s.xk[sl], s.w + PW_WXK, PFMT, true);
                    umma1_commit_pair(&s.full[sl][1]);
                    pmma3_ss1<64, 4>(ts + 128, s.h[sl][0], s.h[sl][1], s.w + PW_WHH + 8192, s.w + PW_WHH + 8192 + 4096);
                    pmma1_ss<64, 1>(ts + 128, s.xk[sl], s.w + PW_WXK + 1024, PFMT, true);
                    umma1_commit_pair(&s.full[sl][2]);
                }
                __syncwarp();
            };
            for (int ub = 2 * pair; ub < n_units; ub += 2 * n_pairs) {
                // (both slots always run: the slot of a unit or tile beyond the batch works on zero rows -- at most one tile per
                //  pair at the very end -- which keeps every "is the slot active" test out of the step loops)
#pragma unroll 1
                for (int sl = 0; sl < 2; ++sl) issue_hoist(sl);
#pragma unroll 1
                for (int sl = 0; sl < 2; ++sl) issue_l1(sl);           // layer 1 of step 0
                // Step loop.  The slots run HALF A STEP apart (slot 1 behind): the epilogue warps visit
                //   L1(0,t)  cell(1,t-1)  L2(0,t)  L1(1,t)  cell(0,t)  L2(1,t)
                // so that the long MMA group (layer 2 + the h part of gates half 0: ~2.5 K clk) of one slot runs under the long
                // epilogue phase (the cell update, ~5 K clk) of the other.  This warp issues in the order the arrivals come:
                // one rolled loop over the six groups of a step, slot = e & 1, group = e % 3.
#pragma unroll 1
                for (int t = 0; t < n_next; ++t) {
                    const bool feed_back = t + 1 < n_next;
#pragma unroll 1
                    for (int e = 0; e < 6; ++e) {
                        const int sl = e & 1, kind = e % 3;
                        if (kind == 0) issue_l2(sl, feed_back);
                        else if (kind == 1) { if (sl == 1 ? t > 0 : feed_back) issue_l1(sl); }
                        else if (feed_back) issue_x(sl);
                    }
                }
            }
        }
    } else {
        // =========================== the 16 epilogue warps ===========================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        const int etid = r + (cq << 7), ewarp = etid >> 5;                // = threadIdx.x / warp index of these warps, from the pinned registers
        const uint32_t tl = tmem + ((uint32_t)(lq * 32) << 16);           // this thread's lane, slot 0, column 0
        uint32_t ph = 0;                    // barrier parities, one register: bit sl = full[sl][0], bit 2 + sl = gates, bit 4 + sl = noise block, bit 6 + sl = vfree
        // every warp: "my operand writes are done" -> one arrival on the leader CTA's barrier
        // (address of ready[0] in the LEADER CTA's shared memory, computed once: ready[1], ready_x[0], ready_x[1] follow at +8 ...)
        uint32_t ready0_leader;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ready0_leader) : "r"((uint32_t)__cvta_generic_to_shared(&s.ready[0])), "r"(0));
        auto arrive = [&](unsigned long long* bar) {
            ptx::tcgen05_fence_before_thread_sync();
            __syncwarp();
            if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];"
                             :: "r"(ready0_leader + (uint32_t)((const char*)bar - (const char*)&s.ready[0])) : "memory");
        };
        // quarter -> layer-1 K blocks: slot 0: quarters 0,1 own 3 blocks, 2,3 own 2; slot 1 mirrored
        int kb0[2], fin[2];
        {
            const int q0 = cq, q1 = 3 - cq;
            kb0[0] = q0 < 2 ? Q_NB * q0 : 2 * Q_NB + Q_NS * (q0 - 2);
            kb0[1] = q1 < 2 ? Q_NB * q1 : 2 * Q_NB + Q_NS * (q1 - 2);
            fin[0] = cq == 2;               // the quarter that finishes the rows of the slot (velocity, integration, emit)
            fin[1] = cq == 1;
        }
        const bool three[2] = {cq < 2, cq >= 2};
        // this thread's scratch lines (float4 index; fixed for the whole kernel): pinned, or ptxas re-derives them from S2R ctaid /
        // tid (~30 instructions) at every c1 prefetch
        uint32_t sc_idx[2];
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) sc_idx[sl] = (blockIdx.x * 2 + sl) * P_SCRATCH_F4_PER_SLOT + kb0[sl] * 4 * P_ROWS + r;
        asm volatile("" : "+r"(sc_idx[0]), "+r"(sc_idx[1]));

        // agent of the first row of each slot's tile, carried from tile to tile (the tiles of a slot advance by 4 n_pairs tiles): one
        // 64-bit modulo per kernel; rows inside a tile wrap with a compare (the prologue spent ~3 K clk per tile pair in `%`)
        int abase[2];
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) abase[sl] = (int)(((long long)(2 * (2 * pair + sl) + (int)cta) * P_ROWS) % n_agents);
        const int astep = (int)(((long long)4 * n_pairs * P_ROWS) % n_agents);
        auto wrap = [&](int v) {            // v < n_agents + 128: one round unless the batch has fewer than 128 agents
            while (v >= n_agents) v -= n_agents;
            return v;
        };
        // cell state of this thread's units: c[sl][0..7] = units 8 cq .. (gates half 0), c[sl][8..15] = units 32 + 8 cq .. (half 1).
        // The rows of a tile pair are requested during the LAST step of the pair before it (after the final cell update the
        // registers are dead; the loads land under the remaining layer-1 / layer-2 phases instead of in the next prologue, whose
        // ~190 KB of row loads are bound by the L2 -> SM bandwidth); abase[] must already refer to the tiles being loaded.
        float c[2][16];
        auto load_c0 = [&](int ubx) {
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                const int tile = 2 * (ubx + sl) + (int)cta;
                const long long left = n_rows - (long long)tile * P_ROWS;
                const bool ok = tile < n_tiles && r < left;
                const float* src = c0 + (size_t)wrap(abase[sl] + r) * SW_H + cq * 8;
#pragma unroll
                for (int half = 0; half < 2; ++half) {                    // units 8 cq .. and 32 + 8 cq ..: 32 contiguous bytes each
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                    if (ok) ldg256(src + half * 32, a, b);
                    c[sl][8 * half] = a.x; c[sl][8 * half + 1] = a.y; c[sl][8 * half + 2] = a.z; c[sl][8 * half + 3] = a.w;
                    c[sl][8 * half + 4] = b.x; c[sl][8 * half + 5] = b.y; c[sl][8 * half + 6] = b.z; c[sl][8 * half + 7] = b.w;
                }
            }
        };
        load_c0(2 * pair);
        for (int ub = 2 * pair; ub < n_units; ub += 2 * n_pairs) {
#ifdef SW_PAIR_TRACE
            const bool tr_on = blockIdx.x == 0 && etid == 0 && ub == 2 * pair + 2 * (2 * n_pairs);
#endif
            SW_TR(0);
            bool has_tile[2], valid[2];
            long long row0[2];
            int rows_here[2], agent[2];     // rows of the tile inside the batch (0: no tile, < 128: last tile)
            float2 xl = make_float2(0.f, 0.f);
            float* out_row = out;           // where this thread emits (the rows of the slot it finishes; one slot at most)
            // ---------------- tile prologue, both slots: every global load coalesced and issued up front ----------------
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                const int tile = 2 * (ub + sl) + (int)cta;
                has_tile[sl] = tile < n_tiles;                            // a slot without a tile (end of the batch) runs on zero rows
                row0[sl] = (long long)tile * P_ROWS;
                rows_here[sl] = has_tile[sl] ? (int)(n_rows - row0[sl] < P_ROWS ? n_rows - row0[sl] : P_ROWS) : 0;
                valid[sl] = r < rows_here[sl];
                agent[sl] = valid[sl] ? wrap(abase[sl] + r) : 0;
                SW_TR(49 + 4 * sl);
                SW_TR(50 + 4 * sl);
                {   // S tile [128 rows][16 pieces] -> the slot's (still unused) h operand region, piece' = piece ^ (row & 7), by cp.async:
                    // global -> shared without a register in between.  (Through registers the 8 loaded float4 of both slots did not
                    // fit beside the cell state: ptxas spilled each one right behind its load, and every spill store waited for
                    // its load -- the loads of the prologue ran one after the other, ~17 K clk per tile pair.)
                    float4* sS = reinterpret_cast<float4*>(s.h[sl][0]);
                    // piece g = etid + 512 i: row (etid >> 4) + 32 i, the same 16-byte piece and swizzle term for every i; rows outside the
                    // batch (and every row without a pooled term) are zero-filled by the copy itself (src-size 0): no branch
                    const int row_a = etid >> 4, piece = etid & 15;
                    const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(sS + row_a * 16 + (piece ^ (row_a & 7)));
                    int a = wrap(abase[sl] + row_a);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const bool ok = pooled && row_a + 32 * i < rows_here[sl];
                        const float* src = ok ? pooled + (size_t)a * SW_H + piece * 4 : c0;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst0 + (uint32_t)(i * 32 * 16 * 16)), "l"(src), "r"(ok ? 16u : 0u) : "memory");
                        a = wrap(a + 32);
                    }
                }
                SW_TR(51 + 4 * sl);
                if (fin[sl] && valid[sl]) {
                    xl = __ldg(reinterpret_cast<const float2*>(x_last + (size_t)agent[sl] * 4));
                    out_row = out + (size_t)(row0[sl] + r) * n_next * 4;
                }
            }
            // h0 items: (row, 8-column chunk = (lane >> 3) + 4 i), 8 rows x 128 B per instruction; requested here, consumed at the end
            // of each slot's staging below (the S tiles no longer pass through registers, so both slots' items fit)
            SW_TR(1);
            float4 hreg[2][2][2];
            const int hrow = ewarp * 8 + (lane & 7);
#pragma unroll
            for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    hreg[sl][i][0] = hreg[sl][i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (hrow < rows_here[sl]) {
                        ldg256(h0 + (size_t)wrap(abase[sl] + hrow) * SW_H + ((lane >> 3) + 4 * i) * 8, hreg[sl][i][0], hreg[sl][i][1]);
                    }
                }
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                float4* sS = reinterpret_cast<float4*>(s.h[sl][0]);       // [128 rows][16 pieces], piece' = piece ^ (row & 7); 32 KB = h hi|lo
                const float4* sZ = reinterpret_cast<const float4*>(s.zst[sl]);   // [128 rows][8 pieces], TMA swizzle: piece' = piece ^ (row & 7)
                SW_TR(2 + 8 * sl);
                asm volatile("cp.async.wait_all;" ::: "memory");          // this thread's pieces of the S tile(s) have landed
                SW_TR(3 + 8 * sl);
                if (has_tile[sl]) { mbar_wait(&s.bar_z[sl], (ph >> (4 + sl)) & 1u); ph ^= 16u << sl; }
                SW_TR(4 + 8 * sl);
                epi_sync();
                SW_TR(5 + 8 * sl);
                // [S ; z] (K = 96 = 24 pieces): this thread owns pieces 6 cq .. 6 cq + 5 of its row -> hi|lo TMEM A operand
                uint32_t hi[12], lo[12];
#pragma unroll
                for (int e = 0; e < 6; ++e) {
                    const int piece = cq * 6 + e;
                    float4 v;
                    if (piece < 16) v = sS[r * 16 + (piece ^ (r & 7))];
                    else            v = has_tile[sl] ? sZ[r * 8 + ((piece - 16) ^ (r & 7))] : make_float4(0.f, 0.f, 0.f, 0.f);
                    psplit2(v.x, v.y, hi[2 * e], lo[2 * e]);
                    psplit2(v.z, v.w, hi[2 * e + 1], lo[2 * e + 1]);
                }
                const uint32_t tls = tl + (uint32_t)(sl * 256);
                tmem_st<12>(tls + PC_AHI + cq * 12, hi);
#if !SW_PAIR_BF16
                tmem_st<12>(tls + PC_ALO + cq * 12, lo);
#endif
                ptx::tcgen05_wait_st();
                SW_TR(6 + 8 * sl);
                epi_sync();                                               // staging consumed: h region and noise buffer are free
                SW_TR(7 + 8 * sl);
                {   // next tile's noise block: 12 steps ahead of its use
                    const int un = ub + sl + 2 * n_pairs, tn = 2 * un + (int)cta;
                    if (etid == 0 && un < n_units && tn < n_tiles) prefetch_noise(sl, tn);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {                             // h0 -> hi|lo operand chunks [chunk][row][8]
                    uint32_t hh[4], ll[4];
                    psplit2(hreg[sl][i][0].x, hreg[sl][i][0].y, hh[0], ll[0]);
                    psplit2(hreg[sl][i][0].z, hreg[sl][i][0].w, hh[1], ll[1]);
                    psplit2(hreg[sl][i][1].x, hreg[sl][i][1].y, hh[2], ll[2]);
                    psplit2(hreg[sl][i][1].z, hreg[sl][i][1].w, hh[3], ll[3]);
                    const size_t off = ((size_t)((lane >> 3) + 4 * i) * P_ROWS + hrow) * 8;
                    *reinterpret_cast<uint4*>(s.h[sl][0] + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
#if !SW_PAIR_BF16
                    *reinterpret_cast<uint4*>(s.h[sl][1] + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
#endif
                }
                ptx::fence_proxy_async(ptx::space_shared);
                arrive(&s.ready[sl]);                                     // -> hoist MMAs of the slot
                SW_TR(8 + 8 * sl);
            }
            SW_TR(18);
            bool out_of_range = false;
            float p0 = xl.x, p1 = xl.y;                                   // state of the rows this thread finishes (one slot at most)
            float4* sc[2];
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                sc[sl] = scratch + sc_idx[sl];
                const uint32_t ta = tl + (uint32_t)(sl * 256) + PC_R1 + kb0[sl] * 16;
                wait_full3(&s.full[sl][0], (ph >> sl) & 1u); ph ^= 1u << sl;      // c1 + b1 -> scratch (the K blocks this thread re-reads)
                SW_TR(19 + 2 * sl);
                c1_to_scratch(ta, sc[sl], s.f32 + PF_B1 + kb0[sl] * 16, three[sl]);
                arrive(&s.ready[sl]);                                     // -> layer 1 of step 0
                SW_TR(20 + 2 * sl);
            }
            {   // the NEXT tiles' rows of pooled / h0 / c0 -> L2, twelve steps ahead of their use (no registers held: prefetch only);
                // issued here, under the layer-1 MMAs of step 0 (before the c1 pass it delayed the first step by ~1.2 K clk)
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    const int un = ub + sl + 2 * n_pairs, tn = 2 * un + (int)cta;
                    abase[sl] += astep;                                   // -> the slot's next tile
                    if (abase[sl] >= n_agents) abase[sl] -= n_agents;
                    if (un < n_units && tn < n_tiles && (long long)tn * P_ROWS + (etid >> 2) < n_rows) {
                        // thread -> (row = etid / 4, 64-byte quarter of the row's 256 B)
                        const size_t off = (size_t)wrap(abase[sl] + (etid >> 2)) * SW_H + (etid & 3) * 16;
                        if (pooled) asm volatile("prefetch.global.L2 [%0];" :: "l"(pooled + off));
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(h0 + off));
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(c0 + off));
                    }
                }
            }

            // ---------------- layer 1 epilogue: a1 = lrelu(acc + c1) -> hi|lo in place ----------------
            float4 cpre[4];             // c1 of the first layer-1 block of the NEXT layer-1 phase (one pending at a time)
            auto c1_prefetch = [&](auto slc) {
                constexpr int sl = decltype(slc)::value;
#pragma unroll
                for (int q = 0; q < 4; ++q) cpre[q] = __ldcg(sc[sl] + q * P_ROWS);
            };
            auto phase_l1 = [&](auto slc) {
                constexpr int sl = decltype(slc)::value;
                const uint32_t ta = tl + (uint32_t)(sl * 256) + PC_R1 + kb0[sl] * 16;
                l1_epilogue(ta, sc[sl], &s.full[sl][0], (ph >> sl) & 1u, cpre, three[sl]);
                ph ^= 1u << sl;
                arrive(&s.ready[sl]);                                     // -> layer 2 (+ h part of gates half 0)
            };
            // ---------------- layer-2 epilogue + folded layers 3+4 (80 -> 2), row finish ----------------
            auto phase_l2 = [&](auto slc, int t, bool feed_back) {
                constexpr int sl = decltype(slc)::value;
                const float4* b2 = reinterpret_cast<const float4*>(s.f32 + PF_B2 + cq * 20);
                const float4* w34 = reinterpret_cast<const float4*>(s.f32 + PF_W34 + cq * 40);
                wait_full3(&s.full[sl][0], (ph >> sl) & 1u); ph ^= 1u << sl;
                float v0 = 0.0f, v1 = 0.0f;
                {
                    uint32_t acc[20];
                    tmem_ld<20>(tl + (uint32_t)(sl * 256) + PC_R2 + cq * 20, acc);
                    ptx::tcgen05_wait_ld();
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        const float4 b = b2[j], wa = w34[2 * j], wb = w34[2 * j + 1];
                        float y0, y1, y2, y3;
                        lrelu_add2(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]), b.x, b.y, y0, y1);
                        lrelu_add2(__uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]), b.z, b.w, y2, y3);
                        v0 = fmaf(y0, wa.x, v0); v1 = fmaf(y0, wa.y, v1);
                        v0 = fmaf(y1, wa.z, v0); v1 = fmaf(y1, wa.w, v1);
                        v0 = fmaf(y2, wb.x, v0); v1 = fmaf(y2, wb.y, v1);
                        v0 = fmaf(y3, wb.z, v0); v1 = fmaf(y3, wb.w, v1);
                    }
                }
                ptx::tcgen05_fence_before_thread_sync();
                ph ^= 64u << sl;            // parity of vfree[sl]: one phase per step
                if (!fin[sl]) {
                    mbar_wait(&s.vfree[sl], (ph >> (6 + sl)) & 1u);     // phase of the PREVIOUS step (a fresh barrier passes parity 1: first step)
                    s.vpart[sl][(cq * 2) * P_ROWS + r] = v0; s.vpart[sl][(cq * 2 + 1) * P_ROWS + r] = v1;
                    vel_arrive(sl);     // (bar.arrive orders the shared-memory writes above before the finishing warps' reads)
                } else {            // velocity, integration, emit; (p, v) -> hi|lo x block of the gate MMA
                    vel_sync(sl);
                    const float* vp = s.vpart[sl];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (q != (sl == 0 ? 2 : 1)) { v0 += vp[(q * 2) * P_ROWS + r]; v1 += vp[(q * 2 + 1) * P_ROWS + r]; }
                    __syncwarp();
                    if (lane == 0) ptx::mbarrier_arrive(reinterpret_cast<uint64_t*>(&s.vfree[sl]));
                    v0 += s.f32[PF_B34]; v1 += s.f32[PF_B34 + 1];
                    p0 += v0; p1 += v1;
                    out_of_range |= !(fmaxf(fmaxf(fabsf(p0), fabsf(p1)), fmaxf(fabsf(v0), fabsf(v1))) <= 6.0e4f);   // fp16 range guard
                    if (feed_back) {
                        uint32_t hp, lp, hv, lv;
                        psplit2(p0, p1, hp, lp);
                        psplit2(v0, v1, hv, lv);
                        *reinterpret_cast<uint4*>(s.xk[sl] + (size_t)r * 8) = make_uint4(hp, hv, lp, lv);                      // k 0..7
                        *reinterpret_cast<uint4*>(s.xk[sl] + (size_t)(P_ROWS + r) * 8) = make_uint4(hp, hv, P_ONE2, 0u);  // k 8..15
                        ptx::fence_proxy_async(ptx::space_shared);
                        arrive(&s.ready_x[sl]);                           // -> x blocks of the gates, h part of half 1
                    }
                    if (valid[sl])
                        *reinterpret_cast<float4*>(out_row + (size_t)t * 4) = make_float4(p0, p1, v0, v1);
                }
            };
            // ---------------- LSTM cell: 8 units of gates half 0 (units 8 cq ..), then 8 units of half 1 (units 32 + 8 cq ..) ----------------
            auto phase_cell = [&](auto slc) {
                constexpr int sl = decltype(slc)::value;
                const uint32_t tls = tl + (uint32_t)(sl * 256);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (half == 0) wait_full3(&s.full[sl][1], (ph >> (2 + sl)) & 1u);
                    uint32_t a[32];
                    tmem_ld<32>(tls + half * 128 + cq * 32, a);
                    ptx::tcgen05_wait_ld();
                    float hv[8];
#pragma unroll
                    for (int uu = 0; uu < 8; uu += 2) {
                        float g[2][4];
#pragma unroll
                        for (int w2 = 0; w2 < 2; ++w2)
#pragma unroll
                            for (int q = 0; q < 4; ++q) g[w2][q] = __uint_as_float(a[(uu + w2) * 4 + q]);
                        lstm_cell_pair_prescaled_x2(g[0], g[1], c[sl][half * 8 + uu], c[sl][half * 8 + uu + 1], hv[uu], hv[uu + 1]);
                    }
                    uint32_t hh[4], ll[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) psplit2(hv[2 * e], hv[2 * e + 1], hh[e], ll[e]);
                    // h is an operand of the half-1 gate MMAs: nothing may overwrite it before they have completed
                    if (half == 0) wait_full3(&s.full[sl][2], (ph >> (2 + sl)) & 1u);
                    const size_t off = ((size_t)(half * 4 + cq) * P_ROWS + r) * 8;
                    *reinterpret_cast<uint4*>(s.h[sl][0] + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
#if !SW_PAIR_BF16
                    *reinterpret_cast<uint4*>(s.h[sl][1] + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
#endif
                }
                ph ^= 4u << sl;
                ptx::fence_proxy_async(ptx::space_shared);
                arrive(&s.ready[sl]);                                     // -> next step's layer 1
            };
            using S0 = std::integral_constant<int, 0>;
            using S1 = std::integral_constant<int, 1>;
            // The slots run half a step apart (slot 1 behind): the long MMA group of one slot (layer 2 + the h part of gates half
            // 0, ~2.5 K clk) runs under the long epilogue phase of the other (the cell update, ~5 K clk); see the issuing warp.
            c1_prefetch(S0{});
            for (int t = 0; t < n_next; ++t) {
                const bool feed_back = t + 1 < n_next;
                SW_TR(24 + (t < 2 ? 0 : t == n_next - 1 ? 16 : 8));
                phase_l1(S0{});
                SW_TR(25 + (t < 2 ? 0 : t == n_next - 1 ? 16 : 8));
                if (t > 0) phase_cell(S1{});
                c1_prefetch(S1{});
                if (!feed_back) load_c0(ub + 2 * n_pairs);               // (the cell updates of this tile pair are done)
                SW_TR(26 + (t < 2 ? 0 : t == n_next - 1 ? 16 : 8));
                phase_l2(S0{}, t, feed_back);
                SW_TR(27 + (t < 2 ? 0 : t == n_next - 1 ? 16 : 8));
                phase_l1(S1{});
                SW_TR(28 + (t < 2 ? 0 : t == n_next - 1 ? 16 : 8));
                if (feed_back) phase_cell(S0{});
                if (feed_back) c1_prefetch(S0{});
                SW_TR(29 + (t < 2 ? 0 : t == n_next - 1 ? 16 : 8));
                phase_l2(S1{}, t, feed_back);
                SW_TR(30 + (t < 2 ? 0 : t == n_next - 1 ? 16 : 8));
            }
            SW_TR(48);
            if (out_of_range && status) {
                if ((fin[0] && valid[0]) || (fin[1] && valid[1])) atomicOr(status, 1);
            }
        }
    }
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

}  // namespace sw

static int pair_grid(long long tiles, int sm_count) {
    const long long units = (tiles + 1) / 2;            // two tiles (one per CTA of a pair) per unit, two unit slots per pair
    long long pairs = (units + 1) / 2;
    if (pairs > sm_count / 2) pairs = sm_count / 2;
    if (pairs < 1) pairs = 1;
    return (int)(2 * pairs);
}

#if !SW_PAIR_BF16
#ifdef SW_PAIR_TRACE
extern "C" int sw_pair_trace_read(long long* out64) {
    SW_CUDA_TRY(cudaMemcpyFromSymbol(out64, sw::g_pair_trace, sizeof(long long) * 64));
    return SW_OK;
}
#endif

extern "C" long long sw_decode_pair_scratch_bytes(int sm_count) {
    if (sm_count < 2) return 0;
    return (long long)(sm_count / 2) * 2 * 2 * sw::P_SCRATCH_F4_PER_SLOT * 16;
}

extern "C" int sw_decode_pair_pack_sizes(int* n_w16, int* n_f32) {
    if (!n_w16 || !n_f32) return SW_ERR_ARG;
    *n_w16 = 2 * sw::PW_TOTAL;
    *n_f32 = sw::PF_TOTAL;
    return SW_OK;
}

#else
extern "C" long long sw_decode_pair_scratch_bytes(int sm_count);
#endif

extern "C" int PAIR_ENTRY(const void* pair_w16, const float* pair_f32, const float* h0, const float* c0,
                                  const float* pooled, const float* noise, const float* x_last, float* out, void* scratch,
                                  long long scratch_bytes, int* status, int n_agents, int n_samples, int n_next, int sm_count,
                                  void* stream) {
    if (!pair_w16 || !pair_f32 || !h0 || !c0 || !noise || !x_last || !out || !scratch) return SW_ERR_ARG;
    if (n_agents <= 0 || n_samples <= 0 || n_next <= 0 || sm_count < 2) return SW_ERR_ARG;
    if (scratch_bytes < sw_decode_pair_scratch_bytes(sm_count) || ((uintptr_t)scratch & 15u) != 0) return SW_ERR_ARG;
    const long long n_rows = (long long)n_agents * n_samples;
    const long long tiles = (n_rows + sw::P_ROWS - 1) / sw::P_ROWS;
    if (tiles > 0x3fffffffLL) return SW_ERR_UNSUPPORTED;
    if (((uintptr_t)noise & 15u) != 0) return SW_ERR_ARG;
    if ((((uintptr_t)h0 | (uintptr_t)c0) & 31u) != 0 || (pooled && ((uintptr_t)pooled & 15u) != 0)) return SW_ERR_ARG;    // 32-byte row pieces (LDG.256), 16-byte cp.async
    CUtensorMap noise_map;
    const int rc = encode_noise_map2(&noise_map, noise, n_rows);
    if (rc != SW_OK) return rc;
    const int smem = (int)sizeof(sw::PairSmem);
    SW_SET_MAX_SMEM(sw::PAIR_KERNEL, smem);
    const int grid = pair_grid(tiles, sm_count);
    sw::PAIR_KERNEL<<<grid, sw::Q_THREADS, smem, (cudaStream_t)stream>>>(
        noise_map, (const __half*)pair_w16, pair_f32, h0, c0, pooled, x_last, out, (float4*)scratch, status, n_agents, n_rows, n_next,
        (int)tiles);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
