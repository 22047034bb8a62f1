// fp32-faithful tensor-core decode kernel, TWO 128-row tiles in flight per SM, warps SPECIALISED BY PIPE.
//
// Same computation, arithmetic, weight image, TMEM column map and c1 scratch as decode_fwd_pair.cu (CTA pairs issuing
// tcgen05 cta_group::2 MMAs, two tile slots of 256 TMEM columns per CTA).  What changes is the division of labour.  In
// decode_fwd_pair.cu all 16 epilogue warps walk through the phases together, so at any moment the SM runs ONE kind of code: the
// LSTM cell update (bound by the MUFU pipe: 104 ex2 / rcp per thread-step, issue slots 3/4 idle) or the layer-1 / layer-2
// epilogues (bound by instruction issue and latency, MUFU pipe idle) -- ncu: cell update 48 % of the step at 75 % of the MUFU
// roof, the epilogues the rest at IPC ~0.45.  Here the two kinds of work run AT THE SAME TIME on different warps:
//   * warps 0-7  ("X", thread = (row, column half)): the tile prologue, the layer-1 epilogue (a1 = lrelu(acc + c1) -> hi|lo in
//     place), the layer-2 epilogue with the folded output layers, and the row finish (velocity, integration, emit, x block) --
//     for BOTH slots, in the fixed order L1(0) L2(0) L1(1) L2(1);
//   * warps 8-15 ("Y", thread = (row, unit half)): the LSTM cell update of BOTH slots (cell state of both tiles in registers),
//     in the fixed order cell(0) cell(1);
//   * warps 16 / 17 of the leader CTA: the issuing warps, one per slot (a slot's MMA chain is strictly serial:
//     L2 + gates-h(half 0) -> x blocks + gates(half 1) -> next L1), one elected lane issuing back to back.
// A slot's chain (L1 epilogue -> MMA -> L2 epilogue -> MMA -> cell -> MMA -> ...) is serial, the two slots fill each other's
// gaps on the X warps, the Y warps and the tensor pipe.  Registers follow the work (setmaxnreg): X 96, Y 128, issuing group 32.
#include "decode_pair.cuh"

namespace sw {

constexpr int R_GROUP = 256;          // threads of the X group / of the Y group
constexpr int R_THREADS = 640;        // X (warps 0-7) + Y (8-15) + the issuing warps' warpgroup (16-19)

struct Pair2Smem {
    float zst[2][P_ROWS * SW_Z];             // noise block of each slot's tile (TMA, 128-byte swizzle; 1024-byte aligned)
    __half w[PW_TOTAL];                      // this rank's half of every weight matrix (113 664 B)
    __half h[2][2][8 * P_ROWS * 8];          // [slot][hi|lo][8 chunks][128][8]
    __half xk[2][2 * P_ROWS * 8];            // [slot] x-feedback A operand, one K block
    float f32[PF_TOTAL];
    float vpart[2][2 * P_ROWS];              // [slot][component][row]: partial velocity of column half 1
    unsigned long long bar_x[2];             // X group: operands written (16 arrivals: 8 warps x 2 CTAs; used in the leader CTA)
    unsigned long long bar_f[2];             // row finish: x block written (8 arrivals: 4 warps x 2 CTAs)
    unsigned long long bar_y[2];             // Y group: h written (16 arrivals)
    unsigned long long full[2][3];           // MMA completion: hoist / L1 / L2 | gates half 0 | gates half 1
    unsigned long long bar_z[2];             // TMA: the slot's noise block
    unsigned long long bar_w;                // TMA: weights
    uint32_t tmem_base;
};
static_assert(sizeof(Pair2Smem) <= 227 * 1024, "shared memory of one CTA");

__device__ __forceinline__ void x_sync() { asm volatile("bar.sync 1, %0;" :: "n"(R_GROUP) : "memory"); }
// partial-velocity exchange of a slot (named barrier 2 + slot): column half 1 only arrives, column half 0 (row finish) waits
__device__ __forceinline__ void vel_arrive2(int sl) { asm volatile("bar.arrive %0, %1;" :: "r"(2 + sl), "n"(R_GROUP) : "memory"); }
__device__ __forceinline__ void vel_sync2(int sl) { asm volatile("bar.sync %0, %1;" :: "r"(2 + sl), "n"(R_GROUP) : "memory"); }
__device__ __forceinline__ void wait_full2(unsigned long long* bar, uint32_t parity) {
    mbar_wait(bar, parity);
    ptx::tcgen05_fence_after_thread_sync();
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(R_THREADS, 1)
decode_fwd_pair2_kernel(const __grid_constant__ CUtensorMap noise_map /* [n_rows][32] fp32, box 128 x 32, 128-byte swizzle */,
                        const __half* __restrict__ w16 /* [2 ranks][PW_TOTAL] */, const float* __restrict__ wf32,
                        const float* __restrict__ h0, const float* __restrict__ c0, const float* __restrict__ pooled,
                        const float* __restrict__ x_last, float* __restrict__ out, float4* __restrict__ scratch,
                        int* __restrict__ status, int n_agents, long long n_rows, int n_next, int n_tiles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Pair2Smem& s = *reinterpret_cast<Pair2Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int lq = warp & 3;                        // TMEM lane quarter (= warp % 4)
    int lane = tid & 31, hf = (warp >> 2) & 1;      // column half (X) / unit half (Y)
    int r = lq * 32 + lane;
    // opaque: kept in registers (ptxas otherwise re-derives them from S2R tid, a ~25 clk special-register read, at every use)
    asm volatile("" : "+r"(lane), "+r"(hf), "+r"(r));
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
            ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.bar_x[sl]), 16);
            ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.bar_f[sl]), 8);
            ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&s.bar_y[sl]), 16);
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
        // =========================== the issuing warps (leader CTA: warp 16 -> slot 0, warp 17 -> slot 1) ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (cta == 0 && warp < 18) {
            const int sl = warp - 16;
            const uint32_t ts = tmem + (uint32_t)(sl * 256);
            uint32_t ph_x = 0, ph_f = 0, ph_y = 0;
            const __half* const h_hi = s.h[sl][0];
            const __half* const h_lo = s.h[sl][1];
            const __half* const xk = s.xk[sl];
            auto wait_bar = [&](unsigned long long* bar, uint32_t& ph) {
                mbar_wait(bar, ph); ph ^= 1;
                ptx::tcgen05_fence_after_thread_sync();
            };
            for (int u = sl + 2 * pair; u < n_units; u += 2 * n_pairs) {
                wait_bar(&s.bar_x[sl], ph_x);           // c1 = [S ; z] . W1[S,z rows]^T -> [0,160)
                if (elect_one()) {
                    pmma3_ts1<80, 6, 8>(ts + PC_R1, ts + PC_AHI, ts + PC_ALO, s.w + PW_WSZ_HI, s.w + PW_WSZ_LO);
                    umma1_commit_pair(&s.full[sl][0]);
                }
                __syncwarp();
                wait_bar(&s.bar_x[sl], ph_x);           // layer 1 of step 0
                if (elect_one()) {
                    pmma3_ss1<80, 4>(ts + PC_R1, h_hi, h_lo, s.w + PW_W1H_HI, s.w + PW_W1H_LO);
                    umma1_commit_pair(&s.full[sl][0]);
                }
                __syncwarp();
#pragma unroll 1
                for (int t = 0; t < n_next; ++t) {
                    const bool feed_back = t + 1 < n_next;
                    wait_bar(&s.bar_x[sl], ph_x);       // layer 2: a1 (K = 160, TMEM) -> [160,240); the h part of gates half 0 behind it
                    if (elect_one()) {
                        pmma3_ts1<P_L2NL, 10, 16>(ts + PC_R2, ts + PC_R1, ts + PC_R1 + 8, s.w + PW_W2_HI, s.w + PW_W2_LO);
                        umma1_commit_pair(&s.full[sl][0]);
                        if (feed_back) pmma3_ss1<64, 4>(ts, h_hi, h_lo, s.w + PW_WHH, s.w + PW_WHH + 4096);
                    }
                    __syncwarp();
                    if (!feed_back) break;
                    wait_bar(&s.bar_f[sl], ph_f);       // gates: x block of half 0 -> full[1]; half 1 (h part + x block) -> full[2]
                    if (elect_one()) {
                        pmma1_ss<64, 1>(ts, xk, s.w + PW_WXK, PFMT, true);
                        umma1_commit_pair(&s.full[sl][1]);
                        pmma3_ss1<64, 4>(ts + 128, h_hi, h_lo, s.w + PW_WHH + 8192, s.w + PW_WHH + 8192 + 4096);
                        pmma1_ss<64, 1>(ts + 128, xk, s.w + PW_WXK + 1024, PFMT, true);
                        umma1_commit_pair(&s.full[sl][2]);
                    }
                    __syncwarp();
                    wait_bar(&s.bar_y[sl], ph_y);       // next step's layer 1
                    if (elect_one()) {
                        pmma3_ss1<80, 4>(ts + PC_R1, h_hi, h_lo, s.w + PW_W1H_HI, s.w + PW_W1H_LO);
                        umma1_commit_pair(&s.full[sl][0]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 8) {
        // =========================== Y: the LSTM cell update of both slots ===========================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
        const uint32_t tl = tmem + ((uint32_t)(lq * 32) << 16);
        uint32_t bar_y_leader;              // bar_y[0] in the LEADER CTA's shared memory (bar_y[1] follows 8 bytes later)
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bar_y_leader) : "r"((uint32_t)__cvta_generic_to_shared(&s.bar_y[0])), "r"(0));
        uint32_t ph_g = 0;                  // bit sl: parity of the slot's gate barriers
        for (int ub = 2 * pair; ub < n_units; ub += 2 * n_pairs) {
            const int n_act = ub + 1 < n_units ? 2 : 1;
            // cell state of this thread's units: c[sl][0..15] = units 16 hf .. (gates half 0), c[sl][16..31] = units 32 + 16 hf .. (half 1)
            float c[2][32];
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                const int tile = 2 * (ub + sl) + (int)cta;
                const long long row0 = (long long)tile * P_ROWS;
                const bool valid = sl < n_act && tile < n_tiles && row0 + r < n_rows;
                const int agent = valid ? (int)((row0 + r) % n_agents) : 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 v = valid ? __ldg(reinterpret_cast<const float4*>(c0 + (size_t)agent * SW_H + (q >> 2) * 32 + hf * 16) + (q & 3))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
                    c[sl][4 * q] = v.x; c[sl][4 * q + 1] = v.y; c[sl][4 * q + 2] = v.z; c[sl][4 * q + 3] = v.w;
                }
            }
            for (int t = 0; t + 1 < n_next; ++t) {
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    if (sl >= n_act) continue;
                    const uint32_t tls = tl + (uint32_t)(sl * 256);
                    __half* const h_hi = s.h[sl][0];
                    __half* const h_lo = s.h[sl][1];
                    const uint32_t parity = (ph_g >> sl) & 1u;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        if (half == 0) wait_full2(&s.full[sl][1], parity);
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
                            uint32_t a[32];
                            tmem_ld<32>(tls + half * 128 + hf * 64 + ch * 32, a);
                            ptx::tcgen05_wait_ld();
                            float hv[8];
#pragma unroll
                            for (int uu = 0; uu < 8; uu += 2) {
                                float g[2][4];
#pragma unroll
                                for (int w2 = 0; w2 < 2; ++w2)
#pragma unroll
                                    for (int q = 0; q < 4; ++q) g[w2][q] = __uint_as_float(a[(uu + w2) * 4 + q]);
                                lstm_cell_pair_prescaled(g[0], g[1], c[sl][half * 16 + ch * 8 + uu], c[sl][half * 16 + ch * 8 + uu + 1], hv[uu], hv[uu + 1]);
                            }
#pragma unroll
                            for (int e = 0; e < 4; ++e) psplit2(hv[2 * e], hv[2 * e + 1], hi[ch * 4 + e], lo[ch * 4 + e]);
                        }
                        // h is an operand of the half-1 gate MMAs: nothing may overwrite it before they have completed
                        if (half == 0) wait_full2(&s.full[sl][2], parity);
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
                            const size_t off = ((size_t)(half * 4 + hf * 2 + ch) * P_ROWS + r) * 8;
                            *reinterpret_cast<uint4*>(h_hi + off) = make_uint4(hi[ch * 4], hi[ch * 4 + 1], hi[ch * 4 + 2], hi[ch * 4 + 3]);
                            *reinterpret_cast<uint4*>(h_lo + off) = make_uint4(lo[ch * 4], lo[ch * 4 + 1], lo[ch * 4 + 2], lo[ch * 4 + 3]);
                        }
                    }
                    ph_g ^= 1u << sl;
                    ptx::fence_proxy_async(ptx::space_shared);
                    ptx::tcgen05_fence_before_thread_sync();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(bar_y_leader + 8u * sl) : "memory");
                }
            }
        }
    } else {
        // =========================== X: tile prologue, layer-1 / layer-2 epilogues, row finish of both slots ===========================
        // (stays at the launch allocation of 96 registers: the issuing group's 64 x 128 go to the Y warps)
        const uint32_t tl = tmem + ((uint32_t)(lq * 32) << 16);
        uint32_t bar_x_leader;              // bar_x[0] in the LEADER CTA's shared memory (bar_x[1] at +8, bar_f[sl] at +16 + 8 sl)
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bar_x_leader) : "r"((uint32_t)__cvta_generic_to_shared(&s.bar_x[0])), "r"(0));
        auto arrive_x = [&](int sl) {
            ptx::tcgen05_fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(bar_x_leader + 8u * sl) : "memory");
        };
        uint32_t ph_l[2] = {0, 0}, ph_z[2] = {0, 0};
        for (int ub = 2 * pair; ub < n_units; ub += 2 * n_pairs) {
            const int n_act = ub + 1 < n_units ? 2 : 1;
            bool has_tile[2], valid[2];
            long long row0[2];
            float p0[2], p1[2];
            bool out_of_range = false;
            float4* sc[2];
            // ---------------- tile prologue, slot by slot: every global load coalesced and issued up front ----------------
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                const int tile = 2 * (ub + sl) + (int)cta;
                has_tile[sl] = sl < n_act && tile < n_tiles;              // the odd last unit has one tile only
                row0[sl] = (long long)tile * P_ROWS;
                valid[sl] = has_tile[sl] && row0[sl] + r < n_rows;
                sc[sl] = scratch + ((size_t)blockIdx.x * 2 + sl) * P_SCRATCH_F4_PER_SLOT + (size_t)(hf * 5) * 4 * P_ROWS + r;
                p0[sl] = p1[sl] = 0.0f;
                if (sl >= n_act) continue;
                const int abase = has_tile[sl] ? (int)(row0[sl] % n_agents) : 0;
                const int agent = valid[sl] ? (abase + r) % n_agents : 0;
                float4 sreg[8], hreg[4][2];
#pragma unroll
                for (int i = 0; i < 8; ++i) {                             // S tile [128][16 pieces]: piece g = tid + 256 i
                    const int g = tid + i * R_GROUP, row = g >> 4, piece = g & 15;
                    sreg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (pooled && has_tile[sl] && row0[sl] + row < n_rows)
                        sreg[i] = __ldg(reinterpret_cast<const float4*>(pooled + (size_t)((abase + row) % n_agents) * SW_H) + piece);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {                         // h0 items: (row, 8-column chunk), 8 rows x 128 B per instruction
                        const int hrow = warp * 8 + (lane & 7) + 64 * j, chunk = (lane >> 3) + 4 * i;
                        hreg[j * 2 + i][0] = hreg[j * 2 + i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (has_tile[sl] && row0[sl] + hrow < n_rows) {
                            const float4* src = reinterpret_cast<const float4*>(h0 + (size_t)((abase + hrow) % n_agents) * SW_H) + chunk * 2;
                            hreg[j * 2 + i][0] = __ldg(src);
                            hreg[j * 2 + i][1] = __ldg(src + 1);
                        }
                    }
                if (hf == 0 && valid[sl]) {
                    const float2 xl = __ldg(reinterpret_cast<const float2*>(x_last + (size_t)agent * 4));
                    p0[sl] = xl.x; p1[sl] = xl.y;
                }
                float4* sS = reinterpret_cast<float4*>(s.h[sl][0]);       // [128 rows][16 pieces], piece' = piece ^ (row & 7); 32 KB = h hi|lo
                const float4* sZ = reinterpret_cast<const float4*>(s.zst[sl]);   // [128 rows][8 pieces], TMA swizzle: piece' = piece ^ (row & 7)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int g = tid + i * R_GROUP, row = g >> 4, piece = g & 15;
                    sS[row * 16 + (piece ^ (row & 7))] = sreg[i];
                }
                if (has_tile[sl]) { mbar_wait(&s.bar_z[sl], ph_z[sl]); ph_z[sl] ^= 1; }
                x_sync();
                {   // [S ; z] (K = 96 = 24 pieces): this thread owns pieces 12 hf .. 12 hf + 11 of its row = K blocks 3 hf .. 3 hf + 2
                    uint32_t hi[24], lo[24];
#pragma unroll
                    for (int e = 0; e < 12; ++e) {
                        const int piece = hf * 12 + e;
                        float4 v;
                        if (piece < 16) v = sS[r * 16 + (piece ^ (r & 7))];
                        else            v = has_tile[sl] ? sZ[r * 8 + ((piece - 16) ^ (r & 7))] : make_float4(0.f, 0.f, 0.f, 0.f);
                        psplit2(v.x, v.y, hi[2 * e], lo[2 * e]);
                        psplit2(v.z, v.w, hi[2 * e + 1], lo[2 * e + 1]);
                    }
                    const uint32_t tls = tl + (uint32_t)(sl * 256);
                    tmem_st<24>(tls + PC_AHI + hf * 24, hi);
                    tmem_st<24>(tls + PC_ALO + hf * 24, lo);
                    ptx::tcgen05_wait_st();
                }
                x_sync();                                                 // staging consumed: h region and noise buffer are free
                {   // next tile's noise block: 12 steps ahead of its use
                    const int un = ub + sl + 2 * n_pairs, tn = 2 * un + (int)cta;
                    if (tid == 0 && un < n_units && tn < n_tiles) prefetch_noise(sl, tn);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {                             // h0 -> hi|lo operand chunks [chunk][row][8]
                    const int hrow = warp * 8 + (lane & 7) + 64 * (q >> 1), chunk = (lane >> 3) + 4 * (q & 1);
                    uint32_t hh[4], ll[4];
                    psplit2(hreg[q][0].x, hreg[q][0].y, hh[0], ll[0]);
                    psplit2(hreg[q][0].z, hreg[q][0].w, hh[1], ll[1]);
                    psplit2(hreg[q][1].x, hreg[q][1].y, hh[2], ll[2]);
                    psplit2(hreg[q][1].z, hreg[q][1].w, hh[3], ll[3]);
                    const size_t off = ((size_t)chunk * P_ROWS + hrow) * 8;
                    *reinterpret_cast<uint4*>(s.h[sl][0] + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                    *reinterpret_cast<uint4*>(s.h[sl][1] + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                }
                ptx::fence_proxy_async(ptx::space_shared);
                arrive_x(sl);                                             // -> hoist MMAs of the slot
            }
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                if (sl >= n_act) continue;
                const uint32_t tls = tl + (uint32_t)(sl * 256);
                wait_full2(&s.full[sl][0], ph_l[sl]); ph_l[sl] ^= 1;
#pragma unroll
                for (int kb = 0; kb < 5; ++kb) {       // c1 + b1 -> scratch (this thread's 5 K blocks of 16 columns)
                    const int col0 = (hf * 5 + kb) * 16;
                    uint32_t v[16];
                    tmem_ld<16>(tls + PC_R1 + col0, v);
                    ptx::tcgen05_wait_ld();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 b = *reinterpret_cast<const float4*>(s.f32 + PF_B1 + col0 + 4 * q);
                        __stcg(sc[sl] + (kb * 4 + q) * P_ROWS, make_float4(__uint_as_float(v[4 * q]) + b.x, __uint_as_float(v[4 * q + 1]) + b.y,
                                                                           __uint_as_float(v[4 * q + 2]) + b.z, __uint_as_float(v[4 * q + 3]) + b.w));
                    }
                }
                arrive_x(sl);                                             // -> layer 1 of step 0
            }

            for (int t = 0; t < n_next; ++t) {
                const bool feed_back = t + 1 < n_next;
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    if (sl >= n_act) continue;
                    const uint32_t tls = tl + (uint32_t)(sl * 256);
                    // ---------------- layer 1 epilogue: a1 = lrelu(acc + c1) -> hi|lo in place; c1 read one K block ahead ----------------
                    {
                        float4 cn[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) cn[q] = __ldcg(sc[sl] + q * P_ROWS);
                        wait_full2(&s.full[sl][0], ph_l[sl]); ph_l[sl] ^= 1;
#pragma unroll
                        for (int kb = 0; kb < 5; ++kb) {
                            const float4 cc[4] = {cn[0], cn[1], cn[2], cn[3]};
                            if (kb < 4) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) cn[q] = __ldcg(sc[sl] + ((kb + 1) * 4 + q) * P_ROWS);
                            }
                            uint32_t acc[16], pc[16];
                            const uint32_t ta = tls + PC_R1 + (hf * 5 + kb) * 16;
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
                        arrive_x(sl);                                     // -> layer 2 (+ h part of gates half 0)
                    }
                    // ---------------- layer-2 epilogue + folded layers 3+4 (80 -> 2): partial velocity over this thread's 40 columns ----------------
                    wait_full2(&s.full[sl][0], ph_l[sl]); ph_l[sl] ^= 1;
                    float v0 = 0.0f, v1 = 0.0f;
                    {
                        uint32_t acc[40];
                        tmem_ld<40>(tls + PC_R2 + hf * 40, acc);
                        ptx::tcgen05_wait_ld();
                        const float4* b2 = reinterpret_cast<const float4*>(s.f32 + PF_B2 + hf * 40);
                        const float4* w34 = reinterpret_cast<const float4*>(s.f32 + PF_W34 + hf * 80);
#pragma unroll
                        for (int j = 0; j < 10; ++j) {
                            const float4 b = b2[j], wa = w34[2 * j], wb = w34[2 * j + 1];
                            const float y0 = lrelu02(__uint_as_float(acc[4 * j]) + b.x), y1 = lrelu02(__uint_as_float(acc[4 * j + 1]) + b.y);
                            const float y2 = lrelu02(__uint_as_float(acc[4 * j + 2]) + b.z), y3 = lrelu02(__uint_as_float(acc[4 * j + 3]) + b.w);
                            v0 = fmaf(y0, wa.x, v0); v1 = fmaf(y0, wa.y, v1);
                            v0 = fmaf(y1, wa.z, v0); v1 = fmaf(y1, wa.w, v1);
                            v0 = fmaf(y2, wb.x, v0); v1 = fmaf(y2, wb.y, v1);
                            v0 = fmaf(y3, wb.z, v0); v1 = fmaf(y3, wb.w, v1);
                        }
                    }
                    ptx::tcgen05_fence_before_thread_sync();
                    if (hf == 1) { s.vpart[sl][r] = v0; s.vpart[sl][P_ROWS + r] = v1; vel_arrive2(sl); }
                    else {              // column half 0 finishes the row: velocity, integration, emit; (p, v) -> hi|lo x block of the gate MMA
                        vel_sync2(sl);
                        v0 += s.vpart[sl][r] + s.f32[PF_B34];
                        v1 += s.vpart[sl][P_ROWS + r] + s.f32[PF_B34 + 1];
                        p0[sl] += v0; p1[sl] += v1;
                        out_of_range |= !(fmaxf(fmaxf(fabsf(p0[sl]), fabsf(p1[sl])), fmaxf(fabsf(v0), fabsf(v1))) <= 6.0e4f);   // fp16 range guard
                        if (feed_back) {
                            uint32_t hp, lp, hv, lv;
                            psplit2(p0[sl], p1[sl], hp, lp);
                            psplit2(v0, v1, hv, lv);
                            *reinterpret_cast<uint4*>(s.xk[sl] + (size_t)r * 8) = make_uint4(hp, hv, lp, lv);                      // k 0..7
                            *reinterpret_cast<uint4*>(s.xk[sl] + (size_t)(P_ROWS + r) * 8) = make_uint4(hp, hv, 0x3C003C00u, 0u);  // k 8..15
                            ptx::fence_proxy_async(ptx::space_shared);
                            ptx::tcgen05_fence_before_thread_sync();
                            __syncwarp();
                            if (lane == 0) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(bar_x_leader + 16u + 8u * sl) : "memory");
                        }
                        if (valid[sl])
                            *reinterpret_cast<float4*>(out + ((size_t)(row0[sl] + r) * n_next + t) * 4) = make_float4(p0[sl], p1[sl], v0, v1);
                    }
                }
            }
            if (out_of_range && status && hf == 0 && (valid[0] || valid[1])) atomicOr(status, 1);
        }
    }
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

}  // namespace sw

static int pair2_grid(long long tiles, int sm_count) {
    const long long units = (tiles + 1) / 2;            // two tiles (one per CTA of a pair) per unit, two unit slots per pair
    long long pairs = (units + 1) / 2;
    if (pairs > sm_count / 2) pairs = sm_count / 2;
    if (pairs < 1) pairs = 1;
    return (int)(2 * pairs);
}

extern "C" long long sw_decode_pair_scratch_bytes(int sm_count);

extern "C" int sw_decode_fwd_pair2(const void* pair_w16, const float* pair_f32, const float* h0, const float* c0,
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
    CUtensorMap noise_map;
    const int rc = encode_noise_map2(&noise_map, noise, n_rows);
    if (rc != SW_OK) return rc;
    const int smem = (int)sizeof(sw::Pair2Smem);
    SW_SET_MAX_SMEM(sw::decode_fwd_pair2_kernel, smem);
    const int grid = pair2_grid(tiles, sm_count);
    sw::decode_fwd_pair2_kernel<<<grid, sw::R_THREADS, smem, (cudaStream_t)stream>>>(
        noise_map, (const __half*)pair_w16, pair_f32, h0, c0, pooled, x_last, out, (float4*)scratch, status, n_agents, n_rows, n_next,
        (int)tiles);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
