// Tensor-core version of the fused pairwise social features + embedding MLP + attention pooling kernel (pool_fwd.cu;
// reference train.py:208-241, :178-189, :153-175) for the inference path.
//
// pool_fwd_kernel spends 2 048 of its ~2 200 FMAs per ordered pair (i, j) in layer 2 of EmbedSocialFeatures
// (32 -> 64), with every thread re-reading the same weights from shared memory: it is bound by instruction issue
// (~3.5 K instructions per pair), an order of magnitude below what the 788 B of HBM traffic per agent would allow.
// Here layer 2 is a real dense contraction on the 5th-gen tensor cores: 128 pairs = the 128 rows of one tcgen05.mma
// tile, a1 (the 32 layer-1 activations of each pair) written by the pair's thread as fp16 hi|lo operand rows in the
// canonical K-major layout, W2 hi|lo resident in shared memory, three MMAs per product (hi.hi + hi.lo + lo.hi, fp32
// accumulate in TMEM: ~1e-6 of the FFMA result, same scheme as decode_fwd_tcx.cu).  The CUDA cores keep what is not a
// contraction: the 3 features and layer 1 per pair (~150 instructions), the layer-3/attention fold
// sigma_ij = relu(a2_ij + b2) . u_j + beta_j read straight from the accumulator's TMEM lane, the -1000 self mask,
// the per-row softmax and the weighted sum of the raw h.
//
// Work unit = a run of consecutive agent rows from a host-built table (ops.SceneIndex.pool_units): whole scenes packed up to
// 64 rows while scenes are small, <= 32 (16) rows of ONE scene when a scene has more than 64 agents -- so a unit's span (the
// agents its rows attend to) is the unit itself or that one scene.  x and (u | beta) of the span are staged in shared memory;
// h too when it fits (scenes up to ~256 agents), else the weighted sum reads it through L1/L2 (every row of the unit reads
// the same scene).  The unit's ordered pairs are enumerated row-major ([row][j]) and processed 128 at a time.  Without a
// table (NULL) units are 64 consecutive rows, scenes <= 64 agents (the round-1 scheme).
#include <cuda_fp16.h>
#include <cstdlib>

#include "sw_common.cuh"
#include "sw_umma.cuh"

namespace sw {

constexpr int PT_GROUP = 128;        // one 128-pair tile per thread group: thread p of the group = pair p of its tile = TMEM lane p
constexpr int PT_THREADS = 256;      // TWO groups per CTA work on two tiles at a time (own operand buffer, TMEM columns and barrier)
                                     // against ONE staged span: the tile chain (features -> MMA -> epilogue) is latency-bound, and
                                     // the second group doubles the warps per SM at the same shared-memory footprint
#ifndef SW_PT_ROWS
#define SW_PT_ROWS 64     // agent rows per work unit (A/B-tested with -DSW_PT_ROWS=32 / 128)
#endif
constexpr int PT_ROWS = SW_PT_ROWS;
constexpr int PT_LD = 65;
constexpr int PT_A_MAX = 512;        // largest scene this kernel takes (larger ones go to pool_fwd_kernel)
constexpr int PT_A_LEGACY = 64;      // largest scene of the table-less unit scheme
constexpr int PT_PP_P1 = 0, PT_PP_B2 = 128 + 64 * 32;      // offsets inside pool_pack (pool_fwd.cu)

__device__ __forceinline__ void split2_pt(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h2 = __floats2half2_rn(a, b);
    const float2 back = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(a - back.x, b - back.y);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}

template <int G>      // lanes per row in the softmax / weighted-sum stage (8, 16 or 32 by the largest scene)
__global__ void __launch_bounds__(PT_THREADS)
pool_fwd_tcx_kernel(const float* __restrict__ pool_pack, const __half* __restrict__ w2_16 /* canonical [4][64][8] hi | lo */,
                    const float* __restrict__ x_last, const float* __restrict__ h, const float* __restrict__ ub,
                    const int* __restrict__ scene_offsets, const int* __restrict__ agent_scene, float* __restrict__ pooled,
                    const int* __restrict__ units /*[n_units][2] = (first row, rows) or null*/, int stage_h,
                    int n_agents, int span_cap, int pair_cap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __half* w2 = reinterpret_cast<__half*>(smem_raw);                       // [2][2048]           8 KB
    __half* a1s_all = w2 + 2 * 2048;                                        // [group][2][4][128][8]   2 x 16 KB
    float* p1 = reinterpret_cast<float*>(a1s_all + 2 * 2 * 4096);           // [32][4]
    float* b2 = p1 + 128;                                                   // [64]
    float* sx = b2 + 64;                                                    // [span][4]
    float* sh = sx + span_cap * 4;                                          // [span][64] (read lane <-> column only), if staged
    float* su_raw = sh + (stage_h ? span_cap * SW_H : 0);                   // [span][65] (col 64 = beta) + 4 floats of slack
    float* sig = su_raw + span_cap * PT_LD + 4;                             // [pair_cap]
    int* off = reinterpret_cast<int*>(sig + pair_cap);                      // [PT_ROWS + 1] pair offsets of the unit's rows
    int* sstart = off + PT_ROWS + 1;                                        // [PT_ROWS] first agent of the row's scene
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sstart + PT_ROWS + 1);  // [2], 8 B aligned (even int count)
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = tid >> 7, gt = tid & (PT_GROUP - 1);                    // tile group, pair slot inside the group's tile
    __half* a1s = a1s_all + grp * 2 * 4096;
    unsigned long long* bar = bars + grp;

    const int row0 = units ? units[2 * blockIdx.x] : blockIdx.x * PT_ROWS;
    const int row1 = units ? row0 + units[2 * blockIdx.x + 1] : min(row0 + PT_ROWS, n_agents);
    const int R = row1 - row0;
    const int span0 = scene_offsets[agent_scene[row0]];
    const int span1 = scene_offsets[agent_scene[row1 - 1] + 1];
    const int span = span1 - span0;

    for (int i = tid; i < 2 * 2048 / 8; i += PT_THREADS)
        reinterpret_cast<uint4*>(w2)[i] = __ldg(reinterpret_cast<const uint4*>(w2_16) + i);
    for (int i = tid; i < 128; i += PT_THREADS) p1[i] = __ldg(pool_pack + PT_PP_P1 + i);
    if (tid < 64) b2[tid] = __ldg(pool_pack + PT_PP_B2 + tid);
    for (int i = tid; i < span; i += PT_THREADS)
        *reinterpret_cast<float4*>(sx + i * 4) = __ldg(reinterpret_cast<const float4*>(x_last) + span0 + i);
    if (stage_h)
        for (int i = tid; i < span * 16; i += PT_THREADS)                   // flat copy: same [.][64] layout on both sides
            reinterpret_cast<float4*>(sh)[i] = __ldg(reinterpret_cast<const float4*>(h) + (size_t)span0 * 16 + i);
    // (u | beta) rows keep the global [.][65] layout; the copy is flat too.  The shared copy starts at the same offset
    // modulo 4 floats as the global one, so the 16-byte aligned body moves as float4 on both sides.
    const float* ug = ub + (size_t)span0 * 65;
    const int phase = (int)(((size_t)span0 * 65) & 3);
    float* su = su_raw + phase;
    {
        const int total = span * 65, head = min((4 - phase) & 3, total), body4 = (total - head) >> 2;
        for (int i = tid; i < head; i += PT_THREADS) su[i] = __ldg(ug + i);
        for (int i = tid; i < body4; i += PT_THREADS)
            *reinterpret_cast<float4*>(su + head + i * 4) = __ldg(reinterpret_cast<const float4*>(ug + head) + i);
        for (int i = head + body4 * 4 + tid; i < total; i += PT_THREADS) su[i] = __ldg(ug + i);
    }
    if (tid < R) {
        const int sc = agent_scene[row0 + tid];
        sstart[tid] = scene_offsets[sc];
        off[tid + 1] = scene_offsets[sc + 1] - scene_offsets[sc];         // scene size, prefix-summed below
    }
    if (warp == 0) {
        ptx::tcgen05_alloc(ptx::cta_group_1, tmem_base_s, 128u);
        ptx::tcgen05_relinquish_alloc_permit(ptx::cta_group_1);
    }
    if (tid == 0) {
        ptx::mbarrier_init(reinterpret_cast<uint64_t*>(bars), 1);
        ptx::mbarrier_init(reinterpret_cast<uint64_t*>(bars + 1), 1);
        ptx::fence_mbarrier_init(ptx::sem_release, ptx::scope_cluster);
    }
    __syncthreads();
    if (tid == 0) {
        off[0] = 0;
        for (int r = 0; r < R; ++r) off[r + 1] += off[r];
    }
    ptx::fence_proxy_async(ptx::space_shared);
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    ptx::tcgen05_fence_after_thread_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_base_s, 0) + (uint32_t)(grp * 64);     // this group's 64 columns
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int P = off[R];
    uint32_t ph = 0;
    auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" :: "r"(1 + grp), "n"(PT_GROUP) : "memory"); };

    for (int t0 = grp * PT_GROUP; t0 < P; t0 += PT_THREADS) {     // the groups take alternate tiles; a group's loop is its own
        const int q = t0 + gt;
        const bool valid = q < P;
        int i = 0, j = 0;
        float a1[32];
        if (valid) {
            int lo = 0, hi = R;                         // row r with off[r] <= q < off[r + 1]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (off[mid] <= q) lo = mid; else hi = mid;
            }
            i = row0 + lo;
            j = sstart[lo] + (q - off[lo]);
            const float4 xi = *reinterpret_cast<const float4*>(sx + (size_t)(i - span0) * 4);
            const float4 xj = *reinterpret_cast<const float4*>(sx + (size_t)(j - span0) * 4);
            // the three features, term for term as pool_fwd.cu::pair_score (train.py:208-241)
            const float dpx = xi.x - xj.x, dpy = xi.y - xj.y, dvx = xi.z - xj.z, dvy = xi.w - xj.w;
            const float dist = sqrtf(dpx * dpx + dpy * dpy);
            const float vnorm = sqrtf(xi.z * xi.z + xi.w * xi.w);
            const float bearing = (dpx * xi.z + dpy * xi.w) / (dist * vnorm + 1e-6f);
            const float ttca = -((dpx * dvx + dpy * dvy) / (dvx * dvx + dvy * dvy + 1e-6f));
            const float cx = dpx + ttca * dvx, cy = dpy + ttca * dvy;
            const float dca = sqrtf(cx * cx + cy * cy);
#pragma unroll
            for (int n = 0; n < 32; ++n) {
                const float4 w = *reinterpret_cast<const float4*>(p1 + n * 4);
                a1[n] = fmaxf(fmaf(w.x, dist, fmaf(w.y, bearing, fmaf(w.z, dca, w.w))), 0.0f);
            }
        } else {
#pragma unroll
            for (int n = 0; n < 32; ++n) a1[n] = 0.0f;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {                   // operand row of this pair: chunk c = features 8c .. 8c+7
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split2_pt(a1[c * 8 + 2 * e], a1[c * 8 + 2 * e + 1], hi[e], lo[e]);
            *reinterpret_cast<uint4*>(a1s + ((size_t)c * PT_GROUP + gt) * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(a1s + 4096 + ((size_t)c * PT_GROUP + gt) * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        ptx::fence_proxy_async(ptx::space_shared);
        ptx::tcgen05_fence_before_thread_sync();
        group_sync();
        if ((warp & 3) == 0) {                          // a2 = a1 . W2^T : [128 pairs] x [64] x K = 32, three passes
            ptx::tcgen05_fence_after_thread_sync();
            if (elect_one()) {                          // single-lane issue (sw_umma.cuh)
                umma1_ss<64, 64, 2>(tmem, a1s, w2, 0u, false);
                umma1_ss<64, 64, 2>(tmem, a1s, w2 + 2048, 0u, true);
                umma1_ss<64, 64, 2>(tmem, a1s + 4096, w2, 0u, true);
                umma1_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, ph); ph ^= 1;
        ptx::tcgen05_fence_after_thread_sync();
        {   // sigma_ij = relu(a2 + b2) . u_j + beta_j from this pair's TMEM lane
            const float* uj = su + (size_t)(valid ? j - span0 : 0) * PT_LD;
            float sigma = uj[64];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t acc[32];
                tmem_ld<32>(tl + half * 32, acc);
                ptx::tcgen05_wait_ld();
#pragma unroll
                for (int n = 0; n < 32; ++n)
                    sigma = fmaf(fmaxf(__uint_as_float(acc[n]) + b2[half * 32 + n], 0.0f), uj[half * 32 + n], sigma);
            }
            if (valid) sig[q] = (j == i) ? -1000.0f : sigma;              // train.py:170
        }
        ptx::tcgen05_fence_before_thread_sync();
        group_sync();
    }
    __syncthreads();                                    // both groups' scores are in shared memory

    // softmax over the scene (train.py:172) and S_i = sum_j a_ij h_j on the RAW h (:173); a group of G lanes per row
    {
        constexpr int SLOTS = (PT_THREADS / 32) * (32 / G);
        const int gl = lane % G, slot = warp * (32 / G) + lane / G;
        const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
        for (int r = slot; r < R; r += SLOTS) {
            const int A = off[r + 1] - off[r];
            float* s = sig + off[r];
            float* dst = pooled + (size_t)(row0 + r) * SW_H;
            if (A == 1) {                                    // train.py:165
#pragma unroll
                for (int k = 0; k < SW_H / G; ++k) dst[gl + k * G] = 0.0f;
                continue;
            }
            float mx = -3.0e38f;
            for (int jj = gl; jj < A; jj += G) mx = fmaxf(mx, s[jj]);
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, o));
            float sum = 0.0f;
            for (int jj = gl; jj < A; jj += G) {
                const float e = expf(s[jj] - mx);
                s[jj] = e;
                sum += e;
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(gmask, sum, o);
            __syncwarp(gmask);
            for (int jj = gl; jj < A; jj += G) s[jj] = s[jj] / sum;
            __syncwarp(gmask);
            const float* hp = (stage_h ? sh + (size_t)(sstart[r] - span0) * SW_H : h + (size_t)sstart[r] * SW_H) + gl;
            float acc[SW_H / G];
#pragma unroll
            for (int k = 0; k < SW_H / G; ++k) acc[k] = 0.0f;
            for (int jj = 0; jj < A; ++jj) {
                const float w = s[jj];
#pragma unroll
                for (int k = 0; k < SW_H / G; ++k) acc[k] = fmaf(w, hp[(size_t)jj * SW_H + k * G], acc[k]);
            }
#pragma unroll
            for (int k = 0; k < SW_H / G; ++k) dst[gl + k * G] = acc[k];
        }
    }
    ptx::tcgen05_fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) ptx::tcgen05_dealloc(ptx::cta_group_1, __shfl_sync(0xffffffffu, *tmem_base_s, 0), 128u);
}

}  // namespace sw

// Same contract as sw_pool_fwd (inference: no attention record) plus the fp16 hi|lo operand pack of layer 2
// (packing.pack_pool_tcx: fc.2.weight [64][32] as canonical [4][64][8] hi block, then lo block = 4096 halves) and the work-unit
// table: units [n_units][2] = (first row, rows <= 64), every unit's rows consecutive, its span (all agents of the scenes it
// touches) <= max_unit_span agents and its ordered pairs <= max_unit_pairs (ops.SceneIndex.pool_units builds it: whole scenes
// packed to <= 64 rows, scenes above 64 agents cut into units of <= 32 / 16 rows).  units == NULL: 64 consecutive rows per
// unit, scenes of up to 64 agents.  Takes scenes of up to sw_pool_tcx_max_scene() agents; the caller uses sw_pool_fwd beyond.
extern "C" int sw_pool_tcx_max_scene(void) { return sw::PT_A_MAX; }

extern "C" int sw_pool_fwd_tcx(const float* pool_pack, const void* pool_w16, const float* x_last, const float* h, const float* ub,
                               const int* scene_offsets, const int* agent_scene, float* pooled, const int* units, int n_units,
                               int max_unit_span, int max_unit_pairs, int n_agents, int max_scene, void* stream) {
    if (!pool_pack || !pool_w16 || !x_last || !h || !ub || !scene_offsets || !agent_scene || !pooled) return SW_ERR_ARG;
    if (n_agents <= 0 || max_scene <= 0) return SW_ERR_ARG;
    if (max_scene > sw::PT_A_MAX || (!units && max_scene > sw::PT_A_LEGACY)) return SW_ERR_UNSUPPORTED;
    if (units && (n_units <= 0 || max_unit_span <= 0 || max_unit_pairs <= 0)) return SW_ERR_ARG;
    // multiple of 4: keeps every shared-memory array behind the (u | beta) rows (65 floats each) 16-byte aligned, the mbarrier 8
    const int span_cap = units ? (max_unit_span + 3) & ~3 : sw::PT_ROWS + 2 * (max_scene - 1);
    const int pair_cap = ((units ? max_unit_pairs : sw::PT_ROWS * max_scene) + 3) & ~3;
    const size_t fixed = 2 * 2048 * 2 + 2 * 2 * 4096 * 2 + (size_t)(128 + 64 + 4) * 4 + (size_t)(2 * sw::PT_ROWS + 2) * 4 + 24;
    const size_t no_h = fixed + (size_t)(span_cap * (4 + sw::PT_LD) + pair_cap) * 4;
    const size_t with_h = no_h + (size_t)span_cap * SW_H * 4;
    const char* knob = getenv("SW_POOL_STAGE_H");                   // A/B knob (profiles/): default = measured best
    const int stage_h = knob ? (atoi(knob) != 0 && with_h <= 200 * 1024) : (with_h <= 200 * 1024 ? 1 : 0);
    const size_t smem = stage_h ? with_h : no_h;
    if (smem > 227 * 1024) return SW_ERR_UNSUPPORTED;
    const int grid = units ? n_units : (n_agents + sw::PT_ROWS - 1) / sw::PT_ROWS;
    cudaStream_t st = (cudaStream_t)stream;
#define SW_POOL_TCX_LAUNCH(GG)                                                                                              \
    do {                                                                                                                    \
        SW_SET_MAX_SMEM(sw::pool_fwd_tcx_kernel<GG>, (int)smem);                                                            \
        sw::pool_fwd_tcx_kernel<GG><<<grid, sw::PT_THREADS, smem, st>>>(pool_pack, (const __half*)pool_w16, x_last, h, ub,  \
                                                                        scene_offsets, agent_scene, pooled, units, stage_h, \
                                                                        n_agents, span_cap, pair_cap);                      \
    } while (0)
    if (max_scene <= 8) SW_POOL_TCX_LAUNCH(8);
    else if (max_scene <= 16) SW_POOL_TCX_LAUNCH(16);
    else SW_POOL_TCX_LAUNCH(32);
#undef SW_POOL_TCX_LAUNCH
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
