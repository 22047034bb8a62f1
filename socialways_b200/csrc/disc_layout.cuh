// Working layout of the Discriminator's 8 Linear layers (reference train.py:281-292) for csrc/disc_step.cu, produced
// by sw_disc_pack (csrc/pack.cu) from the 16 head tensors in Discriminator.parameters() order:
//   0 Wo1 [32][64]  1 bo1   2 Wo2 [32][32]  3 bo2   4 Wp1 [32][P]  5 bp1   6 Wp2 [32][32]  7 bp2
//   8 Wc1 [32][64]  9 bc1  10 Wc2 [1][32]  11 bc2  12 Wl1 [32][64] 13 bl1  14 Wl2 [2][32]  15 bl2
// Forward operands are k-major W^T[k][n] with padded rows (36 floats: the two k-split lanes of a thread group hit
// different banks); backward operands are the torch [out j][in k] matrices (contraction over j), rows padded the same way.
#pragma once
#define SW_DISC_PMAX 128

namespace sw {

constexpr int HW_LD32 = 36, HW_LD64 = 68;

struct HeadsWork {
    int P, P4;
    int f_wo1t, f_wo2t, f_wp1t, f_wp2t, f_wc1t, f_wl1t;       // forward, k-major [K][36]
    int b_wcl, b_wo2, b_wp2, b_wo1, b_wp1;                     // backward, [j][ld]
    int v_bo1, v_bo2, v_bp1, v_bp2, v_bc1, v_bl1, v_wc2, v_wl2, v_bc2, v_bl2;
    int total;
    __host__ __device__ explicit HeadsWork(int p) : P(p), P4((p + 3) & ~3) {
        int o = 0;
        f_wo1t = o; o += 64 * HW_LD32;
        f_wo2t = o; o += 32 * HW_LD32;
        f_wp1t = o; o += P * HW_LD32;
        f_wp2t = o; o += 32 * HW_LD32;
        f_wc1t = o; o += 64 * HW_LD32;
        f_wl1t = o; o += 64 * HW_LD32;
        b_wcl = o; o += 64 * HW_LD64;
        b_wo2 = o; o += 32 * HW_LD32;
        b_wp2 = o; o += 32 * HW_LD32;
        b_wo1 = o; o += 32 * HW_LD64;
        b_wp1 = o; o += 32 * P4;
        v_bo1 = o; o += 32; v_bo2 = o; o += 32; v_bp1 = o; o += 32; v_bp2 = o; o += 32; v_bc1 = o; o += 32; v_bl1 = o; o += 32;
        v_wc2 = o; o += 32; v_wl2 = o; o += 64; v_bc2 = o; o += 1; v_bl2 = o; o += 3;
        total = (o + 3) & ~3;
    }
    // value of work element e from the 16 head tensors
    __device__ float value(const float* const* T, int e) const {
        auto kmajor = [&](const float* W, int k_in, int q) {           // [K][36] <- W[n][k]
            const int k = q / HW_LD32, n = q % HW_LD32;
            return n < 32 ? __ldg(W + n * k_in + k) : 0.0f;
        };
        auto rows = [&](const float* W, int k_in, int ld, int q) {     // [j][ld] <- W[j][k]
            const int j = q / ld, k = q % ld;
            return k < k_in ? __ldg(W + j * k_in + k) : 0.0f;
        };
        if (e < f_wo2t) return kmajor(T[0], 64, e - f_wo1t);
        if (e < f_wp1t) return kmajor(T[2], 32, e - f_wo2t);
        if (e < f_wp2t) return kmajor(T[4], P, e - f_wp1t);
        if (e < f_wc1t) return kmajor(T[6], 32, e - f_wp2t);
        if (e < f_wl1t) return kmajor(T[8], 64, e - f_wc1t);
        if (e < b_wcl) return kmajor(T[12], 64, e - f_wl1t);
        if (e < b_wo2) { const int q = e - b_wcl; return q < 32 * HW_LD64 ? rows(T[8], 64, HW_LD64, q) : rows(T[12], 64, HW_LD64, q - 32 * HW_LD64); }
        if (e < b_wp2) return rows(T[2], 32, HW_LD32, e - b_wo2);
        if (e < b_wo1) return rows(T[6], 32, HW_LD32, e - b_wp2);
        if (e < b_wp1) return rows(T[0], 64, HW_LD64, e - b_wo1);
        if (e < v_bo1) return rows(T[4], P, P4, e - b_wp1);
        if (e < v_bo2) return __ldg(T[1] + e - v_bo1);
        if (e < v_bp1) return __ldg(T[3] + e - v_bo2);
        if (e < v_bp2) return __ldg(T[5] + e - v_bp1);
        if (e < v_bc1) return __ldg(T[7] + e - v_bp2);
        if (e < v_bl1) return __ldg(T[9] + e - v_bc1);
        if (e < v_wc2) return __ldg(T[13] + e - v_bl1);
        if (e < v_wl2) return __ldg(T[10] + e - v_wc2);
        if (e < v_bc2) return __ldg(T[14] + e - v_wl2);
        if (e < v_bl2) return __ldg(T[11]);
        if (e < v_bl2 + 2) return __ldg(T[15] + e - v_bl2);
        return 0.0f;
    }
};

// Activation / gradient records of one 32-row tile, written as tile images [rows][32] for sw_contract:
//   X image rows: h 64 | o1 32 | pred P | p1 32 | both 64 | c1 32 | l1 32                      (256 + P rows)
//   G image rows: do1 32 | doc 32 | dp1 32 | dpc 32 | dc1 32 | dlabel 1 | dl1 32 | dcode 2     (195 rows, padded to 196)
constexpr int HX_H = 0, HX_O1 = 64, HX_PRED = 96;
__host__ __device__ constexpr int hx_p1(int P) { return 96 + P; }
__host__ __device__ constexpr int hx_both(int P) { return 128 + P; }
__host__ __device__ constexpr int hx_c1(int P) { return 192 + P; }
__host__ __device__ constexpr int hx_l1(int P) { return 224 + P; }
__host__ __device__ constexpr int hx_rows(int P) { return 256 + P; }
constexpr int HG_DO1 = 0, HG_DOC = 32, HG_DP1 = 64, HG_DPC = 96, HG_DC1 = 128, HG_DLABEL = 160, HG_DL1 = 161, HG_DCODE = 193,
              HG_ROWS = 196;

}  // namespace sw
