/* Job descriptor of sw_contract / sw_contract_tc (socialways_b200.h): one weight-gradient contraction
 *     G[k][n] = sum_rows A[row][a_k0 + k] * B[row][b_n0 + n],   k < K (+ 1 if ones_row), n < N <= 256
 * replacing the `grad_weight = grad_output^T . input` (and grad_bias) nodes of the reference's autograd graph
 * (train.py:495 d_loss.backward(), :538 g_loss.backward()).  Row K of G, present when ones_row != 0, is the product with an
 * all-ones column, i.e. the bias gradient sum_rows B[row][n].
 *
 * G is scattered into up to SW_CONTRACT_MAX_SEGS destinations: segment s owns rows k_begin <= k < k_begin + k_count and
 * stores  out[(k - k_begin) * out_sk + perm(n) * out_sn] = G[k][n]  (and the same into out2 when given -- e.g. b_ih and
 * b_hh of an LSTM).  Rows not covered by a segment are dropped.  With n_perm = SW_CONTRACT_PERM_GATES (N == 256) the
 * gate-interleaved column n' = 4*unit + gate is stored at gate*64 + unit (torch's LSTM row order).
 *
 * Operand kinds
 *   SW_CONTRACT_IMAGE  tile images: element (row = image*32 + r, k) at base[image * stride + k * 32 + r];
 *                      `stride` = floats between consecutive images.  Padding rows of a gradient image are zero.
 *   SW_CONTRACT_ROWS   row-major records: element (row, k) at base[row * stride + k]; `stride` = floats per record;
 *                      rows >= n_rows are masked.
 * n_images = ceil(n_rows / 32) groups of 32 batch rows are summed.
 */
#ifndef SW_CONTRACT_H
#define SW_CONTRACT_H

#define SW_CONTRACT_IMAGE 0
#define SW_CONTRACT_ROWS 1
#define SW_CONTRACT_PERM_NONE 0
#define SW_CONTRACT_PERM_GATES 1
#define SW_CONTRACT_MAX_JOBS 16
#define SW_CONTRACT_MAX_SEGS 3
#define SW_CONTRACT_MAX_N 256

typedef struct sw_contract_seg {
    float* out;
    float* out2;          /* optional second destination, same strides, or NULL */
    int k_begin, k_count;
    int out_sk, out_sn;
} sw_contract_seg;

typedef struct sw_contract_job {
    const float* a;
    const float* b;
    long long a_stride;
    long long b_stride;
    int a_k0, K;
    int b_n0, N;
    int n_images, n_rows;
    int a_kind, b_kind;
    int ones_row, n_perm;
    int n_segs, reserved;
    sw_contract_seg seg[SW_CONTRACT_MAX_SEGS];
} sw_contract_job;

#endif
