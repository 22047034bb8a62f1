/* Job descriptor of sw_contract (socialways_b200.h): one weight-gradient contraction
 *     out[k * out_sk + n * out_sn] (+)= scale * sum_rows A[row][a_k0 + k] * B[row][b_n0 + n],   k < K, n < N
 * replacing one `grad_weight = grad_output^T . input` node of the reference's autograd graph
 * (train.py:495 d_loss.backward(), :538 g_loss.backward()).
 *
 * Operand kinds
 *   SW_CONTRACT_IMAGE  tile images [image][rows_per_image_matrix][32]: element (row = image*32 + r, k) at
 *                      base[image * stride + k * 32 + r]; `stride` = floats between consecutive images
 *   SW_CONTRACT_ROWS   row-major records: element (row, k) at base[row * stride + k]; `stride` = floats per record
 *   SW_CONTRACT_ONES   (A only) the all-ones column: K must be 1 -- bias gradients
 * n_images = ceil(n_rows / 32) batch-row groups are summed; rows >= n_rows contribute nothing (ROWS / ONES operands are
 * masked; IMAGE gradient operands carry zeros in their padding rows).
 */
#ifndef SW_CONTRACT_H
#define SW_CONTRACT_H

#define SW_CONTRACT_IMAGE 0
#define SW_CONTRACT_ROWS 1
#define SW_CONTRACT_ONES 2
#define SW_CONTRACT_PERM_NONE 0
#define SW_CONTRACT_PERM_GATES 1 /* N == 256 gate-interleaved columns n' = 4*unit + gate are stored at gate*64 + unit */
#define SW_CONTRACT_MAX_JOBS 24

typedef struct sw_contract_job {
    const float* a;
    const float* b;
    float* out;
    float* out2;          /* optional second destination, same strides (e.g. b_ih and b_hh of an LSTM), or NULL */
    long long a_stride;
    long long b_stride;
    int a_k0, K;
    int b_n0, N;
    int n_images, n_rows;
    int a_kind, b_kind;
    int out_sk, out_sn;
    int n_perm, accumulate;
    float scale;
    int reserved;
} sw_contract_job;

#endif
