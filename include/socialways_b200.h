/* socialways_b200 C-ABI -- the drop-in boundary of the B200-native Social Ways hot path.
 *
 * The reference (crowdbotp/socialways, /root/reference, commit 0b13f2d) is pure Python over PyTorch
 * and has no FFI: its operator surface for this path is the Python functions/classes of train.py.
 * Each entry point below names the reference code it replaces (file:line); the Python host
 * (socialways_b200/reference_api.py) mirrors the reference names on top of these calls and
 * INTEGRATION.md shows the ctypes stub a maintainer would add to train.py.
 *
 * Conventions
 *   - all pointers are DEVICE pointers (fp32 / int32), owned and sized by the caller; no allocation,
 *     no host synchronisation, no host<->device copy happens inside the library;
 *   - `stream` is a cudaStream_t (CUstream) passed as void*; calls on different streams may overlap;
 *   - return value: SW_OK (0) or a negative code; sw_error_string(code) describes it, and for
 *     SW_ERR_CUDA sw_last_cuda_error() holds the cudaError_t of the calling thread;
 *   - `sm_count` is the number of SMs to size persistent grids for (148 on B200);
 *   - hidden size 64, noise length 32, 3 social features are compile-time constants of the path
 *     (train.py:43-45,79-81 defaults; every BASELINE.json config).
 *
 * Packed weights (`*_pack`) are produced by socialways_b200/packing.py from the reference
 * state_dict tensors; the layouts are documented there and in INTEGRATION.md:
 *   lstm_pack  [69][256]   rows: Wx[4] | Whh[64] | bias ; columns gate-interleaved n' = 4*unit+gate
 *   dec_pack   sw_decode_pack_floats() floats: W1[160][160] k-major | b1 | W2[160][80] | b2 | W34[80][2] | b34
 *   pool_pack  sw_pool_pack_floats() floats:   P1[32][4] (w_dist,w_bearing,w_dca,bias) | P2[64][32] | b2[64]
 */
#ifndef SOCIALWAYS_B200_H
#define SOCIALWAYS_B200_H

#include "sw_contract.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SW_OK 0
#define SW_ERR_ARG (-1)
#define SW_ERR_CUDA (-2)
#define SW_ERR_UNSUPPORTED (-3)

int sw_abi_version(void);
int sw_last_cuda_error(void);
const char* sw_error_string(int code);
int sw_decode_pack_floats(void);
int sw_pool_pack_floats(void);
int sw_decode_pack_t_floats(void);

/* LSTM over a sequence.  Replaces: get_traj_4d (train.py:130-134, when in_dim == 2) + EncoderLstm.forward
 * (train.py:262-269; observation pass at :404) and Discriminator.obsv_encoder_lstm (train.py:296-299).
 *   x        [n_rows][n_steps][in_dim]  in_dim 2 = positions (velocities formed on the fly), 4 = (p, v) states
 *   h_in,c_in [n_rows][64] initial state, both NULL = zeros (train.py:399-401)
 *   y_out    [n_rows][n_steps][64] every h_t, or NULL
 *   h_out,c_out [n_rows][64] state after the last step
 *   x_last   [n_rows][4] last 4-d state (train.py:416), or NULL
 *   stash_gates [T][tiles][5][64][32] (i,f,g,o,c per unit) and stash_xh [T][tiles][68][32] (the {x4;h} operand
 *            of every step): backward-pass stash in tile-image layout, both NULL for inference */
int sw_lstm_seq_fwd(const float* lstm_pack, const float* x, int in_dim, int n_rows, int n_steps,
                    const float* h_in, const float* c_in, float* y_out, float* h_out, float* c_out,
                    float* x_last, float* stash_gates, float* stash_xh, int sm_count, void* stream);

/* Tensor-core variant of sw_lstm_seq_fwd for the inference path (zero initial state, no stash, no y_out): the recurrent
 * projection h . Whh^T as tcgen05.mma on fp16 hi/lo split operands (fp32 accumulate in TMEM); Wx . x4 + b enters as one extra
 * K block of the same MMA ([x_hi | x_lo | x_hi | 1 | 1 | 0 0] . [Wx_hi | Wx_hi | Wx_lo | b_hi | b_lo | 0 0]^T), gate rows
 * pre-scaled by -log2(e) / -2 log2(e) so that the accumulator is the ex2 argument of the cell update.
 * enc_w16 from packing.pack_encoder_tcx: Whh hi | lo fp16 canonical [2][8][256][8], then the x block [2][256][8]. */
int sw_lstm_seq_fwd_tcx(const void* enc_w16, const float* x, int in_dim, int n_rows,
                        int n_steps, float* h_out, float* c_out, float* x_last, int sm_count, void* stream);

/* Backward of sw_lstm_seq_fwd from a zero initial state.  Replaces autograd through nn.LSTM
 * (train.py:254,268 for the generator, :278,299 for D) as walked by d_loss.backward() / g_loss.backward()
 * (train.py:495,538).  Tile-image layout = [..][tiles = ceil(N/32)][k][32 rows].
 *   lstm_pack_t [256][68] = transpose of lstm_pack rows 0..67; stash_gates from the forward call
 *   dh_last,dc_last [N][64] gradients w.r.t. the final state (NULL = zero)
 *   g_gates out [T][tiles][256][32] pre-activation gate gradients; the weight gradient is the plain GEMM
 *           d(lstm_pack)[0:68] = sum stash_xh[..][k][r] * g_gates[..][n][r], d(lstm_pack)[68] = sum g_gates
 *   dx out [N][T][4] gradient w.r.t. the 4-d inputs, or NULL */
int sw_lstm_seq_bwd(const float* lstm_pack_t, const float* stash_gates, const float* dh_last,
                    const float* dc_last, float* g_gates, float* dx, int n_rows, int n_steps,
                    int sm_count, void* stream);

/* Fused pairwise social features + embedding MLP + attention pooling.  Replaces: SocialFeatures,
 * BearingMTX, DCA_MTX (train.py:208-241), EmbedSocialFeatures.forward (train.py:178-189) and
 * AttentionPooling.forward (train.py:153-175) as called from predict() (train.py:408-411).
 *   x_last [N][4], h [N][64] post-observation encoder state, ub [N][65] = (u | beta) (packing.py)
 *   scene_offsets [n_scenes+1] ascending agent offsets (`sub_batches`, train.py:461), agent_scene [N]
 *   pooled [N][64] out; attn [N][round_up(max_scene,4)] softmax weights out, or NULL */
int sw_pool_fwd(const float* pool_pack, const float* x_last, const float* h, const float* ub,
                const int* scene_offsets, const int* agent_scene, float* pooled, float* attn,
                int n_agents, int max_scene, void* stream);
/* Same contract, inference only (no attention record), with layer 2 of the pair MLP (32 -> 64, 93 % of the pair's
 * arithmetic) on the tcgen05 tensor cores: 128 ordered pairs per MMA tile, fp16 hi|lo split operands, fp32 accumulate in
 * TMEM (~1e-6 of sw_pool_fwd).  pool_w16 = packing.pack_pool_tcx (fp16 [4096]).
 * units [n_units][2] = (first row, rows <= 64): work units of consecutive rows whose span (all agents of the scenes they touch)
 * is at most max_unit_span agents and whose ordered pairs number at most max_unit_pairs -- whole scenes packed up to 64 rows,
 * scenes above 64 agents cut into units of <= 32 / 16 rows (ops.SceneIndex.pool_units).  units == NULL: 64 consecutive rows per
 * unit, scenes of up to 64 agents.  Scenes of up to sw_pool_tcx_max_scene() agents; larger scenes use sw_pool_fwd. */
int sw_pool_fwd_tcx(const float* pool_pack, const void* pool_w16, const float* x_last, const float* h, const float* ub,
                    const int* scene_offsets, const int* agent_scene, float* pooled, const int* units, int n_units,
                    int max_unit_span, int max_unit_pairs, int n_agents, int max_scene, void* stream);
int sw_pool_tcx_max_scene(void);

/* Backward of sw_pool_fwd.  Replaces autograd through AttentionPooling.forward / EmbedSocialFeatures.fc
 * (train.py:160-175,183-188) inside g_loss.backward() (train.py:538).
 *   dS [N][64] gradient of the pooled vector, tdot [N] = dS_i . S_i (or NULL: evaluated in the kernel from `pooled`
 *   [N][64], the forward output), attn from the forward call,
 *   pair_offsets [n_scenes+1] (int64) start of every scene's A*A block of ordered pairs
 *   dub out [N][65] gradient of (u | beta); dh_direct out [N][64] = sum_i a_ij dS_i
 *   st_a1 [P][32], st_g2 [P][64], st_g1 [P][32], st_f [P][4] out: per-pair layer-1 activations, layer-2 /
 *   layer-1 pre-activation gradients and (features, 1); the MLP weight gradients are plain GEMMs of these */
int sw_pool_bwd(const float* pool_pack, const float* x_last, const float* h, const float* ub,
                const float* dS, const float* tdot, const float* pooled, const float* attn, const int* scene_offsets,
                const int* agent_scene, const long long* pair_offsets, float* dub, float* dh_direct,
                float* st_a1, float* st_g2, float* st_g1, float* st_f, int n_agents, int max_scene,
                void* stream);

/* K-sample autoregressive decode in one launch.  Replaces the loop of predict() (train.py:418-432:
 * DecoderFC.forward :330-335 + integration :423-425 + one EncoderLstm step :430 per predicted step) and
 * the serial best-of-K loop around it in test() (train.py:583-585): row = k * n_agents + n.
 *   h0,c0 [N][64], pooled [N][64] or NULL (use_social False, train.py:413), noise [K][N][32],
 *   x_last [N][4]; out [K][N][n_next][4] = (p, v) per step (train.py:425,432)
 *   stash_xh [T][tiles][68][32], stash_gates [T-1][tiles][5][64][32], stash_a1 [T][tiles][160][32],
 *   stash_a2 [T][tiles][80][32]: backward-pass stash (tile images), all NULL for inference;
 *   stash_sz [tiles][96][32] (optional, with the stash only): the step-invariant layer-1 operand [S ; z] as a tile image
 *   (zeros in the S rows when pooled is NULL) -- the A operand of the W1[S,z] weight gradient */
int sw_decode_fwd(const float* lstm_pack, const float* dec_pack, const float* h0, const float* c0,
                  const float* pooled, const float* noise, const float* x_last, float* out,
                  float* stash_xh, float* stash_gates, float* stash_a1, float* stash_a2, float* stash_sz,
                  int n_agents, int n_samples, int n_next, int sm_count, void* stream);

/* Backward of sw_decode_fwd.  Replaces autograd through the decode loop (train.py:418-430) inside
 * g_loss.backward() (train.py:538).  dec_pack_t (sw_decode_pack_t_floats() floats) =
 * W1h^T [160][64] | W2^T [80][160] | W34 [80][2].  d_out [K*N][T][4] gradient of the emitted (p, v).
 * Outputs (tile images): g_gates [T-1][tiles][256][32], g_a1 [T][tiles][160][32], g_a2 [T][tiles][80][32],
 * g_v [T][tiles][2][32]; dh0, dc0 [K*N][64] gradients of the initial state.  Weight gradients = contractions of the
 * forward stash images against these (sw_contract).  Optional (NULL to skip): g_a1sum [tiles][160][32] = sum over the
 * steps of g_a1 (gradient operand of the step-invariant [S ; z] columns of layer 1 and of b1); d_pooled [K*N][64] =
 * dL/dS per row, which needs w1_torch = DecoderFC.fc1[0].weight [160][160] in torch layout. */
int sw_decode_bwd(const float* lstm_pack_t, const float* dec_pack_t, const float* c0,
                  const float* stash_gates, const float* stash_a1, const float* stash_a2, const float* d_out,
                  float* g_gates, float* g_a1, float* g_a2, float* g_v, float* dh0, float* dc0,
                  float* g_a1sum, const float* w1_torch, float* d_pooled,
                  int n_agents, int n_samples, int n_next, int sm_count, void* stream);

/* Tensor-core variant of sw_decode_fwd: tcgen05.mma with BF16 operands in shared memory and FP32 accumulators
 * in TMEM (fast mode, not the fp32 parity mode).  tc_w16 / tc_f32 from packing.pack_decoder_tc;
 * sizes via sw_decode_tc_pack_sizes.  Other arguments as sw_decode_fwd (no backward stash). */
int sw_decode_fwd_tc(const void* tc_w16, const float* tc_f32, const float* h0, const float* c0,
                     const float* pooled, const float* noise, const float* x_last, float* out,
                     int n_agents, int n_samples, int n_next, int sm_count, void* stream);
int sw_decode_tc_pack_sizes(int* n_bf16, int* n_f32);

/* fp32-faithful tensor-core variant of sw_decode_fwd: tcgen05.mma on fp16 hi/lo split operands (x = hi + lo,
 * A.B ~= Ahi.Bhi + Ahi.Blo + Alo.Bhi, fp32 accumulate in TMEM; a1/a2 live in TMEM as the A operand of the next
 * layer).  Packs from packing.pack_decoder_tcx; sizes via sw_decode_tcx_pack_sizes.
 * status (device int, may be NULL): bit 0 is set when an emitted (p, v) leaves fp16's exponent range or is not finite --
 * i.e. when an operand of the split overflowed (layer-1 activations > 65 504, states > 6e4): the result is then not
 * trustworthy and the caller should use sw_decode_fwd (fp32 FFMA) for these inputs / weights. */
int sw_decode_fwd_tcx(const void* tcx_w16, const void* tcx_wsz16, const float* tcx_f32, const float* h0,
                      const float* c0, const float* pooled, const float* noise, const float* x_last, float* out,
                      int* status, int n_agents, int n_samples, int n_next, int sm_count, void* stream);
int sw_decode_tcx_pack_sizes(int* n_w16, int* n_wsz16, int* n_f32);

/* The same fp16 hi/lo split decode with TWO 128-row tiles in flight per SM (csrc/decode_fwd_pair.cu; the default decode of the
 * inference path): clusters of two CTAs issue every MMA as one tcgen05 cta_group::2 instruction (each CTA holds half of every
 * weight matrix), all 16 epilogue warps of a CTA serve two tile slots alternately (while they work on one slot's epilogue the
 * tensor pipe runs the other slot's MMAs), a dedicated warp issues every MMA of the pair, and the hoisted layer-1 term lives
 * in `scratch` (device memory, sw_decode_pair_scratch_bytes(sm_count) bytes, 16-byte aligned; written and re-read inside the
 * launch, contents meaningless afterwards; one buffer per concurrently running launch).  h0 and c0 must be 32-byte aligned
 * (their rows are read in 32-byte pieces), pooled and noise 16-byte aligned: SW_ERR_ARG otherwise.
 * Same inputs / outputs / status word / arithmetic as sw_decode_fwd_tcx; pack from packing.pack_decoder_pair
 * (sizes via sw_decode_pair_pack_sizes).  Replaces the loop of predict(), reference train.py:418-430, x K samples. */
int sw_decode_fwd_pair(const void* pair_w16, const float* pair_f32, const float* h0, const float* c0, const float* pooled,
                       const float* noise, const float* x_last, float* out, void* scratch, long long scratch_bytes,
                       int* status, int n_agents, int n_samples, int n_next, int sm_count, void* stream);
int sw_decode_pair_pack_sizes(int* n_w16, int* n_f32);

/* The bf16 build of the same kernel (csrc/decode_fwd_pair_bf16.cu): single bf16 operands, ONE tcgen05.mma per product instead of
 * the three of the hi/lo split (the x-feedback block keeps hi + lo).  Fast mode (BASELINE configs[2] names bf16): ~3e-3 from the
 * fp32 path in normalised coordinates, not inside the 1e-4 parity bar.  Same arguments; pack from
 * packing.pack_decoder_pair(..., bf16=True) (same sizes and layout, bf16 bit patterns). */
int sw_decode_fwd_pair_bf16(const void* pair_w16, const float* pair_f32, const float* h0, const float* c0, const float* pooled,
                            const float* noise, const float* x_last, float* out, void* scratch, long long scratch_bytes,
                            int* status, int n_agents, int n_samples, int n_next, int sm_count, void* stream);
long long sw_decode_pair_scratch_bytes(int sm_count);

/* Discriminator FC heads (train.py:281-292, 300-309), one thread per trajectory, all 8 Linear layers fused.
 *   pack: sw_disc_heads_pack_floats(P, L) floats = Wo1[32][64] bo1 Wo2[32][32] bo2 Wp1[32][P] bp1 Wp2[32][32] bp2
 *         Wc1[32][64] bc1 Wc2[1][32] bc2 Wl1[32][64] bl1 Wl2[L][32] bl2 (torch [out][in] layouts), P = n_next*4, L = 2
 *   h [N][64] last hidden state of the observation LSTM, pred [N][P]; label [N] and code [N][L] out
 *   xrec [N][257+P] out (or NULL): per-row record [h|o1|pred|p1|both|c1|l1|1] for the backward pass
 * Backward: d_label [N], d_code [N][L] in (NULL = 0); d_h [N][64], d_pred [N][P] out (NULL = skip);
 *   grec [N][193+L] out = [d_o1|d_oc|d_p1|d_pc|d_c1|d_label|d_l1|d_code]; every parameter gradient is a block of
 *   the plain GEMM xrec^T . grec (sw_disc_heads_record_dims gives the two widths). */
int sw_disc_heads_pack_floats(int pred_dim, int n_latent);
int sw_disc_heads_record_dims(int pred_dim, int n_latent, int* x_dim, int* g_dim);
int sw_disc_heads_fwd(const float* pack, const float* h, const float* pred, int pred_dim, int n_latent,
                      int n_rows, float* label, float* code, float* xrec, void* stream);
int sw_disc_heads_bwd(const float* pack, const float* xrec, int pred_dim, int n_latent, int n_rows,
                      const float* d_label, const float* d_code, float* d_h, float* d_pred, float* grec,
                      void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * The training iteration as ~30 launches (socialways_b200/native_step.py): what torch.autograd, ATen and cuBLAS do
 * around the forward / backward kernels in train() (reference train.py:470-551) -- parameter folding, the heads of the
 * discriminator with their losses, every weight gradient, the loss / ADE / FDE scalars -- as entry points of this library.
 * ------------------------------------------------------------------------------------------------------------------ */

/* Generator parameter folds (packing.py as ONE kernel).  params22 = HOST array of 22 device pointers in
 * Generator.optimizer_parameters() order (train.py:379-380):
 *   attention.W.{weight,bias}, feature_embedder.fc.{0,2,4}.{weight,bias}, encoder.embed.{weight,bias},
 *   encoder.lstm.{weight_ih_l0,weight_hh_l0,bias_ih_l0,bias_hh_l0}, decoder.fc1.{0,2,4,5}.{weight,bias}.
 * Outputs (sizes from sw_gen_pack_sizes): enc_pack [69][256] and its transpose rows 0..67 [256][68], dec_pack, dec_pack_t,
 * pool_pack, pool_m = M [64][65] | m0 [65] with (u | beta) = h.M + m0, pool_mt = M^T [65][64].
 * sw_gen_pack_bwd: the adjoint of the folds -- from d_enc [69][256] (pack layout), d_w34 [80][2] | d_b34 [2] and
 * d_m [64][65] | d_m0 [65] (written by sw_contract) to the gradients of attention.W, fc.4, embed, the LSTM and fc1.4/fc1.5
 * (grads22 = host array of 22 device pointers, same order; the other tensors' gradients are written by sw_contract directly).
 * have_pool = 0 (use_social False, train.py:83): the four pooling tensors get zero gradients. */
int sw_gen_pack_sizes(int* enc_pack, int* enc_pack_t, int* dec_pack, int* dec_pack_t, int* pool_pack, int* pool_m, int* pool_mt);
int sw_gen_pack(const float* const* params22, float* enc_pack, float* enc_pack_t, float* dec_pack, float* dec_pack_t,
                float* pool_pack, float* pool_m, float* pool_mt, void* stream);
int sw_gen_pack_bwd(const float* const* params22, float* const* grads22, const float* d_enc, const float* d_w34,
                    const float* d_m, int have_pool, void* stream);

/* Discriminator packs.  params20 = host array of 20 device pointers in Discriminator.parameters() order (train.py:278-292):
 * obsv_encoder_lstm.{weight_ih_l0,weight_hh_l0,bias_ih_l0,bias_hh_l0}, obsv_encoder_fc.{0,2}, pred_encoder.{0,2},
 * classifier.{0,2}, latent_decoder.{0,2} (.weight, .bias each).  pred_dim = n_next * 4.
 * Outputs: lstm_pack [69][256], lstm_pack_t [256][68], heads_work (layout csrc/disc_layout.cuh; sw_disc_pack_sizes). */
int sw_disc_pack_sizes(int pred_dim, int* lstm_pack, int* lstm_pack_t, int* heads_work);
int sw_disc_pack(const float* const* params20, int pred_dim, float* lstm_pack, float* lstm_pack_t, float* heads_work,
                 void* stream);

/* Discriminator heads + losses + their backward pass in one launch.  Replaces Discriminator.forward after the LSTM
 * (train.py:300-309), the nn.MSELoss terms (train.py:484-493 D step, :514-521 G step) and the part of d_loss.backward() /
 * g_loss.backward() (train.py:495,538) that runs through the heads.
 *   mode 0 (D step): rows = every agent's fake trajectory pred_fake [N][pred_dim] AND its real one, formed on the fly from
 *     pred_pos [N][n_next][2] and the last observed position of obsv_pos [N][n_past][2] (get_traj_4d, train.py:135-137).
 *     Targets: fake -> targets[0] (`zeros`, train.py:471), real -> targets[1] (`ones`, :472); info loss on the fake rows.
 *     Outputs: d_h [N][64] (gradient of the observation code, both branches summed), x_img [ceil(N/16)][x_rows][32] and
 *     g_img [ceil(N/16)][g_rows][32] tile-image records for sw_contract (rows: csrc/disc_layout.cuh, sw_disc_step_image_rows).
 *   mode 1 (G step): fake rows only, target targets[1]; output d_pred [N][pred_dim] = dL/d(pred_hat_4d).
 *   noise [N][noise_ld]: columns 0,1 are the latent codes (train.py:486,518); inv_n = 1 / (GLOBAL number of agents of the
 *   mini-batch) so that per-rank gradients sum to nn.MSELoss's over the global batch; info_w = loss_info_w.
 *   loss_part [tiles][4] = per-tile sums of squared errors (fake-or-fooling label, real label, info, 0).
 *   label_out [2N or N], code_out [N][2]: optional raw outputs. */
int sw_disc_step_image_rows(int pred_dim, int* x_rows, int* g_rows);
int sw_disc_step(const float* heads_work, int pred_dim, int mode, const float* obsv_h, const float* pred_fake,
                 const float* pred_pos, const float* obsv_pos, int n_past, const float* noise, int noise_ld,
                 const float* targets, float inv_n, float info_w, float* d_h, float* d_pred, float* x_img, float* g_img,
                 float* loss_part, float* label_out, float* code_out, int n_agents, int sm_count, void* stream);

/* All weight-gradient contractions of one backward pass in one launch (job descriptor: sw_contract.h).  Replaces the
 * grad_weight / grad_bias halves of autograd's Linear / LSTM nodes (train.py:495,538).
 *   sw_contract_tc: tcgen05.mma kind::tf32 on tf32-split fp32 operands (3 MMAs per product, fp32 accumulate in TMEM);
 *   sw_contract:    fp32 FFMA register tiles (arithmetic reference of the former, A/B timing).
 * `workspace` (sw_contract_plan floats, for the kernel chosen by `tensor_cores`) holds the partial slabs of jobs that are split
 * over CTAs, `counters` (zero-initialised, restored by the kernel) one word per output slab.  Deterministic: the summation
 * order does not depend on scheduling. */
int sw_contract_plan(const sw_contract_job* jobs, int n_jobs, int sm_count, int tensor_cores, long long* workspace_floats,
                     int* n_counters);
int sw_contract(const sw_contract_job* jobs, int n_jobs, float* workspace, long long workspace_floats, unsigned* counters,
                int n_counters, int sm_count, void* stream);
int sw_contract_tc(const sw_contract_job* jobs, int n_jobs, float* workspace, long long workspace_floats, unsigned* counters,
                   int n_counters, int sm_count, void* stream);

/* out[row][n] = bias[n] + add1[row][n] + add2[row][n] + sum_k x[row][k] w[k][n]  (k_in, n_out <= 80; bias/add1/add2 may
 * be NULL; add1/add2 have row stride ldo).  (u | beta) = h.M + m0 (train.py:158,167-169,185 folded) and its adjoint. */
int sw_rows_linear(const float* x, int ldx, const float* w, const float* bias, const float* add1, const float* add2,
                   float* out, int ldo, int n_rows, int k_in, int n_out, void* stream);

/* Scalars of one iteration (train.py:546-551 and the mse_loss values): stats[8] = (sum ADE terms / n_next, sum FDE terms,
 * d_loss, d_fake, d_real, d_info, g_fooling, g_info) from pred_hat [N][n_next][4], pred [N][n_next][2], ss = Scale.sx and
 * the loss_part arrays of the last D pass / the G pass.  partial [sm_count][2] and counter (zero-initialised) are scratch. */
int sw_train_stats(const float* pred_hat, const float* pred, int n_rows, int n_next, float ss, const float* d_parts,
                   int d_tiles, const float* g_parts, int g_tiles, float inv_n, float info_w, float* partial,
                   unsigned* counter, float* stats, int sm_count, void* stream);

/* Device-side latent noise (optional replacement of the host-side torch.rand + upload of train.py:584,473): n uniform
 * [0, 1) floats from Philox4x32-10; group g = first_group + e / 4 of the output = Philox(counter = (g, offset), key = seed),
 * values (x >> 8) * 2^-24.  A different stream from torch's CPU generator: parity runs keep caller-supplied noise. */
int sw_noise_uniform(float* out, long long n, unsigned long long seed, unsigned long long offset,
                     unsigned long long first_group, int sm_count, void* stream);

/* Best-of-K error metrics.  Replaces train.py:587 and :602-607 of test().
 *   pred [K][N][T][4], gt [N][T][2] (normalised), ss = Scale.sx (train.py:121)
 *   out [N][4] = (avg-K ADE, avg-K FDE, min-K ADE, min-K FDE) per agent */
int sw_bestofk_metrics(const float* pred, const float* gt, float ss, int n_agents, int n_samples,
                       int n_next, float* out, void* stream);

/* Sample-set statistics of calc_statistics.py.  Samples are [n][n_ped][t_len][2] arrays of fp32
 * (dtype_bytes 4) or fp64 (8) device data -- numpy's dtype decides the arithmetic there, so it does here; the distance
 *   d(a, b) = mean_{t >= obsv_len} || a_t - b_t ||_2   (calc_statistics.py:31-32, :56-57)
 * is evaluated in numpy's operation order (bit-identical matrices).
 *
 * sw_traj_nn1_counts: replaces compute_1nn's loops (calc_statistics.py:17-45).  counts[4] (device int32, zeroed by the
 *   call) = (Real_pos, Real_neg, Fake_pos, Fake_neg) summed over pedestrians; np.argmin's first-minimum rule.
 * sw_traj_emd_cost: replaces the cost-matrix loop of compute_wasserstein (calc_statistics.py:53-58) including its
 *   mirrored write (the matrix the solver sees is C[max(a,b)][min(a,b)]); cost [n_ped][n][n] fp64.
 * sw_lsap_solve: replaces scipy.optimize.linear_sum_assignment (calc_statistics.py:60; scipy is a third-party
 *   dependency of the reference, version unpinned) for square problems: cost [n_problems][n][n] fp64,
 *   col4row [n_problems][n] out (column assigned to each row, scipy's tie-breaking), *status = 1 if any problem
 *   is infeasible.  One warp per problem; sw_lsap_smem_bytes(n) of shared memory bounds n (<= ~4800). */
int sw_traj_nn1_counts(const void* reals, const void* fakes, int dtype_bytes, int n_reals, int n_fakes, int n_ped,
                       int t_len, int obsv_len, int* counts, void* stream);
int sw_traj_emd_cost(const void* reals, const void* fakes, int dtype_bytes, int n, int n_ped, int t_len, int obsv_len,
                     double* cost, void* stream);
int sw_lsap_smem_bytes(int n);
int sw_lsap_solve(const double* cost, int n, int n_problems, int* col4row, int* status, void* stream);

/* Optimiser step on ONE flat fp32 buffer.  Replaces torch.optim.Adam.step() of the optimisers
 * built at train.py:381,385 (called at :496, :539); same update rule and operation order; `step` = TWO device words:
 * [0] the step count as a float (advanced by the call: safe to replay from a CUDA graph), [1] a zero-initialised
 * scratch word (sw_adam_flat's finished-CTA counter); hyper-parameters are the python doubles torch receives.
 *   sw_adam_flat: params, grads, exp_avg, exp_avg_sq [n].
 *   sw_allreduce_adam: additionally replaces the gradient all-reduce of the sharded step (one NCCL call per optimiser step
 *     otherwise).  peer_bufs_dev = device array of `world` pointers to every rank's SYMMETRIC buffer laid out
 *     [n_pad floats of gradient | uint32 ready[world] | uint32 done[world]] (flags zero-initialised, n_pad % 32 == 0);
 *     params / exp_avg / exp_avg_sq are n_pad floats (zero padded); seq = 4 device uint32 (zero-initialised; seq[2] is a
 *     status word: 0 = ok, 1 = a peer never published its gradients within the wait bound and the step was NOT applied,
 *     2 = a peer never acknowledged reading).  Gradients are summed with peer loads over NVLink in rank order on every rank
 *     (bit-identical replicas); the kernel returns once every peer has finished reading this rank's gradients.  Every rank
 *     must make the call (same order).  sw_set_peer_wait_timeout_ms bounds every wait in time (default 60 000 ms). */
int sw_adam_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* step, int n, double lr,
                 double beta1, double beta2, double eps, int sm_count, void* stream);
int sw_set_peer_wait_timeout_ms(int ms);
int sw_allreduce_adam(const void* peer_bufs_dev, int rank, int world, int n, int n_pad, float* params, float* exp_avg,
                      float* exp_avg_sq, float* step, unsigned* seq, double lr, double beta1, double beta2, double eps,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif
