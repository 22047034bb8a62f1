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
 * projection h . Whh^T as tcgen05.mma on fp16 hi/lo split operands (fp32 accumulate in TMEM), Wx . x4 + b as fp32 FMAs.
 * enc_w16 / enc_f32 from packing.pack_encoder_tcx: Whh hi | lo fp16 canonical [2][8][256][8]; wx4 [256][4] | bL [256]. */
int sw_lstm_seq_fwd_tcx(const void* enc_w16, const float* enc_f32, const float* x, int in_dim, int n_rows,
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
 * TMEM (~1e-6 of sw_pool_fwd).  pool_w16 = packing.pack_pool_tcx (fp16 [4096]).  Scenes of up to
 * sw_pool_tcx_max_scene() agents; larger scenes use sw_pool_fwd. */
int sw_pool_fwd_tcx(const float* pool_pack, const void* pool_w16, const float* x_last, const float* h, const float* ub,
                    const int* scene_offsets, const int* agent_scene, float* pooled, int n_agents, int max_scene,
                    void* stream);
int sw_pool_tcx_max_scene(void);

/* Backward of sw_pool_fwd.  Replaces autograd through AttentionPooling.forward / EmbedSocialFeatures.fc
 * (train.py:160-175,183-188) inside g_loss.backward() (train.py:538).
 *   dS [N][64] gradient of the pooled vector, tdot [N] = dS_i . S_i, attn from the forward call,
 *   pair_offsets [n_scenes+1] (int64) start of every scene's A*A block of ordered pairs
 *   dub out [N][65] gradient of (u | beta); dh_direct out [N][64] = sum_i a_ij dS_i
 *   st_a1 [P][32], st_g2 [P][64], st_g1 [P][32], st_f [P][4] out: per-pair layer-1 activations, layer-2 /
 *   layer-1 pre-activation gradients and (features, 1); the MLP weight gradients are plain GEMMs of these */
int sw_pool_bwd(const float* pool_pack, const float* x_last, const float* h, const float* ub,
                const float* dS, const float* tdot, const float* attn, const int* scene_offsets,
                const int* agent_scene, const long long* pair_offsets, float* dub, float* dh_direct,
                float* st_a1, float* st_g2, float* st_g1, float* st_f, int n_agents, int max_scene,
                void* stream);

/* K-sample autoregressive decode in one launch.  Replaces the loop of predict() (train.py:418-432:
 * DecoderFC.forward :330-335 + integration :423-425 + one EncoderLstm step :430 per predicted step) and
 * the serial best-of-K loop around it in test() (train.py:583-585): row = k * n_agents + n.
 *   h0,c0 [N][64], pooled [N][64] or NULL (use_social False, train.py:413), noise [K][N][32],
 *   x_last [N][4]; out [K][N][n_next][4] = (p, v) per step (train.py:425,432)
 *   stash_xh [T][tiles][68][32], stash_gates [T-1][tiles][5][64][32], stash_a1 [T][tiles][160][32],
 *   stash_a2 [T][tiles][80][32]: backward-pass stash (tile images), all NULL for inference */
int sw_decode_fwd(const float* lstm_pack, const float* dec_pack, const float* h0, const float* c0,
                  const float* pooled, const float* noise, const float* x_last, float* out,
                  float* stash_xh, float* stash_gates, float* stash_a1, float* stash_a2,
                  int n_agents, int n_samples, int n_next, int sm_count, void* stream);

/* Backward of sw_decode_fwd.  Replaces autograd through the decode loop (train.py:418-430) inside
 * g_loss.backward() (train.py:538).  dec_pack_t (sw_decode_pack_t_floats() floats) =
 * W1h^T [160][64] | W2^T [80][160] | W34 [80][2].  d_out [K*N][T][4] gradient of the emitted (p, v).
 * Outputs (tile images): g_gates [T-1][tiles][256][32], g_a1 [T][tiles][160][32], g_a2 [T][tiles][80][32],
 * g_v [T][tiles][2][32]; dh0, dc0 [K*N][64] gradients of the initial state.  Weight gradients = plain GEMMs
 * of the forward stash images against these (packing.py / autograd_path.py). */
int sw_decode_bwd(const float* lstm_pack_t, const float* dec_pack_t, const float* c0,
                  const float* stash_gates, const float* stash_a1, const float* stash_a2, const float* d_out,
                  float* g_gates, float* g_a1, float* g_a2, float* g_v, float* dh0, float* dc0,
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
 * layer).  Packs from packing.pack_decoder_tcx; sizes via sw_decode_tcx_pack_sizes. */
int sw_decode_fwd_tcx(const void* tcx_w16, const void* tcx_wsz16, const float* tcx_f32, const float* h0,
                      const float* c0, const float* pooled, const float* noise, const float* x_last, float* out,
                      int n_agents, int n_samples, int n_next, int sm_count, void* stream);
int sw_decode_tcx_pack_sizes(int* n_w16, int* n_wsz16, int* n_f32);

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
 * built at train.py:381,385 (called at :496, :539); same update rule and operation order, step count `step` (one device
 * float, advanced by the call: safe to replay from a CUDA graph); hyper-parameters are the python doubles torch receives.
 *   sw_adam_flat: params, grads, exp_avg, exp_avg_sq [n].
 *   sw_allreduce_adam: additionally replaces the gradient all-reduce of the sharded step (one NCCL call per optimiser step
 *     otherwise).  peer_bufs_dev = device array of `world` pointers to every rank's SYMMETRIC buffer laid out
 *     [n_pad floats of gradient | uint32 ready[world] | uint32 done[world]] (flags zero-initialised, n_pad % 32 == 0);
 *     params / exp_avg / exp_avg_sq are n_pad floats (zero padded); seq = 2 device uint32 (zero-initialised).  Gradients are
 *     summed with peer loads over NVLink in rank order on every rank (bit-identical replicas); the call returns once
 *     every peer has finished reading this rank's gradients.  Every rank must make the call (same order). */
int sw_adam_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* step, int n, double lr,
                 double beta1, double beta2, double eps, int sm_count, void* stream);
int sw_allreduce_adam(const void* peer_bufs_dev, int rank, int world, int n, int n_pad, float* params, float* exp_avg,
                      float* exp_avg_sq, float* step, unsigned* seq, double lr, double beta1, double beta2, double eps,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif
