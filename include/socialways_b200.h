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

/* LSTM over a sequence.  Replaces: get_traj_4d (train.py:130-134, when in_dim == 2) + EncoderLstm.forward
 * (train.py:262-269; observation pass at :404) and Discriminator.obsv_encoder_lstm (train.py:296-299).
 *   x        [n_rows][n_steps][in_dim]  in_dim 2 = positions (velocities formed on the fly), 4 = (p, v) states
 *   h_in,c_in [n_rows][64] initial state, both NULL = zeros (train.py:399-401)
 *   y_out    [n_rows][n_steps][64] every h_t, or NULL
 *   h_out,c_out [n_rows][64] state after the last step
 *   x_last   [n_rows][4] last 4-d state (train.py:416), or NULL
 *   stash_*  backward-pass stash ([T][N][64][5] gates+cell, [T][N][64] h, [T][N][4] inputs), all NULL for inference */
int sw_lstm_seq_fwd(const float* lstm_pack, const float* x, int in_dim, int n_rows, int n_steps,
                    const float* h_in, const float* c_in, float* y_out, float* h_out, float* c_out,
                    float* x_last, float* stash_gates, float* stash_h, float* stash_x4, int sm_count, void* stream);

/* Fused pairwise social features + embedding MLP + attention pooling.  Replaces: SocialFeatures,
 * BearingMTX, DCA_MTX (train.py:208-241), EmbedSocialFeatures.forward (train.py:178-189) and
 * AttentionPooling.forward (train.py:153-175) as called from predict() (train.py:408-411).
 *   x_last [N][4], h [N][64] post-observation encoder state, ub [N][65] = (u | beta) (packing.py)
 *   scene_offsets [n_scenes+1] ascending agent offsets (`sub_batches`, train.py:461), agent_scene [N]
 *   pooled [N][64] out; attn [N][round_up(max_scene,4)] softmax weights out, or NULL */
int sw_pool_fwd(const float* pool_pack, const float* x_last, const float* h, const float* ub,
                const int* scene_offsets, const int* agent_scene, float* pooled, float* attn,
                int n_agents, int max_scene, void* stream);

/* K-sample autoregressive decode in one launch.  Replaces the loop of predict() (train.py:418-432:
 * DecoderFC.forward :330-335 + integration :423-425 + one EncoderLstm step :430 per predicted step) and
 * the serial best-of-K loop around it in test() (train.py:583-585): row = k * n_agents + n.
 *   h0,c0 [N][64], pooled [N][64] or NULL (use_social False, train.py:413), noise [K][N][32],
 *   x_last [N][4]; out [K][N][n_next][4] = (p, v) per step (train.py:425,432) */
int sw_decode_fwd(const float* lstm_pack, const float* dec_pack, const float* h0, const float* c0,
                  const float* pooled, const float* noise, const float* x_last, float* out,
                  int n_agents, int n_samples, int n_next, int sm_count, void* stream);

/* Best-of-K error metrics.  Replaces train.py:587 and :602-607 of test().
 *   pred [K][N][T][4], gt [N][T][2] (normalised), ss = Scale.sx (train.py:121)
 *   out [N][4] = (avg-K ADE, avg-K FDE, min-K ADE, min-K FDE) per agent */
int sw_bestofk_metrics(const float* pred, const float* gt, float ss, int n_agents, int n_samples,
                       int n_next, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
