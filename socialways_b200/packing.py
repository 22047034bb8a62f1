"""Weight folding / packing for the sm_100a kernels (differentiable torch ops, tiny).

All folds are exact algebra on the reference parameters (fp32 re-association only):

* ``pack_encoder``: EncoderLstm.embed (Linear 4->64, train.py:251) is folded into the LSTM input
  projection:  gates = W_ih (W_e x + b_e) + b_ih + W_hh h + b_hh = (W_ih W_e) x + W_hh h + b.
* ``pack_decoder``: DecoderFC.fc1 (train.py:324-328): layer-1 weight stored k-major with the
  rows ordered {h, S, z}; the last two Linear layers (no activation between) become one 80->2.
* ``pack_pool``: EmbedSocialFeatures.fc.{0,2} as-is; fc.4 and AttentionPooling.W folded into the
  per-agent vectors  u_j = W3^T (W h_j + b_W),  beta_j = b3 . (W h_j + b_W)  (``pool_agent_terms``), so
  that sigma_ij = a2_ij . u_j + beta_j  (train.py:169 with emb = fc(features)).

lstm_pack layout [69][256]: rows 0..3 Wx (k-major), 4..67 Whh (k-major), 68 bias; column
n' = 4*unit + gate with torch's gate order (i, f, g, o).
"""
import torch


def _interleave(m):
    """[256, K] with rows gate*64+unit  ->  [K, 256] k-major with columns unit*4+gate."""
    k = m.shape[1]
    return m.view(4, 64, k).permute(2, 1, 0).reshape(k, 256)


def pack_lstm(w_x, w_hh, bias):
    return torch.cat([_interleave(w_x), _interleave(w_hh), _interleave(bias.view(256, 1))], dim=0).contiguous()


def pack_encoder(embed_w, embed_b, w_ih, w_hh, b_ih, b_hh):
    return pack_lstm(w_ih @ embed_w, w_hh, w_ih @ embed_b + b_ih + b_hh)


def pack_disc_lstm(w_ih, w_hh, b_ih, b_hh):
    return pack_lstm(w_ih, w_hh, b_ih + b_hh)


def pack_decoder(w1, b1, w2, b2, w3, b3, w4, b4):
    w34 = w4 @ w3                       # [2, 80]
    b34 = w4 @ b3 + b4
    return torch.cat([w1.t().reshape(-1), b1, w2.t().reshape(-1), b2, w34.t().reshape(-1), b34]).contiguous()


def pack_pool(fc0_w, fc0_b, fc2_w, fc2_b):
    return torch.cat([torch.cat([fc0_w, fc0_b.unsqueeze(1)], dim=1).reshape(-1), fc2_w.reshape(-1), fc2_b]).contiguous()


def pool_agent_matrix(att_w, att_b, fc4_w, fc4_b):
    """[64, 65] matrix M and [65] offset m0 with  (u | beta) = h @ M + m0."""
    tail = torch.cat([fc4_w, fc4_b.unsqueeze(1)], dim=1)    # [64(n), 65]: Wh @ tail = (u | beta)
    return att_w.t() @ tail, att_b @ tail
