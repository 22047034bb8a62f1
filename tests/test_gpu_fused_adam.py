"""Flat fused Adam (csrc/flat_adam.cu) vs torch.optim.Adam -- the optimiser the reference builds at
train.py:381,385 -- and the trainer with fused_adam=True (eager and CUDA-graph replay) vs the plain trainer."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_flat_adam_matches_torch_adam():
    from socialways_b200.fused_optim import FlatAdam
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(64, 4), (64,), (256, 64), (2, 40), (7,)]
    ref = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes]
    mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    o_ref = torch.optim.Adam(ref, lr=1e-3, betas=(0.9, 0.999))
    o_mine = FlatAdam(mine, lr=1e-3, betas=(0.9, 0.999))
    for step in range(25):
        o_ref.zero_grad()
        o_mine.zero_grad()
        for a, b in zip(ref, mine):
            gr = torch.randn(a.shape, device="cuda", generator=g) * (10.0 ** ((step % 5) - 3))
            a.grad = gr.clone()
            b.grad.add_(gr)                      # autograd accumulates into the existing view the same way
        o_ref.step()
        o_mine.step()
        for a, b in zip(ref, mine):
            # one Adam step moves a weight by <= lr; the two implementations may differ by an ulp of that move
            assert (a - b).abs().max().item() <= 2e-7 * (step + 1) + 1e-9, step
    sd = o_mine.state_dict()
    assert float(sd["state"][0]["step"]) == 25.0 and sd["param_groups"][0]["betas"] == (0.9, 0.999)
    ref_sd = o_ref.state_dict()
    for i in range(len(shapes)):
        # moments of O(1) gradients: fp32 rounding of 25 accumulations (ATen contracts some of these into FMAs, this
        # kernel rounds every operation; entries that cancel to ~0 only have absolute accuracy)
        assert torch.allclose(sd["state"][i]["exp_avg"], ref_sd["state"][i]["exp_avg"], rtol=1e-5, atol=2e-6)
        assert torch.allclose(sd["state"][i]["exp_avg_sq"], ref_sd["state"][i]["exp_avg_sq"], rtol=1e-5, atol=1e-7)
    # round trip through the torch-layout state dict
    again = FlatAdam([torch.nn.Parameter(p.detach().clone()) for p in mine], lr=5e-4)
    again.load_state_dict(sd)
    assert again.lr == 1e-3 and float(again.step_t[0].item()) == 25.0
    assert torch.equal(again.exp_avg[:again.n], o_mine.exp_avg[:o_mine.n])


@pytest.mark.parametrize("graph", [False, True])
def test_trainer_with_fused_adam_matches_plain_trainer(graph):
    from oracle import socialways_oracle as so
    from socialways_b200.trainer import SocialWaysTrainer
    data = so.toy_samples(216, 6)
    W = so.init_weights(seed=4, n_next=2)
    epochs = 3 if graph else 2

    def run(fused, graph):
        tr = SocialWaysTrainer(data, batch_size=64, use_social=True, weights=W, cuda_graph=graph, fused_adam=fused)
        np.random.seed(9)
        torch.manual_seed(9)
        for _ in range(epochs):
            ade, fde = (tr.train_graphed if graph else tr.train)(verbose=False)
        return tr.reference_weights(), ade, fde, tr

    w0, ade0, fde0, _ = run(False, False)
    w1, ade1, fde1, tr = run(True, graph)
    worst = max((w0[k] - w1[k]).abs().max().item() for k in w0)
    print(f"fused_adam (graph={graph}): max |w - w_plain| = {worst:.3e}; ADE {ade1:.6f} vs {ade0:.6f}")
    assert worst < 5e-5 and abs(ade0 - ade1) < 1e-4 and abs(fde0 - fde1) < 1e-4
    # the packed-weight cache of the inference path must see the fused updates (version counters advanced)
    m = tr.test(5, verbose=False)
    assert np.isfinite(m["ade_avg"])
    sd = tr.state()
    assert set(sd) == {'epoch', 'attentioner_dict', 'feature_embedder_dict', 'encoder_dict', 'decoder_dict', 'pred_optimizer',
                       'D_dict', 'D_optimizer'}
