"""Oracle self-checks that need no reference: closed-form head gradients vs autograd, AdamW rule vs
torch.optim.AdamW, selection tie-break contract.  CPU only."""
import math

import pytest
import torch

from oracle import ttl_oracle as O


@pytest.mark.parametrize("C", [10, 200, 1000])
def test_avg_entropy_grad_closed_form(C):
    torch.manual_seed(C)
    x = (torch.randn(6, C, dtype=torch.float64) * 3).requires_grad_(True)
    O.avg_entropy(x).backward()
    assert torch.allclose(x.grad, O.avg_entropy_grad(x.detach()), atol=1e-12)


@pytest.mark.parametrize("C", [10, 1000])
def test_deyo_grad_closed_form(C):
    torch.manual_seed(C)
    x = (torch.randn(64, C, dtype=torch.float64) * 2).requires_grad_(True)
    O.deyo_loss(x).backward()
    assert torch.allclose(x.grad, O.deyo_loss_grad(x.detach()), atol=1e-12)


def test_adamw_matches_torch():
    torch.manual_seed(0)
    p0 = torch.randn(16, 48)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=5e-3)
    mine = {0: [p0.clone()]}
    st = O.AdamWState()
    for _ in range(3):
        g = torch.randn_like(p0)
        ref.grad = g.clone()
        opt.step()
        O.adamw_step(mine, {0: [g]}, st)
    assert torch.allclose(ref.detach(), mine[0][0], atol=1e-7)


def test_selection_lowest_index_first_on_ties():
    logits = torch.zeros(64, 10)
    logits[:, 0] = 1.0                       # all entropies identical
    logits[40, 0] = 5.0                      # strictly most confident
    _, idx = O.select_confident_samples(logits, 0.1)
    assert idx.tolist() == [40, 0, 1, 2, 3, 4]
    assert int(8 * 0.1) == 0                 # V < 10 selects nothing (ttl.py:52)


def test_tiny_arch_runs():
    arch = O.ARCHS["ViT-tiny"]
    spec = O.LoraSpec(rank=4, alpha=8.0, layer_lo=2, layer_hi=3)
    w = O.make_synthetic_weights(arch, 1)
    lora0 = O.lora_init(arch, spec, 0)
    imgs = O.make_synthetic_views(16, arch.image_size, 3)
    text = O.make_text_features(7, arch.proj)
    res = O.adapt_and_predict(arch, w, imgs, text, math.log(100.0), lora0, spec, head="tpt", selection_p=0.25)
    assert res.idx.numel() == 4 and res.pred_logits.shape == (1, 7)
    assert all(float(res.grads[i][0].abs().max()) == 0 for i in spec.layers())   # dA == 0 at step 1
    assert all(float(res.grads[i][1].abs().max()) > 0 for i in spec.layers())
