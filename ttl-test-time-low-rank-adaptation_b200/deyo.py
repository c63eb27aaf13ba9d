"""Drop-in for the reference's `deyo.py` on the TTL path: the weighted-entropy head that `--deyo_selection` (truthy by
default, ttl.py:380,408) routes to.  Only the script-default flag set is on the B200 path (filter_ent=0, filter_plpd=0,
reweight_ent=1, reweight_plpd=0; deyo.py:103-108,159-181); the PLPD patch-shuffle second forward (deyo.py:115-151) is
out of scope.  Loss and gradient come from one CUDA kernel (ttl_op_deyo_loss)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ttl_b200 import functional as F_ttl


def softmax_entropy(x: torch.Tensor) -> torch.Tensor:
    """deyo.py:85-90."""
    return F_ttl.softmax_entropy(x)


def _check_flags(args):
    if getattr(args, "filter_ent", 0) or getattr(args, "filter_plpd", 0) or getattr(args, "reweight_plpd", 0) \
            or not getattr(args, "reweight_ent", 1):
        raise NotImplementedError("DeYO filter_ent / filter_plpd / reweight_plpd branches are outside the B200 path")


@torch.enable_grad()
def forward_and_adapt_sar(x, iter_, model, args, optimizer, scaler, deyo_margin, margin, targets=None, flag=True,
                          group=None):
    """deyo.py:93-196 (default flags): forward, weighted entropy over the views with H <= ln 1000, one optimiser step."""
    _check_flags(args)
    outputs = model(x)
    if not flag:
        return outputs
    entropys = softmax_entropy(outputs)
    backward = int((entropys <= math.log(1000)).sum())
    if backward == 0:
        return outputs, 0, 0
    loss = F_ttl.deyo_weighted_entropy(outputs, margin)
    optimizer.zero_grad()
    scaler.scale(loss).backward()
    scaler.step(optimizer)
    scaler.update()
    return outputs, backward, backward


class DeYO(nn.Module):
    """deyo.py:17-75: calls forward_and_adapt_sar `steps` times per forward (the reference wraps this in another
    tta_steps loop, ttl.py:78-81, hence tta_steps**2 optimiser steps -- reproduced)."""

    def __init__(self, model, args, optimizer, scaler, steps=1, episodic=False, deyo_margin=0.5 * math.log(1000),
                 margin_e0=0.4 * math.log(1000)):
        super().__init__()
        self.model, self.optimizer, self.scaler, self.args = model, optimizer, scaler, args
        self.steps, self.episodic = steps, episodic
        self.deyo_margin, self.margin_e0 = deyo_margin, margin_e0

    def forward(self, x, iter_=None, targets=None, flag=True, group=None):
        out = None
        for _ in range(self.steps):
            out = forward_and_adapt_sar(x, iter_, self.model, self.args, self.optimizer, self.scaler, self.deyo_margin,
                                        self.margin_e0, targets, flag, group)
        return out
