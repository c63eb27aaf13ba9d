"""Drop-in for the reference's `deyo.py` on the TTL path: the weighted-entropy head that `--deyo_selection` (truthy by
default, ttl.py:380,408) routes to.
  * script-default flags (filter_ent=0, filter_plpd=0, reweight_ent=1, reweight_plpd=0; deyo.py:103-108,159-181): loss and
    gradient come from one CUDA kernel (ttl_op_deyo_loss); this is also what the fused library call implements;
  * the optional branches -- `filter_ent` top-p selection inside this head (deyo.py:103-105) and
    `filter_plpd`, the second forward on a structure-destroyed copy of the kept views (deyo.py:115-151) -- run in compat
    mode only: the encoder forward/backward are the library's, the few [V,C]-sized head operations are torch ops on the
    device (they are off the north-star path and not worth a kernel)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ttl_b200 import functional as F_ttl


def softmax_entropy(x: torch.Tensor) -> torch.Tensor:
    """deyo.py:85-90."""
    return F_ttl.softmax_entropy(x)


def _default_flags(args) -> bool:
    return not (getattr(args, "filter_ent", 0) or getattr(args, "filter_plpd", 0) or getattr(args, "reweight_plpd", 0)
                or getattr(args, "reweight_ent", 1) != 1)


def destroy_structure(x: torch.Tensor, args) -> torch.Tensor:
    """x' of deyo.py:116-136: the kept views with their object structure destroyed, by `--aug_type`
    occ   : a occlusion_size^2 window at (row_start, column_start) replaced by the per-channel mean of the view;
    patch : resize to a multiple of patch_len, shuffle the patch_len^2 tiles of every view independently, resize back
            (torchvision Resize on tensors = antialiased bilinear); the permutations come from the CPU torch RNG;
    pixel : one random permutation of the pixel positions shared by all views and channels."""
    x = x.detach().clone()
    B, Cc, Hh, Ww = x.shape
    kind = getattr(args, "aug_type", "patch")
    if kind == "occ":
        fill = x.reshape(B, Cc, -1).mean(dim=2)[:, :, None, None]
        r0, c0, n = args.row_start, args.column_start, args.occlusion_size
        x[:, :, r0:r0 + n, c0:c0 + n] = fill.expand(-1, -1, n, n)
        return x
    if kind == "patch":
        from torchvision.transforms import Resize
        pl = args.patch_len
        side = (Ww // pl) * pl
        t = Resize((side, side))(x)
        ph = side // pl
        tiles = t.reshape(B, Cc, pl, ph, pl, ph).permute(0, 2, 4, 1, 3, 5).reshape(B, pl * pl, Cc, ph, ph)
        order = torch.argsort(torch.rand(B, pl * pl), dim=-1).to(x.device)
        tiles = tiles[torch.arange(B, device=x.device)[:, None], order]
        t = tiles.reshape(B, pl, pl, Cc, ph, ph).permute(0, 3, 1, 4, 2, 5).reshape(B, Cc, side, side)
        return Resize((Ww, Ww))(t)
    if kind == "pixel":
        flat = x.reshape(B, Cc, Hh * Ww)
        return flat[:, :, torch.randperm(Hh * Ww).to(x.device)].reshape(B, Cc, Ww, Ww)
    raise ValueError(f"unknown --aug_type {kind}")


def _adapt_general(outputs, x, model, args, optimizer, scaler, margin):
    """deyo.py:102-188 with any flag combination (compat mode; [V,C]-sized torch ops around the library's forward/backward)."""
    ent = -(outputs.softmax(1) * outputs.log_softmax(1)).sum(1)          # softmax_entropy with its autograd graph
    if args.filter_ent:
        keep = torch.argsort(ent, descending=False)[:int(ent.size(0) * args.selection_p)]
    else:
        keep = torch.nonzero(ent <= math.log(1000)).flatten()
    ent = ent[keep]
    backward = int(ent.numel())
    if backward == 0:
        return outputs, 0, 0
    if args.filter_plpd:
        with torch.no_grad():       # x' only decides which views stay; no gradient flows through it (deyo.py:137-148)
            out_prime = model(destroy_structure(x[keep], args))
            p, p_prime = outputs[keep].softmax(1), out_prime.softmax(1)
            top = p.argmax(dim=1, keepdim=True)
            plpd = (p.gather(1, top) - p_prime.gather(1, top)).flatten()
            keep2 = torch.nonzero(plpd > args.plpd_threshold).flatten()
        ent = ent[keep2]
    final_backward = int(ent.numel())
    if args.reweight_ent or args.reweight_plpd:
        ent = ent * (args.reweight_ent / torch.exp(ent.detach() - margin))   # the PLPD re-weighting term is disabled in the reference
    if final_backward != 0:
        loss = ent.mean(0)
        optimizer.zero_grad()
        scaler.scale(loss).backward()
        scaler.step(optimizer)
        scaler.update()
    return outputs, backward, final_backward


@torch.enable_grad()
def forward_and_adapt_sar(x, iter_, model, args, optimizer, scaler, deyo_margin, margin, targets=None, flag=True,
                          group=None):
    """deyo.py:93-196: forward, (filtered, re-weighted) entropy over the views, one optimiser step."""
    outputs = model(x)
    if not flag:
        return outputs
    if not _default_flags(args):
        return _adapt_general(outputs, x, model, args, optimizer, scaler, margin)
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
