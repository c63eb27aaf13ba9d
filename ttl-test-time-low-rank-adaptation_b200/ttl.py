"""Drop-in for the reference's `ttl.py`: same command line (every flag of ttl.py:383-424, same names/defaults, prefix
abbreviations still resolve), same function surface (`select_confident_samples`, `avg_entropy`, `test_time_tuning`,
`test_time_adapt_eval`, `main_worker`, `main`), executed on B200 by libttl_b200.

    python ttl.py DATA --test_sets A --deyo_selection ''            # north-star head on a real dataset tree (needs the
                                                                     # reference's data/ package on PYTHONPATH)
    python ttl.py --synthetic 256 --test_sets A --deyo_selection '' # seeded synthetic evaluation set, no files needed
    torchrun --nproc-per-node 8 ttl.py --synthetic 8192 ...         # samples sharded across the GPUs of one box

New flags (names chosen so that the launcher's `--data`/`--b` abbreviations stay unambiguous):
  --synthetic N   evaluate on N seeded synthetic samples instead of a dataset on disk
  --compat        drive the module through autograd + torch.optim.AdamW (the reference's control flow) instead of
                  the fused per-sample call
  --views_on_host the reference's input route: the DataLoader workers produce 64 fp32 views per sample (host -> device copy
                  of 38.5 MB per sample) instead of the default uint8 image + crop boxes
  --concurrent_samples S  adapt S test samples per library call (default 9; 1 = strictly one at a time)
  --precision fp32  validation mode: every activation/contraction in fp32 (held to 1e-4 against the reference), samples one at a time
  --vision_checkpoint F  load image tower + text tower + logit_scale from one CLIP checkpoint file (HF or OpenAI format)
                  instead of the local HF cache
  --random_init   allow seeded random-init weights / random class features when no checkpoint exists (implied by --synthetic)
  --merges F      CLIP BPE merge table for the class prompts
  --views_on_device  (default) ship the decoded uint8 image + the drawn crop boxes and generate the 64 views on the GPU
                  (bit-exact with the reference's PIL/torchvision AugMixAugmenter) instead of 64 fp32 views per sample
"""
from __future__ import annotations

import argparse
import math
import os
import time
from copy import deepcopy

import torch

from ttl_b200 import dist as tdist
from ttl_b200 import functional as F_ttl


def list_of_ints(arg):
    return list(map(int, arg.split(',')))


# ------------------------------------------------------------------------------------- loss head (kernel-backed)
def select_confident_samples(logits, top):
    """ttl.py:50-54."""
    return F_ttl.select_confident_samples(logits, top)


def avg_entropy(outputs, plot=True):
    """ttl.py:56-61."""
    return F_ttl.avg_entropy(outputs)


# ------------------------------------------------------------------------------------- meters (utils/tools.py:26-102)
class AverageMeter:
    def __init__(self, name, fmt=':f'):
        self.name, self.fmt = name, fmt
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def __str__(self):
        return ('{name} {val' + self.fmt + '} ({avg' + self.fmt + '})').format(**self.__dict__)


def accuracy(output, target, topk=(1,)):
    """Percentage of rows whose target is within the top-k logits (k <= C)."""
    with torch.no_grad():
        maxk = min(max(topk), output.size(1))
        pred = output.topk(maxk, 1, True, True).indices.t()
        correct = pred.eq(target.view(1, -1).expand_as(pred))
        return [correct[:min(k, maxk)].reshape(-1).float().sum(0, keepdim=True) * (100.0 / target.size(0)) for k in topk]


# ------------------------------------------------------------------------------------- adaptation
def test_time_tuning(model, inputs, optimizer, scaler, args):
    """ttl.py:70-110 in the reference's own control flow (compat mode): forward with grad -> head -> backward -> step."""
    if getattr(args, "cocoop", False):
        raise NotImplementedError("--cocoop is outside the TTL path (and broken in the reference, SURVEY.md Q15)")
    if args.deyo_selection and args.lora_encoder != 'prompt':
        import deyo
        for j in range(args.tta_steps):
            d = deyo.DeYO(model, args, optimizer, scaler, steps=args.tta_steps, deyo_margin=args.deyo_margin,
                          margin_e0=args.deyo_margin_e0)
            d(inputs)
        return
    selected_idx = None
    for j in range(args.tta_steps):
        output = model(inputs)
        if selected_idx is not None:
            output = output[selected_idx]
        else:
            output, selected_idx = select_confident_samples(output, top=args.selection_p)
        if output.shape[0] == 0:
            return
        loss = avg_entropy(output.float())
        optimizer.zero_grad()
        scaler.scale(loss).backward()
        scaler.step(optimizer)
        scaler.update()
    return


class SyntheticViews(torch.utils.data.Dataset):
    """Seeded stand-in for AugMixAugmenter(ImageFolder) (data/datautils.py:129-157): item i is (list of `n_views`
    tensors [3,S,S] -- view 0 the centre crop, the rest random-resized-crop + flip of one smooth base image --, label)."""

    def __init__(self, n_samples, n_views=64, size=224, n_classes=1000, seed=0):
        self.n, self.v, self.size, self.c, self.seed = n_samples, n_views, size, n_classes, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        import torch.nn.functional as F
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        big = int(self.size * 1.25)
        base = (F.interpolate(torch.randn(1, 3, 7, 7, generator=g), size=(big, big), mode="bicubic")
                + 0.5 * F.interpolate(torch.randn(1, 3, 56, 56, generator=g), size=(big, big), mode="bilinear"))
        off = (big - self.size) // 2
        views = [base[0, :, off:off + self.size, off:off + self.size].clone()]
        r = torch.rand(self.v - 1, 5, generator=g)
        for k in range(self.v - 1):
            s = (0.08 + 0.92 * float(r[k, 0])) * big * big
            ar = math.exp(math.log(3 / 4) + float(r[k, 1]) * (math.log(4 / 3) - math.log(3 / 4)))
            cw = min(big, max(8, int(round(math.sqrt(s * ar)))))
            ch = min(big, max(8, int(round(math.sqrt(s / ar)))))
            top, left = int(float(r[k, 2]) * (big - ch)), int(float(r[k, 3]) * (big - cw))
            v = F.interpolate(base[:, :, top:top + ch, left:left + cw], size=(self.size, self.size), mode="bilinear")[0]
            views.append(v.flip(-1) if float(r[k, 4]) < 0.5 else v)
        label = int(torch.randint(0, self.c, (1,), generator=g))
        return views, label


class SyntheticImages(torch.utils.data.Dataset):
    """Seeded stand-in for ImageFolder + ViewSpecSampler (--views_on_device): item i is (uint8 image [H,W,3], int32 view
    specs [n_views, 6], label) -- what the loader ships when the views are generated on the GPU."""

    def __init__(self, n_samples, n_views=64, n_classes=1000, seed=0):
        from ttl_b200.views import ViewSpecSampler
        self.n, self.c, self.seed = n_samples, n_classes, seed
        self.sampler = ViewSpecSampler(n_views - 1)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        import numpy as np
        g = np.random.default_rng(self.seed * 1000003 + i)
        h, w = int(g.integers(256, 513)), int(g.integers(256, 513))
        lo = g.integers(0, 256, size=(h // 32 + 2, w // 32 + 2, 3)).astype(np.float32)
        ys, xs = np.arange(h) / 32.0, np.arange(w) / 32.0
        y0, x0 = ys.astype(int), xs.astype(int)
        fy, fx = (ys - y0)[:, None, None], (xs - x0)[None, :, None]
        img = ((lo[y0][:, x0] * (1 - fx) + lo[y0][:, x0 + 1] * fx) * (1 - fy)
               + (lo[y0 + 1][:, x0] * (1 - fx) + lo[y0 + 1][:, x0 + 1] * fx) * fy)
        img = np.clip(img + g.normal(0, 12, size=img.shape), 0, 255).astype(np.uint8)
        st = torch.random.get_rng_state()
        torch.manual_seed(self.seed * 7919 + i)          # the sampler draws from the global torch RNG like torchvision
        arr, specs = self.sampler(img)
        torch.random.set_rng_state(st)
        return torch.from_numpy(arr), torch.from_numpy(specs), int(g.integers(0, self.c))


class _ImageSpecTransform:
    """`transform=` for the reference's build_dataset when --views_on_device is set: PIL image -> (uint8, specs)."""

    def __init__(self, n_views):
        from ttl_b200.views import ViewSpecSampler
        self.sampler = ViewSpecSampler(n_views)

    def __call__(self, img):
        arr, specs = self.sampler(img)
        return torch.from_numpy(arr), torch.from_numpy(specs)


def _collate_samples(items):
    """DataLoader collate for the fused route: keep the S dataset items of a batch as a list (images of a batch differ in size)."""
    return list(items)


def _iter_samples(val_loader):
    """(payload, target[1]) per test sample from either loader flavour: the reference's (batch_size=1, default collate: a list of
    64 tensors [1,3,S,S] + target [1], ttl.py:274-279,322-336) or this file's (batch_size=S, `_collate_samples`).
    payload = fp32 views [V,3,S,S]  |  (uint8 image [H,W,3], int32 view specs [V,6])."""
    for item in val_loader:
        for it in (item if isinstance(item, list) and item and isinstance(item[0], tuple) else [item]):
            if len(it) == 3:                                   # SyntheticImages: image, specs, label
                images, target = (it[0], it[1]), it[2]
            else:
                images, target = it
            if isinstance(images, (list, tuple)) and len(images) == 2 and torch.is_tensor(images[0]) and images[0].dtype == torch.uint8:
                im, sp = images                                # build_dataset(transform=_ImageSpecTransform)
                images = (im[0] if im.dim() == 4 else im, sp[0] if sp.dim() == 3 else sp)
            elif isinstance(images, (list, tuple)):
                images = torch.cat([v if v.dim() == 4 else v[None] for v in images], dim=0)
            elif images.dim() > 4:
                images = images.squeeze(0)
            yield images, torch.as_tensor(target).view(-1)[:1]


class _Ready:
    """Device-resident results behind the same .wait() as ttl_b200.engine.Pending."""

    def __init__(self, outs):
        self.outs = outs

    def wait(self):
        return {k: v.cpu() for k, v in self.outs.items()}


@torch.enable_grad()
def test_time_adapt_eval(val_loader, model, model_state, optimizer, optim_state, scaler, args):
    """ttl.py:300-363.  With default flags S test samples are ONE fused library call (each: reset -> adapt -> predict,
    ttl.py:338-352) submitted without synchronising: the host->device copy and the view generation of batch i+1 overlap the
    kernels of batch i and the predictions of batch i are scored while batch i+1 runs (depth-2 pipeline, the role of the
    reference's pin_memory + non_blocking loader).  `--compat` keeps the reference's explicit per-sample sequence
    LoRA_reset / load_state_dict / test_time_tuning / model(image)."""
    from collections import deque
    batch_time = AverageMeter('Time', ':6.3f')
    top1 = AverageMeter('Acc@1', ':6.2f')
    top5 = AverageMeter('Acc@5', ':6.2f')
    model.eval()
    with torch.no_grad():
        model.LoRA_reset()
    fused = (not getattr(args, "compat", False)) and model.fast_path_ok(args)
    rank, world = getattr(args, "rank_id", 0), getattr(args, "world_size", 1)
    tally = [0, 0, 0]                       # top-1 hits, top-5 hits, n  (utils/tools.py:40-44 semantics: sums, not averages)
    S = max(1, int(getattr(args, "concurrent_samples", 1))) if fused else 1
    n_total = len(val_loader.dataset) if hasattr(val_loader, "dataset") else None
    t_start = end = time.time()
    done_at = []                            # (wall time, samples scored so far) after every scored batch

    def score(output, target):
        output, target = output.float().cpu(), target.cpu()
        acc1, acc5 = accuracy(output, target, topk=(1, 5))
        n = target.numel()
        top1.update(float(acc1[0]), n)
        top5.update(float(acc5[0]), n)
        tally[0] += int(round(float(acc1[0]) * n / 100.0))
        tally[1] += int(round(float(acc5[0]) * n / 100.0))
        tally[2] += n

    inflight = deque()
    staging = {}                            # two pinned fp32 staging batches for host-resident views

    def retire(k):
        nonlocal end
        while len(inflight) > k:
            pending, targets = inflight.popleft()
            score(pending.wait()["pred_logits"], torch.cat(targets))
            now = time.time()
            done_at.append((now, tally[2]))
            batch_time.update((now - end) / len(targets), len(targets))
            end = now
            if rank == 0 and (tally[2] // args.print_freq) != ((tally[2] - len(targets)) // args.print_freq):
                print(f"Test: [{tally[2]}/{n_total if n_total is not None else '?'}]\t{batch_time}\t{top1}\t{top5}")

    def submit(payloads, targets, seq):
        if isinstance(payloads[0], tuple):      # uint8 image + view specs per sample: views are generated on the device
            pend = model.adapt_and_predict_images([im.numpy() for im, _ in payloads], [sp.numpy() for _, sp in payloads],
                                                  args, sync=False)
        elif payloads[0].is_cuda or model.deyo_general(args) or model.lora_encoder == "text":
            # the optional DeYO branches build x' from device-resident views; the text route takes image features view by view
            pend = _Ready(model.adapt_and_predict_batch(torch.stack(payloads).to(model.device, non_blocking=True), args))
        else:                                   # fp32 views from the loader: stage in pinned memory, copy asynchronously
            shape = (S,) + tuple(payloads[0].shape)
            if staging.get("shape") != shape:
                staging["shape"] = shape
                staging["buf"] = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(2)]
            buf = staging["buf"][seq % 2]
            for j, pl in enumerate(payloads):
                buf[j].copy_(pl)
            pend = model.adapt_and_predict_batch(buf[:len(payloads)], args, sync=False)
        inflight.append((pend, targets))
        retire(1)                               # keep one batch in flight behind the one just submitted

    pend_p, pend_t, seq = [], [], 0
    for images, target in _iter_samples(val_loader):
        if fused:
            pend_p.append(images)
            pend_t.append(target)
            if len(pend_p) == S:
                submit(pend_p, pend_t, seq)
                pend_p, pend_t, seq = [], [], seq + 1
        else:
            if isinstance(images, tuple):
                raise NotImplementedError("device-generated views need the fused path (default flags, no --compat): "
                                          "pass --views_on_host")
            images = images.to(model.device, non_blocking=True)
            image = images[:1]
            if args.tta_steps > 0:
                with torch.no_grad():
                    model.LoRA_reset()
            optimizer.load_state_dict(optim_state)
            test_time_tuning(model, images, optimizer, scaler, args)
            with torch.no_grad():
                output = model(image)
            inflight.append((_Ready({"pred_logits": output}), [target]))
            retire(0)
    if pend_p:
        submit(pend_p, pend_t, seq)
    retire(0)
    counts = torch.tensor(tally, dtype=torch.int64, device=model.device)
    tot = tdist.reduce_counts(counts, world)   # the only collective of the path: 3 int64 (utils/tools.py:40-44 semantics)
    n = max(tot[2], 1)
    # throughput of this rank's loop: whole loop, and steady state (after the first two batches: eager pass + graph capture)
    elapsed = max(time.time() - t_start, 1e-9)
    stats = {"samples": tally[2], "seconds": elapsed, "samples_per_s": tally[2] / elapsed, "steady_samples_per_s": None,
             "concurrent_samples": S, "fused": fused}
    if len(done_at) > 3:
        (t0, n0), (t1, n1) = done_at[1], done_at[-1]
        stats["steady_samples_per_s"] = (n1 - n0) / max(t1 - t0, 1e-9)
    test_time_adapt_eval.last_stats = stats
    if rank == 0:
        steady = stats["steady_samples_per_s"]
        print(f" *  Acc@1 {100.0 * tot[0] / n:.3f} Acc@5 {100.0 * tot[1] / n:.3f}  (n={tot[2]}, {world} rank(s))")
        print(f" *  {stats['samples_per_s']:.1f} adapted samples/s on this rank over the whole loop"
              + (f", {steady:.1f} in steady state (after the first two batches)" if steady else ""))
    return [100.0 * tot[0] / n, 100.0 * tot[1] / n]


def _build_dataset(set_id, transform, args):
    """The reference's data package when it is importable (same tree and class tables), else ttl_b200/datasets.py."""
    try:
        from data.datautils import build_dataset
    except ImportError:
        from ttl_b200.datasets import build_dataset
    return build_dataset(set_id=set_id, transform=transform, args=args)


def _classnames_for(set_id, args, dataset=None):
    try:   # class lists live in the reference's data/ package (inputs to the path, not vendored here)
        from data.imagnet_prompts import imagenet_classes
        from data.imagenet_variants import imagenet_a_mask, imagenet_r_mask, imagenet_v_mask  # noqa: F401
        import data.cls_to_names as c2n
        if len(set_id) > 1:
            return getattr(c2n, f"{set_id.lower()}_classes")
        if set_id in ('A', 'V'):
            mask = {'A': imagenet_a_mask, 'V': imagenet_v_mask}[set_id]
            return [imagenet_classes[i] for i in mask]
        if set_id == 'R':
            return [imagenet_classes[i] for i, m in enumerate(imagenet_r_mask) if m]
        return imagenet_classes
    except Exception:
        if dataset is not None and hasattr(dataset, "classes"):    # folder-per-class test set: names from its own folders
            from ttl_b200.datasets import classnames_for_folders
            return classnames_for_folders(dataset.root, dataset.classes)
        n = {'A': 200, 'R': 200}.get(set_id, 1000)
        return [f"class {i}" for i in range(n)]


def main_worker(gpu, args):
    """ttl.py:121-297."""
    args.gpu = gpu
    torch.manual_seed(args.seed)
    if args.cocoop:
        raise NotImplementedError("--cocoop is outside the TTL path")
    if args.lora_encoder == 'prompt':
        raise NotImplementedError("--lora_encoder prompt does not run in the reference either (clip/custom_clip.py:680); "
                                  "the B200 path implements image and text")
    from clip.custom_clip import get_coop
    rank, world = args.rank_id, args.world_size
    first = args.test_sets.split("/")[0]
    extra = {"precision": args.precision}
    if args.precision == "fp32":
        args.concurrent_samples = 1
    # Input route.  Default: the host ships the decoded uint8 image + the drawn crop boxes and the 64 views are generated on
    # the GPU, bit-exactly as the reference's PIL/torchvision pipeline would (csrc/views.cu).  --views_on_host, --compat and
    # --precision fp32 take the reference's route: 64 fp32 views per sample from the DataLoader workers (ttl.py:232-241).
    args.views_on_device = not (args.views_on_host or args.compat or args.precision == "fp32")
    if args.lora_encoder == 'text':
        args.views_on_device = False      # the image tower only supplies frozen features here: it takes the fp32 views
    if args.deyo_selection and (args.filter_ent or args.filter_plpd or args.reweight_plpd or args.reweight_ent != 1):
        args.views_on_device = False      # filter_plpd destroys the structure of the fp32 views (deyo.py:116-136): they must exist
    extra["allow_synthetic"] = args.synthetic > 0 or args.random_init
    if args.vision_checkpoint:      # HF safetensors/bin or OpenAI-format .pt (ttl_b200/weights.py); default: local HF cache
        # one file = what CLIPModel.from_pretrained yields (clip/custom_clip.py:581,619): image tower, the text tower that
        # builds the class features, logit_scale
        from ttl_b200.weights import load_clip_checkpoint
        ck = load_clip_checkpoint(args.vision_checkpoint)
        extra["weights"] = ck.vision
        if ck.text is not None:
            extra["text_weights"] = ck.text
        elif not extra["allow_synthetic"]:
            raise RuntimeError(f"{args.vision_checkpoint} has no text tower: class features cannot be built "
                               "(give a full CLIP checkpoint, or --random_init for random class features)")
        if ck.logit_scale is not None:
            extra["logit_scale"] = ck.logit_scale
        bpe = args.bpe_path or ck.bpe_path
        if bpe:
            extra["bpe_path"] = bpe
    elif args.bpe_path:
        extra["bpe_path"] = args.bpe_path
    model = get_coop(args.arch, args.test_sets, args.gpu, args.n_ctx, args.ctx_init, layer_range=args.layer_range,
                     init_method=args.init_method, lora_encoder=args.lora_encoder, rank=args.rank,
                     classnames=_classnames_for(first, args), max_views=args.batch_size,
                     max_samples=max(1, args.concurrent_samples), **extra)
    # requires-grad filter by parameter NAME, exactly the reference's rule (ttl.py:151-163)
    enc_name = 'text_encoder' if args.lora_encoder == 'text' else 'image_encoder'      # ttl.py:145-149
    for name, param in model.named_parameters():
        ok = (enc_name in name and ("lora_A" in name or "lora_B" in name)
              and any(f"layers.{i}." in name for i in range(args.layer_range[0], args.layer_range[1] + 1)))
        param.requires_grad_(ok)
    # optimizer groups walked like ttl.py:189-218
    groups = []
    lora_layers = (model.text_encoder.text_model.encoder.layers if args.lora_encoder == 'text'
                   else model.image_encoder.vision_model.encoder.layers)                 # ttl.py:190-193
    for i, layer in enumerate(lora_layers):
        if args.layer_range[0] <= i <= args.layer_range[1]:
            groups.extend([{'params': layer.self_attn.q_proj.lora_A.parameters()},
                           {'params': layer.self_attn.q_proj.lora_B.parameters()},
                           {'params': layer.self_attn.v_proj.lora_A.parameters()},
                           {'params': layer.self_attn.v_proj.lora_B.parameters()}])
    optimizer = torch.optim.AdamW(groups, lr=args.lr)
    optim_state = deepcopy(optimizer.state_dict())
    scaler = torch.amp.GradScaler("cuda", init_scale=1000, enabled=False)   # bf16 path: no loss scaling
    results = {}
    fused = (not args.compat) and model.fast_path_ok(args)
    args.views_on_device = args.views_on_device and fused
    for set_id in args.test_sets.split("/"):
        ds = None
        if args.synthetic <= 0 and args.views_on_device:
            ds = _build_dataset(set_id, _ImageSpecTransform(args.batch_size - 1), args)
        elif args.synthetic <= 0:
            # [clean view] + (batch_size - 1) x (RandomResizedCrop + flip), normalised: data/datautils.py:98-157, ttl.py:226-241
            from ttl_b200.datasets import default_host_views
            ds = _build_dataset(set_id, default_host_views(args.batch_size - 1, args.resolution), args)
        classnames = _classnames_for(set_id, args, ds)
        if ds is not None and hasattr(ds, "classes") and len(ds.classes) != len(classnames):
            raise RuntimeError(f"test set {set_id}: {len(ds.classes)} class folders but {len(classnames)} class names")
        model.reset_classnames(classnames, args.arch)
        if args.synthetic > 0 and args.views_on_device:
            ds = SyntheticImages(args.synthetic, args.batch_size, len(classnames), seed=args.seed)
        elif args.synthetic > 0:
            ds = SyntheticViews(args.synthetic, args.batch_size, args.resolution, len(classnames), seed=args.seed)
        # sample-sharding: rank r evaluates samples r, r+W, ... of the seeded order
        g = torch.Generator().manual_seed(args.seed)
        order = torch.randperm(len(ds), generator=g).tolist()
        shard = [order[j] for j in tdist.shard_indices(len(order), rank, world)]
        if fused:   # S samples per item, kept as a list (image sizes differ); pinning happens in the staging buffers
            loader = torch.utils.data.DataLoader(torch.utils.data.Subset(ds, shard), batch_size=max(1, args.concurrent_samples),
                                                 shuffle=False, num_workers=args.workers, collate_fn=_collate_samples,
                                                 persistent_workers=False, prefetch_factor=4 if args.workers > 0 else None)
        else:
            loader = torch.utils.data.DataLoader(torch.utils.data.Subset(ds, shard), batch_size=1, shuffle=False,
                                                 num_workers=args.workers, pin_memory=True)
        t0 = time.time()
        results[set_id] = test_time_adapt_eval(loader, model, None, optimizer, optim_state, scaler, args)
        if rank == 0:
            print("=> Acc. on testset [{}]: @1 {}/ @5 {}   ({:.1f} s)".format(set_id, results[set_id][0], results[set_id][1],
                                                                               time.time() - t0))
    if rank == 0:
        print("======== Result Summary ========")
        print("params: nstep\tlr\tbs")
        print("params: {}\t{}\t{}".format(args.tta_steps, args.lr, args.batch_size))
        for k in results:
            print(k, end="\t")
        print()
        for k in results:
            print("{:.2f}".format(results[k][0]), end="\t")
        print()
    return results


def build_parser():
    """Flags of ttl.py:367-424 with identical names, defaults and types (including the always-true --tpt and the untyped
    --deyo_selection, for which any non-empty string is truthy)."""
    p = argparse.ArgumentParser(description='Test-time Prompt Tuning')
    p.add_argument('data', metavar='DIR', nargs="?", default='/home/raza.imam/Documents/TPT/datasets', help='path to dataset root')
    p.add_argument('--test_sets', type=str, default='A')
    p.add_argument('--dataset_mode', type=str, default='test')
    p.add_argument('-a', '--arch', metavar='ARCH', default='ViT-B/16')
    p.add_argument('--resolution', default=224, type=int)
    p.add_argument('-j', '--workers', default=4, type=int, metavar='N')
    p.add_argument('-b', '--batch-size', default=64, type=int, metavar='N')
    p.add_argument('--lr', '--learning-rate', default=5e-3, type=float, metavar='LR', dest='lr')
    p.add_argument('-p', '--print_freq', default=10, type=int, metavar='N')
    p.add_argument('--gpu', default=1, type=int)
    p.add_argument('--tpt', action='store_true', default=True)
    p.add_argument('--selection_p', default=0.1, type=float)
    p.add_argument('--tta_steps', default=1, type=int)
    p.add_argument('--n_ctx', default=4, type=int)
    p.add_argument('--ctx_init', default='a_photo_of_a', type=str)
    p.add_argument('--cocoop', action='store_true', default=False)
    p.add_argument('--load', default=None, type=str)
    p.add_argument('--seed', type=int, default=0)
    p.add_argument('--images_per_class', default=None, type=int)
    p.add_argument('--layer_range', type=list_of_ints, default=(9, 11))
    p.add_argument('--init_method', default='xavier', choices=['xavier', 'gaussian', 'kaiming', 'pretrained', None])
    p.add_argument('--lora_encoder', default='image', choices=['text', 'image', 'prompt'])
    p.add_argument('--rank', default=16, type=int)
    p.add_argument('--deyo_selection', default=True)
    p.add_argument('--aug_type', default='patch', type=str)
    p.add_argument('--occlusion_size', default=112, type=int)
    p.add_argument('--patch_len', default=6, type=int)
    p.add_argument('--row_start', default=56, type=int)
    p.add_argument('--column_start', default=56, type=int)
    p.add_argument('--deyo_margin', default=0.5, type=float)
    p.add_argument('--deyo_margin_e0', default=0.4, type=float)
    p.add_argument('--plpd_threshold', default=0.2, type=float)
    p.add_argument('--fishers', default=0, type=int)
    p.add_argument('--filter_ent', default=0, type=int)
    p.add_argument('--filter_plpd', default=0, type=int)
    p.add_argument('--reweight_ent', default=1, type=int)
    p.add_argument('--reweight_plpd', default=0, type=int)
    # additions of this implementation (none starts with --data / --b)
    p.add_argument('--synthetic', default=0, type=int, help='evaluate on N seeded synthetic samples')
    p.add_argument('--compat', action='store_true', default=False, help='autograd + torch.optim.AdamW control flow')
    p.add_argument('--views_on_host', action='store_true', default=False,
                   help="the reference's input route: 64 fp32 views per sample from the DataLoader workers, copied host->device")
    p.add_argument('--views_on_device', action='store_true', default=False,
                   help='(default on the fused path) generate the views on the GPU from the uint8 image, bit-exact with '
                        'PIL/torchvision; kept as an explicit flag for older command lines')
    p.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'],
                   help='bf16 = tensor-core path (default); fp32 = validation mode (fp32 everywhere, samples one at a time)')
    p.add_argument('--vision_checkpoint', default=None, type=str,
                   help='CLIP checkpoint file for the image tower: HF model.safetensors / pytorch_model.bin or OpenAI ViT-*.pt')
    p.add_argument('--random_init', action='store_true', default=False,
                   help='allow seeded random-init towers / random class features when no checkpoint is available '
                        '(implied by --synthetic); without it a missing checkpoint is an error, as in the reference')
    p.add_argument('--merges', dest='bpe_path', default=None, type=str,
                   help='CLIP BPE merge table (bpe_simple_vocab_16e6.txt.gz or HF merges.txt); default: next to the '
                        'checkpoint, else $TTL_BPE_PATH')
    p.add_argument('--concurrent_samples', default=9, type=int,
                   help='test samples adapted concurrently per library call (each keeps its own adapter/optimiser state)')
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.world_size, args.rank_id = world, int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        args.gpu = int(os.environ.get("LOCAL_RANK", "0"))     # one process per GPU
        torch.cuda.set_device(args.gpu)
        dist.init_process_group("nccl", device_id=torch.device("cuda", args.gpu))
    elif args.gpu >= torch.cuda.device_count():
        args.gpu = 0    # the reference defaults to --gpu 1 (ttl.py:375); fall back to the only device
    try:
        return main_worker(args.gpu, args)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
