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
  --views_on_host keep the synthetic views in pinned host memory (exercises the H2D path)
  --concurrent_samples S  adapt S test samples per library call (default 9; 1 = strictly one at a time)
  --precision fp32  validation mode: every activation/contraction in fp32 (held to 1e-4 against the reference), one sample per call
  --vision_checkpoint F  load the image tower from a checkpoint file (HF or OpenAI format) instead of the local HF cache
  --views_on_device  ship the decoded uint8 image + the drawn crop boxes and generate the 64 views on the GPU
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


@torch.enable_grad()
def test_time_adapt_eval(val_loader, model, model_state, optimizer, optim_state, scaler, args):
    """ttl.py:300-363.  With default flags each sample is ONE fused library call (reset -> adapt -> predict); `--compat`
    keeps the reference's explicit sequence LoRA_reset / load_state_dict / test_time_tuning / model(image)."""
    batch_time = AverageMeter('Time', ':6.3f')
    top1 = AverageMeter('Acc@1', ':6.2f')
    top5 = AverageMeter('Acc@5', ':6.2f')
    model.eval()
    with torch.no_grad():
        model.LoRA_reset()
    fused = (not getattr(args, "compat", False)) and model.fast_path_ok(args)
    rank, world = getattr(args, "rank_id", 0), getattr(args, "world_size", 1)
    counts = torch.zeros(3, dtype=torch.int64, device=model.device)
    end = time.time()
    S = max(1, int(getattr(args, "concurrent_samples", 1))) if fused else 1
    pend_imgs, pend_tgt, seen = [], [], 0

    def score(output, target):
        acc1, acc5 = accuracy(output, target, topk=(1, 5))
        n = target.numel()
        top1.update(float(acc1[0]), n)
        top5.update(float(acc5[0]), n)
        k1 = torch.round(acc1[0] * n / 100.0).long()
        k5 = torch.round(acc5[0] * n / 100.0).long()
        counts.add_(torch.stack([k1, k5, torch.full_like(k1, n)]).to(counts.device))

    def flush():
        """one fused library call for the pending samples (each: reset -> adapt -> predict, ttl.py:338-352)"""
        nonlocal pend_imgs, pend_tgt
        if not pend_imgs:
            return
        if isinstance(pend_imgs[0], tuple):     # --views_on_device: (uint8 image, view specs) per sample
            out = model.adapt_and_predict_images([im.numpy() for im, _ in pend_imgs], [sp.numpy() for _, sp in pend_imgs],
                                                 args)["pred_logits"].to(model.device)
            score(out, torch.cat(pend_tgt))
            pend_imgs, pend_tgt = [], []
            return
        batch = torch.stack(pend_imgs)
        if not batch.is_cuda and not getattr(args, "views_on_host", False):
            batch = batch.to(model.device, non_blocking=True)
        out = model.adapt_and_predict_batch(batch, args)["pred_logits"].to(model.device)
        score(out, torch.cat(pend_tgt))
        pend_imgs, pend_tgt = [], []

    for i, item in enumerate(val_loader):
        if len(item) == 3:                      # --views_on_device: uint8 image [1,H,W,3], specs [1,V,6], label
            if not fused:
                raise NotImplementedError("--views_on_device needs the fused path (default flags, no --compat)")
            images, target = (item[0][0], item[1][0]), item[2]
        else:
            images, target = item
            if isinstance(images, (list, tuple)) and len(images) == 2 and images[0].dtype == torch.uint8:
                images = (images[0][0], images[1][0])   # build_dataset(transform=_ImageSpecTransform) item
        if isinstance(images, tuple):
            pass
        elif isinstance(images, list):
            images = torch.cat([im if im.dim() == 4 else im[None] for im in images], dim=0)
        elif images.dim() > 4:
            images = images.squeeze(0)
        target = torch.as_tensor(target).view(-1)[:1].to(model.device)
        if fused:
            pend_imgs.append(images)
            pend_tgt.append(target)
            if len(pend_imgs) == S:
                flush()
        else:
            images = images.to(model.device, non_blocking=True)
            image = images[:1]
            if args.tta_steps > 0:
                with torch.no_grad():
                    model.LoRA_reset()
            optimizer.load_state_dict(optim_state)
            test_time_tuning(model, images, optimizer, scaler, args)
            with torch.no_grad():
                output = model(image)
            score(output, target)
        seen += 1
        batch_time.update(time.time() - end)
        end = time.time()
        if (i + 1) % args.print_freq == 0 and rank == 0:
            print(f"Test: [{i + 1}/{len(val_loader)}]\t{batch_time}\t{top1}\t{top5}")
    flush()
    tot = tdist.reduce_counts(counts, world)   # the only collective of the path: 3 int64 (utils/tools.py:40-44 semantics)
    n = max(tot[2], 1)
    if rank == 0:
        print(f" *  Acc@1 {100.0 * tot[0] / n:.3f} Acc@5 {100.0 * tot[1] / n:.3f}  (n={tot[2]}, {world} rank(s))")
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
    if args.lora_encoder != 'image':
        raise NotImplementedError("the B200 path implements --lora_encoder image")
    from clip.custom_clip import get_coop
    rank, world = args.rank_id, args.world_size
    first = args.test_sets.split("/")[0]
    extra = {"precision": args.precision}
    if args.precision == "fp32":
        args.concurrent_samples = 1
        if args.views_on_device:
            raise NotImplementedError("--views_on_device feeds the bf16 patch operand; the fp32 validation mode takes fp32 views")
    if args.vision_checkpoint:      # HF safetensors/bin or OpenAI-format .pt (ttl_b200/weights.py); default: local HF cache
        from ttl_b200.weights import load_vision_checkpoint
        extra["weights"] = load_vision_checkpoint(args.vision_checkpoint)
    model = get_coop(args.arch, args.test_sets, args.gpu, args.n_ctx, args.ctx_init, layer_range=args.layer_range,
                     init_method=args.init_method, lora_encoder=args.lora_encoder, rank=args.rank,
                     classnames=_classnames_for(first, args), max_views=args.batch_size,
                     max_samples=max(1, args.concurrent_samples), **extra)
    # requires-grad filter by parameter NAME, exactly the reference's rule (ttl.py:151-163)
    for name, param in model.named_parameters():
        ok = ('image_encoder' in name and ("lora_A" in name or "lora_B" in name)
              and any(f"layers.{i}." in name for i in range(args.layer_range[0], args.layer_range[1] + 1)))
        param.requires_grad_(ok)
    # optimizer groups walked like ttl.py:189-218
    groups = []
    for i, layer in enumerate(model.image_encoder.vision_model.encoder.layers):
        if args.layer_range[0] <= i <= args.layer_range[1]:
            groups.extend([{'params': layer.self_attn.q_proj.lora_A.parameters()},
                           {'params': layer.self_attn.q_proj.lora_B.parameters()},
                           {'params': layer.self_attn.v_proj.lora_A.parameters()},
                           {'params': layer.self_attn.v_proj.lora_B.parameters()}])
    optimizer = torch.optim.AdamW(groups, lr=args.lr)
    optim_state = deepcopy(optimizer.state_dict())
    scaler = torch.amp.GradScaler("cuda", init_scale=1000, enabled=False)   # bf16 path: no loss scaling
    results = {}
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
    p.add_argument('--views_on_host', action='store_true', default=False)
    p.add_argument('--views_on_device', action='store_true', default=False,
                   help='generate the views on the GPU from the uint8 image (bit-exact with PIL/torchvision)')
    p.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'],
                   help='bf16 = tensor-core path (default); fp32 = validation mode (fp32 everywhere, one sample per call)')
    p.add_argument('--vision_checkpoint', default=None, type=str,
                   help='CLIP checkpoint file for the image tower: HF model.safetensors / pytorch_model.bin or OpenAI ViT-*.pt')
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
