"""Self-contained input side of the path for real images: test-set folders and the host view transform.

The bundled `ttl.py` prefers the reference's own `data` package when it is importable (same tree, same class-name tables);
this module is what runs when it is not, so that `python ttl.py /path/to/datasets --test_sets A` needs nothing but
torchvision.  Reference behaviour it restates (no code shared):

* `data/datautils.py:20-36,38-72` -- a set id names a folder under the data root that is read with `ImageFolder`
  (`I` additionally descends into `val`); `--images_per_class n` keeps the first n files of every class folder.
* `data/datautils.py:98-157` -- one test image becomes `[clean view] + n_views x (RandomResizedCrop(224) +
  RandomHorizontalFlip)` of the *original* image, each normalised; the AugMix operator list is empty there.

The few-shot sets of the reference read JSON split files (`data/fewshot_datasets.py`); here they are accepted only when their
folder is laid out as class sub-folders, and say so otherwise.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence

import torch

# set id -> folder below the data root (the directory convention of data/datautils.py:20-36)
SET_DIRS = {
    "I": os.path.join("ImageNet", "val"),
    "A": os.path.join("imagenet-adversarial", "imagenet-a"),
    "K": "ImageNet-Sketch",
    "R": os.path.join("imagenet-rendition", "imagenet-r"),
    "V": os.path.join("imagenetv2", "imagenetv2-matched-frequency-format-val"),
    "flower102": "oxford_flowers",
    "dtd": "dtd",
    "pets": os.path.join("oxford_pets", "images"),
    "cars": "stanford_cars",
    "ucf101": os.path.join("ucf101", "UCF-101-midframes"),
    "caltech101": os.path.join("caltech-101", "101_ObjectCategories"),
    "food101": "food-101",
    "sun397": os.path.join("sun397", "SUN397"),
    "aircraft": "fgvc_aircraft",
    "eurosat": os.path.join("eurosat", "2750"),
}

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class HostViews:
    """`transform=` producing the list of fp32 views the host-input entry points take: element 0 is
    `preprocess(base_transform(img))`, elements 1..n_views are `preprocess(flip?(random_resized_crop(img)))`.  The torch RNG
    is consumed per view as torchvision's RandomResizedCrop then RandomHorizontalFlip consume it, i.e. exactly as
    `ttl_b200.views.ViewSpecSampler` draws its boxes, so both input routes see the same views under the same seed."""

    def __init__(self, base_transform: Callable, preprocess: Callable, n_views: int = 63, size: int = 224):
        import torchvision.transforms as T
        self.base_transform, self.preprocess, self.n_views = base_transform, preprocess, n_views
        self.draw = T.Compose([T.RandomResizedCrop(size), T.RandomHorizontalFlip()])

    def __call__(self, img) -> List[torch.Tensor]:
        clean = self.preprocess(self.base_transform(img))
        return [clean] + [self.preprocess(self.draw(img)) for _ in range(self.n_views)]


def default_host_views(n_views: int = 63, resolution: int = 224) -> HostViews:
    """The transform `ttl.py:226-241` builds: Resize(BICUBIC) + CenterCrop for the clean view, ToTensor + CLIP normalisation."""
    import torchvision.transforms as T
    base = T.Compose([T.Resize(resolution, interpolation=T.InterpolationMode.BICUBIC, antialias=True),
                      T.CenterCrop(resolution)])
    return HostViews(base, T.Compose([T.ToTensor(), T.Normalize(CLIP_MEAN, CLIP_STD)]), n_views, resolution)


def set_directory(set_id: str, data_root: str) -> str:
    key = set_id if set_id in SET_DIRS else set_id.lower()
    if key not in SET_DIRS:
        raise NotImplementedError(f"unknown test set {set_id!r}; known: {sorted(SET_DIRS)}")
    return os.path.join(data_root, SET_DIRS[key])


def build_dataset(set_id: str, transform: Callable, args, **_ignored):
    """Same call as the reference's `build_dataset(set_id=, transform=, args=)`; reads `args.data` and
    `args.images_per_class`."""
    from torchvision.datasets import ImageFolder
    root = set_directory(set_id, args.data)
    if not os.path.isdir(root):
        raise FileNotFoundError(f"test set {set_id!r}: {root} does not exist (data root {args.data!r})")
    try:
        ds = ImageFolder(root, transform=transform)
    except FileNotFoundError as e:
        raise NotImplementedError(f"test set {set_id!r}: {root} is not laid out as one sub-folder per class; the reference reads "
                                  f"this set through JSON split files (data/fewshot_datasets.py), which this build does not") from e
    n = getattr(args, "images_per_class", None)
    if n is not None:
        kept, seen = [], {}
        for path, cls in ds.samples:                      # ImageFolder lists every class folder in sorted file order
            if seen.get(cls, 0) < n:
                kept.append((path, cls))
                seen[cls] = seen.get(cls, 0) + 1
        ds.samples = ds.imgs = kept
        ds.targets = [c for _, c in kept]
    return ds


def classnames_for_folders(root: str, folders: Sequence[str], table: Optional[str] = None) -> List[str]:
    """Prompt class names for a folder-per-class test set without the reference's name tables: a `classnames.txt` /
    `LOC_synset_mapping.txt`-style file (`<folder> <name>[, synonyms]` per line; looked up in `root`, its parents, or given
    as `table`) maps folder ids such as WordNet ids to names; otherwise the folder names themselves are used."""
    cands = [table] if table else []
    d = os.path.abspath(root)
    for _ in range(3):
        cands += [os.path.join(d, "classnames.txt"), os.path.join(d, "LOC_synset_mapping.txt")]
        d = os.path.dirname(d)
    mapping = {}
    for path in cands:
        if path and os.path.isfile(path):
            with open(path) as f:
                for line in f:
                    key, _, name = line.strip().partition(" ")
                    if key and name:
                        mapping[key] = name.split(",")[0].strip()
            break
    return [mapping.get(f, f.replace("_", " ")) for f in folders]
