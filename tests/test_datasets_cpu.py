"""Input side for real images (ttl_b200/datasets.py, CPU only): set-id folders, --images_per_class, class names from the
folders, and the host view transform against torchvision's own stack and against the device route's spec sampler."""
import os
import types

import numpy as np
import pytest
import torch

PIL = pytest.importorskip("PIL.Image")
T = pytest.importorskip("torchvision.transforms")


def _fake_tree(root, set_dir, classes=("n01", "n02", "n03"), per_class=3, size=(90, 120)):
    g = np.random.default_rng(0)
    for c in classes:
        d = os.path.join(root, set_dir, c)
        os.makedirs(d)
        for i in range(per_class):
            PIL.fromarray(g.integers(0, 256, size=(*size, 3), dtype=np.uint8)).save(os.path.join(d, f"{c}_{i}.png"))


def test_build_dataset_folders_and_images_per_class(tmp_path):
    from ttl_b200 import datasets as D
    _fake_tree(str(tmp_path), D.SET_DIRS["A"])
    args = types.SimpleNamespace(data=str(tmp_path), images_per_class=None)
    ds = D.build_dataset(set_id="A", transform=None, args=args)
    assert len(ds) == 9 and ds.classes == ["n01", "n02", "n03"]
    args.images_per_class = 2
    ds = D.build_dataset(set_id="A", transform=None, args=args)
    assert len(ds) == 6 and sorted(ds.targets) == [0, 0, 1, 1, 2, 2]
    assert [os.path.basename(p) for p, _ in ds.samples[:2]] == ["n01_0.png", "n01_1.png"]
    with pytest.raises(FileNotFoundError):
        D.build_dataset(set_id="R", transform=None, args=args)          # folder absent
    with pytest.raises(NotImplementedError):
        D.build_dataset(set_id="nope", transform=None, args=args)
    assert D.set_directory("I", "/x").endswith(os.path.join("ImageNet", "val"))     # data/datautils.py:41


def test_classnames_from_folders(tmp_path):
    from ttl_b200 import datasets as D
    root = tmp_path / "set"
    root.mkdir()
    assert D.classnames_for_folders(str(root), ["great_white_shark", "n02"]) == ["great white shark", "n02"]
    (tmp_path / "LOC_synset_mapping.txt").write_text("n01 tench, Tinca tinca\nn02 goldfish, Carassius auratus\n")
    assert D.classnames_for_folders(str(root), ["n01", "n02", "n03"]) == ["tench", "goldfish", "n03"]


def test_host_views_equal_torchvision_stack_and_the_spec_sampler():
    """Same seed -> the host transform returns exactly what torchvision's RandomResizedCrop + RandomHorizontalFlip + ToTensor +
    Normalize return (the reference's get_preaugment + preprocess, data/datautils.py:98-127), and its crops are the boxes the
    device route's ViewSpecSampler draws."""
    from ttl_b200 import datasets as D
    from ttl_b200.views import ViewSpecSampler
    import torchvision.transforms.functional as TF
    g = np.random.default_rng(1)
    img = PIL.fromarray(g.integers(0, 256, size=(300, 260, 3), dtype=np.uint8))
    n = 5
    torch.manual_seed(11)
    views = D.default_host_views(n, 224)(img)
    assert len(views) == n + 1 and all(v.shape == (3, 224, 224) and v.dtype == torch.float32 for v in views)
    pre = T.Compose([T.ToTensor(), T.Normalize(D.CLIP_MEAN, D.CLIP_STD)])
    clean = pre(T.CenterCrop(224)(T.Resize(224, interpolation=T.InterpolationMode.BICUBIC, antialias=True)(img)))
    assert torch.equal(views[0], clean)
    torch.manual_seed(11)
    ref = T.Compose([T.RandomResizedCrop(224), T.RandomHorizontalFlip()])
    for v in views[1:]:
        assert torch.equal(v, pre(ref(img)))
    torch.manual_seed(11)
    _, specs = ViewSpecSampler(n)(img)
    for v, (kind, top, left, h, w, flip) in zip(views[1:], specs[1:].tolist()):
        x = TF.resized_crop(img, top, left, h, w, [224, 224])
        assert torch.equal(v, pre(TF.hflip(x) if flip else x))


def test_cli_loader_falls_back_to_the_bundled_datasets(tmp_path, monkeypatch):
    """ttl.py's dataset hook uses the reference's data package when importable, else ttl_b200.datasets."""
    import sys
    import ttl
    from ttl_b200 import datasets as D
    _fake_tree(str(tmp_path), D.SET_DIRS["K"], classes=("a_b", "c"), per_class=1)
    for m in [k for k in sys.modules if k == "data" or k.startswith("data.")]:
        monkeypatch.delitem(sys.modules, m)
    monkeypatch.setattr(sys, "path", [p for p in sys.path if not p.rstrip("/").endswith("reference")])
    args = types.SimpleNamespace(data=str(tmp_path), images_per_class=None)
    ds = ttl._build_dataset("K", D.default_host_views(2), args)
    views, label = ds[0]
    assert len(views) == 3 and label == 0
    assert ttl._classnames_for("K", args, ds) == ["a b", "c"]
