"""Pins oracle/views_oracle.py (the restatement of Pillow's 8-bit antialiased resampler the reference's view pipeline
runs on the CPU, data/datautils.py:98-157) bit-exactly against Pillow and the torchvision transforms themselves."""
import numpy as np
import pytest
import torch

from oracle import views_oracle as VO

PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402
import torchvision.transforms as T  # noqa: E402
import torchvision.transforms.functional as TF  # noqa: E402


def _img(h, w, seed):
    g = np.random.default_rng(seed)
    base = g.integers(0, 256, size=(h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
    big = np.asarray(Image.fromarray(base).resize((w, h), Image.BICUBIC)).astype(np.int16)
    return np.clip(big + g.integers(-20, 21, size=big.shape), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("h,w", [(375, 500), (500, 333), (224, 224), (97, 1011), (640, 480)])
def test_resize_matches_pillow_bit_exactly(h, w):
    img = _img(h, w, h * 1000 + w)
    pil = Image.fromarray(img)
    g = np.random.default_rng(7)
    for _ in range(6):
        ch, cw = int(g.integers(8, h + 1)), int(g.integers(8, w + 1))
        i, j = int(g.integers(0, h - ch + 1)), int(g.integers(0, w - cw + 1))
        ref = np.asarray(pil.crop((j, i, j + cw, i + ch)).resize((224, 224), Image.BILINEAR))
        got = VO.resize_u8(img[i:i + ch, j:j + cw], 224, 224, VO.BILINEAR)
        assert np.array_equal(got, ref), (ch, cw)
    nh, nw = VO.resized_size(h, w, 224)
    ref = np.asarray(pil.resize((nw, nh), Image.BICUBIC))
    assert np.array_equal(VO.resize_u8(img, nh, nw, VO.BICUBIC), ref)


def test_views_match_the_reference_transform_stack():
    """The same torch RNG stream drives torchvision's RandomResizedCrop/RandomHorizontalFlip on the PIL image and the
    oracle's explicit boxes: all 1 + 7 views equal bit for bit (uint8 stage) and to 1 ulp after normalisation."""
    img = _img(375, 500, 3)
    pil = Image.fromarray(img)
    norm = T.Normalize(mean=VO.CLIP_MEAN, std=VO.CLIP_STD)
    base = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224)])
    pre = T.Compose([T.ToTensor(), norm])
    aug = T.Compose([T.RandomResizedCrop(224), T.RandomHorizontalFlip()])
    torch.manual_seed(11)
    ref = [pre(base(pil))] + [pre(aug(pil)) for _ in range(7)]
    torch.manual_seed(11)
    boxes = []
    for _ in range(7):
        i, j, h, w = T.RandomResizedCrop.get_params(pil, scale=[0.08, 1.0], ratio=[3.0 / 4.0, 4.0 / 3.0])
        boxes.append((i, j, h, w, int(torch.rand(1) < 0.5)))
    got = VO.make_views(img, boxes)
    ref = torch.stack(ref).numpy()
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=0, atol=3e-7)


def test_spec_sampler_consumes_the_rng_like_the_reference_transforms():
    """ttl_b200.views.ViewSpecSampler (host half of the GPU view generator) draws exactly what
    RandomResizedCrop.forward + RandomHorizontalFlip.forward draw from the torch RNG (data/datautils.py:98-101)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "ttl-test-time-low-rank-adaptation_b200"))
    from ttl_b200.views import ViewSpecSampler
    pil = Image.fromarray(_img(333, 500, 9))
    torch.manual_seed(3)
    arr, specs = ViewSpecSampler(12)(pil)
    torch.manual_seed(3)
    ref = []
    for _ in range(12):
        i, j, h, w = T.RandomResizedCrop.get_params(pil, scale=[0.08, 1.0], ratio=[3.0 / 4.0, 4.0 / 3.0])
        ref.append([1, i, j, h, w, int(torch.rand(1) < 0.5)])
    assert arr.shape == (333, 500, 3) and arr.dtype == np.uint8
    assert specs[0].tolist() == [0, 0, 0, 333, 500, 0]
    assert specs[1:].tolist() == ref
