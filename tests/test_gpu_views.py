"""GPU view generator (csrc/views.cu) against the CPU oracle of the reference's host pipeline
(oracle/views_oracle.py = Pillow's 8-bit antialiased resampler + torchvision ToTensor/Normalize, pinned to Pillow in
tests/test_views_oracle.py).  Integer/byte work: the bar is BIT-EXACT, including the fp32 normalisation."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import views_oracle as VO  # noqa: E402


def _img(h, w, seed):
    g = np.random.default_rng(seed)
    base = g.integers(0, 256, size=(max(h // 8, 1) + 2, max(w // 8, 1) + 2, 3), dtype=np.uint8)
    ys = (np.arange(h) * base.shape[0] // h)[:, None]
    xs = (np.arange(w) * base.shape[1] // w)[None, :]
    big = base[ys, xs].astype(np.int16)
    return np.clip(big + g.integers(-40, 41, size=big.shape), 0, 255).astype(np.uint8)


def _boxes(h, w, n, seed):
    g = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        ch, cw = int(g.integers(1, h + 1)), int(g.integers(1, w + 1))
        out.append((int(g.integers(0, h - ch + 1)), int(g.integers(0, w - cw + 1)), ch, cw, int(g.integers(0, 2))))
    out[0] = (0, 0, h, w, 1)   # the whole image, flipped
    return out


def _specs(h, w, boxes):
    from ttl_b200 import _lib as L
    s = np.zeros((1 + len(boxes), 6), dtype=np.int32)
    s[0] = (L.VIEW_CLEAN, 0, 0, h, w, 0)
    for k, (i, j, ch, cw, f) in enumerate(boxes):
        s[1 + k] = (L.VIEW_CROP, i, j, ch, cw, f)
    return s


@pytest.fixture(scope="module")
def tiny_engine():
    from ttl_b200 import Engine
    eng = Engine("ViT-B/16", max_views=64, max_classes=16, layer_range=(9, 11), max_samples=3)
    yield eng
    eng.close()


@pytest.mark.parametrize("h,w", [(375, 500), (500, 333), (224, 224), (97, 1011), (640, 480), (31, 47), (1200, 900)])
def test_views_bit_exact_vs_oracle(tiny_engine, h, w):
    img = _img(h, w, 1000 * h + w)
    boxes = _boxes(h, w, 9, h + w)
    ref = VO.make_views(img, boxes)                                   # [10, 3, 224, 224] fp32
    got = tiny_engine.make_views([img], [_specs(h, w, boxes)])[0].cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), float(np.abs(got - ref).max())


def test_views_batch_of_ragged_images(tiny_engine):
    """Several images of different sizes in one call (the concurrent-samples layout): each equals its own oracle."""
    sizes = [(375, 500), (224, 301), (640, 427)]
    imgs = [_img(h, w, 7 + i) for i, (h, w) in enumerate(sizes)]
    boxes = [_boxes(h, w, 5, 11 + i) for i, (h, w) in enumerate(sizes)]
    got = tiny_engine.make_views(imgs, [_specs(h, w, b) for (h, w), b in zip(sizes, boxes)]).cpu().numpy()
    for i in range(3):
        ref = VO.make_views(imgs[i], boxes[i])
        assert np.array_equal(got[i].view(np.uint32), ref.view(np.uint32))
    # the staging buffers alternate between calls: a second call must not disturb the result
    again = tiny_engine.make_views(imgs[:1], [_specs(*sizes[0], boxes[0])]).cpu().numpy()
    assert np.array_equal(again[0], got[0])


def test_views_match_the_reference_transform_stack(tiny_engine):
    """Same torch seed -> the sampler draws what torchvision's RandomResizedCrop/RandomHorizontalFlip draw on the PIL
    image, and the device views equal the reference's AugMixAugmenter output (uint8 stage exact; <= 1 ulp after torch's
    own fp32 normalisation)."""
    Image = pytest.importorskip("PIL.Image")
    T = pytest.importorskip("torchvision.transforms")
    from ttl_b200.views import ViewSpecSampler
    img = _img(375, 500, 3)
    pil = Image.fromarray(img)
    norm = T.Normalize(mean=VO.CLIP_MEAN, std=VO.CLIP_STD)
    base = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224)])   # ttl.py:232-234
    pre = T.Compose([T.ToTensor(), norm])
    aug = T.Compose([T.RandomResizedCrop(224), T.RandomHorizontalFlip()])                            # datautils.py:98-101
    torch.manual_seed(5)
    ref = torch.stack([pre(base(pil))] + [pre(aug(pil)) for _ in range(15)]).numpy()
    torch.manual_seed(5)
    arr, specs = ViewSpecSampler(15)(pil)
    got = tiny_engine.make_views([arr], [specs])[0].cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=0, atol=3e-7)
    u8 = lambda t: np.rint((t * np.asarray(VO.CLIP_STD, np.float32)[None, :, None, None]
                            + np.asarray(VO.CLIP_MEAN, np.float32)[None, :, None, None]) * 255.0)
    assert np.array_equal(u8(got), u8(ref))


def test_views_reject_bad_boxes(tiny_engine):
    img = _img(64, 64, 0)
    bad = _specs(64, 64, [(10, 10, 60, 60, 0)])
    with pytest.raises(RuntimeError):
        tiny_engine.make_views([img], [bad])


def test_adapt_from_images_equals_adapt_from_views():
    """ttl_adapt_predict_images_async (uint8 image -> bf16 patch matrix on the device) gives exactly what the fp32-view
    entry gives on the views the generator itself produces: same patches, same graph."""
    from ttl_b200 import Engine, Hparams
    from ttl_b200.synthetic import synthetic_vit_weights, synthetic_lora_init, synthetic_text_features
    import math
    eng = Engine("ViT-B/16", max_views=16, max_classes=16, layer_range=(9, 11), max_samples=2)
    try:
        eng.load_weights(synthetic_vit_weights("ViT-B/16", seed=1234))
        eng.set_text_features(synthetic_text_features(10, 512, seed=11), math.log(100.0))
        eng.set_lora_init(synthetic_lora_init("ViT-B/16", rank=16, layers=(9, 11), seed=0))
        hp = Hparams(head="tpt", selection_p=0.25)
        sizes = [(300, 400), (256, 256)]
        imgs = [_img(h, w, 21 + i) for i, (h, w) in enumerate(sizes)]
        specs = [_specs(h, w, _boxes(h, w, 15, 31 + i)) for i, (h, w) in enumerate(sizes)]
        want = ("pred_logits", "idx", "entropy")
        for _ in range(3):      # eager, capture, replay
            a = eng.adapt_predict_images(imgs, specs, hp, want=want)
            views = eng.make_views(imgs, specs)
            b = eng.adapt_predict_batch(views, hp, want=want)
            assert torch.equal(a["idx"].cpu(), b["idx"].cpu())
            assert torch.equal(a["entropy"].cpu(), b["entropy"].cpu())
            assert torch.equal(a["pred_logits"].cpu(), b["pred_logits"].cpu())
        p = eng.adapt_predict_images(imgs, specs, hp, want=want, sync=False)
        c = p.wait()
        assert torch.equal(c["pred_logits"], a["pred_logits"])
    finally:
        eng.close()


def test_adapt_from_images_vit_l14_patch_layout():
    """ViT-L/14 (BASELINE config 4 geometry): patch 14 -> 588 patch columns padded to 640 in the GEMM operand; the view
    generator's fused im2col must write the same operand the fp32-view path builds."""
    from ttl_b200 import Engine, Hparams
    from ttl_b200.synthetic import synthetic_vit_weights, synthetic_lora_init, synthetic_text_features
    import math
    eng = Engine("ViT-L/14", max_views=8, max_classes=16, layer_range=(21, 23), max_samples=1)
    try:
        eng.load_weights(synthetic_vit_weights("ViT-L/14", seed=1234))
        eng.set_text_features(synthetic_text_features(10, 768, seed=11), math.log(100.0))
        eng.set_lora_init(synthetic_lora_init("ViT-L/14", rank=16, layers=(21, 23), seed=0))
        hp = Hparams(head="tpt", selection_p=0.25)
        img = _img(333, 500, 77)
        specs = _specs(333, 500, _boxes(333, 500, 7, 5))
        a = eng.adapt_predict_images([img], [specs], hp, want=("pred_logits", "entropy"))
        b = eng.adapt_predict_batch(eng.make_views([img], [specs]), hp, want=("pred_logits", "entropy"))
        assert torch.equal(a["entropy"].cpu(), b["entropy"].cpu())
        assert torch.equal(a["pred_logits"].cpu(), b["pred_logits"].cpu())
    finally:
        eng.close()
