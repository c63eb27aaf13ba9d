"""Small end-to-end exercise of the ViT-B/16 paths for `compute-sanitizer --tool memcheck`: bf16 path (eager, capture,
replay; 2 concurrent samples; images path), fp32 validation mode, text tower.  Not collected by pytest (no test_ prefix); lives
under tests/ because it borrows the text oracle's synthetic weight generator (oracle/ is test infrastructure only):
    compute-sanitizer --tool memcheck python tests/sanitizer_driver_b16.py"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ttl-test-time-low-rank-adaptation_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from ttl_b200 import Engine, Hparams  # noqa: E402
from ttl_b200.synthetic import synthetic_vit_weights, synthetic_lora_init, synthetic_text_features  # noqa: E402
from ttl_b200.views import ViewSpecSampler  # noqa: E402

V = 22          # 2 samples x 22 views x 197 tokens = 8668 rows: the zigzag walk of the frozen pass is active (M >= 8192)
w = synthetic_vit_weights("ViT-B/16", seed=1234)
text = synthetic_text_features(100, 512, seed=11)
lora = synthetic_lora_init("ViT-B/16", rank=16, layers=(9, 11), seed=0)
for prec, S in (("bf16", 2), ("fp32", 1)):
    eng = Engine("ViT-B/16", max_views=V, max_classes=128, max_samples=S, precision=prec)
    eng.load_weights(w)
    eng.set_text_features(text, math.log(100.0))
    eng.set_lora_init(lora)
    for head in ("tpt", "deyo"):
        hp = Hparams(head=head, selection_p=0.25, tta_steps=2 if head == "tpt" else 1)
        x = torch.randn(S, V, 3, 224, 224, device="cuda")
        for _ in range(3 if prec == "bf16" else 1):
            out = eng.adapt_predict_batch(x, hp, want=("pred_logits", "idx", "loss"))
        print(prec, head, out["pred_logits"].float().abs().mean().item())
    if prec == "bf16":
        rng = np.random.default_rng(0)
        imgs = [rng.integers(0, 256, size=(200 + 50 * i, 300, 3), dtype=np.uint8) for i in range(S)]
        torch.manual_seed(0)
        specs = [ViewSpecSampler(V - 1)(im)[1] for im in imgs]
        print("images", eng.adapt_predict_images(imgs, specs, Hparams(selection_p=0.25))["pred_logits"].abs().mean().item())
        # optional DeYO branches (deyo.cu): filter_ent + PLPD with the three structure-destroying transforms
        x = torch.randn(S, V, 3, 224, 224, device="cuda")
        for aug in ("occ", "patch", "pixel"):
            o = eng.adapt_predict_batch_deyo(x, Hparams(head="deyo", selection_p=0.25), filter_ent=1, filter_plpd=1, plpd_threshold=-1.0,
                                             aug_type=aug)
            print("deyo", aug, o["pred_logits"].abs().mean().item(), eng.deyo_last_plpd()[1])
    eng.close()
# adapter on the text tower (text-mode context: causal attention fwd/bwd, EOT pooling, exact-Delta kernel, swapped head)
from ttl_b200.synthetic import synthetic_text_weights, HashTokenizer  # noqa: E402
ev = Engine("ViT-B/16", max_views=V, max_classes=16)
ev.load_weights(w)
ev.set_lora_init(lora)
et = Engine("ViT-B/16", max_views=16, max_classes=V, text_mode=True)
et.load_text_weights(synthetic_text_weights("ViT-B/16"))
g = torch.Generator().manual_seed(0)
et.set_lora_init({i: [torch.randn(16, 512, generator=g) * 0.05, torch.zeros(512, 16), torch.randn(16, 512, generator=g) * 0.05,
                      torch.zeros(512, 16)] for i in (9, 10, 11)})
et.set_prompts(HashTokenizer()([f"a photo of a thing {i}." for i in range(12)]), math.log(100.0))
feats = ev.image_features(torch.randn(V, 3, 224, 224, device="cuda"))
for head in ("tpt", "deyo"):
    o = et.adapt_predict_text(feats, Hparams(head=head, selection_p=0.25, tta_steps=2 if head == "tpt" else 1), want=("pred_logits", "loss"))
    print("text-tower adapter", head, o["pred_logits"].abs().mean().item(), float(o["loss"]))
ev.close()
et.close()
from ttl_b200.text import TextEncoder  # noqa: E402
sys.path.insert(0, ROOT)
from oracle import text_oracle as TO  # noqa: E402
a = TO.TEXT_ARCHS["tiny"]
enc = TextEncoder("tiny", max_prompts=4)
enc.load_weights(TO.make_synthetic_text_weights(a, 5))
print("text", enc.encode(TO.make_synthetic_tokens(6, a, 3)).shape)
enc.close()
torch.cuda.synchronize()
print("done")
