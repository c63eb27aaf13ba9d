"""Per-kernel parity on a real B200, through the C ABI.  Floating-point kernels are compared with a plain PyTorch
fp32 reference of the same op on the same (bf16-rounded) inputs; tolerances are stated per test.  Integer results
(selection indices) must be bit-exact."""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ttl_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def G():
    import gpu_util
    from ttl_b200 import _lib as L
    gpu_util.lib()
    return gpu_util, L


def _bf(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (300, 256, 128, 0), (300, 256, 192, 128), (300, 512, 256, 256),
                                      (1182, 768, 768, 0), (12608, 768, 768, 0), (12608, 2304, 768, 256),
                                      (4000, 3072, 768, 256), (2000, 768, 3072, 128), (197, 768, 768, 64)])
def test_gemm_bias_bf16(G, M, N, K, bn):
    gu, L = G
    a, b = _bf(M, K, seed=1), _bf(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda")
    out, _ = gu.gemm(a, b, L.EPI_BF16, bias=bias, block_n=bn)
    ref = a.float() @ b.float().t() + bias
    # fp32 accumulate, bf16 output rounding: 2^-9 relative per element
    assert gu.rel_err(out, ref) < 4e-3
    assert torch.isfinite(out.float()).all()


@pytest.mark.parametrize("M,N,K,bn", [(12608, 768, 768, 192), (12608, 2304, 768, 256), (1182, 768, 3072, 128),
                                      (12608, 3072, 768, 192), (2100, 768, 256, 256), (300, 384, 128, 128)])
def test_gemm_cta_pair_kernel(G, M, N, K, bn):
    """The cta_group::2 kernel (block_n = 1000 + BLOCK_N): TMA-store epilogues, ragged last 256-row tile."""
    gu, L = G
    a, b = _bf(M, K, seed=11), _bf(N, K, scale=K ** -0.5, seed=12)
    bias = torch.randn(N, device="cuda") * 0.2
    acc = a.float() @ b.float().t()
    out, _ = gu.gemm(a, b, L.EPI_BF16, bias=bias, block_n=1000 + bn)
    assert gu.rel_err(out, acc + bias) < 4e-3 and torch.isfinite(out.float()).all()
    out, z2 = gu.gemm(a, b, L.EPI_GELU, bias=bias, block_n=1000 + bn, want_out2=True)     # + pre-activation copy for the backward
    z = acc + bias
    assert gu.rel_err(out, z * torch.sigmoid(1.702 * z)) < 6e-3
    assert gu.rel_err(z2, z) < 4e-3
    resid = torch.randn(M, N, device="cuda")
    out, _ = gu.gemm(a, b, L.EPI_RESID_F32, bias=bias, resid=resid, block_n=1000 + bn)
    assert gu.rel_err(out, resid + acc + bias) < 1e-5
    out, _ = gu.gemm(a, b, L.EPI_F32, block_n=1000 + bn)
    assert gu.rel_err(out, acc) < 1e-5
    # dQuickGELU epilogue (pre-activations read straight from global memory, 64 bytes per lane and chunk)
    zb = _bf(M, N, seed=13)
    out, _ = gu.gemm(a, b, L.EPI_GELU_BWD, aux=zb, block_n=1000 + bn)
    zf = zb.float()
    sg = torch.sigmoid(1.702 * zf)
    assert gu.rel_err(out, acc * sg * (1 + 1.702 * zf * (1 - sg))) < 6e-3


def test_gemm_cta_pair_lora_pair_and_guard_rows(G):
    gu, L = G
    M, N, K = 2500, 2304, 768
    a, b = _bf(M, K, seed=1), _bf(N, K, scale=K ** -0.5, seed=2)
    a2, b2 = _bf(M, 64, seed=3), _bf(N, 64, scale=0.1, seed=4)
    out, _ = gu.gemm(a, b, L.EPI_BF16, a2=a2, b2=b2, block_n=1256, out_rows=M + 40)
    ref = a.float() @ b.float().t() + a2.float() @ b2.float().t()
    assert gu.rel_err(out[:M], ref) < 4e-3
    assert float(out[M:].float().abs().max()) == 0.0          # TMA clips rows >= M


@pytest.mark.parametrize("M,N,K", [(12608, 2304, 768), (1182, 768, 3072), (2100, 768, 256), (300, 256, 128), (37824, 768, 768)])
def test_gemm_four_cta_cluster_multicast(G, M, N, K):
    """block_n = 2256: two CTA pairs per cluster share every B tile through TMA multicast (quarter tiles, forwarded
    arrivals, slot release counted over both pairs).  Odd numbers of 256-row blocks leave a pair of the last cluster on
    rows >= M (zero-filled loads, clipped stores); the LoRA second operand pair rides along."""
    gu, L = G
    a, b = _bf(M, K, seed=21), _bf(N, K, scale=K ** -0.5, seed=22)
    bias = torch.randn(N, device="cuda") * 0.2
    acc = a.float() @ b.float().t()
    out, _ = gu.gemm(a, b, L.EPI_BF16, bias=bias, block_n=2256, out_rows=M + 8)
    assert gu.rel_err(out[:M], acc + bias) < 4e-3 and float(out[M:].float().abs().max()) == 0.0
    resid = torch.randn(M, N, device="cuda")
    out, _ = gu.gemm(a, b, L.EPI_RESID_F32, bias=bias, resid=resid, block_n=2256)
    assert gu.rel_err(out, resid + acc + bias) < 1e-5
    a2, b2 = _bf(M, 64, seed=3), _bf(N, 64, scale=0.1, seed=4)
    out, _ = gu.gemm(a, b, L.EPI_GELU, bias=bias, a2=a2, b2=b2, block_n=2256)
    z = acc + a2.float() @ b2.float().t() + bias
    assert gu.rel_err(out, z * torch.sigmoid(1.702 * z)) < 6e-3
    pair, _ = gu.gemm(a, b, L.EPI_F32, block_n=1256)
    quad, _ = gu.gemm(a, b, L.EPI_F32, block_n=2256)
    assert torch.equal(pair, quad)      # same MMA order per output tile: bit-identical to the pair kernel


def test_gemm_lora_second_pair(G):
    gu, L = G
    M, N, K = 1182, 2304, 768
    a, b = _bf(M, K, seed=1), _bf(N, K, scale=K ** -0.5, seed=2)
    a2, b2 = _bf(M, 64, seed=3), _bf(N, 64, scale=0.1, seed=4)
    out, _ = gu.gemm(a, b, L.EPI_BF16, a2=a2, b2=b2)
    ref = a.float() @ b.float().t() + a2.float() @ b2.float().t()
    assert gu.rel_err(out, ref) < 4e-3


def test_gemm_epilogues(G):
    gu, L = G
    M, N, K = 700, 768, 256
    a, b = _bf(M, K, seed=5), _bf(N, K, scale=K ** -0.5, seed=6)
    bias = torch.randn(N, device="cuda") * 0.1
    acc = a.float() @ b.float().t()
    # QuickGELU (+ pre-activation copy)
    out, z = gu.gemm(a, b, L.EPI_GELU, bias=bias, want_out2=True)
    zr = acc + bias
    assert gu.rel_err(z, zr) < 4e-3
    assert gu.rel_err(out, zr * torch.sigmoid(1.702 * zr)) < 6e-3
    # residual, fp32 out (+ bf16 copy)
    resid = torch.randn(M, N, device="cuda")
    out, cp = gu.gemm(a, b, L.EPI_RESID_F32, bias=bias, resid=resid, want_out2=True)
    assert gu.rel_err(out, resid + acc + bias) < 1e-5
    assert gu.rel_err(cp, resid + acc + bias) < 4e-3
    # plain fp32
    out, _ = gu.gemm(a, b, L.EPI_F32)
    assert gu.rel_err(out, acc) < 1e-5
    # dQuickGELU
    zb = _bf(M, N, seed=7)
    out, _ = gu.gemm(a, b, L.EPI_GELU_BWD, aux=zb)
    zf = zb.float()
    s = torch.sigmoid(1.702 * zf)
    assert gu.rel_err(out, acc * s * (1 + 1.702 * zf * (1 - s))) < 6e-3


def test_gemm_patch_epilogue(G):
    gu, L = G
    V, T, N, K = 3, 196, 768, 768
    a, b = _bf(V * T, K, seed=8), _bf(N, K, scale=K ** -0.5, seed=9)
    pos = torch.randn(T + 1, N, device="cuda")
    out, _ = gu.gemm(a, b, L.EPI_PATCH_F32, pos=pos, tpv=T, out_rows=V * (T + 1))
    ref = torch.zeros(V, T + 1, N, device="cuda")
    ref[:, 1:] = (a.float() @ b.float().t()).view(V, T, N) + pos[1:]
    assert gu.rel_err(out.view(V, T + 1, N)[:, 1:], ref[:, 1:]) < 1e-5
    assert float(out.view(V, T + 1, N)[:, 0].abs().max()) == 0.0   # CLS rows untouched


# ------------------------------------------------------------------------------------------------ row ops
@pytest.mark.parametrize("rows,d", [(1000, 768), (12608, 768), (257, 1024), (33, 128)])
def test_layernorm_fwd_bwd(G, rows, d):
    gu, L = G
    x = torch.randn(rows, d, device="cuda") * 2 + 0.3
    gam, bet = torch.randn(d, device="cuda") * 0.1 + 1, torch.randn(d, device="cuda") * 0.1
    y = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    gu.ok(gu.lib().ttl_op_layernorm(gu.ptr(x), gu.ptr(y), gu.ptr(gam), gu.ptr(bet), rows, d, 1e-5, gu.stream()))
    xr = x.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (d,), gam, bet, 1e-5)
    assert gu.rel_err(y, ref.detach()) < 4e-3
    dy, dres = torch.randn(rows, d, device="cuda"), torch.randn(rows, d, device="cuda")
    ref.backward(dy)
    dx = torch.empty(rows, d, device="cuda")
    dxb = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    gu.ok(gu.lib().ttl_op_layernorm_bwd(gu.ptr(dy), gu.ptr(x), gu.ptr(gam), gu.ptr(dres), gu.ptr(dx), gu.ptr(dxb), rows, d,
                                        1e-5, gu.stream()))
    torch.cuda.synchronize()
    assert gu.rel_err(dx, xr.grad + dres) < 1e-5       # fp32 kernel vs fp32 torch
    assert gu.rel_err(dxb, xr.grad + dres) < 4e-3


def test_im2col(G):
    gu, L = G
    for (V, S, p) in ((2, 224, 16), (2, 224, 14), (3, 64, 16)):
        img = torch.randn(V, 3, S, S, device="cuda")
        K = 3 * p * p
        Kp = (K + 63) // 64 * 64
        T = (S // p) ** 2
        out = torch.full((V * T, Kp), 7.0, device="cuda", dtype=torch.bfloat16)
        gu.ok(gu.lib().ttl_op_im2col(gu.ptr(img), gu.ptr(out), V, S, p, gu.stream()))
        ref = torch.nn.functional.unfold(img, p, stride=p).transpose(1, 2).reshape(V * T, K)   # (c,i,j) order
        torch.cuda.synchronize()
        assert torch.equal(out[:, :K], ref.to(torch.bfloat16))                                  # exact: pure rounding
        assert float(out[:, K:].float().abs().max() if Kp > K else 0.0) == 0.0


# ------------------------------------------------------------------------------------------------ attention
def _attn_ref(qkv, V, tokens, heads):
    d = heads * 64
    q, k, v = qkv.float().view(V, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    p = torch.softmax(s, -1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(V * tokens, d)
    return o, torch.logsumexp(s, -1)


# 129..208 tokens: tcgen05 backward (two query tiles x two key halves; 129 / 144 / 145 / 176 / 208 walk the 16..80-key second half
# and the partial last query tile; 200 views x 12 heads = 2400 units > 148 CTAs wrap the mbarrier phases many times)
@pytest.mark.parametrize("V,tokens,heads", [(3, 197, 12), (2, 257, 16), (4, 17, 2), (1, 64, 1), (64, 197, 12), (5, 50, 3),
                                            (2, 129, 2), (2, 144, 3), (3, 145, 2), (7, 176, 4), (2, 208, 2), (200, 197, 12),
                                            (3, 193, 2), (2, 200, 3)])      # 193..208: forward with P kept in TMEM (one row in the last active warp at 193)
def test_attention_fwd_bwd(G, V, tokens, heads):
    gu, L = G
    d = heads * 64
    qkv = _bf(V * tokens, 3 * d, scale=1.5, seed=11)
    out = torch.empty(V * tokens, d, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(V, heads, tokens, device="cuda")
    gu.ok(gu.lib().ttl_op_attention_fwd(gu.ptr(qkv), gu.ptr(out), gu.ptr(lse), V, tokens, heads, 0.125, gu.stream()))
    torch.cuda.synchronize()
    qr = qkv.float().requires_grad_(True)
    ref, lse_ref = _attn_ref(qr, V, tokens, heads)
    assert gu.rel_err(out, ref.detach()) < 8e-3      # P is rounded to bf16 before P@V, output rounded to bf16
    assert float((lse - lse_ref.detach()).abs().max()) < 2e-3
    dout = _bf(V * tokens, d, seed=12)
    ref.backward(dout.float())
    dqkv = torch.empty_like(qkv)
    gu.ok(gu.lib().ttl_op_attention_bwd(gu.ptr(qkv), gu.ptr(out), gu.ptr(dout), gu.ptr(lse), gu.ptr(dqkv), V, tokens, heads,
                                        0.125, gu.stream()))
    torch.cuda.synchronize()
    g = qr.grad.view(V * tokens, 3, d)
    got = dqkv.float().view(V * tokens, 3, d)
    for i, nm in enumerate("qkv"):
        assert gu.rel_err(got[:, i], g[:, i]) < 1.5e-2, nm   # bf16 P/dS operands; fp32 accumulation


def test_attention_fwd_257_tokens_many_units_and_extreme_scores(G):
    """257 tokens (ViT-L/14) on the P-in-TMEM kernel with the extra key / query row: 40 views x 16 heads = 640 units > 148 CTAs, so
    every CTA wraps its double-buffered K / V and the mbarrier phases several times; a second input whose later keys -- key 256
    included -- are scaled 40x exercises the exact-max softmax across the MMA keys and the folded-in key, and the CUDA-core row."""
    gu, L = G
    V, tokens, heads = 40, 257, 16
    d = heads * 64
    for extreme in (False, True):
        qkv = torch.randn(V * tokens, 3 * d, device="cuda", generator=torch.Generator("cuda").manual_seed(5 + extreme)) * 1.5
        if extreme:
            qkv.view(V, tokens, 3, heads, 64)[:, 200:, 1] *= 40.0
        qkv = qkv.bfloat16()
        out = torch.empty(V * tokens, d, device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(V, heads, tokens, device="cuda")
        gu.ok(gu.lib().ttl_op_attention_fwd(gu.ptr(qkv), gu.ptr(out), gu.ptr(lse), V, tokens, heads, 0.125, gu.stream()))
        torch.cuda.synchronize()
        ref, lse_ref = _attn_ref(qkv.float(), V, tokens, heads)
        assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
        assert gu.rel_err(out, ref) < 8e-3
        o3, r3 = out.float().view(V, tokens, d), ref.view(V, tokens, d)
        assert gu.rel_err(o3[:, 256], r3[:, 256]) < 8e-3                 # the CUDA-core query row
        assert gu.rel_err(o3[:, 128:256], r3[:, 128:256]) < 8e-3         # the second query tile
        assert float(((lse - lse_ref).abs() / lse_ref.abs().clamp_min(1.0)).max()) < 2e-3


@pytest.mark.parametrize("mode", ["pp", "tc", "mma"])
def test_attention_forward_alternatives_still_agree(mode):
    """TTL_ATTN selects the other forward kernels (read once per process, so each runs in a child): the two-tiles-in-flight
    kernel with P in shared memory, the one-tile-per-item kernel and the mma.sync kernel must pass the same checks at 197 tokens."""
    import os, subprocess, sys
    env = dict(os.environ, TTL_ATTN=mode)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.abspath(__file__), "-k",
                        "test_attention_fwd_bwd and (3-197-12 or 64-197-12) or test_attention_fwd_extreme_scores"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "3 passed" in r.stdout, r.stdout[-500:]


def test_attention_fwd_extreme_scores(G):
    """Scores spanning hundreds of nats (keys 64.. scaled 40x): exact-max softmax must stay finite and accurate, and the
    near one-hot rows exercise the 16-key tail block of P (32B-swizzled) and the last V rows of the MN-major operand."""
    gu, L = G
    V, tokens, heads = 3, 197, 12
    d = heads * 64
    qkv = torch.randn(V * tokens, 3 * d, device="cuda") * 1.5
    kk = qkv.view(V, tokens, 3, heads, 64)
    kk[:, 64:, 1] *= 40.0                      # K rows of the later tokens
    qkv = qkv.bfloat16()
    out = torch.empty(V * tokens, d, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(V, heads, tokens, device="cuda")
    gu.ok(gu.lib().ttl_op_attention_fwd(gu.ptr(qkv), gu.ptr(out), gu.ptr(lse), V, tokens, heads, 0.125, gu.stream()))
    torch.cuda.synchronize()
    ref, lse_ref = _attn_ref(qkv.float(), V, tokens, heads)
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    assert gu.rel_err(out, ref) < 8e-3
    assert float(((lse - lse_ref).abs() / lse_ref.abs().clamp_min(1.0)).max()) < 2e-3


# ------------------------------------------------------------------------------------------------ head
@pytest.mark.parametrize("C_", [10, 200, 1000])
def test_logits_entropy(G, C_):
    gu, L = G
    V, P = 64, 512
    feats = torch.randn(V, P, device="cuda")
    text = torch.nn.functional.normalize(torch.randn(C_, P, device="cuda"), dim=-1)
    logits = torch.empty(V, C_, device="cuda")
    ent = torch.empty(V, device="cuda")
    gu.ok(gu.lib().ttl_op_logits_entropy(gu.ptr(feats), gu.ptr(text), 100.0, gu.ptr(logits), gu.ptr(ent), V, C_, P, gu.stream()))
    torch.cuda.synchronize()
    ref = O.clip_logits(feats.cpu().double(), text.cpu().double(), math.log(100.0))
    assert float((logits.cpu().double() - ref).abs().max()) < 2e-4          # fp32 dot products of length 512, |logit| <= 100
    assert float((ent.cpu().double() - O.softmax_entropy(ref)).abs().max()) < 2e-4


def test_selection_bit_exact_given_entropies(G):
    gu, L = G
    g = torch.Generator().manual_seed(3)
    for trial in range(20):
        V = 64
        ent = torch.rand(V, generator=g)
        if trial % 2:                                   # engineered ties
            ent[torch.randint(0, V, (24,), generator=g)] = float(ent[5])
        K = int(V * 0.1)
        idx = torch.empty(K, dtype=torch.int32, device="cuda")
        e = ent.cuda()
        gu.ok(gu.lib().ttl_op_select(gu.ptr(e), V, K, gu.ptr(idx), gu.stream()))
        torch.cuda.synchronize()
        ref = torch.argsort(ent, descending=False, stable=True)[:K]
        assert idx.cpu().tolist() == ref.tolist()


@pytest.mark.parametrize("C_", [10, 200, 1000])
def test_tpt_loss_and_grad(G, C_):
    gu, L = G
    torch.manual_seed(C_)
    V, K = 64, 6
    logits = torch.randn(V, C_) * 3
    _, idx = O.select_confident_samples(logits, 0.1)
    x = logits[idx].double().requires_grad_(True)
    loss_ref = O.avg_entropy(x)
    loss_ref.backward()
    lg, ix = logits.cuda(), idx.to(torch.int32).cuda()
    loss = torch.empty(1, device="cuda")
    dl = torch.empty(K, C_, device="cuda")
    gu.ok(gu.lib().ttl_op_tpt_loss(gu.ptr(lg), gu.ptr(ix), K, C_, gu.ptr(loss), gu.ptr(dl), gu.stream()))
    torch.cuda.synchronize()
    assert abs(float(loss) - float(loss_ref)) < 1e-5 * max(1.0, abs(float(loss_ref)))
    assert gu.rel_err(dl.cpu(), x.grad) < 1e-5
    # idx == NULL means rows 0..K-1 of the given matrix
    sel = lg[idx.cuda()].contiguous()
    gu.ok(gu.lib().ttl_op_tpt_loss(gu.ptr(sel), None, K, C_, gu.ptr(loss), gu.ptr(dl), gu.stream()))
    torch.cuda.synchronize()
    assert gu.rel_err(dl.cpu(), x.grad) < 1e-5


@pytest.mark.parametrize("C_", [10, 1000])
def test_deyo_loss_and_grad(G, C_):
    gu, L = G
    torch.manual_seed(C_ + 1)
    V = 64
    logits = torch.randn(V, C_) * 2
    x = logits.double().requires_grad_(True)
    loss_ref = O.deyo_loss(x, 0.4)
    loss_ref.backward()
    lg = logits.cuda()
    loss = torch.empty(1, device="cuda")
    dl = torch.empty(V, C_, device="cuda")
    gu.ok(gu.lib().ttl_op_deyo_loss(gu.ptr(lg), V, C_, 0.4, gu.ptr(loss), gu.ptr(dl), gu.stream()))
    torch.cuda.synchronize()
    assert abs(float(loss) - float(loss_ref)) < 1e-5 * max(1.0, abs(float(loss_ref)))
    assert gu.rel_err(dl.cpu(), x.grad) < 1e-5


# ------------------------------------------------------------------------------------------------ LoRA-side
def test_adamw_matches_torch(G):
    gu, L = G
    torch.manual_seed(0)
    n = 147456
    p0 = torch.randn(n, device="cuda") * 0.05
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=5e-3)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        g = torch.randn(n, device="cuda") * 1e-2
        ref.grad = g.clone()
        opt.step()
        gu.ok(gu.lib().ttl_op_adamw(gu.ptr(p), gu.ptr(g), gu.ptr(m), gu.ptr(v), n, step, 5e-3, 0.9, 0.999, 1e-8, 1e-2, gu.stream()))
    torch.cuda.synchronize()
    assert float((p - ref.detach()).abs().max()) < 2e-7


@pytest.mark.parametrize("M", [1182, 12608, 100])
def test_skinny_reduce(G, M):
    gu, L = G
    wide, narrow = _bf(M, 2304, seed=21), _bf(M, 64, seed=22)
    ws = torch.empty((M + 127) // 128 * 768 * 32, device="cuda")
    out = torch.empty(768, 16, device="cuda")
    gu.ok(gu.lib().ttl_op_skinny_reduce(gu.ptr(wide[:, 1536:]), 2304, 768, gu.ptr(narrow[:, 16:]), 64, 16, M, 2.0, gu.ptr(out), 0,
                                        gu.ptr(ws), gu.stream()))
    torch.cuda.synchronize()
    ref = 2.0 * wide[:, 1536:].float().t() @ narrow[:, 16:32].float()
    assert gu.rel_err(out, ref) < 1e-5
    outT = torch.empty(16, 768, device="cuda")
    gu.ok(gu.lib().ttl_op_skinny_reduce(gu.ptr(wide), 2304, 768, gu.ptr(narrow), 64, 16, M, 1.0, gu.ptr(outT), 1, gu.ptr(ws), gu.stream()))
    torch.cuda.synchronize()
    assert gu.rel_err(outT, (wide[:, :768].float().t() @ narrow[:, :16].float()).t()) < 1e-5
