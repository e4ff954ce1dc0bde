"""-m gpu: NaViT through the fused padding-free sm_100a path against the reference golden and the per-image oracle."""
import random

import pytest
import torch

from conftest import load_golden
from oracle import navit_oracle as NO
from oracle import vit_oracle as O
from vit_pytorch_b200 import NaViT, _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _stats(got, ref, rtol=1e-2, atol=1e-3):
    d = (got.float().cpu() - ref).abs()
    return d.max().item(), d.mean().item(), (d <= atol + rtol * ref.abs()).float().mean().item()


def test_navit_golden_fused():
    g = load_golden("navit_tiny")
    m = NaViT(**g["kwargs"]).eval()
    m.load_state_dict(g["state_dict"])
    m = m.to(DEV, torch.bfloat16)
    imgs = [im.to(DEV) for im in g["images"]]
    rows = [[imgs[i] for i in r] for r in g["rows"]]
    _lib.reset_launch_count()
    with torch.inference_mode():
        assert m.fused_reason(rows) is None
        out = m(rows)
        out_grouped = m(imgs, group_images=True, group_max_seq_len=g["group_max_seq_len"])
    assert _lib.launch_count() > 0
    mx, mean, frac = _stats(out, g["logits_fp32"])
    print(f"navit_tiny fused vs reference fp32: max {mx:.5f} mean {mean:.5f} within {frac:.4f}")
    assert out.shape == g["logits_fp32"].shape and mx < 1.5e-2 and frac > 0.85
    assert torch.equal(out, out_grouped)          # packing does not exist on the fused path: identical bits


def test_navit_varied_resolutions_against_oracle():
    """BASELINE.json configs[4] style batch (variable resolutions up to 32x32 patches => up to 1024 tokens per image,
    multi key-block attention), small width/depth so the per-image oracle finishes in seconds."""
    kwargs = dict(image_size=512, patch_size=16, num_classes=100, dim=256, depth=2, heads=4, mlp_dim=512)
    torch.manual_seed(3)
    m = NaViT(**kwargs).eval().bfloat16()
    random.seed(1)
    sizes = [(16 * random.randrange(1, 33), 16 * random.randrange(1, 33)) for _ in range(12)] + [(512, 512), (16, 16)]
    torch.manual_seed(4)
    imgs = [torch.randn(3, h, w).bfloat16() for h, w in sizes]
    ref = NO.navit_forward(O.upcast(m.state_dict()), kwargs, [[im.float() for im in imgs]])
    m = m.to(DEV)
    with torch.inference_mode():
        assert m.fused_reason([im.to(DEV) for im in imgs]) is None
        out = m([im.to(DEV) for im in imgs])
    mx, mean, frac = _stats(out, ref)
    print(f"navit 14 images (1..1024 tokens) fused vs fp32 oracle: max {mx:.5f} mean {mean:.5f} within {frac:.4f}")
    assert out.shape == (14, 100) and torch.isfinite(out.float()).all()
    assert mx < 3e-2 and frac > 0.80


def test_varlen_attention_kernel_against_oracle():
    lengths = [197, 1, 130, 577, 64, 1024, 129]
    H, dh = 3, 64
    T = sum(lengths)
    torch.manual_seed(0)
    qkv = torch.randn(T, 3 * H * dh, device=DEV).bfloat16()
    out = torch.zeros(T, H * dh, device=DEV, dtype=torch.bfloat16)
    cu, tp, tiles = _lib.varlen_index(lengths, DEV)
    _lib.attention_varlen(qkv, out, cu, tp, tiles, H, dh, dh ** -0.5)
    ref = torch.empty(T, H * dh)
    o = 0
    for n in lengths:
        q, k, v = qkv[o:o + n].float().cpu().view(n, 3, H, dh).permute(1, 2, 0, 3)
        ref[o:o + n] = (O.softmax_last((q @ k.transpose(-1, -2)) * dh ** -0.5) @ v).permute(1, 0, 2).reshape(n, H * dh)
        o += n
    mx, mean, frac = _stats(out, ref)
    assert frac > 0.995 and mx < 2e-2, (mx, mean, frac)


def test_qk_rmsnorm_and_attn_pool_kernels():
    torch.manual_seed(1)
    T, H, dh = 300, 4, 64
    I = H * dh
    qkv = torch.randn(T, 3 * I, device=DEV).bfloat16()
    g = torch.randn(2, H, dh, device=DEV)
    ref = qkv.float().clone().view(T, 3, H, dh)
    for s in (0, 1):
        ref[:, s] = NO.rms_norm_heads(ref[:, s].permute(1, 0, 2).cpu(), g[s].cpu()[:, None, :]).permute(1, 0, 2).to(DEV)
    _lib.qk_rmsnorm(qkv, g.reshape(-1).contiguous(), H, dh)
    assert torch.allclose(qkv.float().view(T, 3, H, dh), ref, rtol=1e-2, atol=1e-2)
    lengths = [100, 1, 199]
    cu, _, _ = _lib.varlen_index(lengths, DEV)
    kv = qkv[:, I:].contiguous()
    qn = torch.randn(I, device=DEV)
    out = torch.zeros(3, I, device=DEV, dtype=torch.bfloat16)
    _lib.attn_pool(kv, qn, cu, out, H, dh)
    o = 0
    for i, n in enumerate(lengths):
        k = kv[o:o + n, :I].float().view(n, H, dh)
        v = kv[o:o + n, I:].float().view(n, H, dh)
        sc = torch.einsum("hd,nhd->hn", qn.view(H, dh), k)
        want = torch.einsum("hn,nhd->hd", sc.softmax(-1), v).reshape(-1)
        assert torch.allclose(out[i].float(), want, rtol=2e-2, atol=2e-2)
        o += n
