"""CPU oracle for the ViT / SimpleViT encoder forward  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import this module;
the product path (vit_pytorch_b200) never does.

This is a from-scratch functional restatement (explicit tensor arithmetic on CPU, no nn.Module, no einops) of the
algorithm in lucidrains/vit-pytorch v1.23.6:
    vit_pytorch/vit.py:15-138         (FeedForward, Attention, Transformer, ViT)
    vit_pytorch/simple_vit.py:12-120  (posemb_sincos_2d, FeedForward, Attention, Transformer, SimpleViT)
It consumes a reference-compatible ``state_dict`` (same key names) plus a config dict.

Parity pin: the reference ships no golden vectors for this path (its only test checks a shape,
tests/test_vit.py:19-20), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF generated in the build
container by tests/golden/make_golden.py (imports /root/reference) and committed under tests/golden/*.pt;
tests/test_oracle.py checks this file against those fixtures (and, when /root/reference is present, against the
live reference on fresh seeds).

The arithmetic third-party dependency of the reference on this path is PyTorch itself (torch>=2.4, unpinned;
2.11.0+cu128 here: LayerNorm eps 1e-5 with biased variance, exact-erf GELU, softmax over keys) and einops>=0.8.2
for the two permutations; both are restated below with plain tensor ops.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


def pair(t) -> Tuple[int, int]:
    """vit.py:10-11 / simple_vit.py:9-10."""
    return t if isinstance(t, tuple) else (t, t)


# ----------------------------------------------------------------------------------------------------------------
# primitive ops (what the reference obtains from torch.nn)
# ----------------------------------------------------------------------------------------------------------------
def layer_norm(x: Tensor, weight: Tensor, bias: Optional[Tensor], eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm(dim) as used at vit.py:19,39,69,101,103: biased variance, eps inside the sqrt."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    y = xc * torch.rsqrt(var + eps) * weight
    return y if bias is None else y + bias


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None) -> Tensor:
    """nn.Linear: y = x W^T + b with W stored [out, in] (vit.py:20,23,44,47,102,116)."""
    y = x @ weight.transpose(-1, -2)
    return y if bias is None else y + bias


def gelu_erf(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form (vit.py:21 / simple_vit.py:31)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def softmax_last(x: Tensor) -> Tensor:
    """nn.Softmax(dim=-1) (vit.py:41,59)."""
    m = x.amax(dim=-1, keepdim=True)
    e = torch.exp(x - m)
    return e / e.sum(dim=-1, keepdim=True)


# ----------------------------------------------------------------------------------------------------------------
# patch embedding
# ----------------------------------------------------------------------------------------------------------------
def patchify(img: Tensor, ph: int, pw: int) -> Tensor:
    """Rearrange('b c (h p1) (w p2) -> b (h w) (p1 p2 c)') (vit.py:100 / simple_vit.py:91).

    out[b, h*gw + w, (p1*pw + p2)*C + c] = img[b, c, h*ph + p1, w*pw + p2]
    """
    B, C, H, W = img.shape
    gh, gw = H // ph, W // pw
    t = img.reshape(B, C, gh, ph, gw, pw)          # b c h p1 w p2
    t = t.permute(0, 2, 4, 3, 5, 1)                # b h w p1 p2 c
    return t.reshape(B, gh * gw, ph * pw * C)


def patch_embed(sd: StateDict, img: Tensor, ph: int, pw: int) -> Tensor:
    """to_patch_embedding: Rearrange -> LayerNorm(patch_dim) -> Linear -> LayerNorm(dim) (vit.py:99-104)."""
    p = "to_patch_embedding."
    x = patchify(img, ph, pw)
    x = layer_norm(x, sd[p + "1.weight"], sd[p + "1.bias"])
    x = linear(x, sd[p + "2.weight"], sd[p + "2.bias"])
    return layer_norm(x, sd[p + "3.weight"], sd[p + "3.bias"])


def posemb_sincos_2d(h: int, w: int, dim: int, temperature: float = 10000.0) -> Tensor:
    """simple_vit.py:12-21: fp32 table [h*w, dim] = [sin(x w), cos(x w), sin(y w), cos(y w)], token = y*w + x."""
    assert dim % 4 == 0, "feature dimension must be multiple of 4 for sincos emb"
    q = dim // 4
    omega = 1.0 / (temperature ** (torch.arange(q, dtype=torch.float32) / (q - 1)))
    ys = torch.arange(h, dtype=torch.float32).repeat_interleave(w)   # token -> y
    xs = torch.arange(w, dtype=torch.float32).repeat(h)              # token -> x
    ya = ys[:, None] * omega[None, :]
    xa = xs[:, None] * omega[None, :]
    return torch.cat([xa.sin(), xa.cos(), ya.sin(), ya.cos()], dim=1)


# ----------------------------------------------------------------------------------------------------------------
# transformer block
# ----------------------------------------------------------------------------------------------------------------
def attention(sd: StateDict, prefix: str, x: Tensor, heads: int, out_key: Optional[str],
              return_attn: bool = False):
    """Attention.forward (vit.py:51-64 / simple_vit.py:50-62).

    out_key: 'to_out.0' (ViT: Linear+bias inside a Sequential), 'to_out' (SimpleViT: bias-free Linear) or None
    (ViT with heads == 1 and dim_head == dim: nn.Identity, vit.py:34,46-49).
    """
    B, N, _ = x.shape
    xn = layer_norm(x, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])           # :52
    qkv = linear(xn, sd[prefix + "to_qkv.weight"])                                        # :54  (no bias)
    inner = qkv.shape[-1] // 3
    dh = inner // heads
    scale = dh ** -0.5                                                                    # :37
    qkv = qkv.reshape(B, N, 3, heads, dh).permute(2, 0, 3, 1, 4)                          # :54-55 chunk + split heads
    q, k, v = qkv[0], qkv[1], qkv[2]                                                      # each [B, H, N, dh]
    dots = (q @ k.transpose(-1, -2)) * scale                                              # :57
    attn = softmax_last(dots)                                                             # :59
    out = attn @ v                                                                        # :62
    out = out.permute(0, 2, 1, 3).reshape(B, N, inner)                                    # :63 merge heads
    if out_key is not None:
        out = linear(out, sd[prefix + out_key + ".weight"], sd.get(prefix + out_key + ".bias"))   # :64
    return (out, attn) if return_attn else out


def feed_forward(sd: StateDict, prefix: str, x: Tensor, fc2_index: int) -> Tensor:
    """FeedForward.net (vit.py:18-25: indices 0 LN, 1 Linear, 2 GELU, 3 Dropout, 4 Linear, 5 Dropout;
    simple_vit.py:28-33: 0 LN, 1 Linear, 2 GELU, 3 Linear)."""
    h = layer_norm(x, sd[prefix + "net.0.weight"], sd[prefix + "net.0.bias"])
    h = gelu_erf(linear(h, sd[prefix + "net.1.weight"], sd[prefix + "net.1.bias"]))
    return linear(h, sd[f"{prefix}net.{fc2_index}.weight"], sd[f"{prefix}net.{fc2_index}.bias"])


def transformer(sd: StateDict, x: Tensor, depth: int, heads: int, kind: str, prefix: str = "transformer.") -> Tensor:
    """Transformer.forward (vit.py:78-83 / simple_vit.py:74-78): pre-LN residual blocks, final LayerNorm."""
    for i in range(depth):
        ap = f"{prefix}layers.{i}.0."
        fp = f"{prefix}layers.{i}.1."
        if kind == "vit":
            out_key = "to_out.0" if (ap + "to_out.0.weight") in sd else None
            fc2 = 4
        else:
            out_key, fc2 = "to_out", 3
        x = attention(sd, ap, x, heads, out_key) + x          # vit.py:80
        x = feed_forward(sd, fp, x, fc2) + x                  # vit.py:81
    return layer_norm(x, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])   # vit.py:83


# ----------------------------------------------------------------------------------------------------------------
# full models
# ----------------------------------------------------------------------------------------------------------------
def vit_forward(sd: StateDict, cfg: dict, img: Tensor) -> Tensor:
    """ViT.forward (vit.py:118-138), eval mode (dropout = identity).

    cfg keys: patch_size, depth, heads, pool ('cls'|'mean'); num_classes == 0 <=> 'mlp_head.weight' absent.
    """
    ph, pw = pair(cfg["patch_size"])
    x = patch_embed(sd, img, ph, pw)                                       # :120
    B = x.shape[0]
    cls = sd["cls_token"]                                                  # (1, D) or (0, D)
    x = torch.cat([cls.unsqueeze(0).expand(B, -1, -1), x], dim=1)          # :122-123
    x = x + sd["pos_embedding"][: x.shape[1]]                              # :125-127
    x = transformer(sd, x, cfg["depth"], cfg["heads"], "vit")              # :130
    if "mlp_head.weight" not in sd:                                        # :132-133
        return x
    x = x.mean(dim=1) if cfg.get("pool", "cls") == "mean" else x[:, 0]     # :135
    return linear(x, sd["mlp_head.weight"], sd["mlp_head.bias"])           # :137-138 (to_latent = Identity)


def simple_vit_forward(sd: StateDict, cfg: dict, img: Tensor) -> Tensor:
    """SimpleViT.forward (simple_vit.py:110-120).  cfg keys: patch_size, depth, heads."""
    ph, pw = pair(cfg["patch_size"])
    x = patch_embed(sd, img, ph, pw)                                       # :113
    gh, gw = img.shape[2] // ph, img.shape[3] // pw
    x = x + posemb_sincos_2d(gh, gw, x.shape[-1]).to(x.dtype)              # :114 (table built in fp32, :97-101)
    x = transformer(sd, x, cfg["depth"], cfg["heads"], "simple")           # :116
    x = x.mean(dim=1)                                                      # :117
    return linear(x, sd["linear_head.weight"], sd["linear_head.bias"])     # :119-120


def forward(kind: str, sd: StateDict, cfg: dict, img: Tensor) -> Tensor:
    return vit_forward(sd, cfg, img) if kind == "vit" else simple_vit_forward(sd, cfg, img)


def upcast(sd: StateDict, dtype=torch.float32) -> StateDict:
    return {k: v.detach().to("cpu", dtype) for k, v in sd.items()}


# ----------------------------------------------------------------------------------------------------------------
# FLOP accounting shared by bench.py and tests: 2 x MACs of every contraction at unpadded N
# ----------------------------------------------------------------------------------------------------------------
def flops_per_image(*, image_size, patch_size, dim, depth, heads, mlp_dim, num_classes, dim_head=64, channels=3,
                    cls_tokens=1) -> float:
    ih, iw = pair(image_size)
    ph, pw = pair(patch_size)
    n = (ih // ph) * (iw // pw)
    N = n + cls_tokens
    pd = channels * ph * pw
    inner = heads * dim_head
    per_layer = 2 * N * dim * 3 * inner + 2 * heads * N * N * dim_head * 2 + 2 * N * inner * dim + 2 * 2 * N * dim * mlp_dim
    return float(2 * n * pd * dim + depth * per_layer + 2 * dim * num_classes)
