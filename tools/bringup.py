"""GPU bring-up harness: runs each kernel of libb200vit.so against a torch fp32 restatement of the same op.

Every case runs in its own subprocess with a timeout (a trapped kernel poisons the CUDA context), results are
written to gpurun_out/bringup.json.   Usage:  python tools/bringup.py [case-prefix ...]
"""
from __future__ import annotations

import json
import math
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _err(got, ref, atol=1e-3):
    import torch
    got = got.float()
    ref = ref.float()
    d = (got - ref).abs()
    tol = atol + 1e-2 * ref.abs()
    return {
        "max_abs": float(d.max()),
        "mean_abs": float(d.mean()),
        "ref_absmax": float(ref.abs().max()),
        "within_tol": float((d <= tol).float().mean()),
        "nan": int(torch.isnan(got).sum()),
    }


def case_gemm(M, N, K, bias=False, gelu=False, resid=False, f32=False, both=False, lnfold=False, stats=False,
              lda=None, force=0, inplace=False):
    import torch
    from vit_pytorch_b200 import _lib
    _lib.lib().b200vit_debug_set(4, force)
    torch.manual_seed(0)
    dev = "cuda"
    lda = lda or K
    a_full = torch.randn(M, lda, device=dev).bfloat16()
    a = a_full[:, :K]
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
    ldo = N
    out_bf16 = torch.zeros(M, ldo, device=dev, dtype=torch.bfloat16) if (not f32 or both) else None
    out_f32 = torch.zeros(M, ldo, device=dev) if (f32 or both) else None
    b = torch.randn(N, device=dev) if bias else None
    r = torch.randn(M, N, device=dev) if resid else None
    ref = a.float() @ w.float().t()
    ln_sums = col_s = None
    if lnfold:
        af = a.float()
        ln_sums = torch.stack([af.sum(1), (af * af).sum(1)], 1).contiguous()
        col_s = w.float().sum(1).contiguous()
        mu = af.mean(1, keepdim=True)
        var = (af * af).mean(1, keepdim=True) - mu * mu
        rstd = torch.rsqrt(var + 1e-5)
        ref = rstd * (ref - mu * col_s[None, :])
    if bias:
        ref = ref + b
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    if resid:
        ref = ref + r
    st = torch.zeros(M, _lib.stats_parts(N), 2, device=dev) if stats else None
    if inplace:
        out_f32 = r.clone()
        _lib.gemm(a, w, out_f32=out_f32, bias=b, resid=out_f32)
        torch.cuda.synchronize()
        out_f32_first = out_f32.clone()
        r_arg = out_f32
    else:
        r_arg = r
        _lib.gemm(a, w, out_bf16=out_bf16, out_f32=out_f32, bias=b, resid=r, gelu=gelu, ln_sums=ln_sums, col_s=col_s,
                  stats_out=st)
        torch.cuda.synchronize()
    res = {}
    if out_bf16 is not None:
        res["bf16"] = _err(out_bf16, ref)
    if out_f32 is not None:
        res["f32"] = _err(out_f32_first if inplace else out_f32, ref)
    if stats:
        rb = ref.bfloat16().float()
        # a row sum of N terms cancels: its error scales with sqrt(N) * |term|, not with the (possibly tiny) sum
        res["stats_sum"] = _err(st.sum(1)[:, 0], rb.sum(1), atol=1e-3 * float(rb.abs().max()) * N ** 0.5)
        res["stats_sq"] = _err(st.sum(1)[:, 1], (rb * rb).sum(1))
    # timing
    if M * N * K > 1e9:
        r = r_arg
        for _ in range(3):
            _lib.gemm(a, w, out_bf16=out_bf16, out_f32=out_f32, bias=b, resid=r, gelu=gelu, ln_sums=ln_sums,
                      col_s=col_s)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            _lib.gemm(a, w, out_bf16=out_bf16, out_f32=out_f32, bias=b, resid=r, gelu=gelu, ln_sums=ln_sums,
                      col_s=col_s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K / ms / 1e9
        # cuBLAS bar on the same shape
        wt = w.t().contiguous()
        for _ in range(3):
            torch.matmul(a, wt)
        e0.record()
        for _ in range(10):
            torch.matmul(a, wt)
        e1.record()
        torch.cuda.synchronize()
        res["cublas_tflops"] = 2.0 * M * N * K / (e0.elapsed_time(e1) / 10) / 1e9
    ok = all(v["within_tol"] > 0.999 and v["nan"] == 0 for k, v in res.items() if isinstance(v, dict))
    if not ok and out_bf16 is not None:
        d = (out_bf16.float() - ref).abs()
        bad = (d > 1e-3 + 2e-2 * ref.abs())
        res["bad_rows_head"] = bad.any(1).nonzero().flatten()[:16].tolist()
        res["bad_cols_head"] = bad.any(0).nonzero().flatten()[:16].tolist()
        res["bad_frac"] = float(bad.float().mean())
        res["sample_got"] = out_bf16[:2, :8].float().tolist()
        res["sample_ref"] = ref[:2, :8].tolist()
    res["ok"] = ok
    return res


def case_layernorm(M, D, beta=True, idx=False):
    import torch
    from vit_pytorch_b200 import _lib
    torch.manual_seed(0)
    x = torch.randn(M, D, device="cuda") * 3 + 1
    g = torch.randn(D, device="cuda")
    b = torch.randn(D, device="cuda") if beta else None
    ri = None
    Mo = M
    if idx:
        ri = torch.arange(0, M, 3, device="cuda", dtype=torch.int32)
        Mo = ri.numel()
    ob = torch.zeros(Mo, D, device="cuda", dtype=torch.bfloat16)
    of = torch.zeros(Mo, D, device="cuda")
    _lib.layernorm(x, g, b, out_bf16=ob, out_f32=of, row_index=ri)
    xs = x if ri is None else x[ri.long()]
    ref = torch.nn.functional.layer_norm(xs, (D,), g, b, 1e-5)
    torch.cuda.synchronize()
    res = {"f32": _err(of, ref), "bf16": _err(ob, ref)}
    res["ok"] = res["f32"]["max_abs"] < 1e-4 and res["bf16"]["within_tol"] > 0.999
    return res


def case_patchify(B, C, H, W, p):
    import torch
    from einops import rearrange
    from vit_pytorch_b200 import _lib
    torch.manual_seed(0)
    img = torch.randn(B, C, H, W, device="cuda").bfloat16()
    pd = C * p * p
    ldo = (pd + 63) // 64 * 64
    g = torch.randn(pd, device="cuda")
    b = torch.randn(pd, device="cuda")
    out = torch.full((B * (H // p) * (W // p), ldo), 7.0, device="cuda", dtype=torch.bfloat16)
    _lib.patchify_ln(img, g, b, out, p, p)
    ref = rearrange(img.float(), "b c (h p1) (w p2) -> (b h w) (p1 p2 c)", p1=p, p2=p)
    ref = torch.nn.functional.layer_norm(ref, (pd,), g, b, 1e-5)
    torch.cuda.synchronize()
    res = {"bf16": _err(out[:, :pd], ref)}
    res["pad_zero"] = bool((out[:, pd:] == 0).all())
    res["ok"] = res["bf16"]["within_tol"] > 0.999 and res["pad_zero"]
    return res


def case_embed(B, n, ncls, D):
    import torch
    from vit_pytorch_b200 import _lib
    torch.manual_seed(0)
    y = torch.randn(B * n, D, device="cuda")
    g = torch.randn(D, device="cuda")
    be = torch.randn(D, device="cuda")
    cls = torch.randn(ncls, D, device="cuda") if ncls else None
    pos = torch.randn(n + ncls, D, device="cuda")
    x = torch.zeros(B * (n + ncls), D, device="cuda")
    _lib.embed_tokens(y, g, be, cls, pos, x, B, n, ncls)
    t = torch.nn.functional.layer_norm(y, (D,), g, be, 1e-5).view(B, n, D)
    if ncls:
        t = torch.cat([cls[None].expand(B, -1, -1), t], 1)
    ref = (t + pos[None]).reshape(-1, D)
    torch.cuda.synchronize()
    res = {"f32": _err(x, ref)}
    res["ok"] = res["f32"]["max_abs"] < 1e-4
    return res


def case_pool(B, N, D):
    import torch
    from vit_pytorch_b200 import _lib
    torch.manual_seed(0)
    x = torch.randn(B, N, D, device="cuda")
    o = torch.zeros(B, D, device="cuda")
    _lib.mean_pool(x, o, B, N, D)
    xb = torch.zeros(B * N * D, device="cuda", dtype=torch.bfloat16)
    _lib.cast_f32_bf16(x.view(-1), xb)
    torch.cuda.synchronize()
    res = {"f32": _err(o, x.mean(1)), "cast": _err(xb, x.view(-1).bfloat16())}
    res["ok"] = res["f32"]["max_abs"] < 1e-5 and res["cast"]["max_abs"] == 0
    return res


def case_attention(B, N, H, psmem=0, lbo=1024, sbo=1024, time_it=False, skip_max=0, pv_split=0, exp_emul=0):
    import torch
    from vit_pytorch_b200 import _lib
    L = _lib.lib()
    L.b200vit_debug_set(1, psmem)
    L.b200vit_debug_set(5, skip_max)
    L.b200vit_debug_set(6, pv_split)
    L.b200vit_debug_set(7, exp_emul)
    L.b200vit_debug_set(2, lbo)
    L.b200vit_debug_set(3, sbo)
    torch.manual_seed(0)
    dh = 64
    I = H * dh
    qkv = torch.randn(B * N, 3 * I, device="cuda").bfloat16()
    out = torch.zeros(B * N, I, device="cuda", dtype=torch.bfloat16)
    scale = dh ** -0.5
    _lib.attention(qkv, out, B, N, H, dh, scale)
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * scale
    ref = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B * N, I)
    res = {"bf16": _err(out, ref)}
    res["ok"] = res["bf16"]["within_tol"] > 0.995 and res["bf16"]["nan"] == 0
    if not res["ok"]:
        res["sample_got"] = out[:2, :8].float().tolist()
        res["sample_ref"] = ref[:2, :8].tolist()
        # is it the softmax (uniform-v test) or V addressing?
        d = (out.float() - ref).abs().view(B, N, H, dh)
        res["err_by_dh_head"] = d.mean((0, 1, 2))[:16].tolist()
        res["err_by_row_head"] = d.mean((0, 2, 3))[:8].tolist() + d.mean((0, 2, 3))[-4:].tolist()
    if time_it:
        for _ in range(3):
            _lib.attention(qkv, out, B, N, H, dh, scale)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            _lib.attention(qkv, out, B, N, H, dh, scale)
        e1.record()
        torch.cuda.synchronize()
        res["ms"] = e0.elapsed_time(e1) / 10
        res["gbps"] = (qkv.numel() + out.numel()) * 2 / res["ms"] / 1e6
    return res


def case_varlen(lengths, H, scale=0.125, time_it=False, seed=0):
    import torch
    from vit_pytorch_b200 import _lib
    torch.manual_seed(seed)
    dh = 64
    I = H * dh
    T = sum(lengths)
    qkv = torch.randn(T, 3 * I, device="cuda").bfloat16()
    out = torch.zeros(T, I, device="cuda", dtype=torch.bfloat16)
    cu, tp, tiles = _lib.varlen_index(lengths, "cuda")
    _lib.attention_varlen(qkv, out, cu, tp, tiles, H, dh, scale)
    torch.cuda.synchronize()
    ref = torch.empty(T, I)
    o = 0
    for n in lengths:
        q, k, v = qkv[o:o + n].float().cpu().view(n, 3, H, dh).permute(1, 2, 0, 3)
        ref[o:o + n] = (((q @ k.transpose(-1, -2)) * scale).softmax(-1) @ v).permute(1, 0, 2).reshape(n, I)
        o += n
    res = {"bf16": _err(out.cpu(), ref)}
    res["ok"] = res["bf16"]["within_tol"] > 0.995 and res["bf16"]["nan"] == 0
    if not res["ok"]:
        d = (out.float().cpu() - ref).abs()
        res["err_by_token_head"] = d.mean(1)[:12].tolist()
        res["sample_got"] = out[:2, :6].float().tolist()
        res["sample_ref"] = ref[:2, :6].tolist()
    if time_it:
        for _ in range(3):
            _lib.attention_varlen(qkv, out, cu, tp, tiles, H, dh, scale)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            _lib.attention_varlen(qkv, out, cu, tp, tiles, H, dh, scale)
        e1.record()
        torch.cuda.synchronize()
        res["ms"] = e0.elapsed_time(e1) / 10
        res["gbps"] = (qkv.numel() + out.numel()) * 2 / res["ms"] / 1e6
    return res


def _navit_lengths():
    import random
    random.seed(0)
    return [random.randrange(4, 33) * random.randrange(4, 33) for _ in range(256)]


CASES = {
    "varlen_one_block": lambda: case_varlen([64, 17, 128, 1, 100], 2),
    "varlen_two_blocks": lambda: case_varlen([197, 130, 256, 129], 3),
    "varlen_long": lambda: case_varlen([577, 1024, 300, 50], 2),
    "emul8_197": lambda: case_attention(4, 197, 12, exp_emul=8),
    "emul16_197": lambda: case_attention(4, 197, 12, exp_emul=16),
    "emul16_64": lambda: case_attention(4, 64, 3, exp_emul=16),
    "emul0_big": lambda: case_attention(512, 197, 12, time_it=True, exp_emul=0),
    "emul8_big": lambda: case_attention(512, 197, 12, time_it=True, exp_emul=8),
    "emul16_big": lambda: case_attention(512, 197, 12, time_it=True, exp_emul=16),
    "emul16_1cta_big": lambda: case_attention(512, 197, 12, psmem=1, time_it=True, exp_emul=16),
    "emul8_1cta_big": lambda: case_attention(512, 197, 12, psmem=1, time_it=True, exp_emul=8),
    "split2_197": lambda: case_attention(4, 197, 12, psmem=2),
    "split3_197": lambda: case_attention(4, 197, 12, psmem=3),
    "split2_129": lambda: case_attention(3, 129, 2, psmem=2),
    "split3_256": lambda: case_attention(2, 256, 2, psmem=3),
    "split2_big": lambda: case_attention(512, 197, 12, psmem=2, time_it=True),
    "split3_big": lambda: case_attention(512, 197, 12, psmem=3, time_it=True),
    "cmp257_single": lambda: case_attention(128, 257, 16, time_it=True),
    "cmp257_varlen": lambda: case_varlen([257] * 128, 16, time_it=True),
    "cmp577_varlen": lambda: case_varlen([577] * 64, 12, time_it=True),
    "varlen_vit_b16": lambda: case_varlen([197] * 512, 12, time_it=True),
    "varlen_navit_cfg5": lambda: case_varlen(_navit_lengths(), 16, scale=1.0, time_it=True),
    "gemm_min": lambda: case_gemm(128, 256, 64),
    "gemm_k2": lambda: case_gemm(128, 256, 128),
    "gemm_k768": lambda: case_gemm(256, 512, 768),
    "gemm_bn128": lambda: case_gemm(256, 128, 256),
    "gemm_bias": lambda: case_gemm(384, 768, 768, bias=True),
    "gemm_gelu": lambda: case_gemm(384, 3072, 768, bias=True, gelu=True),
    "gemm_resid_f32": lambda: case_gemm(384, 768, 3072, bias=True, resid=True, f32=True),
    "gemm_both_stats": lambda: case_gemm(384, 768, 768, bias=True, resid=True, both=True, stats=True),
    "gemm_lnfold": lambda: case_gemm(384, 2304, 768, lnfold=True),
    "gemm_ragged": lambda: case_gemm(197 * 3, 1000, 768, bias=True),
    "gemm_tail10": lambda: case_gemm(100, 10, 192, bias=True, f32=True),
    "gemm_k48pad": lambda: case_gemm(256, 192, 48, lda=64),
    "g2_min": lambda: case_gemm(256, 256, 64, force=2),
    "g2_k768": lambda: case_gemm(512, 512, 768, force=2),
    "g2_3pairs": lambda: case_gemm(768, 768, 256, bias=True, force=2),
    "g2_ragged": lambda: case_gemm(591, 1000, 768, bias=True, force=2),
    "g2_gelu": lambda: case_gemm(1024, 3072, 768, bias=True, gelu=True, force=2),
    "g2_lnfold": lambda: case_gemm(1024, 2304, 768, lnfold=True, force=2),
    "g2_f32": lambda: case_gemm(1024, 768, 768, bias=True, f32=True, force=2),
    "g2_resid": lambda: case_gemm(1100, 768, 3072, bias=True, resid=True, f32=True, force=2),
    "g2_resid_inplace": lambda: case_gemm(2048, 768, 768, bias=True, resid=True, f32=True, force=2, inplace=True),
    "g2_many_tiles": lambda: case_gemm(20000, 768, 512, bias=True, resid=True, f32=True, force=2, inplace=True),
    "g2_dual": lambda: case_gemm(2048, 768, 768, bias=True, resid=True, both=True, stats=True, force=2),
    "gemm_big_out_dual": lambda: case_gemm(100864, 768, 768, bias=True, resid=True, both=True, stats=True),
    "gemm_big_fc2_dual": lambda: case_gemm(100864, 768, 3072, bias=True, resid=True, both=True, stats=True),
    "gemm_big_qkv_v1": lambda: case_gemm(100864, 2304, 768, force=1),
    "gemm_big_qkv": lambda: case_gemm(100864, 2304, 768),
    "gemm_big_fc1": lambda: case_gemm(100864, 3072, 768, bias=True, gelu=True),
    "gemm_big_fc2": lambda: case_gemm(100864, 768, 3072, bias=True, resid=True, f32=True, inplace=True),
    "gemm_big_outproj": lambda: case_gemm(100864, 768, 768, bias=True, resid=True, f32=True, inplace=True),
    "gemm_big_out": lambda: case_gemm(100864, 768, 768, bias=True, resid=True, both=True, stats=True),
    "ln_768": lambda: case_layernorm(1000, 768),
    "ln_nobeta_idx": lambda: case_layernorm(999, 1024, beta=False, idx=True),
    "ln_odd": lambda: case_layernorm(77, 50),
    "patchify_16": lambda: case_patchify(4, 3, 224, 224, 16),
    "patchify_4": lambda: case_patchify(4, 3, 32, 32, 4),
    "patchify_14": lambda: case_patchify(2, 3, 224, 224, 14),
    "embed_cls": lambda: case_embed(4, 196, 1, 768),
    "embed_nocls": lambda: case_embed(4, 64, 0, 192),
    "pool_cast": lambda: case_pool(8, 197, 768),
    "attn_tmem_197": lambda: case_attention(4, 197, 12, psmem=0),
    "attn_tmem_64": lambda: case_attention(4, 64, 3, psmem=0),
    "attn_tmem_257": lambda: case_attention(2, 257, 16, psmem=0),
    "attn_tmem_50": lambda: case_attention(3, 50, 4, psmem=0),
    "attn_tmem_big": lambda: case_attention(512, 197, 12, psmem=0, time_it=True),
    "attn_1cta_big": lambda: case_attention(512, 197, 12, psmem=1, time_it=True),
    "attn_y_split_197": lambda: case_attention(4, 197, 12, pv_split=1),
    "attn_y_split_64": lambda: case_attention(4, 64, 3, pv_split=1),
    "attn_y_split_big": lambda: case_attention(512, 197, 12, time_it=True, pv_split=1),
    "attn_y_split_1cta_big": lambda: case_attention(512, 197, 12, psmem=1, time_it=True, pv_split=1),
    "attn_x_skipmax_big": lambda: case_attention(512, 197, 12, psmem=0, time_it=True, skip_max=1),
    "attn_x_skipmax_1cta_big": lambda: case_attention(512, 197, 12, psmem=1, time_it=True, skip_max=1),
    "attn_1cta_197": lambda: case_attention(4, 197, 12, psmem=1),
    "attn_tmem_129": lambda: case_attention(3, 129, 2, psmem=0),
    "attn_tmem_256": lambda: case_attention(2, 256, 2, psmem=0),
}


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        name = sys.argv[2]
        try:
            res = CASES[name]()
        except Exception as e:  # noqa
            res = {"ok": False, "exception": repr(e), "tb": traceback.format_exc()[-1500:]}
        print("RESULT " + json.dumps(res))
        return
    prefixes = sys.argv[1:]
    names = [n for n in CASES if not prefixes or any(n.startswith(p) for p in prefixes)]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results = {}
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--one", n], capture_output=True, text=True, timeout=180)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                res = json.loads(line[-1][7:])
            else:
                res = {"ok": False, "rc": r.returncode, "stderr": r.stderr[-1200:], "stdout": r.stdout[-400:]}
        except subprocess.TimeoutExpired:
            res = {"ok": False, "timeout": True}
        res["secs"] = round(time.time() - t0, 1)
        results[n] = res
        print(("PASS " if res.get("ok") else "FAIL ") + n + " " + json.dumps(res)[:900], flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "bringup.json"), "w") as f:
            json.dump(results, f, indent=1)
    npass = sum(1 for r in results.values() if r.get("ok"))
    print(f"SUMMARY {npass}/{len(results)} passed")


if __name__ == "__main__":
    main()
