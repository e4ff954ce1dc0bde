"""`Recorder` with the interface of `vit_pytorch.recorder.Recorder` (reference recorder.py:10-59): returns the logits
and the attention maps `(B, depth, heads, N, N)` of every layer.

The reference finds its layers with `isinstance(module, vit_pytorch.vit.Attention)`, which cannot match this package's
classes, hence this twin: it hooks the `attend` softmax of every attention module of `vit.transformer`.  Attention
maps only exist in the materialised PyTorch graph (the fused kernel never writes the N x N probabilities), so a
recorded call runs that graph -- the model's dispatch sees the hooks; after `eject()` the model is fused again.
"""
from __future__ import annotations

import torch
from torch import nn


class Recorder(nn.Module):
    def __init__(self, vit, device=None) -> None:
        super().__init__()
        self.vit = vit
        self.data = None
        self.recordings = []
        self.hooks = []
        self.hook_registered = False
        self.ejected = False
        self.device = device

    def _hook(self, _, inputs, output) -> None:
        self.recordings.append(output.clone().detach())

    def _register_hook(self) -> None:
        for layer in self.vit.transformer.layers:
            attn = layer[0]
            assert hasattr(attn, "attend"), "attention module without an `attend` softmax"
            self.hooks.append(attn.attend.register_forward_hook(self._hook))
        self.hook_registered = True

    def eject(self):
        self.ejected = True
        for hook in self.hooks:
            hook.remove()
        self.hooks.clear()
        return self.vit

    def clear(self) -> None:
        self.recordings.clear()

    def record(self, attn: torch.Tensor) -> None:
        self.recordings.append(attn.clone().detach())

    def forward(self, img):
        assert not self.ejected, 'recorder has been ejected, cannot be used anymore'
        self.clear()
        if not self.hook_registered:
            self._register_hook()
        pred = self.vit(img)
        target = self.device if self.device is not None else img.device
        recordings = tuple(t.to(target) for t in self.recordings)
        attns = torch.stack(recordings, dim=1) if len(recordings) > 0 else None
        return pred, attns
