"""CPU baseline of the attention hot path for timing: the same torch op sequence the
reference executes (two bmm around a masked softmax, autograd backward), as opposed to
``oracle/attention.py`` whose closed forms are written for checking, not for speed.
Follows AttnGAN2/code/GlobalAttention.py:82-121.  Checked against oracle.attention in
tests/test_oracle.py.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def attn_fwd_bwd_autograd(x, context, weight, mask, g_c):
    """One forward+backward of the generator attention on CPU with torch autograd.
    x B x idf x ih x iw (requires no grad flag needed), weight [idf, cdf, 1, 1].
    Returns (c_code, attn, dX, dW)."""
    B, idf, ih, iw = x.shape
    Q = ih * iw
    L = context.shape[2]
    x = x.detach().requires_grad_(True)
    weight = weight.detach().reshape(idf, -1, 1, 1).requires_grad_(True)
    tgt = x.view(B, idf, Q).transpose(1, 2).contiguous()            # :92-93
    src = F.conv2d(context.unsqueeze(3), weight).squeeze(3)         # :95-97
    s = torch.bmm(tgt, src).view(B * Q, L)                          # :102-104
    if mask is not None:
        s.data.masked_fill_(mask.repeat(Q, 1), float("-inf"))       # :105-108
    p = torch.softmax(s, dim=1).view(B, Q, L).transpose(1, 2).contiguous()   # :109-113
    c = torch.bmm(src, p).view(B, idf, ih, iw)                      # :117-118
    c.backward(g_c)
    return c.detach(), p.detach().view(B, L, ih, iw), x.grad, weight.grad
