"""Drop-in replacement for AttnGAN2/code/GlobalAttention.py of zhengfei0908/SBA-GAN.

Same names, signatures and state_dict layout, so ``from GlobalAttention import
GlobalAttentionGeneral as ATT_NET`` (model.py:12, model_bert.py:12) and ``from
GlobalAttention import func_attention`` (miscc/losses.py:7) keep working once this module
is first on ``sys.path`` under that name (see ``sba_gan_b200.install``).  The arithmetic
runs in hand-written sm_100a kernels behind libsba_attn.so; there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .functional import word_region_attention
from .losses import func_attention  # noqa: F401  (re-exported: GlobalAttention.py:31-69)


def conv1x1(in_planes, out_planes):
    """1x1 convolution without bias (GlobalAttention.py:25-28)."""
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=1, padding=0, bias=False)


class GlobalAttentionGeneral(nn.Module):
    """Word -> region attention of the generator (GlobalAttention.py:72-121).

    ``conv_context`` stays a real bias-free ``nn.Conv2d(cdf, idf, 1)`` so that reference
    checkpoints load (key ``...att.conv_context.weight`` [idf, cdf, 1, 1]) and
    ``weights_init`` (miscc/utils.py:286-289, matches class names containing "Conv")
    initialises it; only its weight tensor is read by the fused kernel.

    mask_mode "reference" reproduces the reference's ``mask.repeat(queryL, 1)`` row order
    (pixel (b, q) masked with caption (b*Q + q) mod B); "per_sample" uses caption b.
    """

    mask_mode = "reference"
    algo = "auto"        # kernel family of the forward ("auto" | "simt" | "mma" | "tc5"); explicit = strict
    algo_bwd = None      # ... of the backward; None = the same as `algo`

    def __init__(self, idf, cdf):
        super().__init__()
        self.conv_context = conv1x1(cdf, idf)
        self.sm = nn.Softmax(dim=1)   # kept for attribute compatibility; unused
        self.mask = None

    def applyMask(self, mask):
        self.mask = mask  # batch x sourceL, sticky until the next call (GlobalAttention.py:79-80)

    def prepare(self, context, stream=None):
        """Optional (an extension; the reference has no counterpart): run this call's projection
        ``sourceT = conv_context(context)`` now, on ``stream`` (default: the current stream), so that the next
        ``forward(input, context)`` with the same word features, weight and mask launches only the streaming kernel.
        In ``G_NET.forward`` (model_bert.py:580-588) ``word_embs`` and the mask exist before ``h_net1`` has produced
        the first ``h_code``: ``h_net2.att`` / ``h_net3.att`` can be prepared on a side stream under it.  A stale
        preparation (anything changed in between) is ignored.  Call ``applyMask`` first."""
        from .functional import _MASK_MODES, prepare_projection
        mask_u8 = None
        if self.mask is not None:
            m = self.mask.detach()
            mask_u8 = (m.view(torch.uint8) if m.dtype == torch.bool and m.device == context.device and m.is_contiguous()
                       else m.to(device=context.device, dtype=torch.uint8).contiguous())
        self._prepared = (prepare_projection(context, self.conv_context.weight, mask_u8, _MASK_MODES[self.mask_mode], stream)
                          if context.is_cuda and self.algo in ("auto", "tc5") else None)
        self._prepared_mask = (self.mask, mask_u8)
        return self._prepared is not None

    def forward(self, input, context):
        """input: batch x idf x ih x iw (queryL = ih*iw); context: batch x cdf x sourceL
        returns (weightedContext batch x idf x ih x iw, attn batch x sourceL x ih x iw)."""
        prepared, mask = getattr(self, "_prepared", None), self.mask
        if prepared is not None:
            self._prepared = None                                   # one preparation serves one call
            if self._prepared_mask[0] is self.mask:
                mask = self._prepared_mask[1]                       # the very bytes the projection read
            else:
                prepared = None
        return word_region_attention(input, context, self.conv_context.weight, mask,
                                     mask_mode=self.mask_mode, algo=self.algo, algo_bwd=self.algo_bwd, prepared=prepared)
