"""the three lines behind the attention in ``NEXT_STAGE_G.forward``
(AttnGAN2/code/model_bert.py:458-461) as one fused operator::

    self.att.applyMask(mask)
    c_code, att = self.att(h_code, word_embs)          # kernel (a)
    h_code = self.adain2(h_code, w_code)               # ADAIN_NORM.forward, model_bert.py:367-374
    h_c_code = torch.cat((h_code, c_code), 1)

``next_stage_attention`` returns ``(h_c_code, att)`` with the same values: the attention forward writes
``c_code`` straight into channels ``[idf, 2 idf)`` of the concatenated buffer (``sba_attn_fwd_into``), an
AdaIN kernel (``sba_adain_fwd``: one read + one write of ``h_code``) writes channels ``[0, idf)``; backward
reads both gradient halves in place and the AdaIN kernel adds its ``dX`` onto the attention's.  The
reference's sequence moves ~10 x ``B*idf*Q`` elements per stage for these three lines; this moves 2 (+ the
attention's own traffic).

``patch_reference_stage(model_bert)`` rebinds ``NEXT_STAGE_G.forward`` to use it; parameters, state_dict
keys and the returned values are unchanged (``att.conv_context``, ``adain2.style``).  Shapes the tcgen05
family does not cover (and CPU tensors never) take the unfused composition of the drop-in module.
"""
from __future__ import annotations

import torch

from . import _abi
from .functional import (_ALGOS, _DTYPES, _MASK_MODES, _context_fp32, _last_algo, _ptr, _require_cuda, _stream,
                         ALGO_NAMES, bwd_workspace, launch_counter)

INSTANCE_NORM_EPS = 1e-5        # nn.InstanceNorm2d default (model_bert.py:364)


def fused_stage_supported(x, context) -> bool:
    if not (x.is_cuda and x.dtype in _DTYPES and x.dim() == 4 and context.dim() == 3):
        return False
    B, idf, ih, iw = x.shape
    lib = _abi.load()
    ok = lambda which: lib.sba_attn_supported(which, _abi.SBA_ALGO_TCGEN05, B, idf, context.shape[1], context.shape[2], ih * iw,
                                              _DTYPES[x.dtype])  # noqa: E731
    return bool(ok(0) and ok(1))


class _AttnAdainCat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, context, weight, style, mask_u8, mask_mode, eps):
        _require_cuda(x, context, weight, style, mask_u8)
        lib = _abi.load()
        x = x.contiguous()
        B, idf, ih, iw = x.shape
        Q = ih * iw
        cdf, L = context.shape[1], context.shape[2]
        dcode = _DTYPES[x.dtype]
        ctx.set_materialize_grads(False)
        ctx32 = _context_fp32(context)
        w32 = weight.detach().reshape(idf, cdf).to(torch.float32).contiguous()
        style32 = style.detach().to(torch.float32).contiguous()
        buf = torch.empty((B, 2 * idf, ih, iw), dtype=x.dtype, device=x.device)
        attn = torch.empty((B, L, ih, iw), dtype=x.dtype, device=x.device)
        srcT = torch.empty((B, idf, L), dtype=torch.float32, device=x.device)
        scratch = torch.empty((3 * B,), dtype=torch.int32, device=x.device)
        stats = torch.empty((2 * B * idf,), dtype=torch.float32, device=x.device)
        rc = lib.sba_attn_fwd_into(_ptr(x), _ptr(ctx32), _ptr(w32), _ptr(mask_u8), _ptr(buf), 2 * idf, idf, _ptr(attn),
                                   _ptr(srcT), _ptr(scratch), B, idf, cdf, L, Q, dcode, mask_mode, _stream())
        _abi.check(rc, "sba_attn_fwd_into")
        n = _abi.last_launch_count()
        _last_algo["fwd"] = ALGO_NAMES.get(lib.sba_last_algo(), "?")
        rc = lib.sba_adain_fwd(_ptr(x), _ptr(style32), _ptr(buf), 2 * idf, 0, _ptr(stats), B, idf, Q, dcode, eps, _stream())
        _abi.check(rc, "sba_adain_fwd")
        launch_counter["n"] += n + _abi.last_launch_count()
        ctx.save_for_backward(x, ctx32, w32, mask_u8, srcT, scratch, style32, stats)
        ctx.meta = (mask_mode, context.dtype, weight.dtype, tuple(weight.shape), style.dtype)
        return buf, attn

    @staticmethod
    def backward(ctx, g_buf, g_attn):
        x, ctx32, w32, mask_u8, srcT, scratch, style32, stats = ctx.saved_tensors
        mask_mode, ctx_dtype, w_dtype, w_shape, style_dtype = ctx.meta
        lib = _abi.load()
        B, idf, ih, iw = x.shape
        Q = ih * iw
        cdf, L = ctx32.shape[1], ctx32.shape[2]
        dcode = _DTYPES[x.dtype]
        need_ctx, need_w = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if g_buf is None:
            g_buf = torch.zeros((B, 2 * idf, ih, iw), dtype=x.dtype, device=x.device)
        g_buf = g_buf.to(x.dtype).contiguous()
        if g_attn is not None:
            g_attn = g_attn.to(x.dtype).contiguous()
        dX = torch.empty_like(x)
        ws = bwd_workspace(B, idf, cdf, L, x.device)
        dW = torch.empty((idf, cdf), dtype=torch.float32, device=x.device) if need_w else None
        dCtx = torch.empty((B, cdf, L), dtype=torch.float32, device=x.device) if need_ctx else None
        d_style = torch.empty_like(style32)
        # the attention's g_c = channels [idf, 2 idf) of the buffer's gradient, read in place
        rc = lib.sba_attn_bwd_from(_ptr(x), _ptr(ctx32), _ptr(w32), _ptr(mask_u8), _ptr(srcT), _ptr(scratch), _ptr(g_buf),
                                   2 * idf, idf, _ptr(g_attn), _ptr(dX), _ptr(ws), ws.numel(), _ptr(dW), _ptr(dCtx), B, idf,
                                   cdf, L, Q, dcode, mask_mode, _stream())
        _abi.check(rc, "sba_attn_bwd_from")
        n = _abi.last_launch_count()
        _last_algo["bwd"] = ALGO_NAMES.get(lib.sba_last_algo(), "?")
        # AdaIN: channels [0, idf), its dX is added onto the attention's
        rc = lib.sba_adain_bwd(_ptr(x), _ptr(style32), _ptr(stats), _ptr(g_buf), 2 * idf, 0, _ptr(dX), 1, _ptr(d_style), B, idf,
                               Q, dcode, _stream())
        _abi.check(rc, "sba_adain_bwd")
        launch_counter["n"] += n + _abi.last_launch_count()
        return (dX if ctx.needs_input_grad[0] else None,
                dCtx.to(ctx_dtype) if need_ctx else None,
                dW.reshape(w_shape).to(w_dtype) if need_w else None,
                d_style.to(style_dtype) if ctx.needs_input_grad[3] else None,
                None, None, None)


def attention_adain_cat(x, context, weight, style, mask=None, mask_mode="reference", eps=INSTANCE_NORM_EPS):
    """x B x idf x ih x iw, context B x cdf x L, weight = conv_context.weight [idf, cdf, 1, 1],
    style B x 2 idf = ADAIN_NORM.style(w_code) (gamma | beta), mask B x L or None.
    Returns (h_c_code B x 2 idf x ih x iw, attn B x L x ih x iw)."""
    if style.dim() != 2 or style.shape[0] != x.shape[0] or style.shape[1] != 2 * x.shape[1]:
        raise RuntimeError("attention_adain_cat: style must be B x 2*idf (gamma | beta)")
    mask_u8 = None
    if mask is not None:
        mask = mask.detach()
        if mask.dim() != 2 or mask.shape[0] != x.shape[0] or mask.shape[1] != context.shape[2]:
            raise RuntimeError(f"attention_adain_cat: mask {tuple(mask.shape)} does not match batch x sourceL")
        mask_u8 = (mask.view(torch.uint8) if mask.dtype == torch.bool and mask.device == x.device and mask.is_contiguous()
                   else mask.to(device=x.device, dtype=torch.uint8).contiguous())
    return _AttnAdainCat.apply(x, context, weight, style, mask_u8, _MASK_MODES[mask_mode], float(eps))


def next_stage_attention(att, adain, h_code, w_code, word_embs):
    """The fused body of NEXT_STAGE_G.forward between ``applyMask`` and ``residual``: ``att`` is the stage's
    GlobalAttentionGeneral (mask already applied), ``adain`` its ADAIN_NORM (``.style`` Linear, ``.norm``
    InstanceNorm2d).  Falls back to the unfused composition where the tcgen05 family does not cover the shape."""
    if fused_stage_supported(h_code, word_embs) and att.algo in ("auto", "tc5"):
        style = adain.style(w_code)                                  # model_bert.py:368 (tiny Linear, stays in torch)
        eps = getattr(getattr(adain, "norm", None), "eps", INSTANCE_NORM_EPS)
        return attention_adain_cat(h_code, word_embs, att.conv_context.weight, style, att.mask, att.mask_mode, eps)
    c_code, att_map = att(h_code, word_embs)
    return torch.cat((adain(h_code, w_code), c_code), 1), att_map


def patch_reference_stage(module) -> None:
    """Rebind ``module.NEXT_STAGE_G.forward`` (model.py:408-423 / model_bert.py:451-468) to the fused body.
    Same signature, same return value ``(out_code, att)``."""
    cls = module.NEXT_STAGE_G

    def forward(self, h_code, c_code, w_code, word_embs, mask):
        self.att.applyMask(mask)
        h_c_code, att = next_stage_attention(self.att, self.adain2, h_code, w_code, word_embs)
        out_code = self.residual(h_c_code)
        out_code = self.upsample(out_code)
        return out_code, att

    forward.__doc__ = cls.forward.__doc__
    cls.forward = forward
