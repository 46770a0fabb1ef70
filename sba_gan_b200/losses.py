"""Drop-in ``func_attention`` and ``words_loss`` (AttnGAN2/code/GlobalAttention.py:31-69,
AttnGAN2/code/miscc/losses.py:62-132) over the fused sm_100a kernel (c).

``words_loss`` keeps the reference's signature and return value
``(loss0, loss1, att_maps)`` including the ``labels=None`` and ``class_ids=None`` modes
(trainer_bert.py:205-208).  The gammas come from the reference's global ``cfg`` when its
``miscc.config`` module is imported (losses.py:91, 106, 123), else from keyword arguments
(defaults of miscc/config.py:43-45).
"""
from __future__ import annotations

import sys

import torch
import torch.nn.functional as F

from . import _abi
from .functional import _ptr, _require_cuda, _stream, launch_counter

DEFAULT_GAMMAS = (5.0, 5.0, 10.0)   # miscc/config.py:43-45


def _cfg_gammas():
    cfgmod = sys.modules.get("miscc.config")
    if cfgmod is not None and hasattr(cfgmod, "cfg"):
        s = cfgmod.cfg.TRAIN.SMOOTH
        return float(s.GAMMA1), float(s.GAMMA2), float(s.GAMMA3)
    return DEFAULT_GAMMAS


class _WordsSimilarity(torch.autograd.Function):
    """sim[j, i] for local image rows j against all captions i; optionally the diagonal
    attention maps.  Backward gives d_img and (if needed) d_words."""

    @staticmethod
    def forward(ctx, img, words, cap_lens_i32, gammas, eps, row_offset, want_att):
        _require_cuda(img, words, cap_lens_i32)
        lib = _abi.load()
        B_img, nef, ih, iw = img.shape
        B_cap, _, Lw = words.shape
        R = ih * iw
        img32 = img.detach().to(torch.float32).contiguous()
        words32 = words.detach().to(torch.float32).contiguous()
        sim = torch.empty((B_img, B_cap), dtype=torch.float32, device=img.device)
        att = torch.zeros((B_cap, Lw, R), dtype=torch.float32, device=img.device) if want_att else None
        rc = lib.sba_words_sim_fwd(_ptr(img32), _ptr(words32), _ptr(cap_lens_i32), _ptr(sim), _ptr(att), B_img, B_cap,
                                   row_offset, nef, R, Lw, gammas[0], gammas[1], gammas[2], eps, _stream())
        _abi.check(rc, "sba_words_sim_fwd")
        launch_counter["n"] += _abi.last_launch_count()
        ctx.save_for_backward(img32, words32, cap_lens_i32)
        ctx.meta = (gammas, eps, row_offset, img.dtype, words.dtype, tuple(img.shape))
        if want_att:
            ctx.mark_non_differentiable(att)
            return sim, att
        return sim, torch.empty(0, device=img.device)

    @staticmethod
    def backward(ctx, d_sim, _d_att):
        img32, words32, cap_lens_i32 = ctx.saved_tensors
        gammas, eps, row_offset, img_dtype, words_dtype, img_shape = ctx.meta
        lib = _abi.load()
        B_img, nef, R = img32.shape[0], img32.shape[1], img32.shape[2] * img32.shape[3]
        B_cap, _, Lw = words32.shape
        need_img, need_words = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_img or need_words):
            return None, None, None, None, None, None, None
        d_sim = d_sim.to(torch.float32).contiguous()
        nbytes = lib.sba_words_sim_bwd_workspace_bytes(B_img, B_cap, nef, R, Lw)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=img32.device)
        d_img = torch.empty_like(img32)
        d_words = torch.empty_like(words32) if need_words else None
        rc = lib.sba_words_sim_bwd(_ptr(img32), _ptr(words32), _ptr(cap_lens_i32), _ptr(d_sim), _ptr(d_img),
                                   _ptr(d_words), _ptr(ws), B_img, B_cap, row_offset, nef, R, Lw, gammas[0], gammas[1],
                                   gammas[2], eps, _stream())
        _abi.check(rc, "sba_words_sim_bwd")
        launch_counter["n"] += _abi.last_launch_count()
        return (d_img.to(img_dtype) if need_img else None, d_words.to(words_dtype) if need_words else None,
                None, None, None, None, None)


def _lens_tensor(cap_lens, device):
    """cap_lens may be a CUDA/CPU tensor or a list (losses.py:71 calls .data.tolist())."""
    if torch.is_tensor(cap_lens):
        return cap_lens.detach().to(device=device, dtype=torch.int32).contiguous()
    return torch.as_tensor(list(cap_lens), dtype=torch.int32, device=device)


def words_similarity(img_features, words_emb, cap_lens, gamma1, gamma2, gamma3, eps=1e-8, row_offset=0,
                     want_att_maps=False):
    """The (local rows of the) B x B caption-image similarity matrix of words_loss before
    class masking: rows = images, cols = captions (losses.py:115, 123).  ``row_offset`` is the
    global index of local image row 0 when images are sharded by rows across ranks."""
    if img_features.dim() != 4 or words_emb.dim() != 3 or img_features.shape[1] != words_emb.shape[1]:
        raise RuntimeError("words_similarity: img_features must be B x nef x ih x iw and words_emb B x nef x L")
    lens = _lens_tensor(cap_lens, img_features.device)
    if lens.numel() != words_emb.shape[0]:
        raise RuntimeError("words_similarity: cap_lens must have one entry per caption")
    sim, att = _WordsSimilarity.apply(img_features, words_emb, lens, (float(gamma1), float(gamma2), float(gamma3)),
                                      float(eps), int(row_offset), bool(want_att_maps))
    return (sim, att) if want_att_maps else sim


def class_mask(class_ids, device):
    """masks[i, j] = class_ids[i] == class_ids[j] and i != j (losses.py:73-76, 116-121),
    built on the device (the reference builds it in numpy and copies it over)."""
    cid = torch.as_tensor(class_ids).to(device)
    same = cid[:, None] == cid[None, :]
    same.fill_diagonal_(False)
    return same


def words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size, gamma1=None, gamma2=None,
               gamma3=None, eps=1e-8):
    """words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size)
    -> (loss0, loss1, att_maps), same as miscc/losses.py:62-132.

    words_emb: batch x nef x seq_len; img_features: batch x nef x 17 x 17.
    att_maps: list of batch tensors 1 x cap_len_i x 17 x 17 (losses.py:92).
    """
    g = _cfg_gammas()
    g1 = g[0] if gamma1 is None else gamma1
    g2 = g[1] if gamma2 is None else gamma2
    g3 = g[2] if gamma3 is None else gamma3
    B = int(batch_size)
    img = img_features[:B]
    words = words_emb[:B]
    lens_host = cap_lens.detach().tolist() if torch.is_tensor(cap_lens) else list(cap_lens)   # losses.py:71
    lens_host = [int(v) for v in lens_host[:B]]
    ih, iw = img.shape[2], img.shape[3]
    sim, att = words_similarity(img, words, lens_host, g1, g2, g3, eps, 0, want_att_maps=True)
    att_maps = [att[i, :lens_host[i]].reshape(1, lens_host[i], ih, iw) for i in range(B)]
    if labels is None:
        return None, None, att_maps
    if class_ids is not None:
        sim = sim.masked_fill(class_mask(class_ids, sim.device), float("-inf"))     # losses.py:124-125
    loss0 = F.cross_entropy(sim, labels)                    # losses.py:128
    loss1 = F.cross_entropy(sim.transpose(0, 1), labels)    # losses.py:129
    return loss0, loss1, att_maps


def func_attention(query, context, gamma1):
    """func_attention(query, context, gamma1) -> (weightedContext, attn)
    (GlobalAttention.py:31-69).  query: batch x ndf x queryL; context: batch x ndf x ih x iw.
    Forward only (the training path differentiates through ``words_loss``)."""
    _require_cuda(query, context)
    lib = _abi.load()
    B, nef, T = query.shape
    ih, iw = context.shape[2], context.shape[3]
    R = ih * iw
    q32 = query.detach().to(torch.float32).contiguous()
    c32 = context.detach().to(torch.float32).contiguous()
    wc = torch.empty((B, nef, T), dtype=torch.float32, device=query.device)
    attn = torch.empty((B, T, R), dtype=torch.float32, device=query.device)
    rc = lib.sba_func_attention(_ptr(q32), _ptr(c32), _ptr(wc), _ptr(attn), B, nef, T, R, float(gamma1), _stream())
    _abi.check(rc, "sba_func_attention")
    launch_counter["n"] += _abi.last_launch_count()
    return wc.to(query.dtype), attn.reshape(B, T, ih, iw).to(query.dtype)
