"""Drop-in ``func_attention``, ``words_loss`` and ``sent_loss`` (AttnGAN2/code/GlobalAttention.py:31-69,
AttnGAN2/code/miscc/losses.py:20-59, 62-132) over the fused sm_100a kernel (c) and the matching-tail
kernels (csrc/match_loss.cu).

``words_loss`` keeps the reference's signature and return value
``(loss0, loss1, att_maps)`` including the ``labels=None`` and ``class_ids=None`` modes
(trainer_bert.py:205-208).  The gammas come from the reference's global ``cfg`` when its
``miscc.config`` module is imported (losses.py:91, 106, 123), else from keyword arguments
(defaults of miscc/config.py:43-45).

Nothing here synchronises with the host in the training path: the reference's
``cap_lens.data.tolist()`` (losses.py:71) is deferred until somebody actually indexes ``att_maps``
(``LazyAttMaps``), the class mask is derived on the device from the class ids instead of numpy + H2D
(losses.py:73-76, 116-121), and the two cross-entropies run in one fused kernel pair.
"""
from __future__ import annotations

import sys

import torch
import torch.nn.functional as F

from . import _abi
from .functional import _ptr, _require_cuda, _stream, launch_counter

DEFAULT_GAMMAS = (5.0, 5.0, 10.0)   # miscc/config.py:43-45
FORWARD_ALGO = "auto"               # "auto": tensor-core forward where the shape is covered; "simt": always the CUDA-core kernel


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
        ws_bytes = lib.sba_words_sim_fwd_workspace_bytes(B_img, B_cap, nef, R, Lw) if FORWARD_ALGO != "simt" else 0
        if ws_bytes:        # tensor-core forward (tcgen05, 3xTF32)
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=img.device)
            rc = lib.sba_words_sim_fwd_ws(_ptr(img32), _ptr(words32), _ptr(cap_lens_i32), _ptr(sim), _ptr(att), _ptr(ws), ws_bytes,
                                          B_img, B_cap, row_offset, nef, R, Lw, gammas[0], gammas[1], gammas[2], eps, _stream())
            _abi.check(rc, "sba_words_sim_fwd_ws")
        else:               # CUDA-core forward (shapes the tensor-core kernel does not cover)
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
        tc_bytes = (lib.sba_words_sim_bwd_tc_workspace_bytes(B_img, B_cap, nef, R, Lw, int(need_words))
                    if FORWARD_ALGO != "simt" else 0)
        if tc_bytes:        # tensor-core backward: d_img, and d_words unless the words are detached (GAN training)
            ws = torch.empty((tc_bytes,), dtype=torch.uint8, device=img32.device)
            d_img = torch.empty_like(img32)
            d_words = torch.empty_like(words32) if need_words else None
            rc = lib.sba_words_sim_bwd_tc(_ptr(img32), _ptr(words32), _ptr(cap_lens_i32), _ptr(d_sim), _ptr(d_img), _ptr(d_words),
                                          _ptr(ws), tc_bytes, B_img, B_cap, row_offset, nef, R, Lw, gammas[0], gammas[1],
                                          gammas[2], eps, _stream())
            _abi.check(rc, "sba_words_sim_bwd_tc")
            launch_counter["n"] += _abi.last_launch_count()
            return (d_img.to(img_dtype) if need_img else None, d_words.to(words_dtype) if need_words else None,
                    None, None, None, None, None)
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


class LazyAttMaps:
    """The ``att_maps`` list of words_loss (losses.py:92: B tensors 1 x T_i x ih x iw) without the
    device->host synchronisation the slicing by caption length needs: the lengths are fetched
    (``.tolist()``) the first time an element is looked at.  The training loops discard
    ``att_maps`` (trainer_bert.py:294-296 via generator_loss, pretrain_DAMSM_bert.py:80-81), so there
    the step never waits for the GPU; ``save_img_results`` (trainer_bert.py:205-216) indexes it and
    gets ordinary tensors."""

    def __init__(self, att, cap_lens, ih, iw):
        self._att, self._lens, self._hw, self._items = att, cap_lens, (ih, iw), None

    def _materialise(self):
        if self._items is None:
            lens = self._lens.tolist() if torch.is_tensor(self._lens) else list(self._lens)       # losses.py:71
            B, (ih, iw) = self._att.shape[0], self._hw
            self._items = [self._att[i, :int(lens[i])].reshape(1, int(lens[i]), ih, iw) for i in range(B)]
        return self._items

    def __len__(self):
        return self._att.shape[0]

    def __getitem__(self, i):
        return self._materialise()[i]

    def __iter__(self):
        return iter(self._materialise())


class _MatchCrossEntropy(torch.autograd.Function):
    """(loss0, loss1) = (CE(scores, labels), CE(scores^T, labels)) with same-class masking
    (losses.py:53-59 / 124-129), one fused kernel pair; backward one kernel."""

    @staticmethod
    def forward(ctx, scores, class_ids_i32, labels_i64):
        _require_cuda(scores, class_ids_i32, labels_i64)
        lib = _abi.load()
        B = scores.shape[0]
        s32 = scores.detach().to(torch.float32).contiguous()
        losses = torch.empty((2,), dtype=torch.float32, device=scores.device)
        lse = torch.empty((4 * B,), dtype=torch.float32, device=scores.device)
        rc = lib.sba_match_ce_fwd(_ptr(s32), _ptr(class_ids_i32), _ptr(labels_i64), _ptr(losses), _ptr(lse), B, _stream())
        _abi.check(rc, "sba_match_ce_fwd")
        launch_counter["n"] += _abi.last_launch_count()
        ctx.save_for_backward(s32, class_ids_i32, labels_i64, lse)
        ctx.dtype = scores.dtype
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g0, g1):
        s32, cls, labels, lse = ctx.saved_tensors
        lib = _abi.load()
        B = s32.shape[0]
        zero = torch.zeros((), dtype=torch.float32, device=s32.device)
        g = torch.stack([zero if g0 is None else g0.to(torch.float32), zero if g1 is None else g1.to(torch.float32)])
        d = torch.empty_like(s32)
        rc = lib.sba_match_ce_bwd(_ptr(s32), _ptr(cls), _ptr(labels), _ptr(lse), _ptr(g), _ptr(d), B, _stream())
        _abi.check(rc, "sba_match_ce_bwd")
        launch_counter["n"] += _abi.last_launch_count()
        return d.to(ctx.dtype), None, None


def match_cross_entropy(scores, labels, class_ids=None):
    """The tail shared by words_loss and sent_loss: entry (i, j), i != j, of the B x B ``scores`` counts as
    -inf when class_ids[i] == class_ids[j]; returns (CE(scores, labels), CE(scores^T, labels))."""
    if scores.dim() != 2 or scores.shape[0] != scores.shape[1]:
        raise RuntimeError("match_cross_entropy: scores must be B x B")
    cls = None
    if class_ids is not None:
        cls = torch.as_tensor(class_ids).to(device=scores.device, dtype=torch.int32).contiguous()
    lab = labels.detach().to(device=scores.device, dtype=torch.int64).contiguous()
    return _MatchCrossEntropy.apply(scores, cls, lab)


class _SentScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cnn, rnn, gamma3, eps):
        _require_cuda(cnn, rnn)
        lib = _abi.load()
        B, nef = cnn.shape
        a = cnn.detach().to(torch.float32).contiguous()
        b = rnn.detach().to(torch.float32).contiguous()
        scores = torch.empty((B, B), dtype=torch.float32, device=cnn.device)
        norms = torch.empty((2 * B,), dtype=torch.float32, device=cnn.device)
        rc = lib.sba_sent_scores_fwd(_ptr(a), _ptr(b), _ptr(scores), _ptr(norms), B, nef, gamma3, eps, _stream())
        _abi.check(rc, "sba_sent_scores_fwd")
        launch_counter["n"] += _abi.last_launch_count()
        ctx.save_for_backward(a, b, norms, scores)
        ctx.meta = (gamma3, eps, cnn.dtype, rnn.dtype)
        return scores

    @staticmethod
    def backward(ctx, d_scores):
        a, b, norms, scores = ctx.saved_tensors
        gamma3, eps, dt_a, dt_b = ctx.meta
        lib = _abi.load()
        B, nef = a.shape
        d = d_scores.to(torch.float32).contiguous()
        d_a, d_b = torch.empty_like(a), torch.empty_like(b)
        rc = lib.sba_sent_scores_bwd(_ptr(a), _ptr(b), _ptr(norms), _ptr(scores), _ptr(d), _ptr(d_a), _ptr(d_b), B, nef,
                                     gamma3, eps, _stream())
        _abi.check(rc, "sba_sent_scores_bwd")
        launch_counter["n"] += _abi.last_launch_count()
        return d_a.to(dt_a), d_b.to(dt_b), None, None


def sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps=1e-8, gamma3=None):
    """sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps=1e-8) -> (loss0, loss1), same as
    miscc/losses.py:20-59: cosine scores x gamma3, same-class masking, two cross-entropies.
    cnn_code, rnn_code: batch x nef."""
    g3 = _cfg_gammas()[2] if gamma3 is None else gamma3
    if cnn_code.dim() == 3:                 # the reference also accepts 1 x batch x nef (losses.py:37-39)
        cnn_code, rnn_code = cnn_code.squeeze(0), rnn_code.squeeze(0)
    if cnn_code.dim() != 2 or cnn_code.shape != rnn_code.shape:
        raise RuntimeError("sent_loss: cnn_code and rnn_code must both be batch x nef")
    if labels is None:
        return None, None
    scores = _SentScores.apply(cnn_code, rnn_code, float(g3), float(eps))
    return match_cross_entropy(scores, labels, class_ids)


def class_mask(class_ids, device):
    """masks[i, j] = class_ids[i] == class_ids[j] and i != j (losses.py:73-76, 116-121),
    built on the device (the reference builds it in numpy and copies it over)."""
    cid = torch.as_tensor(class_ids).to(device)
    same = cid[:, None] == cid[None, :]
    same.fill_diagonal_(False)
    return same


def words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size, gamma1=None, gamma2=None,
               gamma3=None, eps=1e-8, att_maps=True):
    """words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size)
    -> (loss0, loss1, att_maps), same as miscc/losses.py:62-132.

    words_emb: batch x nef x seq_len; img_features: batch x nef x 17 x 17.
    att_maps: sequence of batch tensors 1 x cap_len_i x 17 x 17 (losses.py:92), sliced lazily
    (``LazyAttMaps``: no host synchronisation unless it is indexed); ``att_maps=False`` (an extension:
    the training loops discard them) skips writing the maps altogether and returns None in their place.
    """
    g = _cfg_gammas()
    g1 = g[0] if gamma1 is None else gamma1
    g2 = g[1] if gamma2 is None else gamma2
    g3 = g[2] if gamma3 is None else gamma3
    B = int(batch_size)
    img = img_features[:B]
    words = words_emb[:B]
    lens = cap_lens[:B] if torch.is_tensor(cap_lens) else list(cap_lens)[:B]
    ih, iw = img.shape[2], img.shape[3]
    want = bool(att_maps) or labels is None
    res = words_similarity(img, words, lens, g1, g2, g3, eps, 0, want_att_maps=want)
    sim, att = res if want else (res, None)
    maps = LazyAttMaps(att, lens, ih, iw) if want else None
    if labels is None:
        return None, None, maps
    loss0, loss1 = match_cross_entropy(sim, labels, class_ids)     # losses.py:116-129
    return loss0, loss1, maps


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
