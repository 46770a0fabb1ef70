"""Closed-form restatement (torch, CPU, any float dtype) of the reference hot path.

Every function cites the reference lines it restates (paths relative to
``/root/reference/AttnGAN2/code``).  Written from the maths, not from the source: einsum
contractions and explicit hand-derived backward formulas instead of the reference's
transpose/bmm/autograd sequence, so that agreement with the reference (checked against
``tests/golden``) is evidence and not tautology.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.
"""
from __future__ import annotations

import torch

NEG_INF = float("-inf")


def _w2d(weight: torch.Tensor) -> torch.Tensor:
    """conv1x1 weight [idf, cdf, 1, 1] (GlobalAttention.py:25-28) as an idf x cdf matrix."""
    return weight.reshape(weight.shape[0], weight.shape[1])


def project_words(context: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """sourceT[b] = W . context[b]: the bias-free 1x1 conv of GlobalAttention.py:95-97.

    context B x cdf x L, weight idf x cdf (x1x1)  ->  B x idf x L
    """
    return torch.einsum("ic,bcl->bil", _w2d(weight), context)


def pixel_mask(mask: torch.Tensor, B: int, Q: int, mask_mode: str = "reference") -> torch.Tensor:
    """Expand a B x L caption padding mask to B x Q x L per-pixel masks.

    ``reference``: GlobalAttention.py:104-108 does ``mask.repeat(queryL, 1)`` on a
    (B*Q) x L score matrix whose row index is r = b*Q + q, so row r receives mask[r mod B]:
    pixel (b, q) is masked with caption ``(b*Q + q) mod B``.
    ``per_sample``: the presumably intended behaviour, pixel (b, q) uses caption b.
    """
    mask = mask.to(torch.bool)
    if mask_mode == "reference":
        rows = (torch.arange(B * Q, device=mask.device) % B).view(B, Q)
        return mask[rows]
    if mask_mode == "per_sample":
        return mask[:, None, :].expand(B, Q, mask.shape[1])
    raise ValueError(f"unknown mask_mode {mask_mode!r}")


def attn_forward(x, context, weight, mask=None, mask_mode="reference"):
    """GlobalAttentionGeneral.forward (GlobalAttention.py:82-121).

    x B x idf x ih x iw, context B x cdf x L, weight idf x cdf (x1x1), mask B x L bool|None
    returns (c_code B x idf x ih x iw, attn B x L x ih x iw, sourceT B x idf x L)
    """
    B, idf, ih, iw = x.shape
    Q = ih * iw
    L = context.shape[2]
    src = project_words(context, weight)                      # :95-97
    xf = x.reshape(B, idf, Q)
    S = torch.einsum("biq,bil->bql", xf, src)                 # :102
    if mask is not None:
        S = S.masked_fill(pixel_mask(mask, B, Q, mask_mode), NEG_INF)   # :104-108
    P = torch.softmax(S, dim=2)                               # :109
    c = torch.einsum("bil,bql->biq", src, P)                  # :117
    return c.reshape(B, idf, ih, iw), P.transpose(1, 2).reshape(B, L, ih, iw), src


def attn_backward(x, context, weight, mask, g_c, g_attn=None, mask_mode="reference"):
    """Hand-derived backward of GlobalAttentionGeneral.forward (what autograd does for
    GlobalAttention.py:92-121; formulas of SURVEY.md §8a-4).

    returns (dX like x, dW idf x cdf, dContext like context, dSrc B x idf x L)
    """
    B, idf, ih, iw = x.shape
    Q = ih * iw
    L = context.shape[2]
    W = _w2d(weight)
    src = project_words(context, weight)
    xf = x.reshape(B, idf, Q)
    S = torch.einsum("biq,bil->bql", xf, src)
    if mask is not None:
        S = S.masked_fill(pixel_mask(mask, B, Q, mask_mode), NEG_INF)
    P = torch.softmax(S, dim=2)                               # B x Q x L
    gc = g_c.reshape(B, idf, Q)
    dP = torch.einsum("bil,biq->bql", src, gc)                # through :117
    if g_attn is not None:
        dP = dP + g_attn.reshape(B, L, Q).transpose(1, 2)     # through :111-119
    dS = P * (dP - (P * dP).sum(dim=2, keepdim=True))         # softmax :109 (masked: P = 0)
    dX = torch.einsum("bql,bil->biq", dS, src)                # through :102
    dSrc = torch.einsum("biq,bql->bil", gc, P) + torch.einsum("biq,bql->bil", xf, dS)
    dW = torch.einsum("bil,bcl->ic", dSrc, context)           # conv_context weight grad
    dCtx = torch.einsum("ic,bil->bcl", W, dSrc)               # conv_context input grad
    return dX.reshape(x.shape), dW.reshape(weight.shape), dCtx, dSrc


def func_attention(query, context, gamma1):
    """func_attention (GlobalAttention.py:31-69).

    query B x nef x T (words), context B x nef x ih x iw (regions)
    returns (weightedContext B x nef x T, attn B x T x ih x iw)
    """
    B, nef, T = query.shape
    ih, iw = context.shape[2], context.shape[3]
    ctx = context.reshape(B, nef, ih * iw)
    s = torch.einsum("bcr,bct->brt", ctx, query)              # :48
    a1 = torch.softmax(s, dim=2)                              # :50-51  over words
    a2 = torch.softmax(gamma1 * a1.transpose(1, 2), dim=2)    # :56-60  over regions, B x T x R
    wc = torch.einsum("bcr,btr->bct", ctx, a2)                # :67
    return wc, a2.reshape(B, T, ih, iw)


def _pair_terms(img, word, gamma1, gamma2, eps):
    """All images against ONE caption's words.  img Bi x C x R, word C x T.
    Returns dict of the intermediates of losses.py:91-108 for that caption."""
    s = torch.einsum("jcr,ct->jrt", img, word)                # GlobalAttention.py:48
    a1 = torch.softmax(s, dim=2)                              # :51
    a2 = torch.softmax(gamma1 * a1.transpose(1, 2), dim=2)    # :59-60   Bi x T x R
    wc = torch.einsum("jcr,jtr->jct", img, a2)                # :67      Bi x C x T
    num = torch.einsum("ct,jct->jt", word, wc)                # losses.py:14
    ww = word.norm(dim=0)                                     # :15   T
    wn = wc.norm(dim=1)                                       # :16   Bi x T
    den = (ww[None, :] * wn).clamp(min=eps)                   # :17
    cos = num / den
    e = torch.exp(gamma2 * cos)                               # :106
    E = e.sum(dim=1)                                          # :107
    return dict(s=s, a1=a1, a2=a2, wc=wc, num=num, ww=ww, wn=wn, den=den, cos=cos, e=e, E=E)


def words_similarity(img_features, words_emb, cap_lens, gamma1, gamma2, gamma3, eps=1e-8,
                     want_att_maps=False, row_offset=0):
    """The B_img x B_cap similarity matrix of words_loss before class masking
    (losses.py:72-123): sim[j, i] = gamma3 * log sum_t exp(gamma2 * cos(word_it, wc_jit)).

    img_features Bi x nef x ih x iw, words_emb Bc x nef x Lmax, cap_lens list[int] (Bc).
    ``row_offset``: global index of local image row 0 (row-sharded evaluation), only used
    to pick the diagonal attention maps.
    """
    Bi, nef, ih, iw = img_features.shape
    Bc = words_emb.shape[0]
    img = img_features.reshape(Bi, nef, ih * iw)
    cols, att_maps = [], []
    for i in range(Bc):
        T = int(cap_lens[i])
        word = words_emb[i, :, :T]                            # losses.py:78-80
        t = _pair_terms(img, word, gamma1, gamma2, eps)
        cols.append(gamma3 * torch.log(t["E"]))               # :108, :123
        if want_att_maps and row_offset <= i < row_offset + Bi:
            att_maps.append(t["a2"][i - row_offset].reshape(1, T, ih, iw))   # :92
    sim = torch.stack(cols, dim=1)                            # :115  rows = images, cols = captions
    return (sim, att_maps) if want_att_maps else sim


def ce_tail(sim, labels, class_ids):
    """Class masking + the two cross-entropies (losses.py:73-76, 116-129).

    Entry (j, i), j != i, is set to -inf when class_ids[j] == class_ids[i]."""
    if class_ids is not None:
        cid = torch.as_tensor(class_ids, device=sim.device)
        same = cid[:, None] == cid[None, :]
        same.fill_diagonal_(False)
        sim = sim.masked_fill(same, NEG_INF)
    if labels is None:
        return None, None, sim
    lse0 = torch.logsumexp(sim, dim=1)
    lse1 = torch.logsumexp(sim, dim=0)
    idx = torch.arange(sim.shape[0], device=sim.device)
    picked = sim[idx, labels]
    loss0 = (lse0 - picked).mean()                            # CE(sim, labels)         :128
    loss1 = (lse1 - sim[labels, idx]).mean()                  # CE(sim^T, labels)       :129
    return loss0, loss1, sim


def adain_cat(h_code, style, c_code, eps=1e-5):
    """ADAIN_NORM.forward (model_bert.py:367-374) followed by the concatenation of NEXT_STAGE_G.forward (:460-461):
    cat(((gamma + 1) * InstanceNorm2d(h_code) + beta, c_code), 1) with style = (gamma | beta) B x 2C."""
    gamma, beta = style[:, :, None, None].chunk(2, 1)                       # :368-369
    mean = h_code.mean(dim=(2, 3), keepdim=True)
    var = h_code.var(dim=(2, 3), unbiased=False, keepdim=True)              # nn.InstanceNorm2d: biased variance
    out = (gamma + 1.0) * ((h_code - mean) / torch.sqrt(var + eps)) + beta  # :371-372
    return torch.cat((out, c_code), 1)                                      # :461


def sent_scores(cnn_code, rnn_code, gamma3=10.0, eps=1e-8):
    """The B x B score matrix of sent_loss (losses.py:42-49) before class masking:
    scores[i, j] = gamma3 * <cnn_i, rnn_j> / max(|cnn_i| |rnn_j|, eps)."""
    n0 = cnn_code.norm(dim=1, keepdim=True)                     # :42
    n1 = rnn_code.norm(dim=1, keepdim=True)                     # :43
    scores = cnn_code @ rnn_code.t()                            # :45
    return scores / (n0 @ n1.t()).clamp(min=eps) * gamma3       # :46-47


def sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps=1e-8, gamma3=10.0):
    """sent_loss (losses.py:20-59) with gamma3 explicit; returns (loss0, loss1)."""
    scores = sent_scores(cnn_code[:batch_size], rnn_code[:batch_size], gamma3, eps)
    loss0, loss1, _ = ce_tail(scores, labels, class_ids)        # :24-34, 51-59
    return loss0, loss1


def words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size,
               gamma1=4.0, gamma2=5.0, gamma3=10.0, eps=1e-8):
    """words_loss (losses.py:62-132) with the gammas as explicit arguments
    (reference reads cfg.TRAIN.SMOOTH.GAMMA1/2/3).  Returns (loss0, loss1, att_maps)."""
    lens = cap_lens.tolist() if hasattr(cap_lens, "tolist") else list(cap_lens)
    sim, att_maps = words_similarity(img_features[:batch_size], words_emb[:batch_size], lens[:batch_size],
                                     gamma1, gamma2, gamma3, eps, want_att_maps=True)
    loss0, loss1, _ = ce_tail(sim, labels, class_ids)
    return loss0, loss1, att_maps


def words_loss_backward(img_features, words_emb, cap_lens, d_sim, gamma1, gamma2, gamma3, eps=1e-8):
    """Hand-derived gradient of ``words_similarity`` w.r.t. both inputs given d_sim
    (Bi x Bc).  This is what autograd produces for losses.py:72-123 +
    GlobalAttention.py:31-69.  Returns (d_img like img_features, d_words like words_emb).
    """
    Bi, nef, ih, iw = img_features.shape
    Bc, _, Lmax = words_emb.shape
    img = img_features.reshape(Bi, nef, ih * iw)
    d_img = torch.zeros_like(img)
    d_words = torch.zeros_like(words_emb)
    for i in range(Bc):
        T = int(cap_lens[i])
        word = words_emb[i, :, :T]
        t = _pair_terms(img, word, gamma1, gamma2, eps)
        g = d_sim[:, i]                                                   # Bi
        gcos = (g * gamma3 / t["E"])[:, None] * gamma2 * t["e"]           # d/dcos   Bi x T
        live = (t["ww"][None, :] * t["wn"]) > eps                         # clamp(min=eps) gate
        inv_den = 1.0 / t["den"]
        d_num = gcos * inv_den
        # d den = -num/den^2 * gcos where the clamp is inactive, den = ww * wn
        d_den = torch.where(live, -gcos * t["num"] * inv_den * inv_den, torch.zeros_like(gcos))
        d_wn = d_den * t["ww"][None, :]
        d_ww = (d_den * t["wn"]).sum(dim=0)                               # T
        safe_wn = torch.where(t["wn"] > 0, t["wn"], torch.ones_like(t["wn"]))
        safe_ww = torch.where(t["ww"] > 0, t["ww"], torch.ones_like(t["ww"]))
        d_wc = d_num[:, None, :] * word[None] + (d_wn / safe_wn)[:, None, :] * t["wc"]   # Bi x C x T
        d_word = torch.einsum("jt,jct->ct", d_num, t["wc"]) + (d_ww / safe_ww)[None, :] * word
        d_a2 = torch.einsum("jcr,jct->jtr", img, d_wc)                    # through :67
        d_img += torch.einsum("jct,jtr->jcr", d_wc, t["a2"])
        d_z = t["a2"] * (d_a2 - (t["a2"] * d_a2).sum(dim=2, keepdim=True))  # softmax over regions
        d_a1 = gamma1 * d_z.transpose(1, 2)                               # Bi x R x T
        d_s = t["a1"] * (d_a1 - (t["a1"] * d_a1).sum(dim=2, keepdim=True))  # softmax over words
        d_img += torch.einsum("ct,jrt->jcr", word, d_s)                   # through :48
        d_word = d_word + torch.einsum("jcr,jrt->ct", img, d_s)
        d_words[i, :, :T] = d_word
    return d_img.reshape(img_features.shape), d_words
