"""Seeded synthetic inputs of SURVEY.md §8(d) and the parity metric.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.
"""
from __future__ import annotations

import numpy as np
import torch


def _orthogonal(rs: np.random.RandomState, rows: int, cols: int) -> np.ndarray:
    """Orthogonal init with gain 1 (what miscc/utils.py:286-289 ``weights_init`` applies to
    conv_context), from a frozen-stream RandomState so fixtures are reproducible."""
    a = rs.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))[None, :]
    return (q.T if rows < cols else q)[:rows, :cols]


def synth_attention_inputs(B, idf, cdf, L, ih, iw, seed=1234, dtype=torch.float32, with_g_attn=False,
                           min_len=5):
    """input ~ N(0,1) NCHW; context = tanh(N(0,1)) (BertEncoder ends in tanh,
    model_bert.py:185); W orthogonal [idf,cdf,1,1]; cap_lens ~ U{min_len..L} sorted
    descending (datasets_bert.py:31-33); mask = arange(L) >= len; g_c ~ N(0,1)."""
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((B, idf, ih, iw))
    ctx = np.tanh(rs.standard_normal((B, cdf, L)))
    W = _orthogonal(rs, idf, cdf).reshape(idf, cdf, 1, 1)
    lens = np.sort(rs.randint(min(min_len, L), L + 1, size=B))[::-1].copy()
    mask = np.arange(L)[None, :] >= lens[:, None]
    g_c = rs.standard_normal((B, idf, ih, iw))
    out = dict(
        x=torch.from_numpy(x).to(dtype), context=torch.from_numpy(ctx).to(dtype),
        weight=torch.from_numpy(W).to(dtype), mask=torch.from_numpy(mask),
        cap_lens=torch.from_numpy(lens.astype(np.int64)), g_c=torch.from_numpy(g_c).to(dtype),
    )
    if with_g_attn:
        out["g_attn"] = torch.from_numpy(rs.standard_normal((B, L, ih, iw))).to(dtype)
    return out


def synth_words_loss_inputs(B, nef=256, L=18, ih=17, iw=17, seed=1234, dtype=torch.float32, min_len=5,
                            n_classes=200):
    """img_features ~ N(0,1) B x nef x 17 x 17; words_emb = tanh(N(0,1)) B x nef x L;
    cap_lens ~ U{min_len..L} sorted descending; class_ids ~ randint(1, n_classes+1);
    labels = arange(B)."""
    rs = np.random.RandomState(seed)
    img = rs.standard_normal((B, nef, ih, iw))
    words = np.tanh(rs.standard_normal((B, nef, L)))
    lens = np.sort(rs.randint(min(min_len, L), L + 1, size=B))[::-1].copy()
    class_ids = rs.randint(1, n_classes + 1, size=B)
    return dict(
        img_features=torch.from_numpy(img).to(dtype), words_emb=torch.from_numpy(words).to(dtype),
        cap_lens=torch.from_numpy(lens.astype(np.int64)), class_ids=class_ids,
        labels=torch.arange(B, dtype=torch.int64),
    )


def normalised_max_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    """The parity metric of SURVEY.md §8 ("within X relative"): max|a-ref| / max|ref|.
    NaNs must coincide; +-inf must coincide."""
    a = a.detach().double().cpu()
    ref = ref.detach().double().cpu()
    if a.shape != ref.shape:
        raise AssertionError(f"shape mismatch {tuple(a.shape)} vs {tuple(ref.shape)}")
    nan_a, nan_r = torch.isnan(a), torch.isnan(ref)
    if not torch.equal(nan_a, nan_r):
        return float("inf")
    inf_a, inf_r = torch.isinf(a), torch.isinf(ref)
    if not torch.equal(inf_a, inf_r) or not torch.equal(a[inf_a], ref[inf_r]):
        return float("inf")
    fin = ~(nan_r | inf_r)
    if fin.sum() == 0:
        return 0.0
    scale = ref[fin].abs().max().item()
    diff = (a[fin] - ref[fin]).abs().max().item()
    return diff / scale if scale > 0 else diff
