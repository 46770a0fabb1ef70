"""sba_gan_b200 - B200-native (sm_100a) word-region attention hot path of SBA-GAN.

Public surface mirrors the reference modules it replaces:
  GlobalAttention.GlobalAttentionGeneral / conv1x1 / func_attention
  losses.words_loss
``install()`` registers them under the reference's import names.
"""
from __future__ import annotations

import sys

from . import _abi  # noqa: F401
from .functional import word_region_attention  # noqa: F401
from .GlobalAttention import GlobalAttentionGeneral, conv1x1  # noqa: F401
from .losses import func_attention, match_cross_entropy, sent_loss, words_loss  # noqa: F401

__all__ = ["GlobalAttentionGeneral", "conv1x1", "func_attention", "words_loss", "sent_loss", "match_cross_entropy",
           "word_region_attention", "install"]


def install(patch_loaded: bool = True) -> None:
    """Make the reference's ``from GlobalAttention import ...`` resolve to this package and,
    if reference modules are already imported, rebind their hot-path names
    (model*.ATT_NET, miscc.losses.func_attention / words_loss) - SURVEY.md §1 import seams."""
    from . import GlobalAttention as ga
    sys.modules["GlobalAttention"] = ga
    if not patch_loaded:
        return
    for name in ("model", "model_bert"):
        m = sys.modules.get(name)
        if m is not None and hasattr(m, "ATT_NET"):
            m.ATT_NET = ga.GlobalAttentionGeneral
    for name in ("miscc.losses",):
        m = sys.modules.get(name)
        if m is not None:
            m.func_attention = func_attention
    for name in ("trainer", "trainer_bert", "pretrain_DAMSM", "pretrain_DAMSM_bert", "miscc.losses"):
        m = sys.modules.get(name)
        if m is not None and hasattr(m, "words_loss"):
            m.words_loss = words_loss
        if m is not None and hasattr(m, "sent_loss"):
            m.sent_loss = sent_loss
