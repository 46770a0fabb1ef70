"""Drop-in ``func_attention`` and ``words_loss`` (AttnGAN2/code/GlobalAttention.py:31-69,
AttnGAN2/code/miscc/losses.py:62-132).  Filled in by the words_loss kernel milestone."""
from __future__ import annotations


def func_attention(query, context, gamma1):
    raise RuntimeError("sba_gan_b200.func_attention: kernel (c) not built yet")


def words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size, **kw):
    raise RuntimeError("sba_gan_b200.words_loss: kernel (c) not built yet")
