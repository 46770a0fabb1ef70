"""Multi-GPU plumbing of the hot path: one process per GPU, torch.distributed
(NCCL over NVLink on the B200 box, gloo in the CPU tests).

* The generator attention shards by sample - no data-path collective; its only exchange is the
  gradient all-reduce of ``conv_context.weight`` together with the rest of the generator
  (``allreduce_gradients``: the manual flattened all-reduce SURVEY.md §7-6 recommends, because the
  reference calls discriminator sub-modules outside ``forward`` and a DDP wrapper breaks that).
* ``words_loss`` shards by image rows: all-gather the (small) word features, caption lengths and
  class ids, compute the local row block ``sim[rows_k, :]`` of the B x B similarity matrix with the
  fused kernel (``row_offset`` in the C ABI), all-gather the row blocks, and evaluate the two
  cross-entropies of miscc/losses.py:126-129 replicated on every rank.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn.functional as F


class _AllGatherRows(torch.autograd.Function):
    """Concatenate equal-sized row blocks of all ranks along dim 0.  Backward: every rank holds the
    gradient of the SAME replicated loss w.r.t. the full matrix, so the gradient of the local block
    is just its slice (no reduction)."""

    @staticmethod
    def forward(ctx, t, group):
        world = dist.get_world_size(group)
        ctx.rank, ctx.rows = dist.get_rank(group), t.shape[0]
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous(), group=group)
        return torch.cat(parts, dim=0)

    @staticmethod
    def backward(ctx, g):
        return g[ctx.rank * ctx.rows:(ctx.rank + 1) * ctx.rows].contiguous(), None


class _AllGatherSumBackward(torch.autograd.Function):
    """All-gather whose backward sums the per-rank gradients of the gathered tensor and returns the
    local slice (word features: every rank's row block depends on ALL captions)."""

    @staticmethod
    def forward(ctx, t, group):
        world = dist.get_world_size(group)
        ctx.rank, ctx.rows, ctx.group = dist.get_rank(group), t.shape[0], group
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous(), group=group)
        return torch.cat(parts, dim=0)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dist.all_reduce(g, group=ctx.group)
        return g[ctx.rank * ctx.rows:(ctx.rank + 1) * ctx.rows].contiguous(), None


def _gather_plain(t, group):
    world = dist.get_world_size(group)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous(), group=group)
    return torch.cat(parts, dim=0)


def sharded_words_loss(img_features, words_emb, cap_lens, class_ids, gamma1, gamma2, gamma3, eps=1e-8, group=None,
                       sim_fn=None):
    """words_loss (miscc/losses.py:62-132) over a batch sharded by image rows.

    Every rank passes ITS shard: img_features b x nef x ih x iw, words_emb b x nef x L, cap_lens [b],
    class_ids [b] (tensor / array / None); the global batch is world x b in rank order and
    labels = arange(world x b).  Returns (loss0, loss1), identical on every rank; backward gives each
    rank the exact gradient of that global loss w.r.t. its img_features shard (no collective) and
    w.r.t. its words_emb shard (one all-reduce, only if words require grad - DAMSM pre-training).

    sim_fn(img, words, lens, g1, g2, g3, eps, row_offset) -> [b, B] defaults to the fused kernel.
    """
    if sim_fn is None:
        from .losses import words_similarity

        def sim_fn(img, words, lens, g1, g2, g3, e, off):
            return words_similarity(img, words, lens, g1, g2, g3, e, off)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b = img_features.shape[0]
    words_all = _AllGatherSumBackward.apply(words_emb, group) if words_emb.requires_grad else _gather_plain(words_emb, group)
    lens = torch.as_tensor(cap_lens, device=img_features.device).to(torch.int64)
    lens_all = _gather_plain(lens, group)
    sim_rows = sim_fn(img_features, words_all, lens_all, gamma1, gamma2, gamma3, eps, rank * b)      # [b, B]
    sim = _AllGatherRows.apply(sim_rows, group)                                                     # [B, B]
    if class_ids is not None:
        cid = _gather_plain(torch.as_tensor(class_ids, device=sim.device).to(torch.int64), group)
        same = cid[:, None] == cid[None, :]
        same.fill_diagonal_(False)
        sim = sim.masked_fill(same, float("-inf"))             # losses.py:116-125
    labels = torch.arange(world * b, device=sim.device)
    return F.cross_entropy(sim, labels), F.cross_entropy(sim.transpose(0, 1), labels)


def allreduce_gradients(params, group=None, average=True):
    """Flattened gradient all-reduce of one network (what DDP's bucket would do) after its
    ``backward()``; parameters without a gradient are skipped.  Call it per network: after each
    ``errD.backward()`` for that discriminator and after ``errG_total.backward()`` for the generator
    only (the discriminators also receive gradients in the generator step - the reference just zeroes
    them next iteration, trainer_bert.py:275 - and must not be reduced there)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    world = dist.get_world_size(group)
    by_type = {}
    for g in grads:
        by_type.setdefault((g.dtype, g.device), []).append(g)
    n = 0
    for bucket in by_type.values():
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, group=group)
        if average:
            flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n += flat.numel()
    return n


def shard_batch(t, rank, world):
    """Rank's contiguous share of a batch-leading tensor (batch divisible by world)."""
    b = t.shape[0] // world
    return t[rank * b:(rank + 1) * b]
