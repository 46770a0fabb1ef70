"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes.
The per-rank row block of the similarity matrix is computed by the CPU oracle here (sim_fn is
injectable exactly so that this test needs no GPU); on the B200 box the default sim_fn is the fused
kernel and tests/test_words_loss_gpu.py covers its row_offset handling."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from sba_gan_b200 import parallel

GAMMAS = (4.0, 5.0, 10.0)


def _oracle_sim(img, words, lens, g1, g2, g3, eps, row_offset):
    """[b, B] row block: local images against all captions (rows = images, cols = captions)."""
    return oracle.words_similarity(img, words, lens, g1, g2, g3, eps)


def _worker(rank, world, port, B, with_words_grad, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    d = oracle.synth_words_loss_inputs(B, 64, 10, 5, 5, seed=11, n_classes=3, dtype=torch.float64)
    img = parallel.shard_batch(d["img_features"], rank, world).clone().requires_grad_(True)
    words = parallel.shard_batch(d["words_emb"], rank, world).clone().requires_grad_(with_words_grad)
    lens = parallel.shard_batch(d["cap_lens"], rank, world)
    cls = parallel.shard_batch(torch.as_tensor(d["class_ids"]), rank, world)
    l0, l1 = parallel.sharded_words_loss(img, words, lens, cls, *GAMMAS, sim_fn=_oracle_sim)
    (l0 + l1).backward()
    # a small "network" whose gradients differ per rank -> flattened all-reduce(avg)
    net = torch.nn.Linear(3, 2).double()
    for p in net.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    n = parallel.allreduce_gradients(net.parameters())
    q.put((rank, l0.item(), l1.item(), img.grad.clone(), words.grad.clone() if with_words_grad else None,
           [p.grad.clone() for p in net.parameters()], n))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("with_words_grad", [False, True])
def test_sharded_words_loss_matches_single_process(with_words_grad):
    B, world, port = 6, 2, 29531 + int(with_words_grad)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, with_words_grad, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = oracle.synth_words_loss_inputs(B, 64, 10, 5, 5, seed=11, n_classes=3, dtype=torch.float64)
    img = d["img_features"].clone().requires_grad_(True)
    words = d["words_emb"].clone().requires_grad_(True)
    r0, r1, _ = oracle.words_loss(img, words, d["labels"], d["cap_lens"], d["class_ids"], B, *GAMMAS)
    (r0 + r1).backward()
    b = B // world
    for rank, l0, l1, d_img, d_words, net_grads, n in res:
        assert abs(l0 - r0.item()) < 1e-10 and abs(l1 - r1.item()) < 1e-10          # replicated loss
        assert torch.allclose(d_img, img.grad[rank * b:(rank + 1) * b], atol=1e-10)  # exact shard gradient
        if with_words_grad:
            assert torch.allclose(d_words, words.grad[rank * b:(rank + 1) * b], atol=1e-10)
        assert n == 8
        for g in net_grads:
            assert torch.allclose(g, torch.full_like(g, 1.5))                       # mean of 1 and 2
