#!/usr/bin/env python
"""words_loss sharded by image rows over N GPUs (NCCL) against the single-GPU fused words_loss:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
        --master-port 29541 tools/check_sharded_words_loss.py [B]
Every rank checks the replicated losses and its shard of d_img / d_words."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from sba_gan_b200 import words_loss, parallel

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
g = torch.Generator().manual_seed(7)                     # same global batch on every rank
img = torch.randn(B, 256, 17, 17, generator=g).to(dev)
words = torch.tanh(torch.randn(B, 256, 18, generator=g)).to(dev)
lens = torch.sort(torch.randint(5, 19, (B,), generator=g), descending=True).values.to(dev)
cls = torch.randint(1, 40, (B,), generator=g).to(dev)
G = (4.0, 5.0, 10.0)
# single-GPU result on the full batch
i0, w0 = img.clone().requires_grad_(True), words.clone().requires_grad_(True)
l0, l1, _ = words_loss(i0, w0, torch.arange(B, device=dev), lens, cls.cpu().numpy(), B, *G)
(l0 + l1).backward()
# sharded
b = B // world
sl = slice(rank * b, (rank + 1) * b)
i1, w1 = img[sl].clone().requires_grad_(True), words[sl].clone().requires_grad_(True)
s0, s1 = parallel.sharded_words_loss(i1, w1, lens[sl], cls[sl], *G)
(s0 + s1).backward()
torch.cuda.synchronize()


def nerr(a, ref):
    return ((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


e = dict(loss0=abs(s0.item() - l0.item()) / abs(l0.item()), loss1=abs(s1.item() - l1.item()) / abs(l1.item()),
         d_img=nerr(i1.grad, i0.grad[sl]), d_words=nerr(w1.grad, w0.grad[sl]))
print(f"rank {rank}/{world} B={B}: " + " ".join(f"{k} {v:.2e}" for k, v in e.items()), flush=True)
assert e["loss0"] < 1e-5 and e["loss1"] < 1e-5 and e["d_img"] < 1e-4 and e["d_words"] < 1e-4, e
dist.barrier()
dist.destroy_process_group()
