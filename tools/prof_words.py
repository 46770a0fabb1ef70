"""Minimal launch sequence for ncu: words_loss forward (tensor-core kernel) and forward+backward at BASELINE
configs[2] size.  Usage: python tools/prof_words.py [B] [iters] [words]   (words: also the word gradients)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sba_gan_b200.losses import words_loss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with_words = len(sys.argv) > 3 and sys.argv[3] == "words"
g = torch.Generator().manual_seed(1)
lens = torch.sort(torch.randint(5, 19, (B,), generator=g), descending=True).values.cuda().int()
cls = torch.randint(1, 201, (B,), generator=g).cuda()
img = torch.randn(B, 256, 17, 17, generator=g).cuda().requires_grad_(True)
words = torch.tanh(torch.randn(B, 256, 18, generator=g)).cuda().requires_grad_(with_words)
labels = torch.arange(B).cuda()
for _ in range(iters):
    l0, l1, _ = words_loss(img, words, labels, lens, cls, B, 4.0, 5.0, 10.0, att_maps=False)
    torch.autograd.grad(l0 + l1, [img, words] if with_words else [img])
torch.cuda.synchronize()
print("ok")
