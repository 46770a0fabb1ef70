"""Minimal launch sequence for ncu: a few fused attention forward/backward calls at the
bench shape.  Usage: python tools/prof_attn.py [algo] [dtype] [hw] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sba_gan_b200 import GlobalAttentionGeneral

algo = sys.argv[1] if len(sys.argv) > 1 else "auto"
dt = torch.bfloat16 if len(sys.argv) > 2 and sys.argv[2] == "bf16" else torch.float32
hw = int(sys.argv[3]) if len(sys.argv) > 3 else 128
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 4
B, idf, cdf, L = 64, 32, 256, 18
torch.manual_seed(0)
m = GlobalAttentionGeneral(idf, cdf).cuda().to(dt)
m.algo = algo
x = torch.randn(B, idf, hw, hw, device="cuda", dtype=dt, requires_grad=True)
gc = torch.randn(B, idf, hw, hw, device="cuda", dtype=dt)
ctx = torch.tanh(torch.randn(B, cdf, L, device="cuda")).to(dt)
lens = torch.randint(5, L + 1, (B,))
m.applyMask((torch.arange(L)[None] >= lens[:, None]).cuda())
for _ in range(iters):
    x.grad = None
    c, a = m(x, ctx)
    c.backward(gc)
torch.cuda.synchronize()
print("ok")
