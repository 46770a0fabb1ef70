"""Small fwd+bwd invocations of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sba_gan_b200 import GlobalAttentionGeneral
torch.manual_seed(0)
for algo in ("tc5", "mma", "simt"):
    for dt in (torch.bfloat16, torch.float32):
        B, idf, cdf, L, hw = 3, 32, 256, 18, 16
        m = GlobalAttentionGeneral(idf, cdf).cuda().to(dt)
        m.algo = algo
        x = torch.randn(B, idf, hw, hw, device="cuda", dtype=dt, requires_grad=True)
        ctx = torch.tanh(torch.randn(B, cdf, L, device="cuda")).to(dt).requires_grad_(True)
        lens = torch.tensor([18, 9, 5])
        m.applyMask((torch.arange(L)[None] >= lens[:, None]).cuda())
        c, a = m(x, ctx)
        (c.float().sum() + (a.float() ** 2).sum()).backward()
        torch.cuda.synchronize()
        print(algo, dt, float(c.float().abs().mean()), float(x.grad.float().abs().mean()), flush=True)
print("ok")
