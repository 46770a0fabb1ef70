"""Development aid: run the tcgen05 backward on rotating (L2-cold) buffers with SBA_TC5_TRACE=1
so that the per-phase clock stamps of CTA 0 are printed for an HBM-resident tile stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sba_gan_b200 import _abi
lib = _abi.load()
dev = "cuda"
B, idf, cdf, L, hw = 64, 32, 256, 18, int(sys.argv[1]) if len(sys.argv) > 1 else 128
Q = hw * hw
dt = torch.bfloat16
sets = [(torch.randn(B, idf, Q, device=dev).to(dt), torch.randn(B, idf, Q, device=dev).to(dt),
         torch.empty(B, idf, Q, device=dev, dtype=dt), torch.empty(B, L, Q, device=dev, dtype=dt),
         torch.empty(B, idf, Q, device=dev, dtype=dt)) for _ in range(3)]
ctx = torch.tanh(torch.randn(B, cdf, L, device=dev)); W = torch.randn(idf, cdf, device=dev) / 16
lens = torch.randint(5, L + 1, (B,)); mask = (torch.arange(L)[None] >= lens[:, None]).to(torch.uint8).to(dev)
srcT = torch.empty(B, idf, L, device=dev); mb = torch.empty(3 * B, dtype=torch.int32, device=dev)
dSrc = torch.empty(B * idf * L + B + 1, device=dev); dW = torch.empty(idf, cdf, device=dev)
st = torch.cuda.current_stream().cuda_stream
for k in range(3):
    x, g, c, a, dx = sets[k]
    _abi.check(lib.sba_attn_fwd(x.data_ptr(), ctx.data_ptr(), W.data_ptr(), mask.data_ptr(), c.data_ptr(), a.data_ptr(),
                                srcT.data_ptr(), mb.data_ptr(), B, idf, cdf, L, Q, 1, 0, 3, st), "fwd")
torch.cuda.synchronize()
for k in range(3):
    x, g, c, a, dx = sets[k]
    print("=== bwd call", k, flush=True)
    _abi.check(lib.sba_attn_bwd(x.data_ptr(), ctx.data_ptr(), W.data_ptr(), mask.data_ptr(), srcT.data_ptr(), mb.data_ptr(),
                                g.data_ptr(), None, dx.data_ptr(), dSrc.data_ptr(), dW.data_ptr(), None, B, idf, cdf, L, Q,
                                1, 0, 3, st), "bwd")
    torch.cuda.synchronize()
# event-timed eager calls (no trace env expected for meaningful numbers)
for rep in range(2):
    for k in range(3):
        x, g, c, a, dx = sets[k]
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        _abi.check(lib.sba_attn_bwd(x.data_ptr(), ctx.data_ptr(), W.data_ptr(), mask.data_ptr(), srcT.data_ptr(), mb.data_ptr(),
                                    g.data_ptr(), None, dx.data_ptr(), dSrc.data_ptr(), dW.data_ptr(), None, B, idf, cdf, L, Q,
                                    1, 0, 3, st), "bwd")
        e1.record()
        torch.cuda.synchronize()
        print(f"eager bwd call: {e0.elapsed_time(e1) * 1e3:.1f} us", flush=True)
# back-to-back without sync
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for k in range(12):
    x, g, c, a, dx = sets[k % 3]
    lib.sba_attn_bwd(x.data_ptr(), ctx.data_ptr(), W.data_ptr(), mask.data_ptr(), srcT.data_ptr(), mb.data_ptr(),
                     g.data_ptr(), None, dx.data_ptr(), dSrc.data_ptr(), dW.data_ptr(), None, B, idf, cdf, L, Q, 1, 0, 3, st)
e1.record()
torch.cuda.synchronize()
print(f"12 back-to-back eager bwd calls: {e0.elapsed_time(e1) * 1e3 / 12:.1f} us each", flush=True)
