"""Development aid: 40 back-to-back tcgen05 backward calls on rotating buffers with SBA_TC5_TIMELINE=1;
the library prints globaltimer stamps of the zero / main / post kernels of calls 20..35."""
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
         torch.empty(B, idf, Q, device=dev, dtype=dt)) for _ in range(3)]
ctx = torch.tanh(torch.randn(B, cdf, L, device=dev)); W = torch.randn(idf, cdf, device=dev) / 16
lens = torch.randint(5, L + 1, (B,)); mask = (torch.arange(L)[None] >= lens[:, None]).to(torch.uint8).to(dev)
srcT = torch.randn(B, idf, L, device=dev); mb = torch.zeros(3 * B, dtype=torch.int32, device=dev)
dSrc = torch.empty(B * idf * L + B + 1, device=dev); dW = torch.empty(idf, cdf, device=dev)
st = torch.cuda.current_stream().cuda_stream
torch.cuda.synchronize()
for k in range(40):
    x, g, dx = sets[k % 3]
    _abi.check(lib.sba_attn_bwd(x.data_ptr(), ctx.data_ptr(), W.data_ptr(), mask.data_ptr(), srcT.data_ptr(), mb.data_ptr(),
                                g.data_ptr(), None, dx.data_ptr(), dSrc.data_ptr(), dW.data_ptr(), None, B, idf, cdf, L, Q,
                                1, 0, 3, st), "bwd")
torch.cuda.synchronize()
