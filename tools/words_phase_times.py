"""Where a CTA of the tensor-core words_loss kernel spends its time (needs `python -m sba_gan_b200.build --dev`):
mean over CTAs of [tile loop, wait for the last G2 MMAs, final epilogue] for the forward and for phase A of the backward.
Usage: python tools/words_phase_times.py [B]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sba_gan_b200 import _abi
from sba_gan_b200.losses import words_loss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
_abi.load()
raw = ctypes.CDLL(_abi.LIB_PATH)
out = (ctypes.c_ulonglong * 8)()
g = torch.Generator().manual_seed(1)
lens = torch.sort(torch.randint(5, 19, (B,), generator=g), descending=True).values.cuda().int()
cls = torch.randint(1, 201, (B,), generator=g).cuda()
img = torch.randn(B, 256, 17, 17, generator=g).cuda().requires_grad_(True)
words = torch.tanh(torch.randn(B, 256, 18, generator=g)).cuda()
labels = torch.arange(B).cuda()
for it in range(2):
    raw.sba_dev_words_times(out, 1)
    l0, l1, _ = words_loss(img, words, labels, lens, cls, B, 4.0, 5.0, 10.0, att_maps=False)
    torch.autograd.grad(l0 + l1, [img])
    raw.sba_dev_words_times(out, 0)
v = list(out)
for name, o in (("forward", 0), ("phase A", 4)):
    n = max(v[o], 1)
    print(f"{name}: {v[o]} CTAs; per CTA: tile loop {v[o+1]/n/1e3:.1f} us, wait for last G2 {v[o+2]/n/1e3:.1f} us, final epilogue {v[o+3]/n/1e3:.1f} us")
