"""Quick device timing of words_loss forward/backward (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sba_gan_b200.losses import words_loss, words_similarity

for B in (48, 256):
    g = torch.Generator().manual_seed(1)        # synthetic CUB-shaped inputs
    d = {"cap_lens": torch.sort(torch.randint(5, 19, (B,), generator=g), descending=True).values,
         "class_ids": torch.randint(1, 201, (B,), generator=g).numpy()}
    img = torch.randn(B, 256, 17, 17, generator=g).cuda().requires_grad_(True)
    words = torch.tanh(torch.randn(B, 256, 18, generator=g)).cuda()
    lens = d["cap_lens"].cuda().int()
    labels = torch.arange(B).cuda()
    Tbar = d["cap_lens"].float().mean().item()
    flops_f = 4.0 * B * B * 289 * Tbar * 256

    def fwd():
        return words_similarity(img, words, lens, 4.0, 5.0, 10.0)

    def fb(words_grad=False):
        w = words.detach().requires_grad_(words_grad)
        img.grad = None
        l0, l1, _ = words_loss(img, w, labels, d["cap_lens"], d["class_ids"], B, 4.0, 5.0, 10.0)
        (l0 + l1).backward()

    def timeit(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e-3
    with torch.no_grad():
        tf = timeit(fwd)
    tb = timeit(fb)
    tbw = timeit(lambda: fb(True))
    print(f"B={B} Tbar={Tbar:.1f}: fwd {tf*1e3:.2f} ms ({flops_f/tf/1e12:.1f} TFLOP/s, {B*B/tf/1e3:.0f} kpairs/s) | "
          f"fwd+bwd(img) {tb*1e3:.2f} ms ({B*B/tb/1e3:.0f} kpairs/s) | fwd+bwd(img+words) {tbw*1e3:.2f} ms", flush=True)
