"""Development check of the tensor-core words_loss forward against the CUDA-core kernel and an fp64 torch restatement
on the GPU.  Usage: python tools/check_words_tc5.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sba_gan_b200 import losses


def ref_sim(img, words, lens, g1, g2, g3, eps=1e-8):
    B, nef = img.shape[:2]
    X = img.reshape(B, nef, -1).double()
    cols = []
    for i in range(words.shape[0]):
        T = int(lens[i])
        w = words[i, :, :T].double()
        s = torch.einsum("jcr,ct->jrt", X, w)
        a1 = torch.softmax(s, dim=2)
        a2 = torch.softmax(g1 * a1, dim=1)
        wc = torch.einsum("jcr,jrt->jct", X, a2)
        num = (w[None] * wc).sum(1)
        den = (w.norm(dim=0)[None] * wc.norm(dim=1)).clamp(min=eps)
        cols.append(g3 * torch.log(torch.exp(g2 * num / den).sum(1)))
    return torch.stack(cols, 1)


for (B, nef, L, hw) in [(6, 256, 18, 17), (48, 256, 18, 17), (5, 64, 12, 6), (7, 32, 25, 5), (9, 256, 7, 17), (256, 256, 18, 17)]:
    g = torch.Generator().manual_seed(B)
    img = torch.randn(B, nef, hw, hw, generator=g).cuda()
    words = torch.tanh(torch.randn(B, nef, L, generator=g)).cuda()
    lens = torch.sort(torch.randint(1, L + 1, (B,), generator=g), descending=True).values.cuda()
    out = {}
    for algo in ("simt", "auto"):
        losses.FORWARD_ALGO = algo
        with torch.no_grad():
            out[algo] = losses.words_similarity(img, words, lens, 4.0, 5.0, 10.0)
        torch.cuda.synchronize()
    rows = slice(0, min(B, 8))
    ref = ref_sim(img[rows], words, lens.cpu(), 4.0, 5.0, 10.0)
    e = lambda a: ((a[rows].double() - ref).abs().max() / ref.abs().max()).item()  # noqa: E731
    print(f"B={B} nef={nef} L={L} R={hw*hw}: simt {e(out['simt']):.2e}  tc5 {e(out['auto']):.2e}  "
          f"tc5 vs simt {((out['auto'] - out['simt']).abs().max() / out['simt'].abs().max()).item():.2e}", flush=True)
    if B >= 48:
        for algo in ("simt", "auto"):
            losses.FORWARD_ALGO = algo
            for _ in range(2):
                losses.words_similarity(img, words, lens, 4.0, 5.0, 10.0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(5):
                with torch.no_grad():
                    losses.words_similarity(img, words, lens, 4.0, 5.0, 10.0)
            e1.record()
            torch.cuda.synchronize()
            print(f"   {algo}: {e0.elapsed_time(e1) / 5:.3f} ms per forward", flush=True)
print("done")

# ---- backward (d_img only): tensor-core path vs CUDA-core path vs fp64 autograd of the restatement
from sba_gan_b200.losses import words_loss  # noqa: E402
for (B, nef, L, hw) in [(6, 256, 18, 17), (48, 256, 18, 17), (5, 64, 12, 6), (9, 256, 7, 17), (128, 256, 18, 17), (256, 256, 18, 17)]:
    g = torch.Generator().manual_seed(B + 1)
    img = torch.randn(B, nef, hw, hw, generator=g).cuda()
    words = torch.tanh(torch.randn(B, nef, L, generator=g)).cuda()
    lens = torch.sort(torch.randint(1, L + 1, (B,), generator=g), descending=True).values.cuda()
    cls = torch.randint(1, 50, (B,), generator=g).cuda()
    labels = torch.arange(B).cuda()
    grads, gw, times, times_w = {}, {}, {}, {}
    for algo in ("simt", "auto"):
        losses.FORWARD_ALGO = algo
        x = img.clone().requires_grad_(True)
        l0, l1, _ = words_loss(x, words, labels, lens, cls, B, 4.0, 5.0, 10.0, att_maps=False)
        (gi,) = torch.autograd.grad(l0 + l1, [x])
        grads[algo] = gi
        w = words.clone().requires_grad_(True)
        l0, l1, _ = words_loss(x, w, labels, lens, cls, B, 4.0, 5.0, 10.0, att_maps=False)
        gi2, gw[algo] = torch.autograd.grad(l0 + l1, [x, w])
        assert torch.equal(gi2, gi) or algo == "simt", "d_img differs between the detached and the full backward"
        for wg, tt in ((False, times), (True, times_w)):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(3):
                ww_ = words.clone().requires_grad_(wg)
                l0, l1, _ = words_loss(x, ww_, labels, lens, cls, B, 4.0, 5.0, 10.0, att_maps=False)
                torch.autograd.grad(l0 + l1, [x, ww_] if wg else [x])
            e1.record()
            torch.cuda.synchronize()
            tt[algo] = e0.elapsed_time(e1) / 3
    rel = lambda a, b: ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()  # noqa: E731
    msg = f"bwd B={B} nef={nef} L={L} R={hw*hw}: tc5 vs simt d_img {rel(grads['auto'], grads['simt']):.2e} d_words {rel(gw['auto'], gw['simt']):.2e}"
    if B <= 9:
        x64 = img.double().clone().requires_grad_(True)
        w64 = words.double().clone().requires_grad_(True)
        sim = ref_sim(x64, w64, lens.cpu(), 4.0, 5.0, 10.0)
        same = (cls[:, None] == cls[None, :]) & ~torch.eye(B, dtype=torch.bool, device="cuda")
        sim = sim.masked_fill(same, float("-inf"))
        loss = torch.nn.functional.cross_entropy(sim, labels) + torch.nn.functional.cross_entropy(sim.t(), labels)
        gr, gwr = torch.autograd.grad(loss, [x64, w64])
        msg += (f" | vs fp64 d_img: simt {rel(grads['simt'], gr):.2e} tc5 {rel(grads['auto'], gr):.2e}"
                f"  d_words: simt {rel(gw['simt'], gwr):.2e} tc5 {rel(gw['auto'], gwr):.2e}")
    print(msg + f" | fwd+bwd(img) simt {times['simt']:.3f} tc5 {times['auto']:.3f} ms | fwd+bwd(img,words) simt {times_w['simt']:.3f} tc5 {times_w['auto']:.3f} ms",
          flush=True)
print("bwd done")
