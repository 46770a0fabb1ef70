import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sba_gan_b200 import GlobalAttentionGeneral
from oracle import synth_attention_inputs

def ref_fwd(x, ctx, w, mask, mode):
    B, idf, ih, iw = x.shape; Q = ih*iw; L = ctx.shape[2]
    src = torch.einsum("ic,bcl->bil", w.reshape(idf, -1).float(), ctx.float())
    S = torch.einsum("biq,bil->bql", x.reshape(B, idf, Q).float(), src)
    if mask is not None:
        if mode == "reference":
            rows = (torch.arange(B*Q, device=x.device) % B).view(B, Q)
            S = S.masked_fill(mask[rows], float("-inf"))
        else:
            S = S.masked_fill(mask[:, None, :], float("-inf"))
    P = torch.softmax(S, 2)
    c = torch.einsum("bil,bql->biq", src, P)
    return c.reshape(B, idf, ih, iw), P.transpose(1, 2).reshape(B, L, ih, iw), src

for dt in (torch.bfloat16, torch.float32):
  for B, hw, mode, masked in ((64, 64, "reference", True), (64, 64, "per_sample", True), (64, 64, "reference", False), (32, 64, "reference", True), (48, 32, "reference", True)):
    d = synth_attention_inputs(B, 32, 256, 18, hw, hw, seed=7, with_g_attn=True)
    m = GlobalAttentionGeneral(32, 256).cuda().to(dt); m.algo = "mma"; m.mask_mode = mode
    with torch.no_grad(): m.conv_context.weight.copy_(d["weight"].to(dt))
    mask = d["mask"].cuda() if masked else None; m.applyMask(mask)
    x = d["x"].cuda().to(dt)
    ctx = d["context"].cuda().to(dt)
    with torch.no_grad():
        c, a = m(x, ctx)
    torch.cuda.synchronize()
    cr, ar, src = ref_fwd(x, ctx, m.conv_context.weight.detach(), mask, mode)
    ec = (c.float()-cr).abs().reshape(B, 32, -1)
    ea = (a.float()-ar).abs().reshape(B, 18, -1)
    per_b = ec.amax((1, 2)) / cr.abs().max()
    badb = (per_b > 0.05).nonzero().flatten().tolist()
    print(dt, B, hw, mode, masked, "c err %.3e attn err %.3e" % (per_b.max().item(), ea.max().item()), "bad samples", badb[:20], flush=True)
    if badb:
        b0 = badb[0]
        px = (ec[b0].amax(0) > 0.05 * cr.abs().max()).nonzero().flatten()
        print("   sample", b0, "bad px count", px.numel(), "first", px[:12].tolist(), "last", px[-4:].tolist())
        apx = (ea[b0].amax(0) > 0.05).nonzero().flatten()
        print("   attn bad px count", apx.numel(), apx[:12].tolist())
