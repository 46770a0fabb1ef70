"""Development check of the tcgen05 attention kernels against an fp64 torch restatement on the
GPU itself (the parity tests in tests/ use the CPU oracle; this is a quick bring-up aid).
Usage: python tools/check_tc5.py [fwd|bwd|all] [algo]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sba_gan_b200 import _abi
from sba_gan_b200.functional import _DTYPES, _ALGOS

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
algo = _ALGOS[sys.argv[2] if len(sys.argv) > 2 else "tc5"]
lib = _abi.load()
dev = "cuda"


def nerr(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max()).item()


def ref_fwd(x, ctx, W, mask, mode):
    B, idf, Q = x.shape
    L = ctx.shape[2]
    src = torch.einsum("ic,bcl->bil", W.double(), ctx.double())
    S = torch.einsum("biq,bil->bql", x.double(), src)
    if mask is not None:
        if mode == 0:
            n = torch.arange(B * Q, device=x.device).view(B, Q)
            m = mask.bool()[n % B]            # [B, Q, L]
        else:
            m = mask.bool()[:, None, :].expand(B, Q, L)
        S = S.masked_fill(m, float("-inf"))
    P = torch.softmax(S, dim=2)
    c = torch.einsum("bil,bql->biq", src, P)
    return c, P.transpose(1, 2).contiguous(), src, P


def ref_bwd(x, ctx, W, mask, mode, g, ga):
    c, attn, src, P = ref_fwd(x, ctx, W, mask, mode)
    dP = torch.einsum("bil,biq->bql", src, g.double())
    if ga is not None:
        dP = dP + ga.double().transpose(1, 2)
    dS = P * (dP - (P * dP).sum(2, keepdim=True))
    dX = torch.einsum("bql,bil->biq", dS, src)
    dSrc = torch.einsum("biq,bql->bil", g.double(), P) + torch.einsum("biq,bql->bil", x.double(), dS)
    dW = torch.einsum("bil,bcl->ic", dSrc, ctx.double())
    dCtx = torch.einsum("ic,bil->bcl", W.double(), dSrc)
    return dX, dSrc, dW, dCtx


def run(B, idf, cdf, L, hw, dt, mode, use_mask, with_ga=False):
    torch.manual_seed(B * 1000 + L * 10 + hw)
    Q = hw * hw
    x = torch.randn(B, idf, Q, device=dev).to(dt)
    g = torch.randn(B, idf, Q, device=dev).to(dt)
    ga = torch.randn(B, L, Q, device=dev).to(dt) if with_ga else None
    ctx = torch.tanh(torch.randn(B, cdf, L, device=dev))
    W = torch.nn.init.orthogonal_(torch.empty(idf, cdf)).to(dev)
    mask = None
    if use_mask:
        lens = torch.randint(max(1, L // 3), L + 1, (B,))
        mask = (torch.arange(L)[None] >= lens[:, None]).to(torch.uint8).to(dev)
    c = torch.empty_like(x)
    a = torch.empty(B, L, Q, device=dev, dtype=dt)
    srcT = torch.empty(B, idf, L, device=dev)
    mb = torch.zeros(3 * B, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    mp = mask.data_ptr() if mask is not None else None
    rc = lib.sba_attn_fwd(x.data_ptr(), ctx.data_ptr(), W.data_ptr(), mp, c.data_ptr(), a.data_ptr(), srcT.data_ptr(),
                          mb.data_ptr(), B, idf, cdf, L, Q, _DTYPES[dt], mode, algo, st)
    _abi.check(rc, "fwd")
    torch.cuda.synchronize()
    cr, ar, _, _ = ref_fwd(x.float(), ctx, W, mask, mode)
    msg = f"B={B} idf={idf} L={L} {hw}x{hw} {str(dt)[6:]} mode={mode} mask={use_mask}: fwd c {nerr(c, cr):.2e} attn {nerr(a, ar):.2e}"
    if which in ("bwd", "all"):
        dX = torch.empty_like(x)
        dSrc = torch.full((lib.sba_attn_bwd_workspace_floats(B, idf, cdf, L),), float("nan"), device=dev)
        dW = torch.empty(idf, cdf, device=dev)
        dCtx = torch.empty(B, cdf, L, device=dev)
        rc = lib.sba_attn_bwd(x.data_ptr(), ctx.data_ptr(), W.data_ptr(), mp, srcT.data_ptr(), mb.data_ptr(), g.data_ptr(),
                              ga.data_ptr() if ga is not None else None, dX.data_ptr(), dSrc.data_ptr(), dSrc.numel(), dW.data_ptr(),
                              dCtx.data_ptr(), B, idf, cdf, L, Q, _DTYPES[dt], mode, algo, st)
        _abi.check(rc, "bwd")
        torch.cuda.synchronize()
        rX, rS, rW, rC = ref_bwd(x.float(), ctx, W, mask, mode, g.float(), ga.float() if ga is not None else None)
        msg += (f" | bwd dX {nerr(dX, rX):.2e} dSrc {nerr(dSrc[:B * idf * L].view(B, idf, L), rS):.2e} "
                f"dW {nerr(dW, rW):.2e} dCtx {nerr(dCtx, rC):.2e}")
    print(msg, flush=True)


cases = [
    (2, 32, 256, 18, 16, torch.bfloat16, 1, False),
    (2, 32, 256, 18, 16, torch.float32, 1, False),
    (4, 32, 256, 18, 32, torch.bfloat16, 0, True),
    (4, 32, 256, 18, 32, torch.float32, 0, True),
    (3, 32, 256, 20, 16, torch.float32, 0, True),
    (5, 48, 256, 12, 16, torch.float32, 0, True),
    (5, 48, 256, 25, 16, torch.bfloat16, 0, True),
    (3, 64, 128, 7, 16, torch.float32, 1, True),
    (64, 32, 256, 18, 64, torch.bfloat16, 0, True),
    (64, 32, 256, 18, 64, torch.float32, 0, True),
    (64, 32, 256, 18, 128, torch.float32, 0, True),
]
for cs in cases:
    run(*cs)
    if which in ("bwd", "all"):
        run(*cs, with_ga=True)
print("done")
