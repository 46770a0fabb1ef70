"""GPU parity tests for kernels (a)/(b): the fused GlobalAttentionGeneral forward/backward
through the drop-in module (-> autograd Function -> C ABI -> sm_100a kernels) against
  * the golden vectors produced by the unmodified reference (tests/golden), and
  * the CPU oracle on the same seeded inputs.
Tolerances are BASELINE.json's, as normalised max error (SURVEY.md §8 parity metric):
attention maps and c_code 1e-5 (fp32) / 2e-2 (bf16); gradients 1e-4 (fp32)."""
import os

import numpy as np
import pytest
import torch

from oracle import attn_forward, attn_backward, synth_attention_inputs, normalised_max_err
from tests.cases import ATTN_CASES, SUBSAMPLE

pytestmark = pytest.mark.gpu

TOL_FWD_F32 = 1e-5
TOL_GRAD_F32 = 1e-4
TOL_BF16 = 2e-2
ALGOS = ["simt", "mma", "tc5"]


def _mma_ok(idf, L, Q, B=1):
    """shapes the tensor-core family covers (sba::mma_supports); others must be refused."""
    return idf in (32, 48) and 9 <= L <= 32 and Q % 128 == 0


def _tc5_ok(idf, L, Q, cdf=256):
    """shapes the tcgen05 family covers (sba::tc5_supports); others must be refused."""
    return idf in (32, 48, 64) and 1 <= L <= 32 and Q % 128 == 0 and cdf % 16 == 0


def _covered(algo, idf, L, Q):
    return {"mma": _mma_ok, "tc5": _tc5_ok}.get(algo, lambda *a: True)(idf, L, Q)


def _module(idf, cdf, weight, dtype=torch.float32, algo="auto", mask_mode="reference"):
    from sba_gan_b200 import GlobalAttentionGeneral
    m = GlobalAttentionGeneral(idf, cdf).cuda().to(dtype)
    with torch.no_grad():
        m.conv_context.weight.copy_(weight.to(dtype))
    m.algo = algo
    m.mask_mode = mask_mode
    return m


def _bwd_family(algo, dtype, idf, L, Q):
    """The family the backward is asked for: an explicit family is strict in the C ABI, and the tcgen05
    backward covers bf16 tensors only - fp32 runs of the tc5 forward take the mma.sync backward where that
    covers the shape, else the CUDA-core one."""
    if algo == "tc5" and dtype == torch.float32:
        return "mma" if _mma_ok(idf, L, Q) else "simt"
    return algo


def _run(d, masked, algo="auto", dtype=torch.float32, mask_mode="reference", ctx_grad=True):
    from sba_gan_b200.functional import last_algo
    B, idf = d["x"].shape[:2]
    cdf = d["context"].shape[1]
    m = _module(idf, cdf, d["weight"], dtype, algo, mask_mode)
    m.algo_bwd = _bwd_family(algo, dtype, idf, d["context"].shape[2], d["x"].shape[2] * d["x"].shape[3])
    x = d["x"].cuda().to(dtype).requires_grad_(True)
    ctx = d["context"].cuda().to(dtype).requires_grad_(ctx_grad)
    m.applyMask(d["mask"].cuda() if masked else None)
    c, attn = m(x, ctx)
    ran_fwd = last_algo("fwd")
    loss = (c.float() * d["g_c"].cuda().float()).sum()
    if "g_attn" in d:
        loss = loss + (attn.float() * d["g_attn"].cuda().float()).sum()
    loss.backward()
    ran_bwd = last_algo("bwd")
    torch.cuda.synchronize()
    if algo != "auto":          # the family that was asked for is the family that ran (never a silent substitute)
        assert ran_fwd == algo and ran_bwd == m.algo_bwd, (algo, ran_fwd, ran_bwd)
    return dict(c_code=c.detach(), attn=attn.detach(), dX=x.grad, dW=m.conv_context.weight.grad,
                dCtx=ctx.grad if ctx_grad else None, ran=(ran_fwd, ran_bwd))


def _sub(name, t):
    step = SUBSAMPLE.get(name, 1)
    flat = t.reshape(-1)
    return flat[::step] if step > 1 and flat.numel() > 100000 else t


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name", list(ATTN_CASES))
def test_matches_reference_golden(golden_dir, name, algo):
    B, idf, cdf, L, ih, iw, seed, masked, with_ga = ATTN_CASES[name]
    if not _covered(algo, idf, L, ih * iw):
        pytest.skip("shape not covered by this kernel family")
    g = np.load(os.path.join(golden_dir, f"{name}_f32.npz"))
    d = synth_attention_inputs(B, idf, cdf, L, ih, iw, seed=seed, with_g_attn=with_ga)
    out = _run(d, masked, algo)
    for key, tol in (("c_code", TOL_FWD_F32), ("attn", TOL_FWD_F32), ("dX", TOL_GRAD_F32), ("dW", TOL_GRAD_F32),
                     ("dCtx", TOL_GRAD_F32)):
        ref = torch.from_numpy(g[key])
        err = normalised_max_err(_sub(name, out[key].cpu()).reshape(ref.shape), ref)
        assert err <= tol, f"{name}/{key}: {err:.3e} > {tol}"


SWEEP = [
    # B, idf, L, ih, iw, masked, mask_mode, with_g_attn
    (1, 32, 18, 64, 64, True, "reference", False),
    (10, 32, 18, 64, 64, True, "reference", True),     # BASELINE configs[0] shape
    (3, 32, 20, 32, 32, True, "per_sample", True),
    (2, 48, 12, 16, 16, True, "reference", False),     # coco GF_DIM
    (2, 128, 25, 8, 8, True, "reference", True),       # default GF_DIM, longest yml caption
    (3, 32, 15, 17, 17, True, "reference", True),      # odd Q: scalar path
    (2, 32, 32, 8, 8, False, "reference", True),       # maximum L
    (5, 32, 1, 4, 4, False, "reference", False),       # single word: attn == 1
    (20, 32, 18, 64, 64, True, "reference", False),    # yml batch size
    (4, 16, 7, 2, 2, True, "reference", True),         # tiny
    (3, 48, 12, 16, 16, True, "reference", True),      # coco GF_DIM, 2 word tiles
    (7, 32, 25, 16, 8, True, "reference", True),       # 4 word tiles, B does not divide Q
    (5, 48, 20, 32, 16, True, "per_sample", False),    # default WORDS_NUM
    (2, 32, 9, 16, 16, False, "reference", True),      # smallest L of the tensor-core family
    (130, 32, 18, 16, 8, True, "reference", False),    # more samples than CTAs-per-sample logic assumes
    (100, 32, 18, 60, 64, True, "reference", False),   # 30 tiles per sample (not a power of two), 3000 tiles: the
                                                       # dynamic tail of the tcgen05 forward's tile schedule (fp32 grid)
]


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("spec", SWEEP)
def test_shape_sweep_vs_oracle(spec, algo):
    B, idf, L, ih, iw, masked, mask_mode, with_ga = spec
    if not _covered(algo, idf, L, ih * iw):
        pytest.skip("shape not covered by this kernel family")
    d = synth_attention_inputs(B, idf, 256, L, ih, iw, seed=100 + B + idf + L, with_g_attn=with_ga, min_len=1)
    out = _run(d, masked, algo, mask_mode=mask_mode)
    d64 = {k: (v.double() if v.is_floating_point() else v) for k, v in d.items()}
    mask = d["mask"] if masked else None
    c, attn, _ = attn_forward(d64["x"], d64["context"], d64["weight"], mask, mask_mode)
    dX, dW, dCtx, _ = attn_backward(d64["x"], d64["context"], d64["weight"], mask, d64["g_c"], d64.get("g_attn"),
                                    mask_mode)
    ref = dict(c_code=c, attn=attn, dX=dX, dW=dW, dCtx=dCtx)
    for key, tol in (("c_code", TOL_FWD_F32), ("attn", TOL_FWD_F32), ("dX", TOL_GRAD_F32), ("dW", TOL_GRAD_F32),
                     ("dCtx", TOL_GRAD_F32)):
        err = normalised_max_err(out[key].cpu(), ref[key])
        assert err <= tol, f"{spec}/{key}: {err:.3e} > {tol}"


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("hw", [64, 128])
def test_full_size_fp32(hw, algo):
    """BASELINE configs[1] shape (B=64, idf 32, L 18, 64x64 / 128x128) against the fp32
    oracle, plus the size-independent properties: attention rows sum to 1, masked words get
    exactly 0 weight, gradients of masked words' scores vanish."""
    B, idf, cdf, L = 64, 32, 256, 18
    d = synth_attention_inputs(B, idf, cdf, L, hw, hw, seed=1234)
    out = _run(d, True, algo, ctx_grad=False)
    c, attn, _ = attn_forward(d["x"], d["context"], d["weight"], d["mask"])
    assert normalised_max_err(out["c_code"].cpu(), c) <= TOL_FWD_F32
    assert normalised_max_err(out["attn"].cpu(), attn) <= TOL_FWD_F32
    a = out["attn"].double()
    assert (a.sum(dim=1) - 1).abs().max().item() < 1e-5
    from oracle.attention import pixel_mask
    pm = pixel_mask(d["mask"], B, hw * hw).cuda()                      # B x Q x L
    assert out["attn"].reshape(B, L, -1).transpose(1, 2)[pm].abs().max().item() == 0.0
    dX, dW, _, _ = attn_backward(d["x"], d["context"], d["weight"], d["mask"], d["g_c"])
    assert normalised_max_err(out["dX"].cpu(), dX) <= TOL_GRAD_F32
    assert normalised_max_err(out["dW"].cpu(), dW) <= TOL_GRAD_F32


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("spec", [(10, 32, 18, 64, 64, True), (3, 48, 12, 16, 16, False), (64, 32, 18, 64, 64, True),
                                  (2, 64, 32, 16, 16, True), (5, 32, 4, 16, 8, True), (130, 32, 20, 16, 8, True),
                                  (160, 32, 18, 60, 64, True)])      # 4800 tiles of 30 per sample: dynamic tile schedule
def test_bf16_io(spec, algo):
    """bf16 tensors, fp32 arithmetic.  Oracle = fp32 reference maths on the bf16-rounded
    inputs (SURVEY.md §8 parity note); tolerance 2e-2."""
    B, idf, L, ih, iw, masked = spec
    if not _covered(algo, idf, L, ih * iw):
        pytest.skip("shape not covered by this kernel family")
    d = synth_attention_inputs(B, idf, 256, L, ih, iw, seed=7, with_g_attn=True)
    out = _run(d, masked, algo, dtype=torch.bfloat16)
    r = {k: (v.to(torch.bfloat16).double() if v.is_floating_point() else v) for k, v in d.items()}
    mask = d["mask"] if masked else None
    c, attn, _ = attn_forward(r["x"], r["context"], r["weight"], mask)
    dX, dW, dCtx, _ = attn_backward(r["x"], r["context"], r["weight"], mask, r["g_c"], r["g_attn"])
    for key, ref in (("c_code", c), ("attn", attn), ("dX", dX), ("dW", dW), ("dCtx", dCtx)):
        assert out[key].dtype == torch.bfloat16
        err = normalised_max_err(out[key].float().cpu(), ref)
        assert err <= TOL_BF16, f"{spec}/{key}: {err:.3e}"


@pytest.mark.parametrize("hw", [64, 128])
def test_headline_config_bf16_vs_oracle(hw):
    """BASELINE configs[1] exactly as bench.py times it - bf16 tensors, B=64, idf 32, 18 words, ragged masks,
    g_attn = None - against the oracle (fp64 maths on the bf16-rounded inputs), and on the product kernels:
    both directions must have run on the tcgen05 family."""
    B, idf, cdf, L = 64, 32, 256, 18
    d = synth_attention_inputs(B, idf, cdf, L, hw, hw, seed=1234)
    out = _run(d, True, "auto", dtype=torch.bfloat16, ctx_grad=False)
    assert out["ran"] == ("tc5", "tc5"), out["ran"]
    r = {k: (v.to(torch.bfloat16).double() if v.is_floating_point() else v) for k, v in d.items()}
    c, attn, _ = attn_forward(r["x"], r["context"], r["weight"], d["mask"])
    dX, dW, _, _ = attn_backward(r["x"], r["context"], r["weight"], d["mask"], r["g_c"])
    for key, ref in (("c_code", c), ("attn", attn), ("dX", dX), ("dW", dW)):
        err = normalised_max_err(out[key].float().cpu(), ref)
        assert err <= TOL_BF16, f"{hw}x{hw}/{key}: {err:.3e}"
    a = out["attn"].double()
    assert (a.sum(dim=1) - 1).abs().max().item() < 2e-2


def test_tc5_backward_is_bit_reproducible():
    """The tcgen05 backward accumulates dSrc / dW in a fixed order (per-CTA partial slots, no atomics): two runs
    on the same inputs agree bit for bit, also for the weight gradient."""
    d = synth_attention_inputs(20, 32, 256, 18, 64, 64, seed=5)
    a = _run(d, True, "tc5", dtype=torch.bfloat16)
    b = _run(d, True, "tc5", dtype=torch.bfloat16)
    for key in ("dX", "dW", "dCtx"):
        assert torch.equal(a[key], b[key]), key


def test_fully_masked_caption_gives_nan_like_reference():
    """softmax over all -inf is NaN in the reference (GlobalAttention.py:108-109)."""
    d = synth_attention_inputs(2, 32, 256, 6, 4, 4, seed=3)
    mask = torch.zeros(2, 6, dtype=torch.bool)
    mask[1, :] = True
    m = _module(32, 256, d["weight"], mask_mode="per_sample")
    m.applyMask(mask.cuda())
    c, attn = m(d["x"].cuda(), d["context"].cuda())
    cr, ar, _ = attn_forward(d["x"], d["context"], d["weight"], mask, "per_sample")
    assert torch.isnan(attn[1]).all() and torch.isnan(ar[1]).all()
    assert normalised_max_err(attn[0].cpu(), ar[0]) <= TOL_FWD_F32
    assert normalised_max_err(c[0].cpu(), cr[0]) <= TOL_FWD_F32


def test_sticky_mask_and_eval_mode():
    """applyMask state persists across calls and the module works under no_grad / eval()
    (trainer_bert.py:366, 422)."""
    d = synth_attention_inputs(4, 32, 256, 18, 8, 8, seed=9)
    m = _module(32, 256, d["weight"]).eval()
    m.applyMask(d["mask"].cuda())
    with torch.no_grad():
        c1, a1 = m(d["x"].cuda(), d["context"].cuda())
        c2, a2 = m(d["x"].cuda(), d["context"].cuda())
    assert torch.equal(c1, c2) and torch.equal(a1, a2)
    m.applyMask(None)
    with torch.no_grad():
        _, a3 = m(d["x"].cuda(), d["context"].cuda())
    assert not torch.equal(a1, a3)
    assert c1.shape == (4, 32, 8, 8) and a1.shape == (4, 18, 8, 8)


@pytest.mark.parametrize("algo", ["mma", "tc5"])
def test_tensor_core_families_refuse_uncovered_shapes(algo):
    from sba_gan_b200 import word_region_attention
    x = torch.zeros(2, 32, 4, 4, device="cuda")
    w = torch.zeros(32, 256, 1, 1, device="cuda")
    with pytest.raises(RuntimeError, match="does not cover"):
        word_region_attention(x, torch.zeros(2, 256, 12, device="cuda"), w, algo=algo)


def test_error_behaviour():
    from sba_gan_b200 import word_region_attention
    x = torch.zeros(2, 32, 4, 4, device="cuda")
    w = torch.zeros(32, 256, 1, 1, device="cuda")
    with pytest.raises(RuntimeError, match="exceeds the supported maximum"):
        word_region_attention(x, torch.zeros(2, 256, 33, device="cuda"), w)
    with pytest.raises(RuntimeError, match="mask"):
        word_region_attention(x, torch.zeros(2, 256, 5, device="cuda"), w, torch.zeros(3, 5, dtype=torch.bool))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        word_region_attention(x.cpu(), torch.zeros(2, 256, 5), w.cpu())
    with pytest.raises(RuntimeError, match="unsupported dtype"):
        word_region_attention(x.half(), torch.zeros(2, 256, 5, device="cuda").half(), w.half())


def test_non_contiguous_input_is_made_contiguous():
    d = synth_attention_inputs(2, 32, 256, 18, 8, 8, seed=10)
    m = _module(32, 256, d["weight"])
    xt = d["x"].cuda().permute(0, 1, 3, 2)          # non-contiguous view
    c, attn = m(xt, d["context"].cuda())
    cr, ar, _ = attn_forward(xt.cpu().contiguous(), d["context"], d["weight"], None)
    assert normalised_max_err(c.cpu(), cr) <= TOL_FWD_F32
    assert normalised_max_err(attn.cpu(), ar) <= TOL_FWD_F32


@pytest.mark.parametrize("algo,dt,hw", [("tc5", torch.bfloat16, 64), ("tc5", torch.bfloat16, 128), ("mma", torch.bfloat16, 64),
                                        ("auto", torch.float32, 64), ("auto", torch.float32, 128)])
def test_sustained_back_to_back_launches_stay_correct(algo, dt, hw):
    """Soak: thousands of back-to-back forward+backward launches (one CUDA graph, replayed) must neither trip the
    kernels' bounded barrier waits nor change the results.  A long sustained run is what exposed a barrier-phase
    race in the tcgen05 backward (one shared `dS ready` barrier for both tile parities); the last replay is
    compared with the first launch, so a late or early pipeline stage shows up as a numerical difference."""
    from sba_gan_b200 import _abi
    from sba_gan_b200.functional import _ALGOS
    lib = _abi.load()
    B, idf, cdf, L, Q = 64, 32, 256, 18, hw * hw       # 128x128: the dynamic part of the forward's tile schedule
    if not _covered(algo, idf, L, Q):
        pytest.skip("shape not covered by this kernel family")
    g = torch.Generator().manual_seed(7)
    dcode = 1 if dt == torch.bfloat16 else 0                # SBA_BF16 / SBA_F32
    x = torch.randn(B, idf, Q, generator=g).cuda().to(dt)
    gc = torch.randn(B, idf, Q, generator=g).cuda().to(dt)
    ctx = torch.tanh(torch.randn(B, cdf, L, generator=g)).cuda()
    W = (torch.randn(idf, cdf, generator=g) / 16).cuda()
    lens = torch.randint(5, L + 1, (B,), generator=g)
    mask = (torch.arange(L)[None] >= lens[:, None]).to(torch.uint8).cuda()
    c, a, dx = torch.empty_like(x), torch.empty(B, L, Q, device="cuda", dtype=dt), torch.empty_like(x)
    srcT = torch.empty(B, idf, L, device="cuda")
    scratch = torch.empty(3 * B, dtype=torch.int32, device="cuda")
    dSrc = torch.empty(lib.sba_attn_bwd_workspace_floats(B, idf, cdf, L), device="cuda")
    dW = torch.empty(idf, cdf, device="cuda")
    code = _ALGOS[algo]

    def step(stream):
        _abi.check(lib.sba_attn_fwd(x.data_ptr(), ctx.data_ptr(), W.data_ptr(), mask.data_ptr(), c.data_ptr(), a.data_ptr(),
                                    srcT.data_ptr(), scratch.data_ptr(), B, idf, cdf, L, Q, dcode, 0, code, stream), "fwd")
        _abi.check(lib.sba_attn_bwd(x.data_ptr(), ctx.data_ptr(), W.data_ptr(), mask.data_ptr(), srcT.data_ptr(),
                                    scratch.data_ptr(), gc.data_ptr(), None, dx.data_ptr(), dSrc.data_ptr(), dSrc.numel(), dW.data_ptr(),
                                    None, B, idf, cdf, L, Q, dcode, 0, code, stream), "bwd")

    step(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    first = [t.clone() for t in (c, a, dx, dW)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for _ in range(20):
                step(torch.cuda.current_stream().cuda_stream)
    for _ in range(150 if hw == 64 else 50):   # 3000 (1000) forward + as many backward launches back to back
        graph.replay()
    torch.cuda.synchronize()                   # a tripped barrier bound surfaces here as a CUDA error
    for name, t0, t1 in zip(("c_code", "attn", "dX"), first, (c, a, dx)):
        assert torch.equal(t0, t1), f"{name} changed under sustained launches"
    if dt == torch.bfloat16 and algo == "tc5":
        assert torch.equal(dW, first[3]), "dW changed under sustained launches (fixed-order reduction)"
    else:
        err = normalised_max_err(dW, first[3])       # fp32 atomics (mma.sync backward): order-dependent rounding only
        assert err < 1e-4, err


@pytest.mark.parametrize("dt,hw", [(torch.bfloat16, 64), (torch.bfloat16, 128), (torch.float32, 64)])
def test_phase_entry_points_equal_the_whole_call(dt, hw):
    """sba_attn_fwd_phase / sba_attn_bwd_phase: the projection (forward) and the finish kernel (backward) launched
    separately - on another stream, ordered by an event - give bit-identical results to the one-call form."""
    from sba_gan_b200 import _abi
    lib = _abi.load()
    B, idf, cdf, L, Q = 64, 32, 256, 18, hw * hw
    g = torch.Generator().manual_seed(11)
    dcode = 1 if dt == torch.bfloat16 else 0
    x = torch.randn(B, idf, Q, generator=g).cuda().to(dt)
    gc = torch.randn(B, idf, Q, generator=g).cuda().to(dt)
    ctx = torch.tanh(torch.randn(B, cdf, L, generator=g)).cuda()
    W = (torch.randn(idf, cdf, generator=g) / 16).cuda()
    lens = torch.randint(5, L + 1, (B,), generator=g)
    mask = (torch.arange(L)[None] >= lens[:, None]).to(torch.uint8).cuda()
    nws = lib.sba_attn_bwd_workspace_floats(B, idf, cdf, L)

    def buffers():
        return dict(c=torch.empty_like(x), a=torch.empty(B, L, Q, device="cuda", dtype=dt), dx=torch.empty_like(x),
                    srcT=torch.empty(B, idf, L, device="cuda"), scratch=torch.empty(3 * B, dtype=torch.int32, device="cuda"),
                    ws=torch.empty(nws, device="cuda"), dW=torch.empty(idf, cdf, device="cuda"),
                    dCtx=torch.empty(B, cdf, L, device="cuda"))

    main = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    one, two = buffers(), buffers()
    bwd = dt == torch.bfloat16                       # the tcgen05 backward covers bf16 tensors
    s = main.cuda_stream
    _abi.check(lib.sba_attn_fwd(x.data_ptr(), ctx.data_ptr(), W.data_ptr(), mask.data_ptr(), one["c"].data_ptr(),
                                one["a"].data_ptr(), one["srcT"].data_ptr(), one["scratch"].data_ptr(), B, idf, cdf, L, Q, dcode, 0,
                                _abi.SBA_ALGO_TCGEN05, s), "fwd")
    if bwd:
        _abi.check(lib.sba_attn_bwd(x.data_ptr(), ctx.data_ptr(), W.data_ptr(), mask.data_ptr(), one["srcT"].data_ptr(),
                                    one["scratch"].data_ptr(), gc.data_ptr(), None, one["dx"].data_ptr(), one["ws"].data_ptr(), nws,
                                    one["dW"].data_ptr(), one["dCtx"].data_ptr(), B, idf, cdf, L, Q, dcode, 0,
                                    _abi.SBA_ALGO_TCGEN05, s), "bwd")
    torch.cuda.synchronize()
    # projection on the side stream, streaming kernel on the main stream behind an event
    with torch.cuda.stream(side):
        _abi.check(lib.sba_attn_fwd_phase(None, ctx.data_ptr(), W.data_ptr(), mask.data_ptr(), None, None, two["srcT"].data_ptr(),
                                          two["scratch"].data_ptr(), B, idf, cdf, L, Q, dcode, 0, _abi.SBA_PHASE_FIRST,
                                          side.cuda_stream), "project")
        assert _abi.last_launch_count() == 1
        ev = torch.cuda.Event()
        ev.record(side)
    main.wait_event(ev)
    _abi.check(lib.sba_attn_fwd_phase(x.data_ptr(), None, None, mask.data_ptr(), two["c"].data_ptr(), two["a"].data_ptr(),
                                      two["srcT"].data_ptr(), two["scratch"].data_ptr(), B, idf, cdf, L, Q, dcode, 0,
                                      _abi.SBA_PHASE_SECOND, s), "stream")
    assert _abi.last_launch_count() == 1
    if bwd:
        _abi.check(lib.sba_attn_bwd_phase(x.data_ptr(), None, None, mask.data_ptr(), two["srcT"].data_ptr(),
                                          two["scratch"].data_ptr(), gc.data_ptr(), None, two["dx"].data_ptr(), two["ws"].data_ptr(),
                                          nws, None, None, B, idf, cdf, L, Q, dcode, 0, _abi.SBA_PHASE_FIRST, s), "bwd stream")
        ev2 = torch.cuda.Event()
        ev2.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ev2)
            _abi.check(lib.sba_attn_bwd_phase(None, ctx.data_ptr(), W.data_ptr(), None, None, None, None, None, None,
                                              two["ws"].data_ptr(), nws, two["dW"].data_ptr(), two["dCtx"].data_ptr(), B, idf, cdf,
                                              L, Q, dcode, 0, _abi.SBA_PHASE_SECOND, side.cuda_stream), "bwd finish")
    torch.cuda.synchronize()
    for k in ("srcT", "c", "a") + (("dx", "dW", "dCtx") if bwd else ()):
        assert torch.equal(one[k], two[k]), k
    # a family that has no phases refuses
    rc = lib.sba_attn_fwd_phase(x.data_ptr(), ctx.data_ptr(), W.data_ptr(), mask.data_ptr(), two["c"].data_ptr(),
                                two["a"].data_ptr(), two["srcT"].data_ptr(), two["scratch"].data_ptr(), B, 40, cdf, L, Q, dcode, 0,
                                _abi.SBA_PHASE_FIRST, s)
    assert rc != 0


def test_module_prepare_runs_the_projection_ahead_on_a_side_stream():
    """GlobalAttentionGeneral.prepare(context, stream): same outputs and gradients as the plain call, one kernel in the
    forward; a preparation that went stale (new mask, new weight) is ignored."""
    from sba_gan_b200 import functional as F
    torch.manual_seed(3)
    B, idf, cdf, L, hw = 8, 32, 256, 18, 64
    W = torch.randn(idf, cdf, 1, 1) / 16
    m = _module(idf, cdf, W, torch.float32)
    x = torch.randn(B, idf, hw, hw, device="cuda", dtype=torch.bfloat16)
    ctx = torch.tanh(torch.randn(B, cdf, L, device="cuda"))
    gc = torch.randn_like(x)
    lens = torch.randint(5, L + 1, (B,))
    mask = (torch.arange(L)[None] >= lens[:, None]).cuda()

    def run(prep, mutate=None):
        xs = x.clone().requires_grad_(True)
        m.zero_grad()
        m.applyMask(mask)
        if prep:
            side = torch.cuda.Stream()
            assert m.prepare(ctx, stream=side)
        if mutate:
            mutate()
        n0 = F.launch_counter["n"]
        c, a = m(xs, ctx)
        n_fwd = F.launch_counter["n"] - n0
        c.backward(gc)
        torch.cuda.synchronize()
        return c.detach(), a.detach(), xs.grad, m.conv_context.weight.grad.clone(), n_fwd

    ref = run(False)
    got = run(True)
    assert ref[4] == 2 and got[4] == 1               # projection + streaming kernel vs streaming kernel only
    for r, t in zip(ref[:4], got[:4]):
        assert torch.equal(r, t)
    # stale: the mask is replaced after prepare -> the whole call runs, with the new mask
    mask2 = (torch.arange(L)[None] >= (lens[:, None] - 2).clamp(min=1)).cuda()
    stale = run(True, mutate=lambda: m.applyMask(mask2))
    assert stale[4] == 2
    m.applyMask(mask2)
    xs = x.clone()
    c2, _ = m(xs, ctx)
    assert torch.equal(stale[0], c2.detach())


def test_bf16_context_under_graph_capture_and_replay():
    """The module converts a bf16 context to fp32 per call (no cache keyed on the tensor object): a captured
    forward+backward replayed after the context's CONTENTS were overwritten in place (no version bump visible to a
    cache) must see the new values, and an eager call after the capture must not read graph-private memory."""
    torch.manual_seed(5)
    B, idf, cdf, L, hw = 8, 32, 256, 18, 64
    m = _module(idf, cdf, torch.randn(idf, cdf, 1, 1) / 16, torch.bfloat16)
    x = torch.randn(B, idf, hw, hw, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    gc = torch.randn(B, idf, hw, hw, device="cuda", dtype=torch.bfloat16)
    ctx_a = torch.tanh(torch.randn(B, cdf, L, device="cuda")).to(torch.bfloat16)
    ctx_b = torch.tanh(torch.randn(B, cdf, L, device="cuda")).to(torch.bfloat16)
    ctx = ctx_a.clone()
    m.applyMask(None)

    def step():
        c, _ = m(x, ctx)
        dx, dw = torch.autograd.grad(c, [x, m.conv_context.weight], gc)
        return c, dx, dw

    for _ in range(3):                     # eager warm-up (what bench.py does before it captures)
        step()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            out = step()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    ref_a = [t.clone() for t in step()]                      # eager, same contents
    for o, r in zip(out, ref_a):
        assert torch.equal(o, r)
    ctx.copy_(ctx_b)                                         # new word features in the same storage
    g.replay()
    torch.cuda.synchronize()
    ref_b = [t.clone() for t in step()]
    for o, r in zip(out, ref_b):
        assert torch.equal(o, r)
    assert not torch.equal(ref_a[0], ref_b[0])
