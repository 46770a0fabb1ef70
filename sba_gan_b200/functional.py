"""torch.autograd wrappers over the C ABI: device memory and streams are PyTorch's, the
arithmetic is libsba_attn.so's.  Mirrors AttnGAN2/code/GlobalAttention.py:82-121."""
from __future__ import annotations

import torch

from . import _abi

_DTYPES = {torch.float32: _abi.SBA_F32, torch.bfloat16: _abi.SBA_BF16}
_MASK_MODES = {"reference": _abi.SBA_MASK_REFERENCE, "per_sample": _abi.SBA_MASK_PER_SAMPLE}
_ALGOS = {"auto": _abi.SBA_ALGO_AUTO, "simt": _abi.SBA_ALGO_SIMT, "mma": _abi.SBA_ALGO_MMA, "tc5": _abi.SBA_ALGO_TCGEN05}

# kernels launched by this process through the ABI (bench.py reports it as gpu_launches)
launch_counter = {"n": 0}


def _ptr(t):
    return None if t is None else t.data_ptr()


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "sba_gan_b200: the word-region attention path runs only on CUDA tensors "
                "(B200, sm_100a); there is no CPU fallback")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _context_fp32(context):
    """fp32, contiguous view / copy of the word features.  Converted per call: the tensor is only
    B x cdf x L elements, and a cache keyed on the tensor object would go stale under CUDA-graph
    replay and ``.data`` writes (neither bumps ``_version``)."""
    if context.dtype == torch.float32 and context.is_contiguous():
        return context.detach()
    return context.detach().to(torch.float32).contiguous()


ALGO_NAMES = {v: k for k, v in _ALGOS.items()}


# kernel family that served the most recent forward / backward call of this process.  (The C ABI keeps it per host
# thread - sba_last_algo() - and autograd runs backward on its own thread, so it is read right behind each call.)
_last_algo = {"fwd": "?", "bwd": "?"}


def last_algo(which: str = "fwd") -> str:
    """Kernel family ("simt" | "mma" | "tc5") that served the last attention forward ("fwd") / backward ("bwd")."""
    return _last_algo[which]


def bwd_workspace(B, idf, cdf, L, device):
    """Workspace of sba_attn_bwd (include/sba_attn.h): uninitialised, nothing needs zeroing."""
    n = _abi.load().sba_attn_bwd_workspace_floats(B, idf, cdf, L)
    return torch.empty((n,), dtype=torch.float32, device=device)


def attn_forward_raw(x, context, weight, mask_u8, mask_mode, algo):
    """x B x idf x ih x iw (fp32|bf16, contiguous), context B x cdf x L, weight idf x cdf.
    Returns (c_code, attn, srcT fp32, mask_bits|None)."""
    lib = _abi.load()
    B, idf, ih, iw = x.shape
    Q = ih * iw
    cdf, L = context.shape[1], context.shape[2]
    ctx32 = _context_fp32(context)
    w32 = weight.detach().reshape(idf, cdf).to(torch.float32).contiguous()
    c_code = torch.empty_like(x)
    attn = torch.empty((B, L, ih, iw), dtype=x.dtype, device=x.device)
    srcT = torch.empty((B, idf, L), dtype=torch.float32, device=x.device)
    mask_bits = torch.empty((3 * B,), dtype=torch.int32, device=x.device)   # scratch words, see sba_attn.h
    rc = lib.sba_attn_fwd(_ptr(x), _ptr(ctx32), _ptr(w32), _ptr(mask_u8), _ptr(c_code), _ptr(attn), _ptr(srcT),
                          _ptr(mask_bits), B, idf, cdf, L, Q, _DTYPES[x.dtype], mask_mode, algo, _stream())
    _abi.check(rc, "sba_attn_fwd")
    launch_counter["n"] += _abi.last_launch_count()
    _last_algo["fwd"] = ALGO_NAMES.get(lib.sba_last_algo(), "?")
    return c_code, attn, srcT, mask_bits, ctx32, w32


class PreparedProjection:
    """sourceT = conv_context(context) (+ the mask words) of one attention call, computed ahead of the call
    (``prepare_projection``): the first of the forward's two kernels needs only the word features and the weight."""
    __slots__ = ("key", "srcT", "scratch", "ctx32", "w32", "event")


def _projection_key(context, weight, mask_u8, mask_mode):
    return (context.data_ptr(), context._version, tuple(context.shape), weight.data_ptr(), weight._version,
            None if mask_u8 is None else (mask_u8.data_ptr(), mask_u8._version), mask_mode)


def prepare_projection(context, weight, mask_u8, mask_mode, stream=None):
    """Run the projection kernel of the tcgen05 forward now (``sba_attn_fwd_phase(.., SBA_PHASE_FIRST)``), on ``stream``
    (a torch.cuda.Stream; default: the current one).  Returns a PreparedProjection for ``word_region_attention(..,
    prepared=)`` or None where the tcgen05 family does not cover the shape.  In ``G_NET.forward`` (model_bert.py:580-588)
    the word features exist before the first stage has produced ``h_code``: both stages' projections can run on a side
    stream under ``h_net1``."""
    _require_cuda(context, weight, mask_u8)
    lib = _abi.load()
    B, cdf, L = context.shape
    idf = weight.numel() // cdf
    if not lib.sba_attn_supported(0, _abi.SBA_ALGO_TCGEN05, B, idf, cdf, L, 128, _abi.SBA_BF16):
        return None
    cur = torch.cuda.current_stream(context.device)
    st = stream if stream is not None else cur
    if st is not cur:
        st.wait_stream(cur)                       # the word features / weight were produced on the caller's stream
    pp = PreparedProjection()
    with torch.cuda.stream(st):
        pp.ctx32 = _context_fp32(context)
        pp.w32 = weight.detach().reshape(idf, cdf).to(torch.float32).contiguous()
        pp.srcT = torch.empty((B, idf, L), dtype=torch.float32, device=context.device)
        pp.scratch = torch.empty((3 * B,), dtype=torch.int32, device=context.device)
        rc = lib.sba_attn_fwd_phase(None, _ptr(pp.ctx32), _ptr(pp.w32), _ptr(mask_u8), None, None, _ptr(pp.srcT),
                                    _ptr(pp.scratch), B, idf, cdf, L, 128, _abi.SBA_BF16, mask_mode, _abi.SBA_PHASE_FIRST,
                                    st.cuda_stream)
        _abi.check(rc, "sba_attn_fwd_phase(projection)")
        launch_counter["n"] += _abi.last_launch_count()
        pp.event = torch.cuda.Event()
        pp.event.record(st)
    pp.key = _projection_key(context, weight, mask_u8, mask_mode)
    return pp


class _WordRegionAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, context, weight, mask_u8, mask_mode, algo, algo_bwd, prepared=None):
        _require_cuda(x, context, weight, mask_u8)
        if x.dtype not in _DTYPES:
            raise RuntimeError(f"sba_gan_b200: unsupported dtype {x.dtype} (float32 or bfloat16)")
        x = x.contiguous()
        # an unused output (attn is discarded in training, trainer_bert.py:267) must reach backward as
        # None, not as a materialised zero tensor the kernel would have to stream
        ctx.set_materialize_grads(False)
        if prepared is not None:
            c_code, attn, srcT, mask_bits, ctx32, w32 = _attn_forward_prepared(x, mask_u8, mask_mode, prepared)
        else:
            c_code, attn, srcT, mask_bits, ctx32, w32 = attn_forward_raw(x, context, weight, mask_u8, mask_mode, algo)
        ctx.save_for_backward(x, ctx32, w32, mask_u8, srcT, mask_bits)
        ctx.meta = (mask_mode, algo_bwd, context.dtype, weight.dtype, tuple(weight.shape))
        return c_code, attn

    @staticmethod
    def backward(ctx, g_c, g_attn):
        x, ctx32, w32, mask_u8, srcT, mask_bits = ctx.saved_tensors
        mask_mode, algo, ctx_dtype, w_dtype, w_shape = ctx.meta
        lib = _abi.load()
        B, idf, ih, iw = x.shape
        Q = ih * iw
        cdf, L = ctx32.shape[1], ctx32.shape[2]
        need_x, need_ctx, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if g_c is None:
            g_c = torch.zeros_like(x)
        g_c = g_c.to(x.dtype).contiguous()
        if g_attn is not None:
            g_attn = g_attn.to(x.dtype).contiguous()
        dX = torch.empty_like(x)
        ws = bwd_workspace(B, idf, cdf, L, x.device)
        dW = torch.empty((idf, cdf), dtype=torch.float32, device=x.device) if need_w else None
        dCtx = torch.empty((B, cdf, L), dtype=torch.float32, device=x.device) if need_ctx else None
        rc = lib.sba_attn_bwd(_ptr(x), _ptr(ctx32), _ptr(w32), _ptr(mask_u8), _ptr(srcT), _ptr(mask_bits), _ptr(g_c),
                              _ptr(g_attn), _ptr(dX), _ptr(ws), ws.numel(), _ptr(dW), _ptr(dCtx), B, idf, cdf, L, Q,
                              _DTYPES[x.dtype], mask_mode, algo, _stream())
        _abi.check(rc, "sba_attn_bwd")
        launch_counter["n"] += _abi.last_launch_count()
        _last_algo["bwd"] = ALGO_NAMES.get(lib.sba_last_algo(), "?")
        return (dX if need_x else None,
                dCtx.to(ctx_dtype) if need_ctx else None,
                dW.reshape(w_shape).to(w_dtype) if need_w else None,
                None, None, None, None, None)


def _attn_forward_prepared(x, mask_u8, mask_mode, pp):
    """The streaming kernel alone, behind a projection that ``prepare_projection`` already ran."""
    lib = _abi.load()
    B, idf, ih, iw = x.shape
    cdf, L = pp.ctx32.shape[1], pp.ctx32.shape[2]
    c_code = torch.empty_like(x)
    attn = torch.empty((B, L, ih, iw), dtype=x.dtype, device=x.device)
    cur = torch.cuda.current_stream(x.device)
    cur.wait_event(pp.event)
    for t in (pp.srcT, pp.scratch, pp.ctx32, pp.w32):
        t.record_stream(cur)
    rc = lib.sba_attn_fwd_phase(_ptr(x), None, None, _ptr(mask_u8), _ptr(c_code), _ptr(attn), _ptr(pp.srcT), _ptr(pp.scratch),
                                B, idf, cdf, L, ih * iw, _DTYPES[x.dtype], mask_mode, _abi.SBA_PHASE_SECOND, _stream())
    _abi.check(rc, "sba_attn_fwd_phase(stream)")
    launch_counter["n"] += _abi.last_launch_count()
    _last_algo["fwd"] = ALGO_NAMES.get(lib.sba_last_algo(), "?")
    return c_code, attn, pp.srcT, pp.scratch, pp.ctx32, pp.w32


def word_region_attention(x, context, weight, mask=None, mask_mode="reference", algo="auto", algo_bwd=None, prepared=None):
    """Fused GlobalAttentionGeneral.forward (GlobalAttention.py:82-121).

    x B x idf x ih x iw; context B x cdf x L; weight [idf, cdf, 1, 1] (conv_context.weight);
    mask B x L bool/uint8 (True = padding word) or None.
    Returns (weightedContext B x idf x ih x iw, attn B x L x ih x iw), differentiable in
    x, context and weight.  ``algo`` picks the kernel family ("auto" | "simt" | "mma" | "tc5");
    an explicit family is honoured strictly (a shape it does not cover raises).  ``algo_bwd``
    picks the backward's family separately (default: the same as ``algo``).
    """
    if x.dim() != 4 or context.dim() != 3:
        raise RuntimeError("word_region_attention: x must be B x idf x ih x iw and context B x cdf x L")
    if context.shape[0] != x.shape[0]:
        raise RuntimeError("word_region_attention: batch sizes of input and context differ")
    if weight.numel() != x.shape[1] * context.shape[1]:
        raise RuntimeError("word_region_attention: conv_context.weight must be [idf, cdf, 1, 1]")
    mask_u8 = None
    if mask is not None:
        if mask.shape[0] != x.shape[0] or mask.shape[-1] != context.shape[2] or mask.dim() != 2:
            # the reference would fail in masked_fill_ on the same mismatch (GlobalAttention.py:107-108)
            raise RuntimeError(f"word_region_attention: mask {tuple(mask.shape)} does not match "
                               f"batch {x.shape[0]} x sourceL {context.shape[2]}")
        mask = mask.detach()
        if mask.dtype == torch.bool and mask.device == x.device and mask.is_contiguous():
            mask_u8 = mask.view(torch.uint8)                      # same bytes, no copy kernel
        else:
            mask_u8 = mask.to(device=x.device, dtype=torch.uint8).contiguous()
    if prepared is not None:
        lib = _abi.load()
        ok = (x.is_cuda and x.dtype in _DTYPES and algo in ("auto", "tc5") and
              prepared.key == _projection_key(context, weight, mask_u8, _MASK_MODES[mask_mode]) and
              lib.sba_attn_supported(0, _abi.SBA_ALGO_TCGEN05, x.shape[0], x.shape[1], context.shape[1], context.shape[2],
                                     x.shape[2] * x.shape[3], _DTYPES[x.dtype]) and x.data_ptr() % 16 == 0)
        if not ok:
            prepared = None                       # stale or not applicable: the whole call, projection included
    return _WordRegionAttention.apply(x, context, weight, mask_u8, _MASK_MODES[mask_mode], _ALGOS[algo],
                                      _ALGOS[algo if algo_bwd is None else algo_bwd], prepared)
