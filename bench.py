#!/usr/bin/env python
"""bench.py - benchmark of the word-region attention hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype bf16|fp32]

metric : region-px/s of the fused GlobalAttentionGeneral forward+backward.
step   : one G_NET training pass over the attention hot path at BASELINE configs[1]:
         stage-2 attention (64x64 regions) fwd+bwd plus stage-3 attention (128x128) fwd+bwd,
         B=64 per GPU, idf 32, cdf 256, 18 words, ragged caption masks, g_attn = None
         (the attention map is discarded in training, trainer_bert.py:267).
value  : inputs resident in HBM, CUDA events, max over ranks.
e2e    : the same step through the public module (GlobalAttentionGeneral.forward + autograd)
         with HOST buffers: pinned H2D of x / context / g_c (prefetched on a copy stream, one step
         ahead, as a training input pipeline would) and D2H of the weight gradients + a scalar.
sub    : the rest of BASELINE.json's metric and configs, each a bounded measurement in the same run:
         fp32 (the same step with fp32 tensors), sustained (>= 2 s of back-to-back steps, own clock sample),
         gpu_eager_baseline (the reference's eager torch op sequence on the same GPU),
         batch_sweep (configs[4]: 128 samples per GPU at 128x128), words_loss (configs[2]: B = 48 / 256,
         row-sharded over the ranks when N > 1), gan_step (configs[3]: full G+D adversarial step,
         imgs/s at N GPUs, fused hot path vs eager reference ops).
--impl reference: the reference's own CPU implementation of the step (oracle/_ref when present, else
         the oracle port), torch CPU fp32, all host threads.
Multi-GPU (torchrun): batch-sharded, weak scaling; the only exchange of the headline step is ONE
flattened all-reduce of the two conv_context weight gradients (what the generator's DDP bucket carries).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_PER_GPU, IDF, CDF, L = 64, 32, 256, 18
STAGES = (64, 128)
WORKLOAD = "AttnGAN2 bird_style G_NET stage-2 (64x64) + stage-3 (128x128) word-region attention fwd+bwd, B=64/GPU, idf32 cdf256 L18"
FFMA_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # CUDA-core fp32 ceiling of a B200: 74.4 TFLOP/s


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled during a timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, mx, pw, reasons = [], 0.0, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows[-3:]]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
                pw.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "power_w": round(statistics.median(pw), 1) if pw else None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()


def make_inputs(hw, seed, dtype, device, pin=False, batch=B_PER_GPU):
    """Synthetic CUB-shaped tensors of SURVEY.md §8(d) for one attention stage."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, IDF, hw, hw, generator=g).to(dtype)
    gc = torch.randn(batch, IDF, hw, hw, generator=g).to(dtype)
    ctx = torch.tanh(torch.randn(batch, CDF, L, generator=g))    # word features stay fp32: the frozen text encoder's output
                                                                 # (trainer_bert.py:256-257), a small per-caption tensor
    lens = torch.sort(torch.randint(5, L + 1, (batch,), generator=g), descending=True).values
    mask = torch.arange(L)[None, :] >= lens[:, None]
    if pin:
        return [t.pin_memory() for t in (x, gc, ctx)] + [mask]
    return [t.to(device) for t in (x, gc, ctx, mask)]


def algorithmic_bytes(px, es, which):
    """fwd (2*idf + L)*s, bwd 3*idf*s per region-pixel (g_attn = None)."""
    per_px = {"fwd": (2 * IDF + L) * es, "bwd": 3 * IDF * es}[which]
    return px * per_px


def ev():
    return torch.cuda.Event(enable_timing=True)


class Ctx:
    """rank / world / device and the collective helpers every leg uses."""

    def __init__(self, rank, world, device):
        self.rank, self.world, self.device = rank, world, device

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([v], device=self.device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return v


# ------------------------------------------------------------------------------------------------
# per-call device times of the C-ABI calls -> roofline of the dominant call
# ------------------------------------------------------------------------------------------------
def time_abi_calls(cx, dtype, algo, batch, stages, nsets=3, reps=12):
    """For each (fwd, bwd) x stage: a CUDA graph of `reps` back-to-back ABI calls over rotating buffer sets
    (inputs larger than L2 in total), bracketed by events on the capture stream."""
    from sba_gan_b200 import _abi
    from sba_gan_b200 import functional as F
    lib = _abi.load()
    device = cx.device
    es = 2 if dtype == torch.bfloat16 else 4
    dcode = _abi.SBA_BF16 if dtype == torch.bfloat16 else _abi.SBA_F32
    code = F._ALGOS[algo]
    calls = {}
    for hw in stages:
        Q = hw * hw
        w32 = torch.nn.init.orthogonal_(torch.empty(IDF, CDF), 1.0).to(device)
        bufs = []
        for s in range(nsets):
            x, gc, ctx, mask = make_inputs(hw, 4321 + cx.rank + 17 * s + hw, dtype, device, batch=batch)
            bufs.append(dict(x=x, gc=gc, ctx=ctx.float().contiguous(), mask=mask.to(torch.uint8).contiguous(),
                             c=torch.empty_like(x), a=torch.empty(batch, L, hw, hw, dtype=dtype, device=device),
                             dx=torch.empty_like(x), srcT=torch.empty(batch, IDF, L, device=device),
                             scr=torch.empty(3 * batch, dtype=torch.int32, device=device),
                             ws=torch.empty(lib.sba_attn_bwd_workspace_floats(batch, IDF, CDF, L), device=device),
                             dw=torch.empty(IDF, CDF, device=device)))

        def fwd_call(s, st):
            b = bufs[s % nsets]
            _abi.check(lib.sba_attn_fwd(b["x"].data_ptr(), b["ctx"].data_ptr(), w32.data_ptr(), b["mask"].data_ptr(),
                                        b["c"].data_ptr(), b["a"].data_ptr(), b["srcT"].data_ptr(), b["scr"].data_ptr(),
                                        batch, IDF, CDF, L, Q, dcode, 0, code, st), "sba_attn_fwd")

        def bwd_call(s, st):
            b = bufs[s % nsets]
            _abi.check(lib.sba_attn_bwd(b["x"].data_ptr(), b["ctx"].data_ptr(), w32.data_ptr(), b["mask"].data_ptr(),
                                        b["srcT"].data_ptr(), b["scr"].data_ptr(), b["gc"].data_ptr(), None,
                                        b["dx"].data_ptr(), b["ws"].data_ptr(), b["ws"].numel(), b["dw"].data_ptr(), None,
                                        batch, IDF, CDF, L, Q, dcode, 0, code, st), "sba_attn_bwd")

        cur = torch.cuda.current_stream().cuda_stream
        for s in range(nsets):
            fwd_call(s, cur)
            bwd_call(s, cur)
        torch.cuda.synchronize()
        for name, fn in (("fwd", fwd_call), ("bwd", bwd_call)):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                st = torch.cuda.current_stream().cuda_stream
                for s in range(reps):
                    fn(s, st)
            g.replay()
            torch.cuda.synchronize()
            best = None
            for _ in range(3):
                e0, e1 = ev(), ev()
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                t = e0.elapsed_time(e1) * 1e-3 / reps
                best = t if best is None else min(best, t)
            nbytes = algorithmic_bytes(batch * hw * hw, es, name)
            calls[f"{name}_{hw}"] = {"us": round(best * 1e6, 2), "gbs": round(nbytes / best / 1e9, 1), "bytes": nbytes}
        del bufs
    return calls


def sub_split_schedule(cx, dtype, nsets=3, reps=12):
    """The two kernels of each call, separately (sba_attn_fwd_phase / sba_attn_bwd_phase; tcgen05 family, bf16):
    (i) per-phase device times -> which part of a call streams the pixels and which is the small kernel around it;
    (ii) the same step (both stages, forward + backward, ABI level) scheduled the way the real generator allows: the
    projections depend only on the word features (available before the first stage has run, model_bert.py:580-588), the
    finish kernels produce parameter / word-feature gradients nothing in the backward chain waits for -> both go to a
    side stream.  All work stays inside the timed region; `serial` is the same step on one stream."""
    from sba_gan_b200 import _abi
    lib = _abi.load()
    device, batch = cx.device, B_PER_GPU
    es = 2 if dtype == torch.bfloat16 else 4
    dcode = _abi.SBA_BF16 if dtype == torch.bfloat16 else _abi.SBA_F32
    peaks, _ = load_peaks()
    w32 = {hw: torch.nn.init.orthogonal_(torch.empty(IDF, CDF), 1.0).to(device) for hw in STAGES}
    bufs = {}
    for hw in STAGES:
        bufs[hw] = []
        for s in range(nsets):
            x, gc, ctx, mask = make_inputs(hw, 9876 + cx.rank + 17 * s + hw, dtype, device, batch=batch)
            bufs[hw].append(dict(x=x, gc=gc, ctx=ctx.float().contiguous(), mask=mask.to(torch.uint8).contiguous(),
                                 c=torch.empty_like(x), a=torch.empty(batch, L, hw, hw, dtype=dtype, device=device),
                                 dx=torch.empty_like(x), srcT=torch.empty(batch, IDF, L, device=device),
                                 scr=torch.empty(3 * batch, dtype=torch.int32, device=device),
                                 ws=torch.empty(lib.sba_attn_bwd_workspace_floats(batch, IDF, CDF, L), device=device),
                                 dw=torch.empty(IDF, CDF, device=device)))

    # a projection serves exactly ONE streaming call (the scratch words carry the call's dynamic tile counter): the
    # per-phase timing of the streaming kernel gives every repetition its own projected (srcT, scratch) pair
    proj = {hw: [(torch.empty(batch, IDF, L, device=device), torch.empty(3 * batch, dtype=torch.int32, device=device))
                 for _ in range(reps)] for hw in STAGES}

    def fwd(hw, s, phase, st, own=False):
        b = bufs[hw][s % nsets]
        srcT, scr = proj[hw][s % reps] if own else (b["srcT"], b["scr"])
        _abi.check(lib.sba_attn_fwd_phase(b["x"].data_ptr(), b["ctx"].data_ptr(), w32[hw].data_ptr(), b["mask"].data_ptr(),
                                          b["c"].data_ptr(), b["a"].data_ptr(), srcT.data_ptr(), scr.data_ptr(),
                                          batch, IDF, CDF, L, hw * hw, dcode, 0, phase, st), "sba_attn_fwd_phase")

    def bwd(hw, s, phase, st):
        b = bufs[hw][s % nsets]
        _abi.check(lib.sba_attn_bwd_phase(b["x"].data_ptr(), b["ctx"].data_ptr(), w32[hw].data_ptr(), b["mask"].data_ptr(),
                                          b["srcT"].data_ptr(), b["scr"].data_ptr(), b["gc"].data_ptr(), None,
                                          b["dx"].data_ptr(), b["ws"].data_ptr(), b["ws"].numel(), b["dw"].data_ptr(), None,
                                          batch, IDF, CDF, L, hw * hw, dcode, 0, phase, st), "sba_attn_bwd_phase")

    cur = torch.cuda.current_stream().cuda_stream
    for hw in STAGES:
        for s in range(nsets):
            fwd(hw, s, _abi.SBA_PHASE_ALL, cur)
            bwd(hw, s, _abi.SBA_PHASE_ALL, cur)
    torch.cuda.synchronize()

    def time_graph(body, n, before=None):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        if before:
            before()
        g.replay()
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            if before:
                before()
                torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) * 1e-3 / n
            best = t if best is None else min(best, t)
        return best

    phases = {}
    for hw in STAGES:
        px = batch * hw * hw
        for name, fn, ph, which in (("fwd_project", fwd, _abi.SBA_PHASE_FIRST, None), ("fwd_stream", fwd, _abi.SBA_PHASE_SECOND, "fwd"),
                                    ("bwd_stream", bwd, _abi.SBA_PHASE_FIRST, "bwd"), ("bwd_finish", bwd, _abi.SBA_PHASE_SECOND, None)):
            own = name == "fwd_stream"

            def body(fn=fn, ph=ph, hw=hw, own=own):
                st = torch.cuda.current_stream().cuda_stream
                for s in range(reps):
                    if own:
                        fn(hw, s, ph, st, True)
                    else:
                        fn(hw, s, ph, st)

            def project_all(hw=hw):
                st = torch.cuda.current_stream().cuda_stream
                for s in range(reps):
                    fwd(hw, s, _abi.SBA_PHASE_FIRST, st, True)
            t = time_graph(body, reps, project_all if own else None)
            rec = {"us": round(t * 1e6, 2)}
            if which:
                nb = algorithmic_bytes(px, es, which)
                rec.update(gbs=round(nb / t / 1e9, 1), frac=round(nb / t / 1e9 / peaks["hbm_gbs"], 4))
            phases[f"{name}_{hw}"] = rec

    side = torch.cuda.Stream()
    lo, hi = STAGES[0], STAGES[-1]

    def step_serial(s):
        st = torch.cuda.current_stream().cuda_stream
        for hw in STAGES:
            fwd(hw, s, _abi.SBA_PHASE_ALL, st)
        for hw in reversed(STAGES):
            bwd(hw, s, _abi.SBA_PHASE_ALL, st)

    def step_split(s):
        m = torch.cuda.current_stream()
        side.wait_stream(m)
        with torch.cuda.stream(side):
            for hw in STAGES:
                fwd(hw, s, _abi.SBA_PHASE_FIRST, side.cuda_stream)
            evp = torch.cuda.Event()
            evp.record(side)
        m.wait_event(evp)                     # one join: the four streaming kernels then form an unbroken chain
        for hw in STAGES:
            fwd(hw, s, _abi.SBA_PHASE_SECOND, m.cuda_stream)
        evb = {}
        for hw in reversed(STAGES):
            bwd(hw, s, _abi.SBA_PHASE_FIRST, m.cuda_stream)
            evb[hw] = torch.cuda.Event()
            evb[hw].record(m)
        with torch.cuda.stream(side):
            for hw in reversed(STAGES):
                side.wait_event(evb[hw])
                bwd(hw, s, _abi.SBA_PHASE_SECOND, side.cuda_stream)
        m.wait_stream(side)

    nrep = 6
    out = {"phases": phases}
    for name, fn in (("serial", step_serial), ("split", step_split)):
        t = time_graph(lambda fn=fn: [fn(s) for s in range(nrep)], nrep)
        t = cx.max_over_ranks(t * 1e3) * 1e-3
        out[name] = {"ms_per_step": round(t * 1e3, 4), "value": px_per_step(cx.world) / t, "unit": "region-px/s"}
    out["speedup_split_vs_serial"] = round(out["serial"]["ms_per_step"] / out["split"]["ms_per_step"], 3)
    out["what"] = ("ABI-level step (fwd 64x64, fwd 128x128, bwd 128x128, bwd 64x64; CUDA-graph replay over rotating buffer sets); "
                   "split: projections and finish kernels on a side stream, ordered by events; serial: one stream")
    del bufs, lo, hi
    return out


# ------------------------------------------------------------------------------------------------
# the headline step: device-resident, through the public module, CUDA-graph replay
# ------------------------------------------------------------------------------------------------
class AttentionStep:
    def __init__(self, cx, dtype, algo, att_cls=None, use_graph=True, nsets=3):
        import torch.distributed as dist
        from sba_gan_b200 import GlobalAttentionGeneral
        from sba_gan_b200 import functional as F
        self.cx, self.dist, self.F, self.nsets = cx, dist, F, nsets
        device = cx.device
        torch.manual_seed(1234 + cx.rank)
        self.mods = []
        for _ in STAGES:
            m = (att_cls or GlobalAttentionGeneral)(IDF, CDF)
            torch.nn.init.orthogonal_(m.conv_context.weight.data, 1.0)      # miscc/utils.py:288-289
            m = m.to(device)        # parameters stay fp32 (autocast-style mixed precision); activations are `dtype`
            if att_cls is None:
                m.algo = algo
            elif dtype != torch.float32:
                m = m.to(dtype)     # the eager reference ops have no mixed-precision path: module and tensors in one dtype
            self.mods.append(m)
        self.sets = [[make_inputs(hw, 1234 + cx.rank + 17 * s + hw, dtype, device) for hw in STAGES] for s in range(nsets)]
        for st_ in self.sets:                         # both generator stages attend over the SAME word features and mask
            if att_cls is not None:
                st_[0][2] = st_[0][2].to(dtype)       # (eager reference ops: everything in one dtype)
            st_[1][2], st_[1][3] = st_[0][2], st_[0][3]    # (model_bert.py:580-588)
        for s in self.sets:
            for st in s:
                st[0].requires_grad_(True)
        self.pending = None
        self.graphs = []
        self.use_graph = use_graph

    def step(self, k):
        """fwd+bwd of both attention stages through the public module.  The gradients are taken with
        autograd.grad: in the generator x is an activation, not a leaf, so there is no .grad accumulation
        pass over it.  Returns the two conv_context weight gradients flattened into ONE buffer (what the
        generator's DDP bucket would carry)."""
        dws = []
        for m, (x, gc, ctx, mask) in zip(self.mods, self.sets[k % self.nsets]):
            m.applyMask(mask)
            c_code, _att = m(x, ctx)
            _dx, dw = torch.autograd.grad(c_code, [x, m.conv_context.weight], gc)
            dws.append(dw.reshape(-1).float())
        return torch.cat(dws)

    def exchange(self, flat):
        """ONE all-reduce of the flattened weight gradients per step, overlapped with the next step the way DDP
        overlaps its buckets with backward: issued asynchronously behind this step's kernels, waited for when
        the next exchange is issued (and at the end of the timed region)."""
        if self.cx.world > 1:
            self.drain()
            self.pending = self.dist.all_reduce(flat, async_op=True)

    def drain(self):
        if self.pending is not None:
            self.pending.wait()
            self.pending = None

    def prepare(self, warmup):
        for k in range(max(warmup, 3)):
            self.exchange(self.step(k))
        self.drain()
        self.cx.barrier()
        if self.use_graph:
            for k in range(self.nsets):
                g = torch.cuda.CUDAGraph()
                self.F.launch_counter["n"] = 0
                with torch.cuda.graph(g):
                    flat = self.step(k)
                self.graphs.append((g, self.F.launch_counter["n"], flat))
            for g, _, _ in self.graphs:
                g.replay()
            self.cx.barrier()

    def run_step(self, k):
        if self.graphs:
            g, _, flat = self.graphs[k % self.nsets]
            g.replay()
        else:
            flat = self.step(k)
        self.exchange(flat)

    def timed(self, steps):
        """EXACTLY `steps` steps between barriers; returns (ms per step max over ranks, launches, wall window)."""
        self.F.launch_counter["n"] = 0
        self.cx.barrier()
        t0 = time.time()
        e0, e1 = ev(), ev()
        e0.record()
        for k in range(steps):
            self.run_step(k)
        self.drain()
        e1.record()
        self.cx.barrier()
        t1 = time.time()
        launches = sum(self.graphs[k % self.nsets][1] for k in range(steps)) if self.graphs else self.F.launch_counter["n"]
        ms = self.cx.max_over_ranks(e0.elapsed_time(e1) / steps)
        return ms, launches, (t0, t1)


def px_per_step(world):
    return B_PER_GPU * sum(hw * hw for hw in STAGES) * world


def roofline_record(calls, ms, dtype_name):
    peaks, peak_kind = load_peaks()
    dom = max(calls, key=lambda k: calls[k]["us"])
    traffic, tsrc = None, None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(f"{dtype_name}_{dom}")
            tsrc = f"profiles/{name}: ncu --set full capture of this kernel (dram__bytes_read+write per launch), not measured in this run"
            if traffic is not None:
                break
    return {
        "bound": "hbm", "kernel": f"sba_attn_{dom.split('_')[0]} @{dom.split('_')[1]}x{dom.split('_')[1]}",
        "achieved": calls[dom]["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": round(calls[dom]["gbs"] / peaks["hbm_gbs"], 4), "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
        "traffic": traffic, "traffic_source": tsrc, "algorithmic_bytes_per_launch": calls[dom]["bytes"],
        "step_frac": round(sum(c["bytes"] for c in calls.values()) / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
        "calls": {k: dict(v, frac=round(v["gbs"] / peaks["hbm_gbs"], 4)) for k, v in calls.items()},
    }


# ------------------------------------------------------------------------------------------------
# e2e: host buffers, copies inside the timed region (prefetched one step ahead on a copy stream)
# ------------------------------------------------------------------------------------------------
def run_e2e(cx, mods, dtype, steps):
    device = cx.device
    host = [make_inputs(hw, 99 + cx.rank + hw, dtype, device, pin=True) for hw in STAGES]
    dev_sets = [[[torch.empty_like(t, device=device) for t in h[:3]] for h in host] for _ in range(2)]
    masks = [h[3].to(device) for h in host]
    res_host = torch.empty(2 * IDF * CDF + 2, dtype=torch.float32).pin_memory()
    h2d = sum(t.numel() * t.element_size() for h in host for t in h[:3])
    d2h = res_host.numel() * 4
    copy_stream = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        """host -> device copy of step i's inputs on the copy stream (its buffers were last read by step i - 2)"""
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            for h, dbuf in zip(host, dev_sets[s]):
                for src, dst in zip(h[:3], dbuf):
                    dst.copy_(src, non_blocking=True)
            copied[s].record(copy_stream)

    def compute(i):
        s = i % 2
        main.wait_event(copied[s])
        outs = []
        for m, dbuf, mask in zip(mods, dev_sets[s], masks):
            x = dbuf[0].requires_grad_(True)
            m.applyMask(mask)
            c_code, _att = m(x, dbuf[2])
            dx, dw = torch.autograd.grad(c_code, [x, m.conv_context.weight], dbuf[1])
            outs += [dw.float().reshape(-1), dx.float().sum().reshape(1)]
            dbuf[0] = x.detach()
        consumed[s].record(main)
        res_host.copy_(torch.cat(outs), non_blocking=True)

    def loop(n):
        for ev_ in consumed:
            ev_.record(main)
        prefetch(0)
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)             # next step's inputs travel while this step computes
            compute(i)
            main.synchronize()              # the step's result has been read on the host
        torch.cuda.synchronize()

    loop(3)
    cx.barrier()
    n = max(3, min(steps, 20))
    t0 = time.perf_counter()
    loop(n)
    cx.barrier()
    # (exactly n input sets are copied for n steps; the first step's copy is not overlapped with anything)
    e2e_s = cx.max_over_ranks((time.perf_counter() - t0) / n)

    # the host->device ceiling of this box for the same buffers, nothing else running on this rank
    def copy_only(m):
        for i in range(m):
            for h, dbuf in zip(host, dev_sets[i % 2]):
                for src, dst in zip(h[:3], dbuf):
                    dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
    copy_only(2)
    cx.barrier()
    t0 = time.perf_counter()
    copy_only(5)
    cx.barrier()
    copy_s = cx.max_over_ranks((time.perf_counter() - t0) / 5)
    return {"value": px_per_step(cx.world) / e2e_s, "unit": "region-px/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": round(e2e_s * 1e3, 3), "h2d_gbs_per_gpu": round(h2d / e2e_s / 1e9, 1),
            "h2d_copy_only_ms": round(copy_s * 1e3, 3), "h2d_copy_only_gbs_per_gpu": round(h2d / copy_s / 1e9, 1),
            "note": "inputs prefetched one step ahead on a copy stream; the step is bound by the host->device copy "
                    "(h2d_copy_only_* = the same copies alone, all ranks at once)"}


# ------------------------------------------------------------------------------------------------
# sub-records
# ------------------------------------------------------------------------------------------------
def sub_sustained(cx, stepper, ms_hint, seconds=2.5):
    """>= `seconds` of back-to-back steps with its own clock sample: clock / power behaviour under a training cadence."""
    n = max(100, int(math.ceil(seconds / (ms_hint * 1e-3))))
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    time.sleep(0.2)
    ms, _, (t0, t1) = stepper.timed(n)
    sampler.stop()
    return {"value": px_per_step(cx.world) / (ms * 1e-3), "unit": "region-px/s", "steps": n, "ms_per_step": round(ms, 4),
            "seconds": round(t1 - t0, 2), "clocks": sampler.window(t0, t1)}


def sub_eager_baseline(cx, dtype, steps):
    """The reference's own op sequence (GlobalAttention.py:92-117) as stock torch / cuBLAS ops on the SAME GPU,
    same step, same inputs: the bar SURVEY.md §2 sets for every new kernel."""
    from harness.gan_step import EagerAttention
    try:
        st = AttentionStep(cx, dtype, "auto", att_cls=EagerAttention, use_graph=True)
        st.prepare(3)
        launch = "cuda-graph replay"
    except Exception as e:  # capture of the eager sequence failed: time it eagerly
        torch.cuda.synchronize()
        st = AttentionStep(cx, dtype, "auto", att_cls=EagerAttention, use_graph=False)
        st.prepare(3)
        launch = f"eager ({type(e).__name__} under capture)"
    ms, _, _ = st.timed(max(5, min(steps, 20)))
    return {"value": px_per_step(cx.world) / (ms * 1e-3), "unit": "region-px/s", "ms_per_step": round(ms, 4), "launch": launch,
            "what": "reference eager op sequence (transpose+bmm+masked_fill+softmax+transpose+bmm, autograd backward) on this GPU"}


def sub_words_loss(cx):
    """BASELINE configs[2]: DAMSM words_loss at B = 48 / 256 (17x17 regions, 18 words, gammas 4/5/10); with N > 1
    ranks the B x B similarity is sharded by image rows (parallel.sharded_words_loss) - strong scaling of one op."""
    from sba_gan_b200 import parallel
    from sba_gan_b200.losses import words_loss, words_similarity
    device, world, rank = cx.device, cx.world, cx.rank
    out = {}
    for B in (48, 256):
        g = torch.Generator().manual_seed(1)
        lens_c = torch.sort(torch.randint(5, 19, (B,), generator=g), descending=True).values
        cls_c = torch.randint(1, 201, (B,), generator=g)
        img_c = torch.randn(B, 256, 17, 17, generator=g)
        words_c = torch.tanh(torch.randn(B, 256, 18, generator=g))
        tbar = lens_c.float().mean().item()
        flops_f = 4.0 * B * B * 289 * tbar * 256
        rec = {"B": B, "mean_len": round(tbar, 2)}

        def timeit(fn, n):
            fn()
            cx.barrier()
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            cx.barrier()
            return cx.max_over_ranks(e0.elapsed_time(e1) / n)

        if world == 1:
            img = img_c.to(device).requires_grad_(True)
            words, lens = words_c.to(device), lens_c.to(device).int()
            labels, cls = torch.arange(B, device=device), cls_c.to(device)

            def fwd():
                with torch.no_grad():
                    words_similarity(img, words, lens, 4.0, 5.0, 10.0)

            def fb(wg=False, maps=False):
                w = words.detach().requires_grad_(wg)
                l0, l1, _ = words_loss(img, w, labels, lens, cls, B, 4.0, 5.0, 10.0, att_maps=maps)
                torch.autograd.grad(l0 + l1, [img, w] if wg else [img])

            n = 20 if B <= 64 else 5
            # fwd_bwd_img: the GAN training call (words detached, trainer_bert.py:257; the maps it returns are dropped);
            # *_attmaps: the same with the reference's third return value written out; *_img_words: DAMSM pre-training
            tf, tb, tw = timeit(fwd, n), timeit(fb, n), timeit(lambda: fb(True), n)
            rec["fwd_bwd_img_attmaps_ms"] = round(timeit(lambda: fb(False, True), n), 3)
            rec.update(fwd_ms=round(tf, 3), fwd_tflops=round(flops_f / (tf * 1e-3) / 1e12, 2),
                       fwd_frac_of_ffma_ceiling=round(flops_f / (tf * 1e-3) / 1e12 / FFMA_TFLOPS, 3),
                       fwd_bwd_img_ms=round(tb, 3), fwd_bwd_img_words_ms=round(tw, 3),
                       pairs_per_s=round(B * B / (tb * 1e-3)))
        else:
            b = B // world
            if b * world != B:
                rec["skipped"] = f"B={B} not divisible by {world} ranks"
                out[f"B{B}"] = rec
                continue
            sl = slice(rank * b, (rank + 1) * b)
            img = img_c[sl].to(device).requires_grad_(True)
            words, lens, cls = words_c[sl].to(device), lens_c[sl].to(device), cls_c[sl].to(device)

            def fb():
                s0, s1 = parallel.sharded_words_loss(img, words, lens, cls, 4.0, 5.0, 10.0)
                torch.autograd.grad(s0 + s1, [img])

            tb = timeit(fb, 20 if B <= 64 else 5)
            rec.update(sharding=f"{world} ranks x {b} image rows, all-gather of word features and of sim rows (NCCL)",
                       fwd_bwd_img_ms=round(tb, 3), pairs_per_s=round(B * B / (tb * 1e-3)))
        out[f"B{B}"] = rec
    out["flops_model"] = "fwd = 4*B^2*289*mean_len*256; ceiling = 74.4 TFLOP/s CUDA-core fp32 (148 SMs x 128 lanes x 2 x 1.965 GHz)"
    return out


def sub_gan_step(cx, steps=6, batch=20):
    """BASELINE configs[3]: the full G+D adversarial step (harness/gan_step.py: trainer_bert.py:251-304 step body,
    bird_style networks restated from scratch, synthetic data, random init), batch-sharded, manual flattened gradient
    all-reduce per network.  Variants: fused = drop-in attention module + words_loss / sent_loss kernels; fused_stage = the
    same with attention + AdaIN + cat of NEXT_STAGE_G.forward as one operator (sba_gan_b200.stage, bf16 tensors);
    eager = the reference's eager ops for the same operators.  fp32 as the reference trains, and bf16 autocast."""
    from harness.gan_step import Trainer
    out = {"batch_per_gpu": batch, "global_batch": batch * cx.world, "steps": steps}
    for name, attention, amp in (("fused", "fused", False), ("eager", "eager", False), ("fused_bf16", "fused", True),
                                 ("fused_stage_bf16", "fused_stage", True), ("eager_bf16", "eager", True)):
        tr = Trainer(batch, cx.device, attention, cx.world, seed=1234 + cx.rank)

        def one():
            if amp:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return tr.step()
            return tr.step()
        for _ in range(3):
            one()
        cx.barrier()
        marks = [ev() for _ in range(steps + 1)]
        marks[0].record()
        for k in range(steps):
            one()
            marks[k + 1].record()
        cx.barrier()
        per_step = sorted(marks[k].elapsed_time(marks[k + 1]) for k in range(steps))
        # median step (a 60 ms step of ~2000 launches through NCCL and cuDNN sees an occasional straggler; the mean of six
        # is reported beside it), max over ranks
        ms = cx.max_over_ranks(per_step[steps // 2])
        ms_mean = cx.max_over_ranks(marks[0].elapsed_time(marks[steps]) / steps)
        out[name] = {"imgs_per_s": round(batch * cx.world / (ms * 1e-3), 1), "ms_per_step": round(ms, 2),
                     "ms_per_step_mean": round(ms_mean, 2)}
        del tr
        torch.cuda.empty_cache()
    out["speedup_fused_vs_eager"] = round(out["fused"]["imgs_per_s"] / out["eager"]["imgs_per_s"], 3)
    out["speedup_fused_stage_vs_eager_bf16"] = round(out["fused_stage_bf16"]["imgs_per_s"] / out["eager_bf16"]["imgs_per_s"], 3)
    return out


def run_ours(args, cx):
    from sba_gan_b200 import _abi
    _abi.load()                                   # fail loudly if the CUDA library is missing
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    world = cx.world
    stepper = AttentionStep(cx, dtype, args.algo, use_graph=not args.no_graph)
    stepper.prepare(args.warmup)
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    time.sleep(0.25)
    ms, launches, (t0, t1) = stepper.timed(args.steps)
    sampler.stop()
    clocks = sampler.window(t0, t1)
    value = px_per_step(world) / (ms * 1e-3)

    calls = time_abi_calls(cx, dtype, args.algo, B_PER_GPU, STAGES)
    roofline = roofline_record(calls, ms, args.dtype)
    e2e = run_e2e(cx, stepper.mods, dtype, args.steps)

    sub = {}
    if not args.no_sub:
        only = set(args.sub.split(",")) if args.sub else None

        state = {"dead": False}

        def leg(name, fn):
            if only is not None and name not in only:
                return
            if state["dead"]:
                sub[name] = {"skipped": "CUDA context lost in an earlier sub-record"}
                return
            t_leg = time.time()
            try:
                sub[name] = fn()
            except Exception as e:      # a failed side measurement must not take the headline line down
                sub[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                try:
                    torch.cuda.synchronize()
                except Exception:       # sticky CUDA error: nothing more can run on this context, keep what we have
                    state["dead"] = True
            if isinstance(sub[name], dict):
                sub[name]["leg_seconds"] = round(time.time() - t_leg, 1)
            if not state["dead"]:
                cx.barrier()

        leg("sustained", lambda: sub_sustained(cx, stepper, ms))
        del stepper
        torch.cuda.empty_cache()
        other = torch.float32 if dtype == torch.bfloat16 else torch.bfloat16
        oname = "fp32" if dtype == torch.bfloat16 else "bf16"

        def other_dtype():
            st = AttentionStep(cx, other, args.algo, use_graph=not args.no_graph)
            st.prepare(3)
            oms, _, _ = st.timed(args.steps)
            del st
            ocalls = time_abi_calls(cx, other, args.algo, B_PER_GPU, STAGES)
            r = roofline_record(ocalls, oms, oname)
            return {"value": px_per_step(world) / (oms * 1e-3), "unit": "region-px/s", "ms_per_step": round(oms, 4), "dtype": oname,
                    "step_frac": r["step_frac"], "calls": r["calls"]}
        leg(oname, other_dtype)
        if dtype == torch.bfloat16 and args.algo in ("auto", "tc5"):
            leg("split_schedule", lambda: sub_split_schedule(cx, dtype))
        leg("gpu_eager_baseline", lambda: sub_eager_baseline(cx, dtype, args.steps))
        leg("gpu_eager_baseline_fp32", lambda: sub_eager_baseline(cx, torch.float32, args.steps))

        def sweep():
            c = time_abi_calls(cx, dtype, args.algo, 128, (128,), nsets=2, reps=8)
            peaks, _ = load_peaks()
            t = (c["fwd_128"]["us"] + c["bwd_128"]["us"]) * 1e-6
            by = c["fwd_128"]["bytes"] + c["bwd_128"]["bytes"]
            return {"config": "BASELINE configs[4]: global batch 1024 over 8 GPUs = 128 per GPU, 128x128 regions", "B_per_gpu": 128,
                    "global_batch": 128 * world, "fwd_us": c["fwd_128"]["us"], "bwd_us": c["bwd_128"]["us"],
                    "region_px_per_s_per_gpu": round(128 * 128 * 128 / t), "hbm_frac": round(by / t / 1e9 / peaks["hbm_gbs"], 4)}
        leg("batch_sweep", sweep)
        leg("words_loss", lambda: sub_words_loss(cx))
        leg("gan_step", lambda: sub_gan_step(cx))

    out = {
        "metric": "word-attn fwd+bwd region-px/s", "value": value, "unit": "region-px/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": WORKLOAD, "B_per_gpu": B_PER_GPU, "regions": [hw * hw for hw in STAGES],
                   "io_dtype": args.dtype, "param_dtype": "fp32", "word_features_dtype": "fp32", "accumulate": "fp32", "algo": args.algo,
                   "launch": "eager" if args.no_graph else "cuda-graph replay", "mask": "ragged, reference mod-B order",
                   "g_attn": None, "exchange": "one flattened 64 KB all-reduce of both conv_context weight gradients per step",
                   "l2": "inputs larger than L2: one step streams %d MB per GPU over 3 rotating buffer sets" %
                   (sum(c["bytes"] for c in calls.values()) // 2 ** 20)},
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "clocks": clocks, "sub": sub,
    }
    if sub and any(isinstance(v, dict) and "skipped" in v for v in sub.values()):
        out["context_lost"] = True
    return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref) or, without it, the oracle port
# ------------------------------------------------------------------------------------------------
class CpuReference:
    def __init__(self):
        self.n_threads = os.cpu_count() or 1
        torch.set_num_threads(self.n_threads)
        self.kind, self.mods = "port", None
        try:
            from oracle.build_ref import load_ref
            ref = load_ref()
        except Exception:
            ref = None
        self.inputs = []
        for hw in STAGES:
            x, gc, ctx, mask = make_inputs(hw, 1234 + hw, torch.float32, "cpu")
            w = torch.nn.init.orthogonal_(torch.empty(IDF, CDF, 1, 1), 1.0)
            self.inputs.append((x, gc, ctx, mask, w))
        if ref is not None:
            self.kind = "reference"
            self.mods = []
            for (_x, _gc, _ctx, _mask, w) in self.inputs:
                m = ref[0].GlobalAttentionGeneral(IDF, CDF)          # the unmodified reference module
                with torch.no_grad():
                    m.conv_context.weight.copy_(w)
                self.mods.append(m)

    def step(self):
        if self.mods is not None:
            for m, (x, gc, ctx, mask, _w) in zip(self.mods, self.inputs):
                xr = x.detach().requires_grad_(True)
                m.conv_context.weight.grad = None
                m.applyMask(mask)
                c, _attn = m(xr, ctx)
                c.backward(gc)
            return
        from oracle.cpu_path import attn_fwd_bwd_autograd
        for x, gc, ctx, mask, w in self.inputs:
            attn_fwd_bwd_autograd(x, ctx, w, mask, gc)

    def describe(self):
        return ("unmodified reference GlobalAttentionGeneral (oracle/_ref, copied from AttnGAN2/code/GlobalAttention.py) + autograd"
                if self.kind == "reference" else "oracle port of the reference op sequence (oracle/cpu_path.py)")


def run_reference(args):
    ref = CpuReference()
    # K and W as asked (a step is ~0.1 s on the box's host cores), bounded so that the run stays within minutes
    warmup = max(1, min(args.warmup, 50))
    for _ in range(warmup):
        ref.step()
    steps = max(1, min(args.steps, 500))
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ref.step()
        ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    value = px_per_step(1) / t
    sample = f"{steps} full steps (B=64, 64x64 + 128x128 fwd+bwd) fp32, median"
    return {
        "impl": "reference", "metric": "word-attn fwd+bwd region-px/s", "value": value, "unit": "region-px/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": round(t * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "B_per_gpu": B_PER_GPU, "regions": [hw * hw for hw in STAGES], "note": ref.describe()},
        "cpu_baseline": {"value": value, "unit": "region-px/s", "cores": ref.n_threads, "kind": ref.kind, "sample": sample},
        "e2e": {"value": value, "unit": "region-px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def cpu_baseline_leg():
    """Bounded CPU sample beside the GPU number (rank 0, N=1 only)."""
    ref = CpuReference()
    ref.step()
    ts = []
    t_end = time.perf_counter() + 12.0
    while len(ts) < 3 or (time.perf_counter() < t_end and len(ts) < 30):
        t0 = time.perf_counter()
        ref.step()
        ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    return {"value": px_per_step(1) / t, "unit": "region-px/s", "cores": ref.n_threads, "kind": ref.kind,
            "sample": f"{len(ts)} full steps of the same workload (fp32, torch CPU, {ref.n_threads} threads), median {t*1e3:.0f} ms; "
                      + ref.describe()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--algo", default="auto", choices=["auto", "simt", "mma", "tc5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="headline only: skip the sub-records (fp32, sustained, words_loss, gan_step ...)")
    ap.add_argument("--sub", default="", help="comma-separated sub-records to run (default: all)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(run_reference(args)), flush=True)
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    out = run_ours(args, Ctx(rank, world, device))
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_leg()
        print(json.dumps(out), flush=True)
    if out.get("context_lost"):
        os._exit(0)          # the headline has been printed; CUDA teardown on a lost context would only hang or abort
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
