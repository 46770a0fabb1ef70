#!/usr/bin/env python
"""bench.py - headline benchmark of the word-region attention hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype bf16|fp32]

metric : region-px/s of the fused GlobalAttentionGeneral forward+backward.
step   : one G_NET training pass over the attention hot path at BASELINE configs[1]:
         stage-2 attention (64x64 regions) fwd+bwd plus stage-3 attention (128x128) fwd+bwd,
         B=64 per GPU, idf 32, cdf 256, 18 words, ragged caption masks, g_attn = None
         (the attention map is discarded in training, trainer_bert.py:267).
value  : inputs resident in HBM, CUDA events, max over ranks.
e2e    : the same step through the public module (GlobalAttentionGeneral.forward + autograd)
         with HOST buffers: pinned H2D of x / context / g_c and D2H of the weight gradients and
         a scalar every step.
--impl reference: the reference's CPU path (oracle port, torch CPU fp32, all host threads).
Multi-GPU (torchrun): batch-sharded, weak scaling; the only exchange is the all-reduce of the
conv_context weight gradients (what the generator's DDP bucket would carry).
"""
from __future__ import annotations

import argparse
import json
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


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled during the timed region."""
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

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows[-3:]]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(hw, seed, dtype, device, pin=False):
    """Synthetic CUB-shaped tensors of SURVEY.md §8(d) for one attention stage."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B_PER_GPU, IDF, hw, hw, generator=g).to(dtype)
    gc = torch.randn(B_PER_GPU, IDF, hw, hw, generator=g).to(dtype)
    ctx = torch.tanh(torch.randn(B_PER_GPU, CDF, L, generator=g)).to(dtype)
    lens = torch.sort(torch.randint(5, L + 1, (B_PER_GPU,), generator=g), descending=True).values
    mask = torch.arange(L)[None, :] >= lens[:, None]
    if pin:
        return [t.pin_memory() for t in (x, gc, ctx)] + [mask]
    return [t.to(device) for t in (x, gc, ctx, mask)]


def algorithmic_bytes(px, es, which):
    """fwd (2*idf + L)*s, bwd 3*idf*s per region-pixel (g_attn = None)."""
    per_px = {"fwd": (2 * IDF + L) * es, "bwd": 3 * IDF * es}[which]
    return px * per_px


def run_ours(args, rank, world, device):
    import torch.distributed as dist
    from sba_gan_b200 import GlobalAttentionGeneral, _abi
    from sba_gan_b200 import functional as F
    _abi.load()                                   # fail loudly if the CUDA library is missing
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    es = 2 if args.dtype == "bf16" else 4
    torch.manual_seed(1234 + rank)
    mods = []
    for _ in STAGES:
        m = GlobalAttentionGeneral(IDF, CDF)
        torch.nn.init.orthogonal_(m.conv_context.weight.data, 1.0)      # miscc/utils.py:288-289
        m = m.to(device)        # parameters stay fp32 (autocast-style mixed precision); activations are `dtype`
        m.algo = args.algo
        mods.append(m)
    nsets = 3                                     # rotate buffers; one step already streams > L2 (126 MB)
    sets = [[make_inputs(hw, 1234 + rank + 17 * s + hw, dtype, device) for hw in STAGES] for s in range(nsets)]
    for st_ in sets:                              # both generator stages attend over the SAME word features and mask
        st_[1][2], st_[1][3] = st_[0][2], st_[0][3]    # (model_bert.py:580-588)
    for s in sets:
        for st in s:
            st[0].requires_grad_(True)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def step(k):
        """fwd+bwd of both attention stages through the public module.  The gradients are taken with
        autograd.grad: in the generator x is an activation, not a leaf, so there is no .grad accumulation
        pass over it.  Returns the conv_context weight gradients (what the DDP bucket would carry)."""
        dws = []
        for m, (x, gc, ctx, mask), hw in zip(mods, sets[k % nsets], STAGES):
            m.applyMask(mask)
            c_code, _att = m(x, ctx)
            _dx, dw = torch.autograd.grad(c_code, [x, m.conv_context.weight], gc)
            dws.append(dw)
        return dws

    pending = []

    def exchange(dws):
        """All-reduce of the conv_context weight gradients, overlapped with the next step the way DDP
        overlaps its buckets with backward: issued asynchronously behind this step's kernels, waited for
        when the next exchange is issued (and at the end of the timed region)."""
        if world > 1:
            for h in pending:
                h.wait()
            pending.clear()
            pending.extend(dist.all_reduce(t, async_op=True) for t in dws)

    def drain():
        for h in pending:
            h.wait()
        pending.clear()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (eager), then capture one CUDA graph per rotating buffer set: the step is a fixed
    # sequence of launches, so replaying it removes the Python/launch latency between kernels.
    for k in range(max(args.warmup, 3)):
        exchange(step(k))
    drain()
    barrier()
    graphs = []
    if not args.no_graph:
        for k in range(nsets):
            g = torch.cuda.CUDAGraph()
            F.launch_counter["n"] = 0
            with torch.cuda.graph(g):
                dws = step(k)
            graphs.append((g, F.launch_counter["n"], dws))
        for g, _, _ in graphs:
            g.replay()
        barrier()

    def run_step(k):
        if graphs:
            g, _, dws = graphs[k % nsets]
            g.replay()
        else:
            dws = step(k)
        exchange(dws)

    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    time.sleep(0.25)
    F.launch_counter["n"] = 0
    barrier()
    t_wall0 = time.time()
    e_start, e_stop = ev(), ev()
    e_start.record()
    for k in range(args.steps):
        run_step(k)
    drain()
    e_stop.record()
    barrier()
    t_wall1 = time.time()
    launches = sum(graphs[k % nsets][1] for k in range(args.steps)) if graphs else F.launch_counter["n"]
    clocks = sampler.stop(t_wall0, t_wall1)
    ms = e_start.elapsed_time(e_stop) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    px_step = B_PER_GPU * sum(hw * hw for hw in STAGES) * world
    value = px_step / (ms * 1e-3)

    # per-call device times -> roofline of the dominant kernel: for each of the four ABI calls of a step, a
    # CUDA graph of `reps` calls over the rotating buffer sets, bracketed by events on the capture stream
    calls = {}
    reps = 12
    lib = _abi.load()
    dcode = _abi.SBA_BF16 if args.dtype == "bf16" else _abi.SBA_F32
    algo = F._ALGOS[args.algo]
    for mi, hw in enumerate(STAGES):
        Q = hw * hw
        w32 = mods[mi].conv_context.weight.detach().reshape(IDF, CDF).float().contiguous()
        bufs = []
        for s in range(nsets):
            x, gc, ctx, mask = sets[s][mi]
            bufs.append(dict(x=x.detach(), gc=gc, ctx=ctx.float().contiguous(), mask=mask.to(torch.uint8).contiguous(),
                             c=torch.empty_like(x), a=torch.empty(B_PER_GPU, L, hw, hw, dtype=dtype, device=device),
                             dx=torch.empty_like(x), srcT=torch.empty(B_PER_GPU, IDF, L, device=device),
                             scr=torch.empty(3 * B_PER_GPU, dtype=torch.int32, device=device),
                             dsrc=torch.empty(B_PER_GPU * IDF * L + B_PER_GPU + 1, device=device),
                             dw=torch.empty(IDF, CDF, device=device)))

        def fwd_call(s, st):
            b = bufs[s % nsets]
            _abi.check(lib.sba_attn_fwd(b["x"].data_ptr(), b["ctx"].data_ptr(), w32.data_ptr(), b["mask"].data_ptr(),
                                        b["c"].data_ptr(), b["a"].data_ptr(), b["srcT"].data_ptr(), b["scr"].data_ptr(),
                                        B_PER_GPU, IDF, CDF, L, Q, dcode, 0, algo, st), "sba_attn_fwd")

        def bwd_call(s, st):
            b = bufs[s % nsets]
            _abi.check(lib.sba_attn_bwd(b["x"].data_ptr(), b["ctx"].data_ptr(), w32.data_ptr(), b["mask"].data_ptr(),
                                        b["srcT"].data_ptr(), b["scr"].data_ptr(), b["gc"].data_ptr(), None,
                                        b["dx"].data_ptr(), b["dsrc"].data_ptr(), b["dw"].data_ptr(), None,
                                        B_PER_GPU, IDF, CDF, L, Q, dcode, 0, algo, st), "sba_attn_bwd")

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
            e0, e1 = ev(), ev()
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) * 1e-3 / reps
            nbytes = algorithmic_bytes(B_PER_GPU * hw * hw, es, name)
            calls[f"{name}_{hw}"] = {"us": round(t * 1e6, 2), "gbs": round(nbytes / t / 1e9, 1), "bytes": nbytes}
        del bufs
    dom = max(calls, key=lambda k: calls[k]["us"])
    peaks, peak_kind = load_peaks()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"{args.dtype}_{dom}")
    roofline = {
        "bound": "hbm", "kernel": f"sba_attn_{dom.split('_')[0]} @{dom.split('_')[1]}x{dom.split('_')[1]}",
        "achieved": calls[dom]["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": round(calls[dom]["gbs"] / peaks["hbm_gbs"], 4), "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
        "traffic": traffic, "algorithmic_bytes_per_launch": calls[dom]["bytes"],
        "step_frac": round(sum(c["bytes"] for c in calls.values()) / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
        "calls": calls,
    }

    # e2e: same step through the public module with HOST buffers (rank-local, all ranks run it)
    host = [make_inputs(hw, 99 + rank + hw, dtype, device, pin=True) for hw in STAGES]
    dev_bufs = [[torch.empty_like(t, device=device) for t in h[:3]] for h in host]
    res_host = torch.empty(2 * IDF * CDF + 2, dtype=torch.float32).pin_memory()
    h2d = sum(t.numel() * t.element_size() for h in host for t in h[:3])
    d2h = res_host.numel() * 4

    def e2e_step():
        outs = []
        for m, h, dbuf, hw in zip(mods, host, dev_bufs, STAGES):
            for src, dst in zip(h[:3], dbuf):
                dst.copy_(src, non_blocking=True)
            x = dbuf[0].requires_grad_(True)
            x.grad = None
            m.conv_context.weight.grad = None
            m.applyMask(h[3].to(device, non_blocking=True))
            c_code, _att = m(x, dbuf[2])
            c_code.backward(dbuf[1])
            outs += [m.conv_context.weight.grad.float().reshape(-1), x.grad.float().sum().reshape(1)]
            dbuf[0] = x.detach()
        res_host.copy_(torch.cat(outs), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        e2e_step()
    barrier()
    n_e2e = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / n_e2e
    if world > 1:
        t = torch.tensor([e2e_s], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e = {"value": px_step / e2e_s, "unit": "region-px/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": round(e2e_s * 1e3, 3)}

    out = {
        "metric": "word-attn fwd+bwd region-px/s", "value": value, "unit": "region-px/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": WORKLOAD, "B_per_gpu": B_PER_GPU, "regions": [hw * hw for hw in STAGES],
                   "io_dtype": args.dtype, "param_dtype": "fp32", "accumulate": "fp32", "algo": args.algo, "launch": "eager" if args.no_graph else "cuda-graph replay", "mask": "ragged, reference mod-B order",
                   "g_attn": None, "l2": "inputs larger than L2: one step streams %d MB per GPU over 3 rotating buffer sets" %
                   (sum(c["bytes"] for c in calls.values()) // 2 ** 20)},
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "clocks": clocks,
    }
    return out


def reference_step_cpu(inputs, n_threads):
    """The reference's CPU path for one step (oracle/cpu_path.py: the torch op sequence of
    GlobalAttention.py:82-121 with autograd backward, torch CPU fp32)."""
    from oracle.cpu_path import attn_fwd_bwd_autograd
    for x, gc, ctx, mask, w in inputs:
        c, _attn, dX, dW = attn_fwd_bwd_autograd(x, ctx, w, mask, gc)
    return c, dX, dW


def run_reference(args):
    n_threads = os.cpu_count() or 1
    torch.set_num_threads(n_threads)
    inputs = []
    for hw in STAGES:
        x, gc, ctx, mask = make_inputs(hw, 1234 + hw, torch.float32, "cpu")
        w = torch.nn.init.orthogonal_(torch.empty(IDF, CDF), 1.0)
        inputs.append((x, gc, ctx, mask, w))
    # K and W as asked (a step is ~0.1 s on the box's host cores), bounded so that the run stays within minutes
    warmup = max(1, min(args.warmup, 50))
    for _ in range(warmup):
        reference_step_cpu(inputs, n_threads)
    steps = max(1, min(args.steps, 500))
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        reference_step_cpu(inputs, n_threads)
        ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    px = B_PER_GPU * sum(hw * hw for hw in STAGES)
    value = px / t
    sample = f"{steps} full steps (B=64, 64x64 + 128x128 fwd+bwd) fp32, median"
    return {
        "impl": "reference", "metric": "word-attn fwd+bwd region-px/s", "value": value, "unit": "region-px/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": round(t * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "B_per_gpu": B_PER_GPU, "regions": [hw * hw for hw in STAGES],
                   "note": "reference CPU path = oracle port (torch CPU); the Python reference cannot travel to the GPU box"},
        "cpu_baseline": {"value": value, "unit": "region-px/s", "cores": n_threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "region-px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def cpu_baseline_leg():
    """Bounded CPU sample beside the GPU number (rank 0, N=1 only)."""
    n_threads = os.cpu_count() or 1
    torch.set_num_threads(n_threads)
    inputs = []
    for hw in STAGES:
        x, gc, ctx, mask = make_inputs(hw, 1234 + hw, torch.float32, "cpu")
        w = torch.nn.init.orthogonal_(torch.empty(IDF, CDF), 1.0)
        inputs.append((x, gc, ctx, mask, w))
    reference_step_cpu(inputs, n_threads)
    ts = []
    t_end = time.perf_counter() + 12.0
    while len(ts) < 3 or (time.perf_counter() < t_end and len(ts) < 30):
        t0 = time.perf_counter()
        reference_step_cpu(inputs, n_threads)
        ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    px = B_PER_GPU * sum(hw * hw for hw in STAGES)
    return {"value": px / t, "unit": "region-px/s", "cores": n_threads, "kind": "port",
            "sample": f"{len(ts)} full steps of the same workload (fp32, torch CPU, {n_threads} threads), median {t*1e3:.0f} ms"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--algo", default="auto", choices=["auto", "simt", "mma"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    out = run_ours(args, rank, world, device)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_leg()
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
