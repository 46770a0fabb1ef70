"""Quick device timing of the attention forward/backward ABI calls (development aid;
bench.py is the contract).  Usage: python tools/time_attn.py [algo] [dtype] [B,B,...]
(the batch list is the bandwidth sweep of BASELINE.json configs[4]: 1024 / 8 GPUs = 128 per GPU)"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sba_gan_b200 import _abi
from sba_gan_b200.functional import _DTYPES, _ALGOS

algo = _ALGOS[sys.argv[1]] if len(sys.argv) > 1 else 0
dt = torch.bfloat16 if len(sys.argv) > 2 and sys.argv[2] == "bf16" else torch.float32
lib = _abi.load()
dev = "cuda"
idf, cdf, L = 32, 256, 18
PEAK = 6553.6
BS = [int(b) for b in sys.argv[3].split(",")] if len(sys.argv) > 3 else [64]
for B, hw in [(b, h) for b in BS for h in (64, 128)]:
    Q = hw * hw
    nset = max(2, int(600e6 // (B * idf * Q * 4 * 5)) + 1)   # rotate > L2
    sets = []
    for s in range(nset):
        x = torch.randn(B, idf, Q, device=dev).to(dt)
        g = torch.randn(B, idf, Q, device=dev).to(dt)
        sets.append((x, g, torch.empty_like(x), torch.empty(B, L, Q, device=dev, dtype=dt), torch.empty_like(x)))
    ctx = torch.tanh(torch.randn(B, cdf, L, device=dev))
    W = torch.randn(idf, cdf, device=dev) / 16
    lens = torch.randint(5, L + 1, (B,))
    mask = (torch.arange(L)[None] >= lens[:, None]).to(torch.uint8).to(dev)
    srcT = torch.empty(B, idf, L, device=dev)
    mb = torch.empty(3 * B, dtype=torch.int32, device=dev)
    dSrc = torch.empty(lib.sba_attn_bwd_workspace_floats(B, idf, cdf, L), device=dev)
    dW = torch.empty(idf, cdf, device=dev)
    st_holder = [torch.cuda.current_stream().cuda_stream]
    dcode = _DTYPES[dt]

    def fwd(k):
        x, g, c, a, dx = sets[k % nset]
        rc = lib.sba_attn_fwd(x.data_ptr(), ctx.data_ptr(), W.data_ptr(), mask.data_ptr(), c.data_ptr(), a.data_ptr(),
                              srcT.data_ptr(), mb.data_ptr(), B, idf, cdf, L, Q, dcode, 0, algo, st_holder[0])
        _abi.check(rc, "fwd")

    def bwd(k):
        x, g, c, a, dx = sets[k % nset]
        rc = lib.sba_attn_bwd(x.data_ptr(), ctx.data_ptr(), W.data_ptr(), mask.data_ptr(), srcT.data_ptr(), mb.data_ptr(),
                              g.data_ptr(), None, dx.data_ptr(), dSrc.data_ptr(), dSrc.numel(), dW.data_ptr(), None, B, idf, cdf, L, Q,
                              dcode, 0, algo, st_holder[0])
        _abi.check(rc, "bwd")

    def timeit(fn, n=30):
        """device time per call: n calls (rotating buffer sets) captured in one CUDA graph"""
        for k in range(5):
            fn(k)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            st_holder[0] = torch.cuda.current_stream().cuda_stream
            with torch.cuda.graph(gr, stream=side):
                st_holder[0] = torch.cuda.current_stream().cuda_stream
                for k in range(n):
                    fn(k)
        st_holder[0] = torch.cuda.current_stream().cuda_stream
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        reps = int(os.environ.get("SBA_REPLAYS", "1"))      # > 1: sustained run (clock / power sampling)
        e0.record()
        for _ in range(reps):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (n * reps) * 1e-3

    es = 4 if dt == torch.float32 else 2
    tf = timeit(fwd)
    tb = timeit(bwd)
    px = B * Q
    bf = px * (2 * idf + L) * es
    bb = px * 3 * idf * es
    print(f"B={B} {hw}x{hw} {dt} algo={algo}: fwd {tf*1e6:.1f} us {bf/tf/1e9:.0f} GB/s ({bf/tf/1e9/PEAK:.2%}) | "
          f"bwd {tb*1e6:.1f} us {bb/tb/1e9:.0f} GB/s ({bb/tb/1e9/PEAK:.2%}) | "
          f"fwd+bwd {px/(tf+tb)/1e9:.2f} Gpx/s {(bf+bb)/(tf+tb)/1e9:.0f} GB/s ({(bf+bb)/(tf+tb)/1e9/PEAK:.2%})", flush=True)
