"""Kernel timeline of back-to-back attention calls (development aid; needs the dev build:
`python -m sba_gan_b200.build --dev`, which adds globaltimer stamps to the tcgen05 kernels).

    python tools/timeline.py [fwd|bwd|step] [hw] [B] [dtype]

Captures 16 calls in one CUDA graph, replays it, and prints for every call (us since the first stamp):
head kernel first..last | streaming kernel entry, past its grid dependency, last tile, exit | finish kernel past its
dependency..exit.  The product library has none of these stamps (rebuild without --dev afterwards)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from sba_gan_b200 import _abi  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "bwd"
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 64
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dt = torch.float32 if len(sys.argv) > 4 and sys.argv[4] == "fp32" else torch.bfloat16
lib = _abi.load()
raw = ctypes.CDLL(_abi.LIB_PATH)
if not hasattr(raw, "sba_dev_timeline_start"):
    raise SystemExit("not a dev build: python -m sba_gan_b200.build --dev")
idf, cdf, L, Q = 32, 256, 18, hw * hw
dev = "cuda"
NCALL = 16
sets = []
for s in range(3):
    x = torch.randn(B, idf, Q, device=dev).to(dt)
    g = torch.randn(B, idf, Q, device=dev).to(dt)
    sets.append(dict(x=x, g=g, c=torch.empty_like(x), a=torch.empty(B, L, Q, device=dev, dtype=dt), dx=torch.empty_like(x),
                     ws=torch.empty(lib.sba_attn_bwd_workspace_floats(B, idf, cdf, L), device=dev),
                     dw=torch.empty(idf, cdf, device=dev), srcT=torch.empty(B, idf, L, device=dev),
                     mb=torch.empty(3 * B, dtype=torch.int32, device=dev)))
ctx = torch.tanh(torch.randn(B, cdf, L, device=dev))
W = torch.randn(idf, cdf, device=dev) / 16
lens = torch.randint(5, L + 1, (B,))
mask = (torch.arange(L)[None] >= lens[:, None]).to(torch.uint8).to(dev)
dcode = 1 if dt == torch.bfloat16 else 0


def fwd(k, st):
    b = sets[k % 3]
    _abi.check(lib.sba_attn_fwd(b["x"].data_ptr(), ctx.data_ptr(), W.data_ptr(), mask.data_ptr(), b["c"].data_ptr(),
                                b["a"].data_ptr(), b["srcT"].data_ptr(), b["mb"].data_ptr(), B, idf, cdf, L, Q, dcode, 0, 0, st), "fwd")


def bwd(k, st):
    b = sets[k % 3]
    _abi.check(lib.sba_attn_bwd(b["x"].data_ptr(), ctx.data_ptr(), W.data_ptr(), mask.data_ptr(), b["srcT"].data_ptr(),
                                b["mb"].data_ptr(), b["g"].data_ptr(), None, b["dx"].data_ptr(), b["ws"].data_ptr(),
                                b["ws"].numel(), b["dw"].data_ptr(), None, B, idf, cdf, L, Q, dcode, 0, 0, st), "bwd")


cur = torch.cuda.current_stream().cuda_stream
for k in range(3):
    fwd(k, cur)
    bwd(k, cur)
torch.cuda.synchronize()
seq = {"fwd": [fwd] * NCALL, "bwd": [bwd] * NCALL, "step": [fwd, bwd] * (NCALL // 2)}[which]
raw.sba_dev_timeline_start()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    st = torch.cuda.current_stream().cuda_stream
    for k, fn in enumerate(seq):
        fn(k, st)
n = raw.sba_dev_timeline_stop()
gr.replay()
torch.cuda.synchronize()
raw.sba_dev_timeline_clear()
gr.replay()
torch.cuda.synchronize()
NS = 16
buf = (ctypes.c_ulonglong * (NS * n))()
raw.sba_dev_timeline_read(buf, n)
v = list(buf)
NONE_MIN, t0 = 2 ** 64 - 1, min(x for x in v if 0 < x < 2 ** 64 - 1)
us = lambda x: "     -  " if x in (0, NONE_MIN) else f"{(x - t0) * 1e-3:8.2f}"  # noqa: E731
print(f"{which} B={B} {hw}x{hw} {dt}: {n} calls, us since the first stamp")
print("call | head first..last      | stream entry  past-dep  last-tile      exit | finish past-dep..exit   | period")
prev = None
for c in range(n):
    r = v[NS * c:NS * c + NS]
    end = max(x for x in (r[1], r[3], r[5]) if x not in (0, NONE_MIN))
    period = "" if prev is None else f"{(end - prev) * 1e-3:7.2f}"
    prev = end
    print(f"{c:4d} | {us(r[0])}..{us(r[1])} | {us(r[2])} {us(r[6])} {us(r[7])} {us(r[3])} | {us(r[4])}..{us(r[5])} | {period}")
    if r[8] not in (0, NONE_MIN) and c >= n - 3:
        print(f"       finish blocks: entry {us(r[8])}..{us(r[9])} past-dep {us(r[10])}..{us(r[11])} staged {us(r[12])}..{us(r[13])} "
              f"FMAs done {us(r[14])}..{us(r[15])} counted {us(r[0])}..{us(r[1])}")
