#!/usr/bin/env python
"""G+D adversarial training step throughput (imgs/s) around the fused hot path, 1..8 GPUs
(BASELINE.json metric, second half).  One process per GPU:

    python tools/e2e_train.py --batch 20 --steps 10                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port 29511 tools/e2e_train.py --batch 20 --steps 10     # N GPUs, batch per GPU (weak scaling)

--attention fused | eager : the repo's kernels, or the reference's eager op sequence for the same two ops
imgs/s = global batch / (max over ranks of the device time of one full step: text encode, G forward,
3 discriminator steps, generator loss incl. DAMSM words/sent losses, backward, gradient all-reduce,
4 Adam steps, EMA).  Synthetic CUB-shaped data, random-init weights (harness/gan_step.py)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from harness.gan_step import Trainer

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=20)          # per GPU; cfg/bird_style.yml:17
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--attention", default="fused", choices=["fused", "eager"])
ap.add_argument("--amp", default="none", choices=["none", "bf16"])
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
tr = Trainer(args.batch, dev, args.attention, world, seed=1234 + rank)


def one_step():
    if args.amp == "bf16":
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return tr.step()
    return tr.step()


for _ in range(args.warmup):
    one_step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    errD, errG = one_step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
if world > 1:
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
if rank == 0:
    print(json.dumps({"metric": "G+D train imgs/s", "value": args.batch * world / (ms * 1e-3), "unit": "imgs/s", "n_gpus": world,
                      "batch_per_gpu": args.batch, "ms_per_step": round(ms, 2), "attention": args.attention, "amp": args.amp,
                      "steps": args.steps, "errD": float(errD), "errG": float(errG), "data": "synthetic", "scaling": "weak"}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
