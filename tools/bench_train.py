#!/usr/bin/env python
"""Training-step timing, BASELINE.json configs[4]: PDA-SSD KITTI config, bf16 autocast, 8 scenes per GPU, data parallel with
torch DDP — ONE collective, the gradient allreduce over NCCL / NVLink (tools/train.py:154 semantics in the reference).

    python tools/bench_train.py [--gpus 1] [--steps 10]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_train.py --gpus N

What runs: the model in train mode — FPS / ball query / group / gather through libpdab.so with their (deterministic) gradient
entry points, the SA / PDA / vote / head modules in the reference's statement order on autograd, BatchNorm in batch-statistics
mode — forward, backward, DDP allreduce, AdamW step.  What does NOT: the reference head's target assignment and loss terms
(pcdet/models/dense_heads/IASSD_head.py:169-1330) are not ported, so the objective is a SURROGATE that reaches every
parameter (mean squares of the class logits, box codes, vote offsets and per-layer confidence logits).  The number times the
compute and communication path of a step, not convergence.  Prints one JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def surrogate_loss(out):
    loss = out["batch_cls_preds"].float().pow(2).mean() + out["batch_box_preds"][..., :6].float().pow(2).mean() * 1e-3
    loss = loss + out["ctr_offsets"][:, 1:].float().pow(2).mean()
    for p in out["sa_ins_preds"]:
        if torch.is_tensor(p) and p.numel():
            loss = loss + p[..., 1:].float().pow(2).mean()
    return loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--points", type=int, default=16384)
    ap.add_argument("--no-autocast", action="store_true")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from pdanet_b200.config import load_config
    from pdanet_b200.iassd import build_model
    from pdanet_b200.synthetic import make_batch
    cfg = load_config("kitti")
    torch.manual_seed(0)
    model = build_model(cfg).to(dev).train()
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01)
    batches = [make_batch(args.batch, args.points, cfg.POINT_CLOUD_RANGE, first_scene=(rank * 4 + r) * args.batch)["points"].to(dev)
               for r in range(4)]

    def step(k):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not args.no_autocast):
            out = net({"batch_size": args.batch, "points": batches[k % 4]})
            loss = surrogate_loss(out)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for k in range(args.warmup):
        loss = step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for k in range(args.steps):
        loss = step(k)
    e.record()
    e.synchronize()
    ms = s.elapsed_time(e)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({
            "metric": "PDA-SSD KITTI training step (surrogate objective), scenes/s", "value": world * args.batch * args.steps / (ms / 1e3),
            "unit": "scenes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "scaling": "weak", "dtype": "fp32 parameters, fp32 native ops, torch layers under bf16 autocast" if not args.no_autocast else "f32",
            "config": {"workload": f"PDA-SSD kitti cfg train-mode forward + backward + AdamW, batch {args.batch} x {args.points} pts per GPU",
                       "parallelism": f"DDP x{world}", "collective": "NCCL allreduce of the gradients (DistributedDataParallel buckets)"
                       if world > 1 else "none (1 GPU)", "gradient_bytes_per_step": n_params * 4,
                       "objective": "surrogate (the reference's target assignment / losses are not ported)"},
            "final_loss": float(loss)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
