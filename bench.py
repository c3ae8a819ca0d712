#!/usr/bin/env python
"""bench.py — PDA-SSD scenes/sec on synthetic KITTI-shape point clouds (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl pdab|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step = one full inference pass (backbone + vote + centroid aggregation + head + 3D NMS) over one batch
of 16 synthetic 16384-point scenes per GPU (BASELINE.json configs[1]); scenes are sharded by GPU with no
data-path collective (weak scaling).  Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement".
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
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "PDA-SSD scenes/sec @16384 pts (KITTI cfg, full inference)"
UNIT = "scenes/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200, help="timed steps (default: ~1.2 s of device time per leg)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="pdab", choices=["pdab", "reference"])
    ap.add_argument("--config", default="kitti", choices=["kitti", "once"])
    ap.add_argument("--batch", type=int, default=None, help="scenes per GPU per step (default 16 kitti / 32 once: BASELINE configs)")
    ap.add_argument("--points", type=int, default=None)
    ap.add_argument("--cpu-scenes", type=int, default=8, help="scenes in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--depth", type=int, default=8, help="batches in flight per GPU (ScenePipeline slots)")
    ap.add_argument("--reserve-sms", type=int, default=None,
                    help="SMs the persistent tensor-core grid leaves free for the other batches' FPS kernels (default: batch)")
    ap.add_argument("--kernels", type=int, default=12, help="how many per-kernel rows to keep in the JSON line")
    ap.add_argument("--no-graphs", action="store_true", help="do not capture the forward in CUDA graphs")
    ap.add_argument("--tc-passes", type=int, default=None, choices=[1, 2, 3, 4],
                    help="tensor-core product mode of our GEMM kernels: 4 = fp16 single pass, fp16 activations between kernels, "
                         "(hi, lo) fp16 residual streams; 3 = 3xTF32 (fp32-level); 2 = split-bf16 (2^-16); 1 = TF32; "
                         "default = the modules' default")
    ap.add_argument("--matmul", default="ieee", choices=["ieee", "tf32"],
                    help="fp32 matmul mode of the unchanged PyTorch layers (PDA transformer); our kernels are fp32 either way")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def tensor_peak():
    """Dense bf16 tensor peak in TFLOP/s: the measured cuBLAS bf16 throughput (sustained figure: the kernel is timed
    inside a long step)."""
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["bf16_tflops_sustained"]), "measured bf16_tflops_sustained (MEASURED_PEAKS.json)"
    return 1590.0, "fallback 1.59 PFLOP/s bf16 (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------- CPU arm

def cpu_model(cfg, name):
    """The CPU arm's model.  Backbone (98 % of the FLOPs, every native op): the REFERENCE's own `IASSD_Backbone` and SA / PDA /
    vote modules, imported unchanged from the copies oracle/build_ref.py stages under oracle/_ref/pyref
    (oracle/ref_python.py), over the CPU oracle's C restatement of FPS / ball query / group / gather.  Head, decode and
    post-processing: our mirror (pinned to the reference's by tests/golden/ref_head_*.npz) over the oracle NMS.  Without the
    staged files the backbone is our mirror as well.  Returns (model, what)."""
    from oracle import ref_python, torch_ops
    from pdanet_b200.config import load_config
    from pdanet_b200.iassd import build_model
    torch.manual_seed(0)
    model = build_model(cfg, ops=torch_ops, nms_utils=torch_ops.nms_utils, batched_post_processing=False).eval()
    what = "oracle port (C ops + our module mirror on torch CPU)"
    if ref_python.reference_root() is not None:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _, _, bb_mod = ref_python.import_reference()
        fresh = load_config(name)   # the reference's constructor edits the mlp specs in place
        ref_bb = bb_mod.IASSD_Backbone(fresh.MODEL.BACKBONE_3D, num_class=len(cfg.CLASS_NAMES),
                                       input_channels=cfg.get("NUM_POINT_FEATURES", 4)).eval()
        ref_bb.load_state_dict(model.backbone_3d.state_dict())
        model.backbone_3d = ref_bb
        model.module_list = [model.backbone_3d, model.point_head]
        what = "the reference's own IASSD_Backbone / SA / PDA / vote modules (oracle/_ref/pyref) over the C oracle ops"
    return model, what


def time_cpu(cfg, n_points, scenes, warmup=1, name="kitti"):
    """The CPU arm on the host cores: `scenes` single-scene batches; returns (scenes/s, threads, per-scene seconds, what ran)."""
    from pdanet_b200.synthetic import make_batch
    model, what = cpu_model(cfg, name)
    torch.set_num_threads(os.cpu_count() or 1)
    times = []
    with torch.no_grad():
        for s in range(warmup + scenes):
            batch = make_batch(1, n_points, cfg.POINT_CLOUD_RANGE, first_scene=s)
            t0 = time.perf_counter()
            model(batch)
            if s >= warmup:
                times.append(time.perf_counter() - t0)
    return len(times) / sum(times), torch.get_num_threads(), times, what


def run_reference_arm(args, cfg, n_points, batch):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    per_step = 1 if args.steps > 20 else max(1, min(2, batch))  # bounded sample: scenes per step (~0.7 s of CPU work each)
    sps, threads, times, what = time_cpu(cfg, n_points, scenes=per_step * args.steps, warmup=max(1, args.warmup),
                                         name=args.config)
    sample = f"{per_step} scene(s)/step x {args.steps} steps, one {n_points}-point scene per forward, batch 1; {what}"
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"PDA-SSD {args.config} cfg full inference, {n_points} pts/scene, CPU oracle port",
                   "scenes_per_step": per_step},
        "cpu_baseline": {"value": sps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------- GPU arm

def algorithmic_bytes(key: str, batch: int):
    """HBM bytes one launch must move (SURVEY.md §8d / DESIGN.md 'Kernels'); key = 'entry(sizes...)'."""
    name, _, rest = key.partition("(")
    a = [int(x) for x in rest.strip(")").split(",") if x.strip()]
    if name == "pdab_fps":
        b, n, m = a[:3]
        return b * (12 * n + 4 * n * 2 + 4 * m)      # xyz read, temp read + written back, idx written
    if name == "pdab_ball_query":
        b, n, m, ns = a[0], a[1], a[2], a[3]
        return b * (12 * n + 12 * m + 4 * m * ns)
    if name in ("pdab_group_points", "pdab_gather_points"):
        b, c, n = a[:3]
        e = a[3] * (a[4] if name == "pdab_group_points" else 1)
        return b * (4 * c * e * 2 + 4 * e)          # gathered elements read + written, idx read
    if name == "pdab_pda_group":
        b, c, n, m, ns = a[:5]
        return b * (12 * n + 4 * c * n + 12 * m + 4 * (7 + c) * m * ns)
    if name == "pdab_pda_group_tokens":
        b, c, n, m, ns = a[:5]
        return b * (12 * n + 4 * c * n + 12 * m + 4 * (8 + c) * m * ns)
    if name == "pdab_sa_fused":
        b, c, n, m, ns = a[:5]
        return b * (12 * n + 4 * c * n + 12 * m)    # + output, added by the caller (needs cout)
    if name == "pdab_topk_ctr":
        b, n, c, k = a[:4]
        return b * (4 * n * c + 4 * k)
    if name == "pdab_tc_linear":
        rows, k, nout, npass, bn, epi = a[:6]
        out_rows = rows if epi not in (3, 4) else rows // 16   # max-pool epilogues write one row per neighbourhood
        out_cols = nout // 3 if epi == 5 else nout             # attention epilogue: only ctx (rows, E) is written
        resid = rows * nout * 4 if epi in (2, 3) else 0
        return rows * k * 4 + out_rows * out_cols * 4 + resid + nout * k * 4 * (2 if npass == 3 else 1)
    if name == "pdab_tc_linear_h":     # fp16 single pass; key ints: rows, k, nout, bn, epilogue, lda (pointers are filtered out)
        rows, k, nout, bn, epi = a[:5]
        # A read once as fp16 (an fp32 A, converted in the kernel, is counted at 2 B too: the smaller, algorithmic figure);
        # residual = (hi, lo) fp16 planes = 4 B; outputs: pooled fp32, LayerNorm (hi, lo) planes 4 B, everything else fp16
        out_rows = rows if epi not in (3, 4) else rows // 16
        out_cols = nout // 3 if epi == 5 else nout
        out_b = 4 if epi in (2, 3, 4) else 2
        resid = rows * nout * 4 if epi in (2, 3) else 0
        return rows * k * 2 + out_rows * out_cols * out_b + resid + nout * k * 2
    if name == "pdab_tc_ffn_h":        # rows, e, nsample, ldc: ctx fp16 + y (hi, lo) read, pooled fp32 written, weights once
        rows, e, ns = a[:3]
        return rows * e * (2 + 4) + rows // ns * e * 4 + 2 * (e * e + e * e)
    if name == "pdab_tc_sa_gather_linear":
        b, c, n, m, ns, nout = a[:6]
        return b * (4 * c * n + 12 * n + 12 * m + 4 * m * ns) + b * m * ns * nout * 4
    if name == "pdab_tc_sa_gather_linear_h":
        b, c, n, m, ns, nout = a[:6]
        return b * (4 * c * n + 12 * n + 12 * m + 4 * m * ns) + b * m * ns * nout * 2
    if name in ("pdab_pda_encode_ln", "pdab_pda_encode_ln_h"):
        b, c, n, m = a[:4]
        ns = a[4] if len(a) > 4 else 16     # (radius is a float and not in the key; nsample is)
        return b * (12 * n + 4 * c * n + 12 * m + 4 * c * m + 16 * c * m * ns)   # output (tokens, 4c) at 4 B dominates
    if name in ("pdab_sa_fused_pair", "pdab_sa_fused_pair_h"):
        b, c, n, m = a[:4]
        return b * (12 * n + 4 * c * n + 12 * m + 4 * 96 * m)   # both scales: 32 + 64 output channels per centre
    if name == "pdab_group_attention":
        groups, ns, heads, hd = a[:4]
        return groups * ns * heads * hd * 4 * 4          # q, k, v read + ctx written
    if name == "pdab_group_attention_h":
        groups, ns, heads, hd = a[:4]
        return groups * ns * heads * hd * 4 * 2          # the same in fp16
    if name == "pdab_nms_batched":
        s, stride = a[:2]
        return s * (28 * stride + 8 * stride * ((stride + 63) // 64) + 8 * stride)
    return None


# C-ABI entry point -> the kernel it launches (launches are judged per kernel: the tcgen05 GEMM runs ~30 times per step at
# different shapes / prologues / epilogues and is ONE kernel template)
KERNEL_OF = {"pdab_tc_linear": "tc_gemm_kernel", "pdab_tc_linear_h": "tc_gemm_kernel", "pdab_tc_sa_gather_linear": "tc_gemm_kernel",
             "pdab_tc_sa_gather_linear_h": "tc_gemm_kernel", "pdab_fps": "fps_kernel", "pdab_fps_with_dist": "fps_kernel",
             "pdab_pda_encode_ln": "pda_encode_ln_kernel", "pdab_pda_encode_ln_h": "pda_encode_ln_kernel",
             "pdab_group_attention": "group_attention_kernel", "pdab_group_attention_h": "group_attention_kernel",
             "pdab_sa_fused_pair": "sa_fused_pair_kernel", "pdab_sa_fused_pair_h": "sa_fused_pair_kernel", "pdab_tc_ffn_h": "ffn_fused_kernel"}

# tensor-pipe work per algorithmic flop, in bf16-MMA flops: split-bf16 ("bf16x3") issues 3 bf16 MMAs per product, 3xTF32
# issues 3 TF32 MMAs (a TF32 MMA occupies the pipe like 2 bf16 MMAs), plain TF32 one
MMA_COST = {2: 3.0, 3: 6.0, 1: 2.0, 4: 1.0}


def algorithmic_flops(key: str):
    """(algorithmic flops, tensor-pipe work in bf16-MMA-equivalent flops) of one launch of a tensor-core entry point."""
    name, _, rest = key.partition("(")
    a = [int(x) for x in rest.strip(")").split(",") if x.strip()]
    if name == "pdab_tc_linear":
        rows, k, nout, npass = a[:4]
        f = 2.0 * rows * k * nout
        return f, f * MMA_COST[npass]
    if name == "pdab_tc_linear_h":
        rows, k, nout = a[:3]
        f = 2.0 * rows * k * nout
        return f, f * MMA_COST[4]
    if name == "pdab_tc_ffn_h":        # out_proj (e x e) + linear1 (e x e/2) + linear2 (e/2 x e)
        rows, e = a[:2]
        f = 2.0 * rows * (e * e + e * e // 2 + e * e // 2)
        return f, f * MMA_COST[4]
    if name == "pdab_tc_sa_gather_linear":
        b, c, n, m, ns, nout, npass = a[:7]
        f = 2.0 * b * m * ns * (c + 3) * nout
        return f, f * MMA_COST[npass]
    if name == "pdab_tc_sa_gather_linear_h":
        b, c, n, m, ns, nout = a[:6]
        f = 2.0 * b * m * ns * (c + 3) * nout
        return f, f * MMA_COST[4]
    if name in ("pdab_sa_fused_pair", "pdab_sa_fused_pair_h"):       # L0: (4 -> 16 -> 16 -> 32) x ns_a + (4 -> 32 -> 32 -> 64) x ns_b rows per centre
        b, c, n, m = a[:4]
        ns_a, ns_b = 16, 32
        f = 2.0 * b * m * (ns_a * ((3 + c) * 16 + 16 * 16 + 16 * 32) + ns_b * ((3 + c) * 32 + 32 * 32 + 32 * 64))
        return f, f * 3.0
    return None


def run_gpu_arm(args, cfg, n_points, batch):
    from pdanet_b200 import _lib
    from pdanet_b200.runner import ScenePipeline, SceneRunner
    from pdanet_b200.synthetic import make_batch

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    assert torch.cuda.is_available(), "bench.py --impl pdab needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    _lib.lib()  # fail loudly if the extension is missing

    torch.backends.cuda.matmul.allow_tf32 = args.matmul == "tf32"
    runner = SceneRunner(cfg, device=dev, batch_size=batch, num_points=n_points, seed=0, tc_passes=args.tc_passes)
    # weak scaling: every rank owns its own scenes (global scene ids rank*batch*R ...); the steps rotate over
    # R distinct synthetic batches so consecutive steps never see the same input
    R = 4
    hosts = [make_batch(batch, n_points, cfg.POINT_CLOUD_RANGE, first_scene=(rank * R + r) * batch)["points"].pin_memory()
             for r in range(R)]
    devs = [h.to(dev) for h in hosts]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    pipe = ScenePipeline(runner, depth=args.depth, graphs=not args.no_graphs, warm_points=hosts[0],
                         reserve_sms=args.reserve_sms)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for w in range(args.warmup):
        runner.infer_device(devs[w % R])
    pipe.run_device([devs[w % R] for w in range(max(args.warmup, args.depth))])
    pipe.run([hosts[w % R] for w in range(max(args.warmup, args.depth))])
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    def timed_pipeline(fn, batches):
        """EXACTLY K steps through the pipelined runner, bracketed by barrier + synchronize; device time between a start
        event every slot stream waits on and an end event that waits for every slot."""
        flush.zero_()
        barrier()
        s = torch.cuda.Event(enable_timing=True)
        s.record()
        fn(batches, start_event=s)
        e = pipe.join_event()
        e.synchronize()
        barrier()
        return s.elapsed_time(e)

    seq_steps = min(args.steps, 20)   # the sequential legs only feed the per-kernel table and the latency figure

    def timed_steps(fn):
        """`seq_steps` sequential steps, each bracketed by its own CUDA events on the current stream; L2 flushed outside them."""
        ms = []
        barrier()
        for k in range(seq_steps):
            flush.zero_()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn(k)
            e.record()
            e.synchronize()
            ms.append(s.elapsed_time(e))
        barrier()
        return ms

    # ---- leg 1: device-resident inputs, K steps pipelined (whole-job throughput)
    total_ms = max_over_ranks(timed_pipeline(pipe.run_device, [devs[k % R] for k in range(args.steps)]))
    value = world * batch * args.steps / (total_ms / 1e3)
    launches = pipe.launches_per_step * args.steps

    # ---- leg 1b: K sequential, un-pipelined steps (latency of one batch), then the same again with CUDA events around
    # every C-ABI call (per-kernel durations for the roofline; kept apart so the event records do not perturb anything)
    seq_ms = timed_steps(lambda k: runner.infer_device(devs[k % R]))
    _lib.enable_timing(True)
    inst_ms = timed_steps(lambda k: runner.infer_device(devs[k % R]))
    kernel_ms = _lib.timings_ms()
    _lib.enable_timing(False)

    # ---- leg 2: end to end through the public API — pinned host input, H2D + forward + D2H of the predictions inside
    # the timed region, K steps pipelined
    e2e_total = max_over_ranks(timed_pipeline(pipe.run, [hosts[k % R] for k in range(args.steps)]))
    e2e_value = world * batch * args.steps / (e2e_total / 1e3)
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel among ours (by time inside the sequential instrumented steps).
    # Launches are grouped by C-ABI entry point (= by kernel): the tensor-core GEMM kernel runs ~25 times per step at
    # different shapes and is judged as one kernel: sum of algorithmic work / sum of launch durations.
    hbm_peak, peak_src = peaks()
    tc_peak, tc_src = tensor_peak()
    per_kernel, groups = [], {}
    for key, ms in kernel_ms.items():
        if not ms:
            continue
        avg = sum(ms) / len(ms)
        nbytes = algorithmic_bytes(key, batch)
        flops = algorithmic_flops(key)
        per_kernel.append({"kernel": key, "calls_per_step": len(ms) / seq_steps, "avg_ms": round(avg, 4),
                           "ms_per_step": round(sum(ms) / seq_steps, 4),
                           "algorithmic_GBps": round(nbytes / avg / 1e6, 2) if nbytes else None,
                           "algorithmic_TFLOPs": round(flops[0] / avg / 1e9, 1) if flops else None})
        g = groups.setdefault(KERNEL_OF.get(key.partition("(")[0], key.partition("(")[0]),
                              {"ms": 0.0, "launches": 0, "bytes": 0.0, "flops": 0.0, "mma": 0.0})
        g["ms"] += sum(ms)
        g["launches"] += len(ms)
        g["bytes"] += (nbytes or 0) * len(ms)
        if flops:
            g["flops"] += flops[0] * len(ms)
            g["mma"] += flops[1] * len(ms)
    per_kernel.sort(key=lambda r: -r["ms_per_step"])
    # BASELINE.json's second figure: FPS + group + MLP microseconds per scene (sequential instrumented steps, our kernels)
    stage_of = {"pdab_fps": "fps", "pdab_fps_with_dist": "fps", "pdab_topk_ctr": "topk_sampling",
                "pdab_ball_query": "group", "pdab_gather_points": "group", "pdab_group_points": "group",
                "pdab_pda_group": "group", "pdab_pda_group_tokens": "group", "pdab_pda_encode_ln": "group_encode_pda",
                "pdab_pda_assemble_ln_split": "group_encode_pda", "pdab_sa_fused": "fused_group_mlp_maxpool",
                "pdab_sa_fused_pair": "fused_group_mlp_maxpool", "pdab_sa_fused_pair_h": "fused_group_mlp_maxpool", "pdab_tc_sa_gather_linear": "fused_group_mlp_maxpool",
                "pdab_tc_linear": "mlp_gemm", "pdab_group_attention": "attention", "pdab_nms_batched": "nms",
                "pdab_tc_linear_h": "mlp_gemm", "pdab_tc_sa_gather_linear_h": "fused_group_mlp_maxpool",
                "pdab_pda_encode_ln_h": "group_encode_pda", "pdab_group_attention_h": "attention",
                "pdab_ball_query_grid": "group", "pdab_tc_ffn_h": "mlp_gemm"}
    stages = {}
    for key, ms in kernel_ms.items():
        st = stage_of.get(key.partition("(")[0], "other")
        stages[st] = stages.get(st, 0.0) + sum(ms) / seq_steps / batch * 1e3
    stages = {k: round(v, 1) for k, v in sorted(stages.items())}
    roofline = None
    if groups:
        # Dominant kernel = largest share of the GPU's work, i.e. of SM-time: a kernel's device time x the fraction of the SMs
        # its grid occupies.  FPS is a latency chain on one CTA (cluster) per scene — 16 of 148 SMs for the KITTI batch, which
        # is why ScenePipeline overlaps it with other batches — every other kernel of ours fills the machine.
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        fps_ctas = batch * max(1, -(-n_points // 16384))
        occupancy = {"fps_kernel": min(1.0, fps_ctas / sms)}
        sm_ms = {k: v["ms"] * occupancy.get(k, 1.0) for k, v in groups.items()}
        name = max(sm_ms, key=sm_ms.get)
        g = groups[name]
        step_ms = sum(inst_ms) / seq_steps
        common = {"kernel": name, "launches_per_step": g["launches"] / seq_steps,
                  "avg_launch_ms": round(g["ms"] / g["launches"], 4), "share_of_step": round(g["ms"] / seq_steps / step_ms, 4),
                  "share_of_sm_time": round(sm_ms[name] / sum(sm_ms.values()), 4),
                  "sm_time_shares": {k: round(v / sum(sm_ms.values()), 4) for k, v in sorted(sm_ms.items(), key=lambda kv: -kv[1])[:6]},
                  "traffic": None}
        tr = ROOT / "profiles" / "r02_tc_gemm_traffic.json"
        if name == "tc_gemm_kernel" and tr.exists() and args.config == "kitti" and batch == 16:
            t = json.loads(tr.read_text())  # ncu dram__bytes_read + write per launch, same workload (see its "source")
            common["traffic"] = round(t["dram_bytes_per_launch"])
            common["traffic_unit"] = "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the step's launches)"
            common["algorithmic_bytes_per_launch"] = round(g["bytes"] / g["launches"])
            common["traffic_source"] = "profiles/r02_tc_gemm_traffic.json"
        if g["flops"] > 0:  # tensor-core kernel: algorithmic flops = 2*rows*k*nout of the fp32 product it computes
            achieved = g["flops"] / g["ms"] / 1e9
            roofline = {"bound": "tensor", "achieved": round(achieved, 1), "peak": tc_peak, "unit": "TFLOP/s",
                        "frac": round(achieved / tc_peak, 4), "peak_source": tc_src,
                        "mma_issued_TFLOPs": round(g["mma"] / g["ms"] / 1e9, 1),
                        "mma_issued_frac": round(g["mma"] / g["ms"] / 1e9 / tc_peak, 4),
                        "hbm_GBps": round(g["bytes"] / g["ms"] / 1e6, 1),
                        "note": "`achieved` = algorithmic flops (2*rows*k*nout per launch, summed over the kernel's launches "
                                "of one step) / their summed durations, against the measured dense bf16 peak; "
                                "`mma_issued_frac` = the same with every product weighted by the MMAs it issues (1 in the "
                                "fp16 single-pass mode, 3 in the split modes); `hbm_GBps` = the launches' algorithmic bytes / "
                                "the same time (half of them are bound by that, not by the tensor pipe)", **common}
        else:
            achieved = g["bytes"] / g["ms"] / 1e6
            roofline = {"bound": "hbm", "achieved": round(achieved, 2), "peak": hbm_peak, "unit": "GB/s",
                        "frac": round(achieved / hbm_peak, 5), "peak_source": peak_src, **common}
            if name == "pdab_fps":
                roofline["note"] = ("FPS is a serial chain bound by on-chip ALU/shared-memory latency, not HBM "
                                    "(SURVEY.md §8d)")

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sps, threads, times, what = time_cpu(cfg, n_points, scenes=args.cpu_scenes, name=args.config)
        cpu = {"value": sps, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.cpu_scenes} of the workload's scenes, one {n_points}-point scene per forward "
                         f"(batch 1), {sum(times):.1f} s of CPU work; {what}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"PDA-SSD {args.config} cfg full inference (backbone+vote+centroid aggregation+head+3D NMS), "
                               f"batch {batch} x {n_points} pts per GPU, random-init weights",
                   "scenes_per_gpu_per_step": batch, "points_per_scene": n_points, "parallelism": f"scene-sharded x{world}",
                   "pipeline": f"{args.depth} batches in flight per GPU (one stream + one CUDA graph each)"
                               if not args.no_graphs else f"{args.depth} batches in flight per GPU (streams, eager launches)",
                   "l2": "256 MiB flush before the timed region; inside it every step streams ~2.5 GB of intermediates "
                         "(>> 126 MB L2) and the steps rotate over 4 distinct input batches",
                   "sequential_ms_per_step": round(sum(seq_ms) / len(seq_ms), 3),
                   "tc_passes": next((m.tc_passes for m in runner.model.modules() if hasattr(m, "tc_passes")), None),
                   "torch_layers": f"fp32 matmul mode {args.matmul}; cuDNN 1x1 convs TF32-allowed (torch default, as the reference)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
                "d2h_bytes_per_step": pipe.d2h_bytes, "ms_per_step": e2e_total / args.steps},
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "stage_us_per_scene": stages, "kernels": per_kernel[:args.kernels],
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout; everything else that writes to fd 1 (NCCL prints its
    version banner there) has been redirected to stderr by main()."""
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


_OUT = sys.stdout


def main():
    global _OUT
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse()
    from pdanet_b200.config import load_config
    cfg = load_config(args.config)
    n_points = args.points or cfg.NUM_POINTS
    batch = args.batch or (16 if args.config == "kitti" else 32)
    if args.config == "once":
        global METRIC
        METRIC = "PDA-SSD scenes/sec @65536 pts (ONCE cfg, full inference)"
        if args.steps == 200:       # the default: ~80 ms per 32-scene step
            args.steps = 20
        if args.depth == 8:
            args.depth = 3
        if args.cpu_scenes == 8:    # ~20 s of CPU work per 65536-point scene
            args.cpu_scenes = 1
    if args.impl == "reference":
        run_reference_arm(args, cfg, n_points, batch)
    else:
        run_gpu_arm(args, cfg, n_points, batch)


if __name__ == "__main__":
    main()
