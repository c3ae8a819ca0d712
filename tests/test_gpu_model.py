"""GPU model-level parity: the PDA-SSD modules running on the CUDA ops (fused kernels included) against the SAME
module code running on the CPU oracle ops with identical weights.  Layer by layer with the oracle's activations as
inputs ("teacher forcing"), so a last-ulp difference in one layer cannot flip a top-k pick in the next and hide or
fake an error.  Sampled indices and gathered coordinates: exact.  Features: 1e-3 relative (north_star)."""
import pytest
import torch

from oracle import torch_ops
from pdanet_b200.config import load_config
from pdanet_b200.iassd import build_model
from pdanet_b200.synthetic import make_batch

pytestmark = pytest.mark.gpu
RTOL = 1e-3


def close(a, b, what):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    scale = b.abs().max().item() + 1e-12
    err = (a - b).abs().max().item()
    assert err <= RTOL * scale, f"{what}: max abs err {err:.3e} vs scale {scale:.3e}"


@pytest.fixture(scope="module")
def models():
    torch.backends.cudnn.allow_tf32 = False          # compare in full fp32 against the CPU oracle path
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = load_config("kitti")
    torch.manual_seed(0)
    gpu = build_model(cfg).cuda().eval()
    cpu = build_model(cfg, ops=torch_ops, nms_utils=torch_ops.nms_utils, batched_post_processing=False).eval()
    # non-trivial BatchNorm statistics so that BN folding is actually exercised
    g = torch.Generator().manual_seed(1)
    sd = gpu.state_dict()
    for k, v in sd.items():
        if k.endswith("running_mean"):
            sd[k] = (torch.randn(v.shape, generator=g) * 0.1).to(v)
        elif k.endswith("running_var"):
            sd[k] = (0.5 + torch.rand(v.shape, generator=g)).to(v)
    gpu.load_state_dict(sd)
    cpu.load_state_dict({k: v.cpu() for k, v in sd.items()})
    return cfg, gpu, cpu


def test_backbone_layers_teacher_forced(models):
    cfg, gpu, cpu = models
    B = 2
    batch = make_batch(B, 16384, cfg.POINT_CLOUD_RANGE, duplicate_frac=0.05)
    with torch.no_grad():
        cd = cpu.backbone_3d({"batch_size": B, "points": batch["points"].clone()})
    bb_g, bb_c = gpu.backbone_3d, cpu.backbone_3d
    xyzs, feats = cd["encoder_xyz"], cd["encoder_features"]
    cls_pred = None
    with torch.no_grad():
        for i, (mg, mc) in enumerate(zip(bb_g.SA_modules, bb_c.SA_modules)):
            src = bb_c.layer_inputs[i]
            x_in, f_in = xyzs[src], feats[src]
            if bb_c.layer_types[i] == "SA_Layer":
                ctr = xyzs[bb_c.ctr_idx_list[i]] if bb_c.ctr_idx_list[i] != -1 else None
                want = mc(x_in, f_in, cls_pred, ctr_xyz=ctr)
                got = mg(x_in.cuda(), f_in.cuda(), None if cls_pred is None else cls_pred.cuda(),
                         ctr_xyz=None if ctr is None else ctr.cuda())
                if ctr is None:
                    assert torch.equal(got[3].cpu(), want[3]), f"layer {i}: sampled indices differ"
                    assert torch.equal(got[0].cpu(), want[0]), f"layer {i}: sampled coordinates differ"
                close(got[1], want[1], f"layer {i} features")
                if want[2] is not None:
                    close(got[2], want[2], f"layer {i} class logits")
                cls_pred = want[2]
            else:
                want = mc(x_in, f_in)
                got = mg(x_in.cuda(), f_in.cuda())
                close(got[0], want[0], "vote xyz")
                close(got[3], want[3], "vote offsets")


def test_head_and_post_processing(models):
    cfg, gpu, cpu = models
    B = 3
    batch = make_batch(B, 16384, cfg.POINT_CLOUD_RANGE, first_scene=7)
    with torch.no_grad():
        cd = cpu.backbone_3d({"batch_size": B, "points": batch["points"].clone()})
        cd = cpu.point_head(cd)
        gd = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in cd.items()
              if k in ("batch_size", "centers_features", "centers", "ctr_offsets", "centers_origin")}
        gd["sa_ins_preds"], gd["sample_list_id"] = [], []
        gd = gpu.point_head(gd)
        close(gd["batch_cls_preds"], cd["batch_cls_preds"], "head class logits")
        close(gd["batch_box_preds"], cd["batch_box_preds"], "decoded boxes")
        # post-processing on IDENTICAL head outputs: reference-semantics path, batched path and the oracle agree
        feed = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in cd.items()
                if k in ("batch_size", "batch_cls_preds", "batch_box_preds", "batch_index", "cls_preds_normalized")}
        want, _ = cpu.post_processing(cd)
        got_ref, _ = gpu.post_processing(feed)
        got_bat, _ = gpu.post_processing_batched(feed)
    for s in range(B):
        for got in (got_ref[s], got_bat[s]):
            assert got["pred_boxes"].shape == want[s]["pred_boxes"].shape
            # scores tie massively with random weights; compare as sets of rows
            gb = got["pred_boxes"].cpu()
            wb = want[s]["pred_boxes"]
            key = lambda t: sorted(map(tuple, t.tolist()))
            assert key(gb) == key(wb), f"scene {s}: kept boxes differ"
            assert sorted(got["pred_labels"].cpu().tolist()) == sorted(want[s]["pred_labels"].tolist())


def test_full_forward_runs_and_is_deterministic(models):
    cfg, gpu, _ = models
    batch = make_batch(4, 16384, cfg.POINT_CLOUD_RANGE)
    pts = batch["points"].cuda()
    with torch.no_grad():
        a, _ = gpu({"batch_size": 4, "points": pts})
        b, _ = gpu({"batch_size": 4, "points": pts})
    assert len(a) == 4
    for x, y in zip(a, b):
        assert torch.equal(x["pred_boxes"], y["pred_boxes"]) and torch.equal(x["pred_scores"], y["pred_scores"])
        assert x["pred_boxes"].shape[1] == 7 and x["pred_boxes"].shape[0] <= 500


def test_scene_sharding_is_a_pure_partition(models):
    """Running scenes in two shards (as two GPUs would) reproduces the one-batch run scene by scene.  Checked up to
    the first class-aware top-k: FPS picks exact, features to rounding (cuBLAS may choose another GEMM kernel for
    another token count, and with random-init weights the class logits are near-ties, so later picks are ill-posed)."""
    from pdanet_b200.runner import shard_scenes
    cfg, gpu, _ = models
    batch = make_batch(4, 16384, cfg.POINT_CLOUD_RANGE)
    pts = batch["points"].cuda().view(4, 16384, 5)
    with torch.no_grad():
        whole = gpu.backbone_3d({"batch_size": 4, "points": pts.reshape(-1, 5).clone()})
        for rank in range(2):
            ids = shard_scenes(4, 2, rank)
            sub = pts[ids.start:ids.stop].clone()
            sub[:, :, 0] -= ids.start
            part = gpu.backbone_3d({"batch_size": len(ids), "points": sub.reshape(-1, 5)})
            for lvl in (1, 2):
                assert torch.equal(part["encoder_xyz"][lvl], whole["encoder_xyz"][lvl][ids.start:ids.stop])
            close(part["encoder_features"][1], whole["encoder_features"][1][ids.start:ids.stop], "L0 features")
            close(part["encoder_features"][2], whole["encoder_features"][2][ids.start:ids.stop], "L1 features")
            assert len(gpu({"batch_size": len(ids), "points": sub.reshape(-1, 5)})[0]) == len(ids)


@pytest.mark.parametrize("name", ["kitti", "once"])
def test_cuda_backbone_matches_reference_modules_golden(name):
    """The CUDA path (fused kernels included) against outputs of the REFERENCE's own Python modules; once = one
    65536-point scene: clustered FPS 65536 -> 16384, cell-list L0, three wide L5 scales (nsample 16 / 32 / 64)."""
    from test_host_cpu import replay_reference_backbone_golden
    from pdanet_b200 import pointnet2_utils
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    replay_reference_backbone_golden(pointnet2_utils, "cuda", rtol=RTOL, name=name)


@pytest.mark.parametrize("name", ["kitti", "once"])
def test_cuda_head_and_post_processing_match_reference_golden(name):
    """The CUDA path's head, box decode and all three post-processing flavours (`post_processing`, `_batched`, `_padded`:
    batched on-device NMS) against outputs of the REFERENCE's own IASSD_Head / decode_torch / post_processing /
    class_agnostic_nms (tests/golden/make_head_golden.py)."""
    from test_host_cpu import replay_reference_head_golden
    from pdanet_b200 import iou3d_nms_utils, pointnet2_utils
    torch.backends.cuda.matmul.allow_tf32 = False
    replay_reference_head_golden(name, pointnet2_utils, iou3d_nms_utils, "cuda", rtol=1e-4)


@pytest.mark.parametrize("tc_passes,tol", [(2, 1e-4), (4, 3e-4)])
def test_pda_fast_path_equals_reference_statement_order(models, tc_passes, tol):
    """pda_block.py (token-major, folded BN, tensor-core projections) vs the module's reference-order forward (fp32 torch),
    same inputs.  tc_passes = 2: split-bf16 products; 4: fp16 single pass with fp16 activations between the kernels and
    (hi, lo) fp16 residual streams — both far inside the 1e-3 feature bar."""
    cfg, gpu, _ = models
    g = torch.Generator().manual_seed(11)
    for layer, (N, C) in ((1, (4096, 64)), (2, (1024, 128))):
        mod = gpu.backbone_3d.SA_modules[layer]
        prev_passes, mod.tc_passes, mod._plans = mod.tc_passes, tc_passes, {}
        xyz = (torch.rand(2, N, 3, generator=g) * torch.tensor([30.0, 30.0, 2.0])).cuda()
        feats = torch.randn(2, C, N, generator=g).cuda()
        cls = torch.randn(2, N, 3, generator=g).cuda()
        with torch.no_grad():
            mod.fast_eval = True
            fast = mod(xyz, feats, cls)
            mod.fast_eval = False
            slow = mod(xyz, feats, cls)
            mod.fast_eval = True
        mod.tc_passes, mod._plans = prev_passes, {}
        assert torch.equal(fast[3], slow[3])
        scale = slow[1].abs().max().item()
        assert (fast[1] - slow[1]).abs().max().item() <= tol * scale
        assert (fast[2] - slow[2]).abs().max().item() <= tol * slow[2].abs().max().item()


@pytest.mark.parametrize("name,B,N", [("kitti", 2, 16384), ("once", 1, 65536)])
def test_no_scale_falls_back_to_the_unfused_grouper(name, B, N):
    """Eval forward of BOTH yamls at their BASELINE point counts: every SA scale — the narrow L0 pair, the PDA scales and
    all wide L5 scales, including ONCE's third one (r = 12.8, nsample 64) — runs on a fused kernel; none reaches
    `groupers[i]` + eager convolutions (tools/cfgs/once_models/PDA-SSD.yaml:26-50, PB/pointnet2_modules.py:1655-1672)."""
    cfg = load_config(name)
    torch.manual_seed(0)
    model = build_model(cfg).cuda().eval()
    hit = []
    for li, mod in enumerate(model.backbone_3d.SA_modules):
        for gi, grouper in enumerate(getattr(mod, "groupers", [])):
            grouper.register_forward_pre_hook(lambda m, a, li=li, gi=gi: hit.append((li, gi)))
    batch = make_batch(B, N, cfg.POINT_CLOUD_RANGE)
    with torch.no_grad():
        preds, _ = model({"batch_size": B, "points": batch["points"].cuda()})
    assert not hit, f"scales that fell back to the unfused grouper (layer, scale): {hit}"
    assert len(preds) == B and all(torch.isfinite(p["pred_boxes"]).all() for p in preds)
    wide = model.backbone_3d.SA_modules[5]
    assert len(wide._wide) == len(wide.groupers) == (2 if name == "kitti" else 3)


@pytest.mark.parametrize("tc_passes", [2, 4])
def test_pda_unfused_encoder_path_equals_fused_encoder(models, tc_passes):
    """pda_block.py keeps an un-fused token encoder (pdab_pda_group_tokens + position MLP GEMMs + DensityNet +
    pdab_pda_assemble_ln_split) for shapes the one-kernel encoder does not cover; with the fused encoder switched off the
    scale must give the same result (both run the same transformer afterwards)."""
    cfg, gpu, _ = models
    g = torch.Generator().manual_seed(13)
    for layer, (N, C) in ((1, (4096, 64)), (2, (1024, 128))):
        mod = gpu.backbone_3d.SA_modules[layer]
        xyz = (torch.rand(2, N, 3, generator=g) * torch.tensor([30.0, 30.0, 2.0])).cuda()
        feats = torch.randn(2, C, N, generator=g).cuda()
        cls = torch.randn(2, N, 3, generator=g).cuda()
        prev = mod.tc_passes
        try:
            mod.tc_passes, mod._plans = tc_passes, {}
            with torch.no_grad():
                fused = mod(xyz, feats, cls)
                for plan in mod._plans.values():
                    assert plan.fused_encode
                    plan.fused_encode = False
                unfused = mod(xyz, feats, cls)
        finally:
            mod.tc_passes, mod._plans = prev, {}
        assert torch.equal(fused[3], unfused[3])
        scale = fused[1].abs().max().item()
        assert (fused[1] - unfused[1]).abs().max().item() <= 2e-4 * scale


def test_pda_group_tokens_matches_channel_major_grouper():
    from pdanet_b200 import pointnet2_utils as ops
    g = torch.Generator().manual_seed(5)
    for B, C, N, M, r, ns in [(2, 64, 4096, 1024, 0.8, 16), (2, 128, 1024, 300, 4.8, 32)]:
        xyz = (torch.rand(B, N, 3, generator=g) * torch.tensor([30.0, 30.0, 2.0])).cuda()
        feats = torch.randn(B, C, N, generator=g).cuda()
        new_xyz = xyz[:, :M].contiguous()
        ref, ref_idx = ops.pda_group(r, ns, xyz, new_xyz, feats, return_idx=True)          # (B, 7+C, M, ns)
        tok, idx = ops.pda_group_tokens(r, ns, xyz, new_xyz, feats.transpose(1, 2).contiguous(), return_idx=True)
        assert torch.equal(idx, ref_idx)
        ref = ref.permute(0, 2, 3, 1)                                                       # (B, M, ns, 7+C)
        assert torch.equal(tok[..., 0:7], ref[..., 0:7])
        assert torch.equal(tok[..., 8:], ref[..., 7:])
        assert (tok[..., 7] == 0).all()


def test_pda_encode_ln_matches_unfused_chain():
    """Fused token encoder (csrc/pda_encode.cu) vs a float64 torch statement of the same chain built on the exact
    grouper kernel's indices: ball query order, density, direction, position MLP, DensityNet, token cat, LayerNorm.
    Ragged M (not a multiple of the CTA's centre count), empty balls (far-away centres) and both widths."""
    from pdanet_b200 import pointnet2_utils as ops
    g = torch.Generator().manual_seed(7)
    for B, C, N, M, r, ns in [(2, 64, 4096, 1000, 0.8, 16), (3, 64, 2048, 77, 1.6, 32), (2, 128, 1024, 300, 4.8, 32),
                              (16, 128, 1024, 512, 2.4, 16)]:
        xyz = (torch.rand(B, N, 3, generator=g) * torch.tensor([30.0, 30.0, 2.0])).cuda()
        feats_t = torch.randn(B, N, C, generator=g).cuda()
        new_xyz = xyz[:, :M].clone()
        new_xyz[:, -1] = xyz[:, 0]                     # an empty ball (above the cloud, farther than r from every
        new_xyz[:, -1, 2] = 2.0 + 1.01 * r             # point): groups point 0, the pre-zeroed idx row
        glob = torch.randn(B * M, C, generator=g).cuda()
        H = C // 2
        w1, b1 = torch.randn(H, 12, generator=g).cuda() * 0.3, torch.randn(H, generator=g).cuda() * 0.1
        w2, b2 = torch.randn(C, H, generator=g).cuda() * 0.2, torch.randn(C, generator=g).cuda() * 0.1
        dens = [(torch.randn(16, 1, generator=g).cuda(), torch.rand(16, generator=g).cuda()),
                (torch.randn(8, 16, generator=g).cuda() * 0.5, torch.rand(8, generator=g).cuda()),
                (torch.randn(1, 8, generator=g).cuda().abs(), torch.rand(1, generator=g).cuda())]
        gamma, beta = torch.rand(4 * C, generator=g).cuda() + 0.5, torch.randn(4 * C, generator=g).cuda() * 0.1
        params = ops.pda_encode_params(w1, b1, w2, b2, dens, gamma, beta)
        y = ops.pda_encode_ln(r, ns, xyz, new_xyz, feats_t, glob, params, 1e-5)

        X, idx = ops.pda_group_tokens(r, ns, xyz, new_xyz, feats_t, return_idx=True)
        assert (idx[:, -1] == 0).all()
        T = B * M * ns
        d = lambda t: t.double()
        Xd = d(X.view(T, 8 + C))
        nbr, den, direction, f = Xd[:, 0:3], Xd[:, 3], Xd[:, 4:7], Xd[:, 8:]
        ctr = d(new_xyz).reshape(B * M, 1, 3).expand(B * M, ns, 3).reshape(T, 3)
        rppe = torch.cat([ctr, nbr, ctr - nbr, direction], dim=1)
        pos = torch.relu(torch.relu(rppe @ d(w1).t() + d(b1)) @ d(w2).t() + d(b2))
        dg = den.view(B * M, ns)
        sc = (dg / dg.max(dim=1, keepdim=True)[0]).reshape(T, 1)
        for w, b in dens:
            sc = torch.relu(sc @ d(w).t() + d(b))
        tok = torch.cat([pos, f * sc, f, d(glob).repeat_interleave(ns, dim=0)], dim=1)
        want = torch.nn.functional.layer_norm(tok, (4 * C,), d(gamma), d(beta), 1e-5)
        err = (d(y) - want).abs().max().item()
        assert err <= 2e-5 * want.abs().max().item(), (B, C, N, M, ns, err)


@pytest.mark.parametrize("graphs", [True, False])
def test_pipelined_runner_equals_sequential_runner(graphs):
    """ScenePipeline (2 batches in flight, CUDA graphs, padded sync-free post-processing) returns exactly what
    SceneRunner.infer returns batch by batch — pipelining is scheduling, not arithmetic."""
    from pdanet_b200 import _lib
    from pdanet_b200.runner import ScenePipeline, SceneRunner
    cfg = load_config("kitti")
    B, N = 2, 16384
    runner = SceneRunner(cfg, device="cuda:0", batch_size=B, num_points=N, seed=0)
    batches = [make_batch(B, N, cfg.POINT_CLOUD_RANGE, first_scene=10 * k)["points"] for k in range(5)]
    want = [runner.infer(b) for b in batches]
    pipe = ScenePipeline(runner, depth=2, graphs=graphs, warm_points=batches[0])
    got = pipe.run(batches)
    got2 = pipe.run(batches[::-1])[::-1]       # slots reused in another order: no state leaks between replays
    if graphs:
        assert pipe.launches_per_step > 20
    for res in (got, got2):
        assert len(res) == len(want)
        for w_b, g_b in zip(want, res):
            assert len(w_b) == len(g_b) == B
            for w, g in zip(w_b, g_b):
                assert w["pred_boxes"].shape == g["pred_boxes"].shape
                assert torch.equal(w["pred_labels"], g["pred_labels"])
                assert torch.allclose(w["pred_boxes"], g["pred_boxes"], rtol=1e-5, atol=1e-5)
                assert torch.allclose(w["pred_scores"], g["pred_scores"], rtol=1e-5, atol=1e-6)


def test_train_mode_forward_backward_reaches_every_parameter():
    """Train mode: the whole path differentiates — native ops through their gradient entry points,
    the modules on autograd, BatchNorm on batch statistics, torch layers under bf16 autocast — and a surrogate objective
    (tools/bench_train.py; the reference's target assignment / losses are not ported) gives every parameter a finite
    gradient, twice the same (deterministic backward)."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    from bench_train import surrogate_loss
    cfg = load_config("kitti")
    torch.manual_seed(0)
    model = build_model(cfg).cuda().train()
    batch = make_batch(2, 16384, cfg.POINT_CLOUD_RANGE)["points"].cuda()

    def grads():
        model.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model({"batch_size": 2, "points": batch})
            loss = surrogate_loss(out)
        loss.backward()
        return {k: p.grad.clone() for k, p in model.named_parameters() if p.requires_grad}, float(loss)

    g1, l1 = grads()
    missing = [k for k, p in model.named_parameters() if p.requires_grad and k not in g1]
    assert not missing
    dead = [k for k, g in g1.items() if g is None or not torch.isfinite(g).all()]
    assert not dead, dead[:5]
    assert sum(float(g.abs().sum()) > 0 for g in g1.values()) > 0.9 * len(g1)
    with pytest.raises(NotImplementedError):
        model.point_head.get_loss()


@pytest.mark.parametrize("name", ["kitti", "once"])
def test_fused_post_processing_equals_torch_statement(name):
    """`post_processing_padded` through csrc/post.cu (two kernels around the batched NMS) == the same steps written in torch,
    bit for bit: scores (same sigmoid expression), labels, the (score desc, index asc) order, counts, kept rows and the
    zero padding.  Inputs with massive score ties (quantised logits), saturated logits and below-threshold rows."""
    cfg = load_config(name)
    torch.manual_seed(0)
    model = build_model(cfg).cuda().eval()
    B = 4
    M = 256 if name == "kitti" else 1024
    nc = len(cfg.CLASS_NAMES)
    g = torch.Generator().manual_seed(31)
    lo, hi = torch.tensor(cfg.POINT_CLOUD_RANGE[:3]), torch.tensor(cfg.POINT_CLOUD_RANGE[3:])
    for quant in (None, 0.25):
        logits = torch.randn(B * M, nc, generator=g) * 3
        if quant:
            logits = torch.round(logits / quant) * quant
        logits[::17] = 40.0            # saturated sigmoid: exact ties at 1.0
        logits[5::23] = -9.0           # below both score thresholds
        padded = torch.zeros(B * M, 4 * ((nc + 3) // 4))
        padded[:, :nc] = logits        # the head hands over a column slice of a padded matrix
        ctr = lo + (hi - lo) * torch.rand(B * M, 3, generator=g)
        ctr[1::2] = ctr[0::2] + 0.4 * torch.randn(B * M // 2, 3, generator=g)
        boxes = torch.cat([ctr, 1.0 + 2.0 * torch.rand(B * M, 3, generator=g), 6.28 * torch.rand(B * M, 1, generator=g) - 3.14], dim=1)
        feed = {"batch_size": B, "batch_cls_preds": padded.cuda()[:, :nc], "batch_box_preds": boxes.cuda(),
                "cls_preds_normalized": False}
        with torch.no_grad():
            model.fused_post_processing = True
            a = model.post_processing_padded(dict(feed))
            model.fused_post_processing = False
            b = model.post_processing_padded(dict(feed))
            model.fused_post_processing = True
        assert set(a) == set(b)
        for k in a:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
            assert torch.equal(a[k], b[k]), f"{name} quant={quant}: {k} differs"
        assert int(a["num"].min()) > 0
