"""TEST INFRASTRUCTURE — imports the REFERENCE's own Python modules of the path, unchanged, over the CPU oracle.

`import_reference(root)` loads PB/pointnet2_utils.py, PB/pointnet2_modules.py, PB/PointFormer.py and
backbones_3d/IASSD_backbone.py from `root` — /root/reference in the build container, or the byte-for-byte copies that
oracle/build_ref.py stages under oracle/_ref/pyref (what exists on the GPU box) — through stub packages: third-party
imports the path never uses (open3d / matplotlib visualiser, torchsparse cluster model) are empty modules, the native
module `pointnet2_batch_cuda` is the CPU oracle (same pybind names and arities), and the six `Function.apply` symbols whose
Python bodies allocate with `torch.cuda.*Tensor` (PB/pointnet2_utils.py:25-26,83,200,246) are replaced by allocation-only
CPU equivalents.  Used by tests/golden/make_module_golden.py (golden vectors) and by bench.py's CPU arm (the reference
backbone modules timed on the host cores).  Never imported by the product.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
PYREF = HERE / "_ref" / "pyref"
REFERENCE = Path("/root/reference")


def reference_root():
    """The reference tree if present (build container), else the staged copies, else None."""
    probe = "pcdet/ops/pointnet2/pointnet2_batch/pointnet2_modules.py"
    for root in (REFERENCE, PYREF):
        if (root / probe).exists():
            return root
    return None


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [str(path)]
    sys.modules[name] = m
    return m


def import_reference(root=None):
    """-> (pointnet2_utils, pointnet2_modules, IASSD_backbone) modules of the reference, running on the CPU oracle."""
    import oracle
    from oracle import torch_ops
    root = Path(root) if root is not None else reference_root()
    if root is None:
        raise FileNotFoundError("neither /root/reference nor oracle/_ref/pyref holds the reference's Python modules")
    _pkg("pcdet", root / "pcdet")
    _pkg("pcdet.ops", root / "pcdet/ops")
    _pkg("pcdet.ops.pointnet2", root / "pcdet/ops/pointnet2")
    pb = "pcdet.ops.pointnet2.pointnet2_batch"
    _pkg(pb, root / "pcdet/ops/pointnet2/pointnet2_batch")
    sys.modules[pb + ".pointnet2_batch_cuda"] = oracle           # same pybind names/arity, CPU tensors
    sys.modules[pb + ".semantic_view"] = types.ModuleType("semantic_view")  # open3d visualiser, unused
    _pkg("pcdet.models", root / "pcdet/models")
    _pkg("pcdet.models.backbones_3d", root / "pcdet/models/backbones_3d")
    _pkg("pcdet.models.backbones_3d.cluster", root / "pcdet/models/backbones_3d/cluster")
    spv = types.ModuleType("spvnas_cluster")
    spv.SPVNAS = None                                             # torchsparse model, imported but never built
    sys.modules["pcdet.models.backbones_3d.cluster.spvnas_cluster"] = spv

    utils = importlib.import_module(pb + ".pointnet2_utils")
    # the reference Functions allocate outputs with torch.cuda.*Tensor: swap in CPU-allocating equivalents
    utils.furthest_point_sample = utils.farthest_point_sample = torch_ops.furthest_point_sample
    utils.furthest_point_sample_with_dist = torch_ops.furthest_point_sample_with_dist
    utils.gather_operation = torch_ops.gather_operation
    utils.grouping_operation = torch_ops.grouping_operation
    utils.ball_query = torch_ops.ball_query
    utils.ball_query_dilated = torch_ops.ball_query_dilated
    modules = importlib.import_module(pb + ".pointnet2_modules")
    backbone = importlib.import_module("pcdet.models.backbones_3d.IASSD_backbone")
    return utils, modules, backbone
