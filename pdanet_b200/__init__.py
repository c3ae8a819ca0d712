"""pdanet_b200 — B200-native (sm_100a) PDA-SSD point-backbone hot path.

Hand-written CUDA kernels behind a C ABI (include/pdab.h -> pdanet_b200/libpdab.so) and the host-side
mirror of the reference's operator interface for this path:

  pointnet2_batch_cuda, iou3d_nms_cuda      drop-ins for the reference's two pybind modules
  pointnet2_utils, iou3d_nms_utils          the reference's Python op API (same names / layouts)
  pointnet2_modules, iassd_backbone,
  iassd_head, iassd                         the PDA-SSD modules that run on top of the ops

There is no CPU or eager fallback: importing works anywhere, calling an op needs libpdab.so and a GPU.
"""
__version__ = "0.1.0"
