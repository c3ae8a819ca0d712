"""TEST INFRASTRUCTURE — builds the *unmodified reference* into oracle/_ref/.

Compiles, from the sources where they lie under /root/reference (never copied
into this repo), for sm_100a:

  oracle/_ref/libpdanet_ref_pointnet2.so
      PB/src/{sampling_gpu,ball_query_gpu,group_points_gpu,interpolate_gpu}.cu and
      pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu verbatim
      + oracle/ref_binding.cu (our C-ABI over the reference launchers; the
      reference's own .cpp wrappers need <THC/THC.h>, absent from torch 2.11).
  oracle/_ref/iou3d_nms_cuda.so
      the reference's whole iou3d_nms extension, as-is
      (IOU/src/{iou3d_nms_api,iou3d_nms,iou3d_cpu}.cpp + iou3d_nms_kernel.cu),
      i.e. the reference's own pybind module `iou3d_nms_cuda`
      (nms_gpu, nms_normal_gpu, boxes_overlap_bev_gpu, boxes_iou_bev_gpu,
      boxes_iou_bev_cpu).

  oracle/_ref/pyref/pcdet/...
      byte-for-byte copies of the reference PYTHON files of the path (PB/pointnet2_utils.py, pointnet2_modules.py,
      PointFormer.py, backbones_3d/IASSD_backbone.py, iou3d_nms/iou3d_nms_utils.py, model_utils/model_nms_utils.py), so
      that a GPU test can run the UNMODIFIED reference modules on libpdab.so (tests/test_gpu_reference_python.py).

oracle/_ref/ is git-ignored but NOT gpurun-ignored: the built .so files travel
to the GPU box, where /root/reference does not exist.  Nothing here runs the
reference's own build system (setup.py); it is plain nvcc / g++ on the files.

Usage:  python oracle/build_ref.py [--force]
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
REF = Path(os.environ.get("PDANET_REFERENCE", "/root/reference"))
PB = REF / "pcdet/ops/pointnet2/pointnet2_batch/src"
IOU = REF / "pcdet/ops/iou3d_nms/src"
ROI = REF / "pcdet/ops/roiaware_pool3d/src"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _torch_paths():
    import torch  # noqa: F401  (only to locate headers/libs)
    from torch.utils import cpp_extension as ce
    import logging
    logging.disable(logging.WARNING)
    inc = ce.include_paths()
    lib = ce.library_paths()
    return inc, lib


def _run(cmd):
    print("+", " ".join(str(c) for c in cmd), flush=True)
    subprocess.check_call([str(c) for c in cmd])


PY_FILES = [
    "pcdet/ops/pointnet2/pointnet2_batch/pointnet2_utils.py",
    "pcdet/ops/pointnet2/pointnet2_batch/pointnet2_modules.py",
    "pcdet/ops/pointnet2/pointnet2_batch/PointFormer.py",
    "pcdet/models/backbones_3d/IASSD_backbone.py",
    "pcdet/ops/iou3d_nms/iou3d_nms_utils.py",
    "pcdet/models/model_utils/model_nms_utils.py",
]


def stage_python() -> bool:
    """Copies the reference's Python files of the path into oracle/_ref/pyref (git-ignored; travels to the GPU box)."""
    import shutil
    dst_root = OUT / "pyref"
    if not all((REF / f).exists() for f in PY_FILES):
        return all((dst_root / f).exists() for f in PY_FILES)
    for f in PY_FILES:
        dst = dst_root / f
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(REF / f, dst)
    return True


def reference_available() -> bool:
    return (PB / "sampling_gpu.cu").exists() and (IOU / "iou3d_nms_kernel.cu").exists()


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref holds both libraries (built now or before)."""
    lib_pn = OUT / "libpdanet_ref_pointnet2.so"
    lib_iou = OUT / "iou3d_nms_cuda.so"
    OUT.mkdir(exist_ok=True)
    stage_python()
    if lib_pn.exists() and lib_iou.exists() and not force:
        return True
    if not reference_available():
        print(f"[build_ref] reference tree not found at {REF}; using prebuilt files only")
        return lib_pn.exists() and lib_iou.exists()

    OUT.mkdir(exist_ok=True)
    obj = OUT / "obj"
    obj.mkdir(exist_ok=True)
    inc, libdirs = _torch_paths()
    pyinc = sysconfig.get_paths()["include"]
    incs = [f"-I{p}" for p in inc] + [f"-I{pyinc}"]
    common_defs = ["-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=1"]
    nv = [NVCC, *ARCH, "-O3", "-std=c++17", "--compiler-options", "-fPIC",
          "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
          "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
          "--expt-relaxed-constexpr", *common_defs, *incs]
    gxx = ["g++", "-O3", "-std=c++17", "-fPIC", *common_defs,
           "-DTORCH_EXTENSION_NAME=iou3d_nms_cuda", *incs, "-I/usr/local/cuda/include"]

    jobs = []
    for name in ("sampling_gpu", "ball_query_gpu", "group_points_gpu", "interpolate_gpu"):
        jobs.append([*nv, f"-I{PB}", "-c", PB / f"{name}.cu", "-o", obj / f"pb_{name}.o"])
    jobs.append([*nv, "-c", ROI / "roiaware_pool3d_kernel.cu", "-o", obj / "roi_kernel.o"])
    jobs.append([*nv, "-c", HERE / "ref_binding.cu", "-o", obj / "ref_binding.o"])
    jobs.append([*nv, "-DTORCH_EXTENSION_NAME=iou3d_nms_cuda", f"-I{IOU}", "-c",
                 IOU / "iou3d_nms_kernel.cu", "-o", obj / "iou_kernel.o"])
    for name in ("iou3d_nms_api", "iou3d_nms", "iou3d_cpu"):
        jobs.append([*gxx, f"-I{IOU}", "-c", IOU / f"{name}.cpp", "-o", obj / f"iou_{name}.o"])
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        list(ex.map(_run, jobs))

    _run([NVCC, *ARCH, "-shared", "-o", lib_pn,
          obj / "pb_sampling_gpu.o", obj / "pb_ball_query_gpu.o", obj / "pb_group_points_gpu.o",
          obj / "pb_interpolate_gpu.o", obj / "roi_kernel.o", obj / "ref_binding.o", "-lcudart"])
    ldirs = [f"-L{p}" for p in libdirs]
    rpaths = [f"-Wl,-rpath,{p}" for p in libdirs]
    _run(["g++", "-shared", "-o", lib_iou,
          obj / "iou_kernel.o", obj / "iou_iou3d_nms_api.o", obj / "iou_iou3d_nms.o", obj / "iou_iou3d_cpu.o",
          *ldirs, *rpaths, "-L/usr/local/cuda/lib64", "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python",
          "-lc10_cuda", "-ltorch_cuda", "-lcudart"])
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ok = build(ap.parse_args().force)
    print("[build_ref] ok" if ok else "[build_ref] reference libraries NOT available")
    sys.exit(0 if ok else 1)
