"""Loader for libpdab.so — the C-ABI product library (include/pdab.h).

There is no fallback of any kind: if the library is missing or a call fails the
caller gets an exception.  `build()` compiles it in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libpdab.so"
CSRC = _PKG / "csrc"

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/pdab.h one to one
_SIGNATURES = {
    "pdab_version": (C.c_char_p, []),
    "pdab_error_string": (C.c_char_p, [_i]),
    "pdab_fps": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "pdab_fps_with_dist": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "pdab_gather_points": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pdab_gather_points_grad": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pdab_ball_query": (_i, [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp]),
    "pdab_ball_query_grid_workspace_bytes": (_sz, [_i, _i, _i]),
    "pdab_ball_query_grid": (_i, [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "pdab_ball_query_dilated": (_i, [_i, _i, _i, _f, _f, _i, _vp, _vp, _vp, _vp]),
    "pdab_group_points": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pdab_group_points_grad": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pdab_segment_sum_grad": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pdab_three_nn": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pdab_three_interpolate": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pdab_three_interpolate_grad": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pdab_points_in_boxes": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "pdab_topk_ctr": (_i, [_i, _i, _i, _i, _vp, _vp, _vp]),
    "pdab_pda_group": (_i, [_i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pdab_pda_group_tokens": (_i, [_i, _i, _i, _i, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pdab_pda_assemble_ln_split": (_i, [C.c_longlong, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp]),
    "pdab_pda_encode_param_floats": (_sz, [_i]),
    "pdab_pda_encode_ln": (_i, [_i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp]),
    "pdab_group_attention": (_i, [C.c_longlong, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pdab_group_attention_h": (_i, [C.c_longlong, _i, _i, _i, _vp, _vp, _vp]),
    "pdab_sa_fused": (_i, [_i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "pdab_sa_fused_pair": (_i, [_i, _i, _i, _i, _f, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pdab_sa_fused_pair_h": (_i, [_i, _i, _i, _i, _f, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pdab_sa_grid_workspace_bytes": (_sz, [_i, _i]),
    "pdab_tc_linear": (_i, [C.c_longlong, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _f, _i, _vp, _i, _vp]),
    "pdab_tc_linear_h": (_i, [C.c_longlong, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _f, _i, _vp, _vp, _i,
                               _i, _vp]),
    "pdab_tc_ffn_h": (_i, [C.c_longlong, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "pdab_tc_sa_gather_linear_h": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "pdab_pda_encode_ln_h": (_i, [_i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp]),
    "pdab_set_persistent_ctas": (_i, [_i]),
    "pdab_set_fps_max_cluster": (_i, [_i]),
    "pdab_set_cta_pairs": (_i, [_i]),
    "pdab_tc_packed_floats": (_sz, [_i, _i, _i, _i]),
    "pdab_tc_pack_weights": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pdab_tc_sa_gather_linear": (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "pdab_nms_workspace_bytes": (_sz, [_i]),
    "pdab_nms_device": (_i, [_vp, _i, _f, _vp, _vp, _vp, _vp]),
    "pdab_nms_batched": (_i, [_vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "pdab_nms_host": (_i, [_vp, _i, _f, _vp, _i, _vp]),
    "pdab_post_front": (_i, [_i, _i, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pdab_post_select": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pdab_boxes_overlap_bev": (_i, [_i, _vp, _i, _vp, _vp, _vp]),
    "pdab_boxes_iou_bev": (_i, [_i, _vp, _i, _vp, _vp, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

PDAB_EUNSUPPORTED = -2


class PdabError(RuntimeError):
    def __init__(self, fn: str, code: int, text: str):
        super().__init__(f"{fn} failed with code {code}: {text}")
        self.code = code


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile pdanet_b200/csrc/*.cu into pdanet_b200/libpdab.so (nvcc, sm_100a, -lineinfo)."""
    if force:
        subprocess.check_call(["make", "-C", str(CSRC), "clean"], stdout=subprocess.DEVNULL)
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", str(CSRC), f"-j{min(8, os.cpu_count() or 2)}"], stdout=out)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """The loaded library; raises if it has not been built (no CPU or eager fallback exists)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). pdanet_b200 has no fallback path.")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(fn: str, code: int) -> int:
    if code != 0:
        raise PdabError(fn, code, lib().pdab_error_string(code).decode())
    return code


# kernels launched by one call of each entry point (for bench.py's `gpu_launches` count)
KERNELS_PER_CALL = {"pdab_nms_device": 2, "pdab_nms_batched": 2, "pdab_nms_host": 2, "pdab_sa_fused_pair": 2, "pdab_sa_fused_pair_h": 2, "pdab_ball_query_grid": 3}

launch_counts: dict = {}      # entry point -> number of kernels launched through `call`
_timing = None                # when enabled: entry point -> list of (start_event, end_event)


def enable_timing(on: bool = True):
    """Bracket every C-ABI call with CUDA events on the launching stream (bench.py roofline leg)."""
    global _timing
    _timing = {} if on else None


def timings_ms() -> dict:
    """'entry_point(sizes...)' -> list of per-call durations in ms (the events must have completed)."""
    out = {}
    for name, pairs in (_timing or {}).items():
        out[name] = [a.elapsed_time(b) for a, b in pairs]
    return out


def call(fn: str, *args) -> int:
    """Invoke an int-returning entry point and raise PdabError on a non-zero code.
    By convention the last argument is the cudaStream_t the kernels go to."""
    f = getattr(lib(), fn)
    launch_counts[fn] = launch_counts.get(fn, 0) + KERNELS_PER_CALL.get(fn, 1)
    if _timing is None:
        return check(fn, f(*args))
    import torch
    stream = torch.cuda.ExternalStream(args[-1]) if args[-1] else torch.cuda.current_stream()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(stream)
    rc = f(*args)
    end.record(stream)
    key = fn + str(tuple(a for a in args[:7] if isinstance(a, int) and 0 <= a < (1 << 24)))  # fn + leading sizes
    _timing.setdefault(key, []).append((start, end))
    return check(fn, rc)
