"""ctypes binding of libdemo_b200.so (the C ABI declared in include/demo_b200.h).

The shared library is built in-tree by ``demo2_b200/csrc/Makefile`` (see
``__graft_entry__.build``).  There is NO CPU fallback: if the library is missing
or no sm_100 device is present every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdemo_b200.so")

# flags (mirror include/demo_b200.h)
DIST_SQ, DIST_SQRT, DIST_COS_SIM, DIST_COS_DIST = 0, 1, 2, 3
FLAG_L2NORM, FLAG_TRIPLET_NORM, FLAG_SIMT, FLAG_HOST_INPUT = 0x10, 0x20, 0x40, 0x80

c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int)
c_u32p = C.POINTER(C.c_uint)
c_i64p = C.POINTER(C.c_int64)
vp, sz, i32, i64 = C.c_void_p, C.c_size_t, C.c_int, C.c_int64

# name -> (restype, argtypes); every symbol declared in include/demo_b200.h
SIGNATURES = {
    "demo_last_error": (C.c_char_p, []),
    "demo_version": (i32, []),
    "demo_device_ok": (i32, []),
    "demo_sqdist_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "demo_sqdist_f32": (i32, [vp, vp, i32, i32, i32, i64, i64, vp, i64, i32, vp, vp, vp, vp, sz, vp]),
    "demo_plan_bytes": (sz, [i32, i32]),
    "demo_eval_plan": (i32, [vp, vp, i32, i32, vp, sz, c_i64p, vp]),
    "demo_plan_pointers": (i32, [vp, sz, i32, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "demo_plan_info": (i32, [vp, sz, i32, i32, C.POINTER(vp)]),
    "demo_eval_workspace_bytes": (sz, [i32, i32, i32, i64]),
    "demo_eval_workspace_bytes_ex": (sz, [i32, i32, i32, i64, i32]),
    "demo_eval_matrix_workspace_bytes": (sz, [i32, i32, i64]),
    "demo_eval_prepare": (i32, [vp, i32, i32, i64, i32, i32, i32, i32, vp, sz, i32, i32, i64, vp, sz, vp, vp]),
    "demo_eval_extract": (i32, [i32, i32, i32, vp, vp, i32, vp, vp, sz, i64, vp, sz, vp, vp, vp, i32, i32, vp]),
    "demo_eval_count_range": (i32, [i32, i32, i32, i64, vp, sz, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    "demo_eval_records": (i32, [vp, vp, i32, i32, i32, i64, i64, i32, vp, vp, i32, vp, sz, i64, vp, sz,
                                vp, vp, vp, vp, vp, vp]),
    "demo_build_thresholds": (i32, [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]),
    "demo_eval_count": (i32, [i32, i32, i32, i64, vp, sz, vp, vp, vp, vp, vp, i32, i32, vp]),
    "demo_cmc_map_finalize": (i32, [vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
    "demo_eval_features": (i32, [vp, vp, i32, i32, i32, i64, i64, i32, vp, vp, vp, sz, i64, i32, i32, vp, sz,
                                 vp, vp, vp, vp, vp, vp, vp, vp]),
    "demo_eval_matrix": (i32, [vp, i32, i32, i64, vp, vp, vp, sz, i64, i32, i32, vp, sz, vp, vp, vp, vp, vp, vp]),
    "demo_eval_ws_pointers": (i32, [vp, sz, i32, i32, i32, i64] + [C.POINTER(vp)] * 13),
    "demo_rerank_workspace_bytes": (sz, [i32, i32, i32, i32, i32]),
    "demo_rerank": (i32, [vp, i32, i32, i32, i64, i32, i32, i32, C.c_double, vp, i64, i32, vp, i64, vp, vp, sz, vp]),
    "demo_rerank_matrix": (i32, [vp, i64, i32, i32, i32, i32, C.c_double, vp, i64, vp, sz, vp]),
    "demo_topk_rows": (i32, [vp, i32, i32, i64, i32, vp, vp, vp]),
    "demo_rerank_dims": (i32, [i32, i32, i32, c_i32p, c_i32p, c_i32p]),
    "demo_rerank_shard_workspace_bytes": (sz, [i32, i32, i32, i32, i32, i32]),
    "demo_rerank_shard_topk": (i32, [vp, i32, i32, i32, i64, i32, i32, i32, i32, i32, i32, vp, vp, vp, sz, vp]),
    "demo_rerank_shard_krecip": (i32, [i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp]),
    "demo_rerank_shard_expand": (i32, [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "demo_rerank_shard_jaccard": (i32, [i32, i32, i32, i32, i32, C.c_double, i32, i32, i32, vp, vp, vp, vp, i64, vp,
                                        sz, vp]),
    "demo_triplet_workspace_bytes": (sz, [i32, i32]),
    "demo_triplet_hard_fwd": (i32, [vp, i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "demo_triplet_hard_bwd": (i32, [vp, i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
    "demo_hard_example_mining": (i32, [vp, i32, i64, vp, vp, vp, vp, vp, vp, vp]),
    "demo_triplet_loss_workspace_bytes": (sz, []),
    "demo_triplet_loss_max_batch": (i32, []),
    "demo_triplet_loss_fwd": (i32, [vp, i32, i32, i32, i64, vp, i32, C.c_float, C.c_float, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "demo_triplet_loss_bwd": (i32, [vp, i32, i32, i32, i64, C.c_float, C.c_float, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
    "demo_comm_available": (i32, []),
    "demo_comm_nccl_version": (i32, []),
    "demo_comm_unique_id": (i32, [vp]),
    "demo_comm_init": (i32, [i32, i32, vp]),
    "demo_comm_destroy": (i32, []),
    "demo_comm_info": (i32, [c_i32p, c_i32p]),
    "demo_comm_check": (i32, []),
    "demo_comm_all_gather": (i32, [vp, vp, sz, vp]),
    "demo_comm_all_reduce_sum_u32": (i32, [vp, sz, vp]),
    "demo_comm_broadcast": (i32, [vp, sz, i32, vp]),
}

_lib = None


class DemoError(RuntimeError):
    pass


def load():
    """Load the shared library (no device needed) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DemoError(
            "libdemo_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C demo2_b200/csrc`; there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def require_device():
    lib = load()
    import torch
    if not torch.cuda.is_available() or not lib.demo_device_ok():
        raise DemoError("demo2_b200 needs an NVIDIA sm_100 (B200) device; there is no CPU fallback")
    return lib


def check(rc: int):
    if rc != 0:
        raise DemoError("libdemo_b200 error %d: %s" % (rc, load().demo_last_error().decode()))


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
